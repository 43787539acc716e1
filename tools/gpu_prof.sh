#!/bin/bash
# ncu captures for the kernels named in $2.. (tag $1)
TAG=$1; shift
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --batches-per-step 8 --bank-size 2000 --no-cpu-baseline --e2e-steps 1"
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_${k}_$TAG $SMALL > /dev/null 2>&1
done
ls -la gpurun_out | tail -5
