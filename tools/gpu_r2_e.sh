#!/bin/bash
# traces + ncu of a mixer variant: usage tools/gpu_r2_e.sh <variant>
mkdir -p gpurun_out
v=$1
export ADTFE_LIB=$PWD/adt_str_b200/libadtfe_$v.so
python tools/render_sweep.py --chunks 64,16,4 --render-only --trace gpurun_out 2>&1 | tail -20
SMALL="python bench.py --steps 2 --warmup 1 --batches-per-step 8 --bank-size 2000 --no-cpu-baseline --e2e-steps 1 --no-long-form --no-library-baseline --no-traffic --chunk-batches 8"
ncu --set full --clock-control none --import-source on -k regex:mix_kernel -s 4 -c 1 -f -o gpurun_out/prof_mix_kernel_$v $SMALL > /dev/null 2>&1
ls gpurun_out | grep prof_
