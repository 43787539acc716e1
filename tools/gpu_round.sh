#!/bin/bash
# One GPU session: tests, bench (+ reference arm), ncu launch list and full captures.  Usage: tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
SMALL="python bench.py --steps 2 --warmup 1 --batches-per-step 8 --bank-size 2000 --no-cpu-baseline --e2e-steps 1 --no-long-form"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > /dev/null 2>&1
for k in ${KERNELS:-logmel6_kernel mix_kernel peak_kernel normalise_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_${k}_$TAG $SMALL > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 2 -c 1 -f -o gpurun_out/prof_resample_kernel_$TAG python tools/bench_longform.py --seconds 120 --cpu-seconds 0 > /dev/null 2>&1
ls -la gpurun_out | tail -12
