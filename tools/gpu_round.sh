#!/bin/bash
# One GPU session at the end of a round: tests, bench (+ reference arm), ncu launch lists and full captures of every
# kernel, summarised into gpurun_out/ (copy the summaries into profiles/).  Usage: tools/gpu_round.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
SMALL="python bench.py --steps 2 --warmup 1 --batches-per-step 8 --bank-size 2000 --no-cpu-baseline --e2e-steps 1 --no-long-form --no-traffic --chunk-batches 8"
# launch list of the whole small run (library comparator included: its cuFFT / cuBLAS / elementwise kernels are the
# reference's GPU pipeline for the log-mel), then the same with DRAM bytes for the library kernels only
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > /dev/null 2>&1
# the launch list at the bench's own size (shares to compare with the bench line's)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_full_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-long-form --no-traffic --no-library-baseline --e2e-steps 1 > /dev/null 2>&1
for k in ${KERNELS:-logmel6_kernel mix_kernel peak_kernel normalise_kernel slice_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_${k}_$TAG $SMALL > /dev/null 2>&1
done
for k in fx_reverb_kernel fx_dynamics_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_$TAG python tools/fx_demo.py > /dev/null 2>&1
done
python tools/fx_demo.py | tee gpurun_out/fx_demo_$TAG.txt
ncu --set full --clock-control none --import-source on -k regex:project_kernel -s 1 -c 1 -f -o gpurun_out/prof_project_kernel_$TAG python tools/bench_projection.py --rows 2015232 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 2 -c 1 -f -o gpurun_out/prof_resample_kernel_$TAG python tools/bench_longform.py --seconds 120 --cpu-seconds 0 > /dev/null 2>&1
for f in gpurun_out/prof_*_$TAG.ncu-rep; do python tools/ncu_summary.py $f > ${f%.ncu-rep}.txt 2>/dev/null; done
ls -la gpurun_out | tail -20
