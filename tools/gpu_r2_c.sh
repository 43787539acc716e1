#!/bin/bash
# round 2: tile mixer variants (ring depth x one-warp CTAs per SM) + ncu of the base build
mkdir -p gpurun_out
python tools/wav_digest.py 2>&1 | tail -2 | tee gpurun_out/digest_new.txt
(cd _r1 && python ../tools/wav_digest.py 2>&1 | tail -2) > gpurun_out/digest_r1.txt
cmp gpurun_out/digest_new.txt gpurun_out/digest_r1.txt && echo "BIT-IDENTICAL to round 1"
export BENCH_ARGS="--no-long-form --no-library-baseline --no-traffic"
bash tools/gpu_variants.sh base "$@"
SMALL="python bench.py --steps 2 --warmup 1 --batches-per-step 8 --bank-size 2000 --no-cpu-baseline --e2e-steps 1 --no-long-form --no-library-baseline --no-traffic --chunk-batches 8"
for k in mix_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_${k}_c $SMALL > /dev/null 2>&1
done
ls -la gpurun_out | tail -5
