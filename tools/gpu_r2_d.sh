#!/bin/bash
# round 2: tile mixer variants (warp tile width x ring depth x one-warp CTAs per SM): digest against round 1 + bench
mkdir -p gpurun_out
(cd _r1 && python ../tools/wav_digest.py 2>&1 | tail -2) > gpurun_out/digest_r1.txt
export BENCH_ARGS="--no-long-form --no-library-baseline --no-traffic"
for v in "$@"; do
  if [ "$v" = base ]; then unset ADTFE_LIB; else export ADTFE_LIB=$PWD/adt_str_b200/libadtfe_$v.so; fi
  python tools/wav_digest.py 2>&1 | tail -2 > gpurun_out/digest_$v.txt
  cmp -s gpurun_out/digest_$v.txt gpurun_out/digest_r1.txt && echo "$v BIT-IDENTICAL to round 1" || { echo "$v DIFFERS"; cat gpurun_out/digest_$v.txt; }
  bash tools/gpu_variants.sh $v
done
