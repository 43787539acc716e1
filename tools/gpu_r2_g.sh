#!/bin/bash
# fused row normalisation: parity suite, digest against round 1, variants x chunk sizes (render only), bench
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_g.log; cat gpurun_out/pytest_g.log
(cd _r1 && python ../tools/wav_digest.py 2>&1 | tail -2) > gpurun_out/digest_r1.txt
for v in base "$@"; do
  if [ "$v" = base ]; then unset ADTFE_LIB; else export ADTFE_LIB=$PWD/adt_str_b200/libadtfe_$v.so; fi
  python tools/wav_digest.py 2>&1 | tail -2 > gpurun_out/digest_$v.txt
  cmp -s gpurun_out/digest_$v.txt gpurun_out/digest_r1.txt && echo "$v BIT-IDENTICAL to round 1" || { echo "$v DIFFERS"; cat gpurun_out/digest_$v.txt; }
  if [ "$v" = base ]; then python tools/render_sweep.py --chunks 64,32,16,8 --render-only --trace gpurun_out 2>&1 | tail -20
  else python tools/render_sweep.py --chunks 64,32,16,8 --render-only 2>&1 | tail -4; fi
done
unset ADTFE_LIB
export BENCH_ARGS="--no-long-form --no-library-baseline --no-traffic"
bash tools/gpu_variants.sh base
