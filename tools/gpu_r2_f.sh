#!/bin/bash
# pipelined render schedule: variants x chunk sizes (render only, with traces for the first variant)
mkdir -p gpurun_out
(cd _r1 && python ../tools/wav_digest.py 2>&1 | tail -2) > gpurun_out/digest_r1.txt
first=1
for v in "$@"; do
  export ADTFE_LIB=$PWD/adt_str_b200/libadtfe_$v.so
  python tools/wav_digest.py 2>&1 | tail -2 > gpurun_out/digest_$v.txt
  cmp -s gpurun_out/digest_$v.txt gpurun_out/digest_r1.txt && echo "$v BIT-IDENTICAL to round 1" || { echo "$v DIFFERS"; cat gpurun_out/digest_$v.txt; }
  if [ $first = 1 ]; then python tools/render_sweep.py --chunks 64,32,16,8,4 --render-only --trace gpurun_out 2>&1 | tail -30; first=0
  else python tools/render_sweep.py --chunks 64,32,16,8,4 --render-only 2>&1 | tail -5; fi
done
