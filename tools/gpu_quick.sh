#!/bin/bash
# Quick GPU check: parity tests, bench, ncu of the kernels named after the tag.  Usage: tools/gpu_quick.sh tag [kernel ...]
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "render", round(r["render_ms_per_step"], 2), "logmel", round(r["logmel_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]))
PY
SMALL="python bench.py --steps 2 --warmup 1 --batches-per-step 8 --bank-size 2000 --no-cpu-baseline --e2e-steps 1"
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_${k}_$TAG $SMALL > /dev/null 2>&1
done
ls gpurun_out | tail -5
