#!/bin/bash
# sanitizer evidence + projection + full bench (1 GPU)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_projection.py -x -q -m gpu 2>&1 | tail -3
python tools/bench_projection.py; python tools/bench_projection.py --rows 15744
bash tools/gpu_sanitize.sh memcheck racecheck synccheck 2>&1 | tail -12
python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -3 gpurun_out/bench_h.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_h.json")); r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "render", round(r["render_ms_per_step"], 2), "logmel", round(r["logmel_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "ceiling", d["e2e"]["d2h_ceiling"])
PY
