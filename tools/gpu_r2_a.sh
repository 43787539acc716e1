#!/bin/bash
# round 2, first GPU call: TMA element-offset probe, parity suite, bench with the library comparator
mkdir -p gpurun_out
(cd tools/microbench && ./tma_shift) > gpurun_out/r02_tma_shift.txt 2>&1; cat gpurun_out/r02_tma_shift.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_a.log; cat gpurun_out/pytest_a.log
python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_a.json")); r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "render", round(r["render_ms_per_step"], 2), "logmel", round(r["logmel_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]))
print("traffic", r["traffic"], r["traffic_source"])
print("library", d["gpu_library_baseline"])
PY
