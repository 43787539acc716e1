#!/bin/bash
# round 2: new tile mixer - parity suite, bit-identity against the round-1 tree (_r1), bench
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_b.log; cat gpurun_out/pytest_b.log
python tools/wav_digest.py 2>&1 | tail -3 | tee gpurun_out/digest_new.txt
(cd _r1 && python ../tools/wav_digest.py 2>&1 | tail -3) | tee gpurun_out/digest_r1.txt
cmp gpurun_out/digest_new.txt gpurun_out/digest_r1.txt && echo "BIT-IDENTICAL to round 1"
python bench.py --no-cpu-baseline --no-long-form --no-library-baseline --no-traffic --e2e-steps 1 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_b.json")); r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "render", round(r["render_ms_per_step"], 2), "logmel", round(r["logmel_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]))
PY
