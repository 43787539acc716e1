#!/bin/bash
# A/B of tuning variants built with `python adt_str_b200/build.py -D... --out=libadtfe_<tag>.so`:
# usage: tools/gpu_variants.sh <tag> [<tag> ...]   ("base" = the product library)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset ADTFE_LIB; else export ADTFE_LIB=$PWD/adt_str_b200/libadtfe_$v.so; fi
  python bench.py --no-cpu-baseline --e2e-steps 1 --steps 4 $BENCH_ARGS > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err || tail -3 gpurun_out/bench_var_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_var_$v.json")); r = d["roofline"]
print("$v", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "render", round(r["render_ms_per_step"], 2), "logmel", round(r["logmel_ms_per_step"], 2))
PY
done
