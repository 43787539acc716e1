#!/bin/bash
# multi-GPU session: the copy ceiling of the host at N ranks, then the bench under torchrun.  Usage: tools/gpu_multi.sh N [tag]
N=${1:-8}; TAG=${2:-r02}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py 2>&1 | grep -v "^W\|^\*\*\*" | tee gpurun_out/d2h_probe_${N}gpu_$TAG.txt | tail -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
tail -2 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu_$TAG.json")); e = d["e2e"]
print("N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(e["value"]), "ceiling", round(e["d2h_ceiling"]["value"]), "frac", round(e["frac_of_d2h_ceiling"], 3), e["cpu_affinity"])
PY
