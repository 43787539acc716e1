#!/bin/bash
# render time against the chunk size (batches rendered together): tools/chunk_sweep.sh
for c in 1 2 4 8 16 64; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-long-form --no-traffic --no-library-baseline --chunk-batches $c 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk', $c, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'render', round(d['roofline']['render_ms_per_step'],3), 'logmel', round(d['roofline']['logmel_ms_per_step'],3))"
done
