mkdir -p gpurun_out/trace
python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python tools/co_sweep.py --groups 0 --chunks 16 --trace gpurun_out/trace 2>&1 | grep -v Warning
