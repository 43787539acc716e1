timeout 300 python tools/co_sweep.py --groups 0 --chunks 8,16,32,64,256 2>&1 | grep -v Warning
