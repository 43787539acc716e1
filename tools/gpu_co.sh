#!/bin/bash
# GPU session for the pipelined front end: parity tests, then tools/co_sweep.py for every library variant given.
# usage: tools/gpu_co.sh tag [variant ...]   ("base" = the product library)
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
: > gpurun_out/co_sweep_$TAG.log
for v in "$@"; do
  if [ "$v" = base ]; then unset ADTFE_LIB; else export ADTFE_LIB=$PWD/adt_str_b200/libadtfe_$v.so; fi
  timeout 300 python tools/co_sweep.py $SWEEP_ARGS 2>&1 | grep -v Warning | tee -a gpurun_out/co_sweep_$TAG.log
done
