#!/bin/bash
# Rebuild the library, run the CPU-side symbol test, then hand the command to gpurun (so a stale .so never reaches the GPU box).
#   tools/gpu.sh <timeout-seconds> '<command>'        (log: gpurun_out/last_call.log)
set -e
cd "$(dirname "$0")/.."
python -c "import importlib; importlib.import_module('3d_multiview_reg_b200').build()" > /dev/null
python -m pytest tests/test_cabi_symbols.py -x -q > /dev/null
mkdir -p gpurun_out
/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2" > gpurun_out/last_call.log 2>&1
tail -n 40 gpurun_out/last_call.log
