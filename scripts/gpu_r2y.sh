#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
K='split_k and (32-2048-2048 or 1-512-2048 or 97-520-2000 or 130-256-2048)'
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$K" > gpurun_out/r2y_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2y_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$K" > gpurun_out/r2y_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2y_racecheck.log | tail -3
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$K" > gpurun_out/r2y_synccheck.log 2>&1; echo "synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2y_synccheck.log | tail -3
