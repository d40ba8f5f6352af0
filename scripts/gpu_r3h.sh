#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_tiny_golden_cuda_graph_conditional_exit or test_tiny_golden_dynamic_exit_fp32" > gpurun_out/r3h_synccheck.log 2>&1; echo "synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|=========     at " gpurun_out/r3h_synccheck.log | sort | uniq -c | tail -6
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tma_store and 5003 and 0" > gpurun_out/r3h_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|and Read access|Write access" gpurun_out/r3h_racecheck.log | sort | uniq -c | tail -6
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; echo "bench exit $?"; cut -c1-150 gpurun_out/r3h_bench.json
