#!/bin/bash
# sanitizer pass over the kernels whose issue warps changed (converged + elect.sync): GEMM variants, attention, a tiny engine rollout
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
K='(half_height and (514-1024-1024 or 130-264-72 or 1-512-2048)) or (split_k and (1-512-2048 or 130-256-2048)) or (tma_store and 5003 and 0) or test_tiny_golden_cuda_graph_conditional_exit or (attention and not masked)'
for tool in memcheck synccheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -q -m gpu -k "$K" > gpurun_out/r3g_$tool.log 2>&1; echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r3g_$tool.log | tail -3
done
