mkdir -p gpurun_out
timeout 120 python scripts/attn_bench.py 2>&1 | head -2 > gpurun_out/attn_bench16.log
timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/gemm_bench16.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x > gpurun_out/tests16.log 2>&1; echo "ops exit $?"
cat gpurun_out/attn_bench16.log; head -13 gpurun_out/gemm_bench16.log; tail -2 gpurun_out/tests16.log
