mkdir -p gpurun_out
for d in 0 1; do echo "=== DEER_EPI_DBG=$d (2-CTA)"; DEER_EPI_DBG=$d timeout 300 python scripts/gemm_bench.py 0 2>&1 | head -13; done > gpurun_out/gemm_bench5.log 2>&1
echo "=== 1-CTA" >> gpurun_out/gemm_bench5.log; timeout 300 python scripts/gemm_bench.py 2 2>&1 | head -13 >> gpurun_out/gemm_bench5.log
cat gpurun_out/gemm_bench5.log
