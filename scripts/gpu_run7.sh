mkdir -p gpurun_out
rm -f gpurun_out/summary7.txt
timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/gemm_bench7.log 2>&1; echo "gemm_bench exit $?" >> gpurun_out/summary7.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/tests7.log 2>&1; echo "tests exit $?" >> gpurun_out/summary7.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench exit $?" >> gpurun_out/summary7.txt
cat gpurun_out/summary7.txt; head -13 gpurun_out/gemm_bench7.log; tail -3 gpurun_out/tests7.log; cut -c1-250 gpurun_out/bench7.json
