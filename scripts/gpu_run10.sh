mkdir -p gpurun_out
rm -f gpurun_out/summary10.txt
timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/gemm_bench10.log 2>&1; echo "gemm_bench exit $?" >> gpurun_out/summary10.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k gemm > gpurun_out/tests10.log 2>&1; echo "tests exit $?" >> gpurun_out/summary10.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench exit $?" >> gpurun_out/summary10.txt
cat gpurun_out/summary10.txt; head -13 gpurun_out/gemm_bench10.log; tail -2 gpurun_out/tests10.log; cut -c1-200 gpurun_out/bench10.json
