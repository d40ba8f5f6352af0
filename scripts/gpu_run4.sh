mkdir -p gpurun_out
rm -f gpurun_out/summary4.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "gemm" > gpurun_out/tests4_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/summary4.txt
timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/gemm_bench4.log 2>&1; echo "gemm_bench exit $?" >> gpurun_out/summary4.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/tests4.log 2>&1; echo "parity exit $?" >> gpurun_out/summary4.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench exit $?" >> gpurun_out/summary4.txt
cat gpurun_out/summary4.txt; tail -5 gpurun_out/tests4_ops.log;  cat gpurun_out/gemm_bench4.log | head -14; tail -3 gpurun_out/tests4.log; cat gpurun_out/bench4.json | cut -c1-300
