mkdir -p gpurun_out
rm -f gpurun_out/summary3.txt
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench exit $?" >> gpurun_out/summary3.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/tests3.log 2>&1; echo "tests exit $?" >> gpurun_out/summary3.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?" >> gpurun_out/summary3.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 700 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 2 --warmup 3 --calib-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary3.txt
cat gpurun_out/summary3.txt; cat gpurun_out/gemm_bench.log | head -14; tail -3 gpurun_out/tests3.log; cat gpurun_out/bench3.json | cut -c1-600
