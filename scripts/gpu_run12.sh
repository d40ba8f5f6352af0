mkdir -p gpurun_out
rm -f gpurun_out/summary12.txt
timeout 120 python scripts/attn_bench.py > gpurun_out/attn_bench12.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_facade.py -m gpu -q --tb=short -x > gpurun_out/tests12.log 2>&1; echo "parity exit $?" >> gpurun_out/summary12.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench exit $?" >> gpurun_out/summary12.txt
cat gpurun_out/summary12.txt gpurun_out/attn_bench12.log; tail -2 gpurun_out/tests12.log; cut -c1-200 gpurun_out/bench12.json
