mkdir -p gpurun_out
rm -f gpurun_out/summary17.txt
timeout 900 python -m pytest tests/ -m gpu -q --tb=short > gpurun_out/tests17.log 2>&1; echo "all gpu tests exit $?" >> gpurun_out/summary17.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench exit $?" >> gpurun_out/summary17.txt
cat gpurun_out/summary17.txt; tail -2 gpurun_out/tests17.log; cut -c1-200 gpurun_out/bench17.json
