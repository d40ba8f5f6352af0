#!/bin/bash
# tests + sweep + 9B bench in one box
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/tests_gpu.log
grep "yardstick\]" gpurun_out/tests_gpu.log | head -12
timeout 900 python scripts/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep exit $?"; wc -l gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err
timeout 900 python bench.py --workload deer_9b --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_9b.json 2> gpurun_out/bench_9b.err; echo "bench9b exit $?"; cut -c1-200 gpurun_out/bench_9b.json; tail -3 gpurun_out/bench_9b.err
