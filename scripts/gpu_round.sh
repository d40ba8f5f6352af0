#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -5 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma.json 2> gpurun_out/bench_mma.err; echo "bench exit $?"; cut -c1-120 gpurun_out/bench_mma.json
DEER_ATTN_LLM_SIMT=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "bench exit $?"; cut -c1-120 gpurun_out/bench_simt.json
