#!/bin/bash
# What a round-end check looks like on the GPU box:  gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-220 gpurun_out/bench.json
