#!/bin/bash
# final sanity on the committed build: smoke + GPU test suite + default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2t_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2t_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/r2t_tests_gpu.log
timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench: $(cut -c48-110 gpurun_out/r2t_bench.json)"; tail -2 gpurun_out/r2t_bench.err
