#!/bin/bash
# round 2, pass o: validation of the final build (q/k LayerNorm in one launch): smoke, full GPU suite, bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2o_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2o_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/r2o_tests_gpu.log
grep -h "^\[full3b" gpurun_out/r2o_tests_gpu.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench: $(cut -c48-110 gpurun_out/r2o_bench.json)"; tail -2 gpurun_out/r2o_bench.err
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_28.json 2> gpurun_out/r2o_bench_28.err; echo "bench28: $(cut -c48-110 gpurun_out/r2o_bench_28.json)"
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_b1.json 2> gpurun_out/r2o_bench_b1.err; echo "b1: $(cut -c48-110 gpurun_out/r2o_bench_b1.json)"
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r2o_bench_b1_graph.json 2> gpurun_out/r2o_bench_b1_graph.err; echo "b1 graph: $(cut -c48-110 gpurun_out/r2o_bench_b1_graph.json)"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2o_bench_reference.json 2> gpurun_out/r2o_bench_reference.err; echo "reference: $(cut -c20-120 gpurun_out/r2o_bench_reference.json)"
