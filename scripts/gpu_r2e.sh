#!/bin/bash
# round 2, pass e: full GPU test suite; fused head v3 timeline; bench 224 / 28 / 1; launch list at 224
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2e_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/r2e_tests_gpu.log
grep -h "^\[full3b\|^\[yardstick" gpurun_out/r2e_tests_gpu.log | head -40
timeout 600 python scripts/head_bench.py > gpurun_out/r2e_head_bench.log 2>&1; echo "head_bench exit $?"; grep -v "^{" gpurun_out/r2e_head_bench.log | tail -14
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_224.json 2> gpurun_out/r2e_bench_224.err; echo "bench224 exit $?"; cut -c1-200 gpurun_out/r2e_bench_224.json; tail -3 gpurun_out/r2e_bench_224.err
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_28.json 2> gpurun_out/r2e_bench_28.err; echo "bench28 exit $?"; cut -c1-200 gpurun_out/r2e_bench_28.json; tail -3 gpurun_out/r2e_bench_28.err
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_b1.json 2> gpurun_out/r2e_bench_b1.err; echo "bench b1 exit $?"; cut -c1-200 gpurun_out/r2e_bench_b1.json
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_b224.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2e_ncu_b224.log 2>&1; echo "ncu b224 exit $?"
python scripts/summarize_launches.py gpurun_out/r2e_launches_b224.csv > gpurun_out/r2e_launches_b224.txt 2>&1; head -30 gpurun_out/r2e_launches_b224.txt
