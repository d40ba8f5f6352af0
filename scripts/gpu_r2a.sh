#!/bin/bash
# round 2, first GPU pass: parity tests + bench at 224 and 28 episodes per GPU + launch list at 28
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1   # no-op when the shipped .so matches the sources
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2a_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/r2a_tests_gpu.log
grep -h "^\[full3b\|^\[yardstick" gpurun_out/r2a_tests_gpu.log | head -40
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_224.json 2> gpurun_out/r2a_bench_224.err; echo "bench224 exit $?"; cut -c1-300 gpurun_out/r2a_bench_224.json; tail -3 gpurun_out/r2a_bench_224.err
DEER_GEMM_OLD_PICK=1 timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_28_oldpick.json 2> gpurun_out/r2a_bench_28_oldpick.err; echo "bench28 oldpick exit $?"; cut -c1-300 gpurun_out/r2a_bench_28_oldpick.json
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_28.json 2> gpurun_out/r2a_bench_28.err; echo "bench28 exit $?"; cut -c1-300 gpurun_out/r2a_bench_28.json; tail -3 gpurun_out/r2a_bench_28.err
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_b28.csv python bench.py --batch 28 --steps 2 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2a_ncu_b28.log 2>&1; echo "ncu b28 exit $?"
python scripts/summarize_launches.py gpurun_out/r2a_launches_b28.csv > gpurun_out/r2a_launches_b28.txt 2>&1; head -30 gpurun_out/r2a_launches_b28.txt
