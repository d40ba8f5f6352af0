#!/bin/bash
# round 2, pass k: full GPU suite on the final build, racecheck of the fused head (cluster sync at entry, long barrier timeout), bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2k_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2k_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/r2k_tests_gpu.log
grep -h "^\[full3b" gpurun_out/r2k_tests_gpu.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench exit $?"; cut -c1-260 gpurun_out/r2k_bench.json; tail -2 gpurun_out/r2k_bench.err
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_28.json 2> gpurun_out/r2k_bench_28.err; echo "bench28 exit $?"; cut -c1-200 gpurun_out/r2k_bench_28.json
timeout 600 python bench.py --workload deer_s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_deer_s.json 2> gpurun_out/r2k_bench_deer_s.err; echo "deer_s exit $?"; cut -c1-200 gpurun_out/r2k_bench_deer_s.json
timeout 900 python bench.py --workload deer_9b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_9b.json 2> gpurun_out/r2k_bench_9b.err; echo "9b exit $?"; cut -c1-200 gpurun_out/r2k_bench_9b.json
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2k_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2k_ncu_bench_metrics.log 2>&1; echo "ncu bench metrics exit $?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "fused_head_probe_vs_per_op_path_and_oracle and (1 or 7)" > gpurun_out/r2k_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/r2k_racecheck.log
