#!/bin/bash
# round 2, pass h (round-end check of the committed build on one B200): smoke, full GPU test suite, both bench arms, DeeR-S / 9B / B=1 lines,
# launch list + per-launch DRAM traffic of the bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2h_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2h_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/r2h_tests_gpu.log
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err; echo "reference arm exit $?"; cut -c1-200 gpurun_out/r2h_bench_reference.json
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"; cut -c1-260 gpurun_out/r2h_bench.json; tail -2 gpurun_out/r2h_bench.err
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_28.json 2> gpurun_out/r2h_bench_28.err; echo "bench28 exit $?"; cut -c1-200 gpurun_out/r2h_bench_28.json
timeout 600 python bench.py --workload deer_s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_deer_s.json 2> gpurun_out/r2h_bench_deer_s.err; echo "deer_s exit $?"; cut -c1-200 gpurun_out/r2h_bench_deer_s.json
timeout 900 python bench.py --workload deer_9b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_9b.json 2> gpurun_out/r2h_bench_9b.err; echo "9b exit $?"; cut -c1-200 gpurun_out/r2h_bench_9b.json; tail -2 gpurun_out/r2h_bench_9b.err
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_b1.json 2> gpurun_out/r2h_bench_b1.err; echo "b1 exit $?"; cut -c1-200 gpurun_out/r2h_bench_b1.json
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r2h_bench_b1_graph.json 2> gpurun_out/r2h_bench_b1_graph.err; echo "b1 graph exit $?"; cut -c1-200 gpurun_out/r2h_bench_b1_graph.json
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2h_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2h_ncu_bench_metrics.log 2>&1; echo "ncu bench metrics exit $?"
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_b224.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2h_ncu_b224.log 2>&1; echo "ncu b224 exit $?"
python scripts/summarize_launches.py gpurun_out/r2h_launches_b224.csv > gpurun_out/r2h_launches_b224.txt 2>&1; head -12 gpurun_out/r2h_launches_b224.txt
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_b28.csv python bench.py --batch 28 --steps 2 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2h_ncu_b28.log 2>&1; echo "ncu b28 exit $?"
python scripts/summarize_launches.py gpurun_out/r2h_launches_b28.csv > gpurun_out/r2h_launches_b28.txt 2>&1; head -12 gpurun_out/r2h_launches_b28.txt
