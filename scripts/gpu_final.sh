#!/bin/bash
# Round-end check of the committed build on one B200:  gpurun --timeout 2400 -- 'bash scripts/gpu_final.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/final_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/final_tests_gpu.log
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm exit $?"; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; cut -c1-260 gpurun_out/final_bench.json
timeout 600 python bench.py --workload deer_s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_deer_s.json 2> gpurun_out/final_bench_deer_s.err; echo "deer_s exit $?"; cut -c1-200 gpurun_out/final_bench_deer_s.json
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --graph --no-cpu-baseline > gpurun_out/final_bench_b1_graph.json 2> gpurun_out/final_bench_b1_graph.err; echo "b1 graph exit $?"; cut -c1-200 gpurun_out/final_bench_b1_graph.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/step_metrics.csv python scripts/one_step.py > gpurun_out/final_one_step.log 2>&1; echo "ncu step exit $?"
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --calib-steps 1 > gpurun_out/final_ncu_bench.log 2>&1; echo "ncu bench exit $?"
