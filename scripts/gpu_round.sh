#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x -s > gpurun_out/tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -5 gpurun_out/tests_gpu.log; grep "\[generate" gpurun_out/tests_gpu.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/step_metrics.csv python scripts/one_step.py > gpurun_out/one_step.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/one_step.log
python scripts/ncu_stage_table.py gpurun_out/step_metrics.csv gpurun_out/one_step.json > gpurun_out/stage_table.txt 2>&1; head -30 gpurun_out/stage_table.txt
