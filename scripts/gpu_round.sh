#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -5 gpurun_out/tests_gpu.log; grep "\[generate" gpurun_out/tests_gpu.log
