#!/bin/bash
# one-MUFU exact GELU epilogue: accuracy tests + GEMM microbench + full GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short -x -k "gemm" 2>&1 | tail -3
timeout 300 python scripts/gemm_bench.py 2>&1 | grep "gelu\|qkv\|down"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/exp8_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/exp8_tests_gpu.log
