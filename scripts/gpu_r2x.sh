#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
DEER_GEMM_SPLITK=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_full3b_bf16_within_rounding_yardstick" 2>&1 | tail -40
