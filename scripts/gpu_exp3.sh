#!/bin/bash
# attention barrier-flavour variants + wait-hint A/B on the GEMMs:  gpurun --timeout 1500 -- 'bash scripts/gpu_exp3.sh'
mkdir -p gpurun_out
L=$PWD/deer_vla_b200
for v in engine var_R var_H var_RH var_R_hint500; do
  echo "== attention lib$v"; DEER_LIB=$L/libdeer_$v.so timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "^$" | head -1
done
for v in var_RH_prof var_R_hint500_prof; do
  echo "== phases lib$v"; DEER_LIB=$L/libdeer_$v.so timeout 300 python scripts/attn_bench.py 2>&1 | tail -1
done
for v in engine var_R_hint500; do
  echo "== gemm lib$v"; DEER_LIB=$L/libdeer_$v.so timeout 300 python scripts/gemm_bench.py 2>&1 | grep TFLOP
done
for v in var_RH var_R_hint500; do
  DEER_LIB=$L/libdeer_$v.so timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short -x > gpurun_out/exp3_ops_$v.log 2>&1; echo "ops $v exit $?"; tail -1 gpurun_out/exp3_ops_$v.log
done
