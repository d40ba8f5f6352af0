#!/bin/bash
# A/B of the chunk-pipelined attention softmax (default) vs the whole-row schedule (DEER_ATTN_V3=1) and of programmatic
# dependent launch (DEER_PDL=0 turns it off):  gpurun --timeout 1500 -- 'bash scripts/gpu_exp1.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short -x > gpurun_out/exp1_ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/exp1_ops.log
for v in v4 v3; do
  [ $v = v3 ] && export DEER_ATTN_V3=1 || unset DEER_ATTN_V3
  echo "== attention $v"; timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "^$" | tail -4
  echo "== attention $v (instrumented build)"; DEER_LIB=$PWD/deer_vla_b200/libdeer_engine_prof.so timeout 300 python scripts/attn_bench.py 2>&1 | tail -1
done
unset DEER_ATTN_V3
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/exp1_bench_$name.json 2> gpurun_out/exp1_bench_$name.err; echo "bench $name exit $?: $(python -c "import json;d=json.load(open('gpurun_out/exp1_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'gemm frac', round(d['roofline']['frac'],3), d['clocks'])" 2>&1 | tail -1)"; }
run new DEER_X=0
run nopdl DEER_PDL=0
run old DEER_PDL=0 DEER_ATTN_V3=1
run new2 DEER_X=0
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/exp1_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/exp1_tests_gpu.log
