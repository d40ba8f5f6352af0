#!/bin/bash
# attention epilogue A/B + PDL at B=1:  gpurun --timeout 1500 -- 'bash scripts/gpu_exp2.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short -x -k attention > gpurun_out/exp2_ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/exp2_ops.log
for v in v4 v3; do
  [ $v = v3 ] && export DEER_ATTN_V3=1 || unset DEER_ATTN_V3
  echo "== attention $v"; timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "^$" | head -1
  echo "== attention $v (instrumented build)"; DEER_LIB=$PWD/deer_vla_b200/libdeer_engine_prof.so timeout 300 python scripts/attn_bench.py 2>&1 | tail -1
done
unset DEER_ATTN_V3
run() { name=$1; shift; extra=$1; shift; env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline $extra > gpurun_out/exp2_bench_$name.json 2> gpurun_out/exp2_bench_$name.err; echo "bench $name exit $?: $(python -c "import json;d=json.load(open('gpurun_out/exp2_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'gemm frac', round(d['roofline']['frac'],3), d['clocks'])" 2>&1 | tail -1)"; }
run b224 "" DEER_X=0
run b1_pdl "--batch 1 --steps 50" DEER_PDL=1
run b1_nopdl "--batch 1 --steps 50" DEER_PDL=0
run b1_graph "--batch 1 --steps 50 --graph" DEER_X=0
run b224_b "" DEER_X=0
