#!/bin/bash
# attention schedule v5 vs v4:  gpurun --timeout 1500 -- 'bash scripts/gpu_exp4.sh'
mkdir -p gpurun_out
L=$PWD/deer_vla_b200
timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short -x -k "attention or pv_probe" > gpurun_out/exp4_ops.log 2>&1; echo "ops exit $?"; tail -5 gpurun_out/exp4_ops.log
echo "== attention v5"; timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "^$" | head -1
echo "== attention v5 (instrumented build)"; DEER_LIB=$L/libdeer_engine_prof.so timeout 300 python scripts/attn_bench.py 2>&1 | tail -1
echo "== attention v4"; DEER_ATTN_V4=1 timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "^$" | head -1
run() { name=$1; shift; extra=$1; shift; env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline $extra > gpurun_out/exp4_bench_$name.json 2> gpurun_out/exp4_bench_$name.err; echo "bench $name exit $?: $(python -c "import json;d=json.load(open('gpurun_out/exp4_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'gemm frac', round(d['roofline']['frac'],3), d['clocks'])" 2>&1 | tail -1)"; }
run v5 "" DEER_X=0
run v4 "" DEER_ATTN_V4=1
run v5b "" DEER_X=0
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/exp4_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/exp4_tests_gpu.log
