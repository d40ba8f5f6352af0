#!/bin/bash
# PDL on the remaining per-layer kernels (B=1 latency), ViT image chunking, attention max tree:  gpurun --timeout 1500 -- 'bash scripts/gpu_exp5.sh'
mkdir -p gpurun_out
echo "== attention v5"; timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "^$" | head -1
run() { name=$1; shift; extra=$1; shift; env "$@" timeout 600 python bench.py --warmup 3 --no-cpu-baseline $extra > gpurun_out/exp5_bench_$name.json 2> gpurun_out/exp5_bench_$name.err; echo "bench $name exit $?: $(python -c "import json;d=json.load(open('gpurun_out/exp5_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'raw', round(d['e2e_raw_frames']['value'],1), 'gemm frac', round(d['roofline']['frac'],3), d['clocks'])" 2>&1 | tail -1)"; }
run b1_pdl "--batch 1 --steps 50" DEER_PDL=1
run b1_nopdl "--batch 1 --steps 50" DEER_PDL=0
run b1_graph "--batch 1 --steps 50 --graph" DEER_X=0
run b224 "--steps 8" DEER_X=0
run b224_chunk112 "--steps 8" DEER_VIT_CHUNK=112
run b224_chunk56 "--steps 8" DEER_VIT_CHUNK=56
run b8_pdl "--batch 8 --steps 30" DEER_PDL=1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/exp5_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/exp5_tests_gpu.log
