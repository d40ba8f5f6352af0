#!/bin/bash
# round 2, pass g: attention phase cycles (instrumented build), bench 224 with the restricted residual prefetch, bench 28 eager vs graph
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
(cd deer_vla_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DDEER_ATTN_PROF engine.cu -o /tmp/libdeer_prof.so) && DEER_LIB=/tmp/libdeer_prof.so timeout 300 python scripts/attn_bench.py > gpurun_out/r2g_attn_phases.log 2>&1; echo "attn phases exit $?"; cat gpurun_out/r2g_attn_phases.log | tail -4
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_224.json 2> gpurun_out/r2g_bench_224.err; echo "bench224 exit $?"; cut -c1-200 gpurun_out/r2g_bench_224.json; tail -3 gpurun_out/r2g_bench_224.err
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_28.json 2> gpurun_out/r2g_bench_28.err; echo "bench28 exit $?"; cut -c1-200 gpurun_out/r2g_bench_28.json
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r2g_bench_28_graph.json 2> gpurun_out/r2g_bench_28_graph.err; echo "bench28 graph exit $?"; cut -c1-200 gpurun_out/r2g_bench_28_graph.json; tail -3 gpurun_out/r2g_bench_28_graph.err
