#!/bin/bash
# converged producer / MMA warps (elect.sync): k-block timeline, micro-benchmarks at the three batch sizes, tests, benches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
H=$(python -c "import __graft_entry__ as g; print(g.source_hash())")
(cd deer_vla_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DDEER_GEMM_PROF -DDEER_SRC_HASH=\"$H\" -Xcompiler -fPIC -shared engine.cu -o ../../scripts/libdeer_engine_prof.so) 2>&1 | tail -3
DEER_LIB=scripts/libdeer_engine_prof.so timeout 300 python scripts/gemm_kblock_prof.py 2>&1 | tee gpurun_out/r3e_kblock_prof.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
for set in b1 b28; do echo "== gemm_bench $set auto"; timeout 300 python scripts/gemm_bench.py 0 all $set 2>&1 | grep -v "^{" | tee gpurun_out/r3e_gemm_$set.txt; done
echo "== gemm_bench 224 auto"; timeout 300 python scripts/gemm_bench.py 0 2>&1 | grep -v "^{" | tee gpurun_out/r3e_gemm_224.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3e_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac'])"
timeout 600 python bench.py --batch 28 --no-cpu-baseline > gpurun_out/r3e_bench28.json 2> gpurun_out/r3e_bench28.err; echo "bench28 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3e_bench28.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r3e_bench_b1_graph.json 2> gpurun_out/r3e_bench_b1_graph.err; echo "b1 graph exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3e_bench_b1_graph.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_bench_b1.json 2> gpurun_out/r3e_bench_b1.err; echo "b1 eager exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3e_bench_b1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
