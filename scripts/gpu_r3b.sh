#!/bin/bash
# 64 x 64 tiles (tcgen05.mma M = 64) for sparse-wave launches: bit-identity tests, micro-benchmark, 1-episode / 28-episode step A/B
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm" 2>&1 | tail -4
echo "== gemm_bench b1 impl 2 (1-CTA kernels, 64x64 tiles allowed)"; timeout 300 python scripts/gemm_bench.py 2 all b1 2>&1 | grep -v "^{" | tee gpurun_out/r3b_gemm_b1_m64.txt
echo "== gemm_bench b1 impl 4 (128-row tiles only)"; timeout 300 python scripts/gemm_bench.py 4 all b1 2>&1 | grep -v "^{" | tee gpurun_out/r3b_gemm_b1_m128.txt
for hm in 1 0; do
  DEER_GEMM_M64=$hm timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r3b_bench_b1_graph_m64_${hm}.json 2> gpurun_out/r3b_bench_b1_graph_m64_${hm}.err; echo "b1 graph m64=$hm exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3b_bench_b1_graph_m64_${hm}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
  DEER_GEMM_M64=$hm timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_b1_m64_${hm}.json 2> gpurun_out/r3b_bench_b1_m64_${hm}.err; echo "b1 eager m64=$hm exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3b_bench_b1_m64_${hm}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
  DEER_GEMM_M64=$hm timeout 600 python bench.py --batch 28 --no-cpu-baseline > gpurun_out/r3b_bench28_m64_${hm}.json 2> gpurun_out/r3b_bench28_m64_${hm}.err; echo "b28 m64=$hm exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3b_bench28_m64_${hm}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
