#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
echo "== gemm_bench b1 impl 3 (split-K allowed)"; timeout 300 python scripts/gemm_bench.py 3 all b1 2>&1 | grep -v "^{" | tee gpurun_out/r2w_gemm_b1_split.txt
echo "== gemm_bench b1 impl 2 (one CTA per tile)"; timeout 300 python scripts/gemm_bench.py 2 all b1 2>&1 | grep -v "^{" | tee gpurun_out/r2w_gemm_b1_nosplit.txt
for sk in 1 0; do
  DEER_GEMM_SPLITK=$sk timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r2w_bench_b1_sk${sk}.json 2> gpurun_out/r2w_bench_b1_sk${sk}.err; echo "b1 graph sk=$sk exit $?"; python -c "
import json,sys; d=json.loads(open('gpurun_out/r2w_bench_b1_sk${sk}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
DEER_GEMM_SPLITK=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -6
