#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm" 2>&1 | tail -5
echo "== gemm_bench b1 (auto incl. split-K)"; timeout 300 python scripts/gemm_bench.py 0 all b1 2>&1 | grep -v "^{" | tee gpurun_out/r2v_gemm_b1_split.txt
echo "== gemm_bench b1 (no split-K)"; timeout 300 python scripts/gemm_bench.py 3 all b1 2>&1 | grep -v "^{" | tee gpurun_out/r2v_gemm_b1_nosplit.txt
for m in "" "--graph"; do
  for sk in 1 0; do
    DEER_GEMM_SPLITK=$sk timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline $m > gpurun_out/r2v_bench_b1_sk${sk}${m}.json 2> gpurun_out/r2v_bench_b1_sk${sk}${m}.err; echo "b1 sk=$sk $m exit $?"; python -c "
import json,sys; d=json.loads(open('gpurun_out/r2v_bench_b1_sk${sk}${m}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
  done
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
