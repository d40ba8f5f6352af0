#!/bin/bash
# round 2, second GPU pass: fused head probe (tests, timeline, step A/B) + small-M GEMM shapes with the old and the new tile pick
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q --tb=short -s -k "fused or tiny or variants or ensemble or facade or generate or pth or batch_invariance" > gpurun_out/r2b_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2b_tests.log
timeout 600 python scripts/head_bench.py > gpurun_out/r2b_head_bench.log 2>&1; echo "head_bench exit $?"; grep -v "^{" gpurun_out/r2b_head_bench.log | tail -12
for set in b28 b1; do
  timeout 300 python scripts/gemm_bench.py 0 all $set > gpurun_out/r2b_gemm_${set}_new.log 2>&1; echo "gemm $set new exit $?"
  DEER_GEMM_OLD_PICK=1 timeout 300 python scripts/gemm_bench.py 0 all $set > gpurun_out/r2b_gemm_${set}_old.log 2>&1; echo "gemm $set old exit $?"
  paste <(grep "M=" gpurun_out/r2b_gemm_${set}_old.log | cut -c1-75) <(grep "M=" gpurun_out/r2b_gemm_${set}_new.log | cut -c46-75)
done
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_b1.json 2> gpurun_out/r2b_bench_b1.err; echo "bench b1 exit $?"; cut -c1-200 gpurun_out/r2b_bench_b1.json
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r2b_bench_b1_graph.json 2> gpurun_out/r2b_bench_b1_graph.err; echo "bench b1 graph exit $?"; cut -c1-200 gpurun_out/r2b_bench_b1_graph.json
