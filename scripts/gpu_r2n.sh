#!/bin/bash
# round 2, pass n: GEMM tail split (A/B inside one box), adaptive calibration windows, new tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_round2.py -m gpu -q -x -k "gemm or path_switch or ragged or batch_invariance" > gpurun_out/r2n_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2n_tests.log
for set in all b28; do
  a=$([ $set = all ] && echo "" || echo "all $set")
  timeout 300 python scripts/gemm_bench.py 0 $a > gpurun_out/r2n_gemm_${set}_split.log 2>&1
  DEER_GEMM_TAIL_SPLIT=0 timeout 300 python scripts/gemm_bench.py 0 $a > gpurun_out/r2n_gemm_${set}_nosplit.log 2>&1
  echo "GEMM $set: no tail split | tail split"; paste <(grep "M=" gpurun_out/r2n_gemm_${set}_nosplit.log | cut -c1-75) <(grep "M=" gpurun_out/r2n_gemm_${set}_split.log | cut -c46-75) | head -10
done
for rep in 1 2; do
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_224_split_$rep.json 2> gpurun_out/r2n_bench_224.err; echo "bench224 split: $(cut -c48-110 gpurun_out/r2n_bench_224_split_$rep.json)"
DEER_GEMM_TAIL_SPLIT=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_224_nosplit_$rep.json 2>> gpurun_out/r2n_bench_224.err; echo "bench224 nosplit: $(cut -c48-110 gpurun_out/r2n_bench_224_nosplit_$rep.json)"
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_28_split_$rep.json 2> gpurun_out/r2n_bench_28.err; echo "bench28 split: $(cut -c48-110 gpurun_out/r2n_bench_28_split_$rep.json)"
DEER_GEMM_TAIL_SPLIT=0 timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_28_nosplit_$rep.json 2>> gpurun_out/r2n_bench_28.err; echo "bench28 nosplit: $(cut -c48-110 gpurun_out/r2n_bench_28_nosplit_$rep.json)"
done
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_b1.json 2> gpurun_out/r2n_bench_b1.err; echo "b1: $(cut -c48-110 gpurun_out/r2n_bench_b1.json)"; tail -2 gpurun_out/r2n_bench_b1.err
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r2n_bench_b1_graph.json 2> gpurun_out/r2n_bench_b1_graph.err; echo "b1 graph: $(cut -c48-110 gpurun_out/r2n_bench_b1_graph.json)"
