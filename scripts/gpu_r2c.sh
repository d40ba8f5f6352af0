#!/bin/bash
# round 2, third GPU pass: fused head after the staging fix; residual-prefetch A/B on the big GEMM shapes; bench 224 / 28
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q --tb=short -s -k "fused or tiny or batch_invariance or overlap" > gpurun_out/r2c_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2c_tests.log
timeout 600 python scripts/head_bench.py > gpurun_out/r2c_head_bench.log 2>&1; echo "head_bench exit $?"; grep -v "^{" gpurun_out/r2c_head_bench.log | tail -12
timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/r2c_gemm_big_prefetch.log 2>&1; echo "gemm big exit $?"
DEER_EPI_DBG=4 timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/r2c_gemm_big_noprefetch.log 2>&1; echo "gemm big noprefetch exit $?"
paste <(grep "M=" gpurun_out/r2c_gemm_big_noprefetch.log | cut -c1-75) <(grep "M=" gpurun_out/r2c_gemm_big_prefetch.log | cut -c46-75)
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_224.json 2> gpurun_out/r2c_bench_224.err; echo "bench224 exit $?"; cut -c1-200 gpurun_out/r2c_bench_224.json; tail -3 gpurun_out/r2c_bench_224.err
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_28.json 2> gpurun_out/r2c_bench_28.err; echo "bench28 exit $?"; cut -c1-200 gpurun_out/r2c_bench_28.json; tail -3 gpurun_out/r2c_bench_28.err
