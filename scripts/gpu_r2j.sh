#!/bin/bash
# round 2, pass j: compute-sanitizer (memcheck + racecheck + synccheck) over the tiny-size tests that exercise the new kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
K='fused_head_probe_vs_per_op_path_and_oracle or tiny_stage_taps or tiny_golden_dynamic_exit or head_variants or action_ensemble'
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/r2j_ops_tests.log 2>&1; echo "ops tests exit $?"; tail -3 gpurun_out/r2j_ops_tests.log
timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/r2j_gemm_tma.log 2>&1; DEER_GEMM_TMA_STORE=0 timeout 300 python scripts/gemm_bench.py 0 > gpurun_out/r2j_gemm_notma.log 2>&1
echo "GEMM micro-benchmark: transposing epilogue | TMA-store epilogue (bf16 outputs)"; paste <(grep "M=" gpurun_out/r2j_gemm_notma.log | cut -c1-75) <(grep "M=" gpurun_out/r2j_gemm_tma.log | cut -c46-75)
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_224_tma.json 2> gpurun_out/r2j_bench_224_tma.err; echo "bench224 tma exit $?"; cut -c1-200 gpurun_out/r2j_bench_224_tma.json; tail -2 gpurun_out/r2j_bench_224_tma.err
DEER_GEMM_TMA_STORE=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_224_notma.json 2> gpurun_out/r2j_bench_224_notma.err; echo "bench224 no tma exit $?"; cut -c1-200 gpurun_out/r2j_bench_224_notma.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -q -x -k "$K or (tma_store and 5003)" > gpurun_out/r2j_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/r2j_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "fused_head_probe_vs_per_op_path_and_oracle" > gpurun_out/r2j_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/r2j_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "fused_head_probe_vs_per_op_path_and_oracle" > gpurun_out/r2j_synccheck.log 2>&1; echo "synccheck exit $?"; tail -4 gpurun_out/r2j_synccheck.log
