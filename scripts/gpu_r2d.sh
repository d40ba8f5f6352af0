#!/bin/bash
# round 2, pass d: cluster-staged fused head (tests, timeline at cluster size 8 / 4 / 1, step A/B)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q --tb=short -s -k "fused or tiny or batch_invariance or overlap" > gpurun_out/r2d_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2d_tests.log
timeout 600 python scripts/head_bench.py > gpurun_out/r2d_head_bench.log 2>&1; echo "head_bench exit $?"; grep -v "^{" gpurun_out/r2d_head_bench.log | tail -14
for cs in 4 1; do DEER_HEAD_CLUSTER=$cs timeout 600 python scripts/head_bench.py --quick > gpurun_out/r2d_head_bench_cs$cs.log 2>&1; echo "head_bench cs=$cs exit $?"; grep -v "^{" gpurun_out/r2d_head_bench_cs$cs.log | tail -8; done
timeout 600 python bench.py --batch 28 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_28.json 2> gpurun_out/r2d_bench_28.err; echo "bench28 exit $?"; cut -c1-200 gpurun_out/r2d_bench_28.json; tail -3 gpurun_out/r2d_bench_28.err
