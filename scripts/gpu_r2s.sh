#!/bin/bash
# round 2, pass s: hierarchical (cluster-leader) grid barrier in the fused head probe
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "fused or tiny or batch_invariance or path_switch or long_golden" > gpurun_out/r2s_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2s_tests.log
timeout 600 python scripts/head_bench.py > gpurun_out/r2s_head_bench.log 2>&1; echo "head_bench exit $?"; grep -v "^{" gpurun_out/r2s_head_bench.log | tail -10
for cs in 4 1; do DEER_HEAD_CLUSTER=$cs timeout 300 python scripts/head_bench.py --quick > gpurun_out/r2s_head_bench_cs$cs.log 2>&1; echo "cs=$cs exit $?"; grep "timeline\|fused=1" gpurun_out/r2s_head_bench_cs$cs.log | tail -4; done
timeout 600 python scripts/soak.py > gpurun_out/r2s_soak.log 2>&1; echo "soak exit $?"; tail -6 gpurun_out/r2s_soak.log
