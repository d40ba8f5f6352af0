#!/bin/bash
# round 2, pass j: compute-sanitizer (memcheck + racecheck + synccheck) over the tiny-size tests that exercise the new kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
K='fused_head_probe_vs_per_op_path_and_oracle or tiny_stage_taps or tiny_golden_dynamic_exit or head_variants or action_ensemble'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/r2j_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/r2j_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "fused_head_probe_vs_per_op_path_and_oracle" > gpurun_out/r2j_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/r2j_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "fused_head_probe_vs_per_op_path_and_oracle" > gpurun_out/r2j_synccheck.log 2>&1; echo "synccheck exit $?"; tail -4 gpurun_out/r2j_synccheck.log
