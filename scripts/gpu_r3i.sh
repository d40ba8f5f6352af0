#!/bin/bash
# final build: per-launch DRAM traffic / tensor-pipe activity of the bench command's timed step + one full-set capture of the 2-CTA GEMM
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r3i_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r3i_ncu_bench_metrics.log 2>&1; echo "ncu metrics exit $?"
python scripts/gemm_traffic.py gpurun_out/r3i_step_metrics.csv deer_b 224 "round 2, final build 4bd6d2cf (scripts/gpu_r3i.sh)" > gpurun_out/r3i_gemm_traffic.json; cut -c1-900 gpurun_out/r3i_gemm_traffic.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:gemm_tc2 -s 3 -c 1 -o gpurun_out/r3i_gemm_vit_qkv -f python scripts/gemm_bench.py 0 vit.qkv > gpurun_out/r3i_ncu_gemm_qkv.log 2>&1; echo "ncu gemm qkv exit $?"
ncu -i gpurun_out/r3i_gemm_vit_qkv.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]; v = rows[-1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "launch__registers_per_thread",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
for k in want:
    for i, name in enumerate(h):
        if name == k:
            print(f"{k:75s} {rows[1][i]:>12s} {v[i]}")
PY
