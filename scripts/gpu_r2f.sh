#!/bin/bash
# round 2, pass f: ncu captures (full sets of the attention kernel, the 2-CTA GEMM on two shapes, the fused head probe) and the
# per-launch DRAM-traffic / tensor-pipe metrics of the bench command itself
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn_tc5 -s 2 -c 1 -o gpurun_out/r2f_attn_tc5 -f python scripts/attn_one.py > gpurun_out/r2f_ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 600 $NCU -k regex:gemm_tc2 -s 3 -c 1 -o gpurun_out/r2f_gemm_vit_outres -f python scripts/gemm_bench.py 0 vit.out+res > gpurun_out/r2f_ncu_gemm_outres.log 2>&1; echo "ncu gemm out+res exit $?"
timeout 600 $NCU -k regex:gemm_tc2 -s 3 -c 1 -o gpurun_out/r2f_gemm_vit_qkv -f python scripts/gemm_bench.py 0 vit.qkv > gpurun_out/r2f_ncu_gemm_qkv.log 2>&1; echo "ncu gemm qkv exit $?"
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2f_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2f_ncu_bench_metrics.log 2>&1; echo "ncu bench metrics exit $?"
timeout 300 $NCU -k regex:head_fused -s 20 -c 1 -o gpurun_out/r2f_head_fused -f python scripts/head_bench.py --quick > gpurun_out/r2f_ncu_head.log 2>&1; echo "ncu head exit $?"
ls -la gpurun_out/r2f_*
