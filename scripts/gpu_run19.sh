mkdir -p gpurun_out
# per-launch DRAM traffic + duration of every 2-CTA GEMM launch of one DeeR-B step (1 ncu pass per metric set)
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tc2 -s 560 -c 280 --csv --log-file gpurun_out/gemm_traffic.csv python bench.py --steps 2 --warmup 3 --calib-steps 1 --no-cpu-baseline > gpurun_out/ncu19.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 600 -c 2 -o gpurun_out/prof_gemm_tc2_step python bench.py --steps 2 --warmup 3 --calib-steps 1 --no-cpu-baseline > gpurun_out/ncu19b.log 2>&1; echo "ncu full exit $?"
