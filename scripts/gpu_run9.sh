mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 5 -c 1 -o gpurun_out/prof_tc2_fc1 python scripts/gemm_bench.py 0 vit.fc1+qgelu > gpurun_out/ncu9a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 5 -c 1 -o gpurun_out/prof_tc2_out python scripts/gemm_bench.py 0 vit.out+res > gpurun_out/ncu9b.log 2>&1
echo done
