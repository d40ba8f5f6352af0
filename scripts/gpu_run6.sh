mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 5 -c 1 -o gpurun_out/prof_tc2_qkv python scripts/gemm_bench.py 0 vit.qkv > gpurun_out/ncu6.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu6.log
