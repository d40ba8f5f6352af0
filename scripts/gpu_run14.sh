mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 1 -o gpurun_out/prof_attn_tc python scripts/attn_bench.py > gpurun_out/ncu14.log 2>&1
echo "exit $?"
