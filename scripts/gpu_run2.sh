mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary2.txt
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_deer_b.json 2> gpurun_out/bench_deer_b.err; echo "bench_b exit $?" >> gpurun_out/summary2.txt
timeout 900 python bench.py --steps 8 --warmup 3 --workload deer_s --no-cpu-baseline > gpurun_out/bench_deer_s.json 2> gpurun_out/bench_deer_s.err; echo "bench_s exit $?" >> gpurun_out/summary2.txt
timeout 600 python bench.py --steps 8 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; echo "bench_b1 exit $?" >> gpurun_out/summary2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --calib-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 700 -c 3 -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --calib-steps 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt; cat gpurun_out/bench_deer_b.json; tail -3 gpurun_out/bench_deer_b.err
