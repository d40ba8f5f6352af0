mkdir -p gpurun_out
rm -f gpurun_out/summary13.txt
timeout 900 python -m pytest tests/ -m gpu -q --tb=short > gpurun_out/tests13.log 2>&1; echo "all gpu tests exit $?" >> gpurun_out/summary13.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 700 --csv --log-file gpurun_out/launches13.csv python bench.py --steps 2 --warmup 3 --calib-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench13.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary13.txt
cat gpurun_out/summary13.txt; tail -3 gpurun_out/tests13.log
