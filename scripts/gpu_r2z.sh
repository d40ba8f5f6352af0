#!/bin/bash
# final-build launch lists (ncu, timed steps only) of the default bench command and of the 28-episode shard
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_b224.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2z_ncu_b224.log 2>&1; echo "ncu b224 exit $?"
python scripts/summarize_launches.py gpurun_out/r2z_launches_b224.csv > gpurun_out/r2z_launches_b224.txt 2>&1; head -12 gpurun_out/r2z_launches_b224.txt
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_b28.csv python bench.py --batch 28 --steps 2 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2z_ncu_b28.log 2>&1; echo "ncu b28 exit $?"
python scripts/summarize_launches.py gpurun_out/r2z_launches_b28.csv > gpurun_out/r2z_launches_b28.txt 2>&1; head -12 gpurun_out/r2z_launches_b28.txt
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"; cut -c1-160 gpurun_out/r2z_bench.json
