#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_b1.csv python bench.py --batch 1 --steps 4 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2u_ncu_b1.log 2>&1; echo "ncu b1 exit $?"
python scripts/summarize_launches.py gpurun_out/r2u_launches_b1.csv > gpurun_out/r2u_launches_b1.txt 2>&1; head -30 gpurun_out/r2u_launches_b1.txt
