#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python scripts/soak.py > gpurun_out/r2r_soak.log 2>&1; echo "soak exit $?"; tail -8 gpurun_out/r2r_soak.log
