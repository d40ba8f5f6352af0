#!/bin/bash
# final sanity of the committed build: smoke() + the default bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3; echo "smoke exit $?"
timeout 900 python bench.py > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3j_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic_source'], d['cpu_baseline']['value'])"
