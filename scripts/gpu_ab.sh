#!/bin/bash
# in-box A/B of whole-step throughput (box-to-box variance is larger than the effects measured):  bash scripts/gpu_ab.sh "ENV_A=1" "ENV_B=1" [rounds]
mkdir -p gpurun_out
A=${1:-DEER_X=0}; B=${2:-DEER_ATTN_V4=1}; R=${3:-2}
for i in $(seq 1 $R); do
  for v in "$A" "$B"; do
    env $v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    echo "[$v] $(python -c "import json;d=json.load(open('gpurun_out/ab.json'));print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'raw', round(d['e2e_raw_frames']['value'],1), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])" 2>&1 | tail -1)"
  done
done
