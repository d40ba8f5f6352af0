#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for hm in 1 0; do
echo "== gemm_bench b28 auto, DEER_GEMM_M64=$hm"; DEER_GEMM_M64=$hm timeout 300 python scripts/gemm_bench.py 0 all b28 2>&1 | grep -v "^{" | tee gpurun_out/r3c_gemm_b28_m64_$hm.txt
done
for hm in 1 0 1 0; do
  DEER_GEMM_M64=$hm timeout 600 python bench.py --batch 28 --no-cpu-baseline > gpurun_out/r3c_bench28_m64_${hm}.json 2> gpurun_out/r3c_bench28_m64_${hm}.err; echo "b28 m64=$hm exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3c_bench28_m64_${hm}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
