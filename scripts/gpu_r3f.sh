#!/bin/bash
# elect.sync issue in the attention kernel and the TMA-store epilogue: tests, attention micro-benchmark, benches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/attn_bench.py 2>&1 | tail -6 | tee gpurun_out/r3f_attn_bench.txt
echo "== gemm_bench 224 auto"; timeout 300 python scripts/gemm_bench.py 0 2>&1 | grep -v "^{" | tee gpurun_out/r3f_gemm_224.txt
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3f_bench_$i.json 2> gpurun_out/r3f_bench_$i.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3f_bench_$i.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac'])"
done
timeout 600 python bench.py --batch 28 --no-cpu-baseline > gpurun_out/r3f_bench28.json 2> gpurun_out/r3f_bench28.err; echo "bench28 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3f_bench28.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
timeout 600 python bench.py --batch 1 --steps 50 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r3f_bench_b1_graph.json 2> gpurun_out/r3f_bench_b1_graph.err; echo "b1 graph exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3f_bench_b1_graph.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
