#!/bin/bash
# launch-overlap bit-identity test; PDL inside captured graphs (experimental) at B=1:  gpurun --timeout 1200 -- 'bash scripts/gpu_exp7.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -x -k "launch_overlap or cuda_graph or batch_invariance" 2>&1 | tail -3
run() { name=$1; shift; extra=$1; shift; env "$@" timeout 600 python bench.py --warmup 3 --no-cpu-baseline $extra > gpurun_out/exp7_bench_$name.json 2> gpurun_out/exp7_bench_$name.err; echo "bench $name exit $?: $(python -c "import json;d=json.load(open('gpurun_out/exp7_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" 2>&1 | tail -1)"; tail -2 gpurun_out/exp7_bench_$name.err | cut -c1-200; }
run b1_graph "--batch 1 --steps 50 --graph" DEER_X=0
run b1_graph_pdl "--batch 1 --steps 50 --graph" DEER_PDL_GRAPH=1
DEER_PDL_GRAPH=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -x -k "cuda_graph" 2>&1 | tail -3
