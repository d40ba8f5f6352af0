#!/bin/bash
# final build on 2 GPUs (torchrun, NCCL): 224 episodes sharded 112 per GPU
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_bench_2gpu.json 2> gpurun_out/r3k_bench_2gpu.err; echo "bench 2gpu exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3k_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['clocks']['sm_mhz'], d['roofline']['frac'])"; tail -2 gpurun_out/r3k_bench_2gpu.err
