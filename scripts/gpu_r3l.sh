#!/bin/bash
# final build on 8 GPUs (torchrun, NCCL): 224 episodes sharded 28 per GPU (BASELINE configs[2] as written)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r3l_bench_8gpu.json 2> gpurun_out/r3l_bench_8gpu.err; echo "bench 8gpu exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3l_bench_8gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['clocks']['sm_mhz'], d['roofline']['frac'])"; tail -2 gpurun_out/r3l_bench_8gpu.err
