#!/bin/bash
# round 2, pass l (8 GPUs): BASELINE configs[2] as written -- 224 episodes sharded 28 per GPU over 8 ranks (NCCL), strong scaling
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 16 --warmup 3 > gpurun_out/r2l_bench_8gpu.json 2> gpurun_out/r2l_bench_8gpu.err; echo "bench 8gpu exit $?"; cut -c1-300 gpurun_out/r2l_bench_8gpu.json; tail -3 gpurun_out/r2l_bench_8gpu.err
