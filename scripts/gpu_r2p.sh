#!/bin/bash
# round 2, pass p (4 GPUs): the N = 4 point of the strong-scaling table (56 episodes per GPU)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 16 --warmup 3 > gpurun_out/r2p_bench_4gpu.json 2> gpurun_out/r2p_bench_4gpu.err; echo "bench 4gpu exit $?"; cut -c1-200 gpurun_out/r2p_bench_4gpu.json; tail -2 gpurun_out/r2p_bench_4gpu.err
