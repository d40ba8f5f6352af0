#!/bin/bash
# weak-scaling check on N GPUs of one box: bash scripts/gpu_scale.sh N
N=${1:-8}
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N exit $?"; cut -c1-200 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 4 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "ref arm N=$N exit $?"; cut -c1-160 gpurun_out/bench_ref_${N}gpu.json
