#!/bin/bash
# round 2, pass i (2 GPUs): configs[2] as written -- 224 episodes sharded over the ranks (112 per GPU), both arms under torchrun
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err; echo "bench 2gpu exit $?"; cut -c1-300 gpurun_out/r2i_bench_2gpu.json; tail -3 gpurun_out/r2i_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/r2i_bench_reference_2gpu.json 2> gpurun_out/r2i_bench_reference_2gpu.err; echo "reference 2gpu exit $?"; cut -c1-200 gpurun_out/r2i_bench_reference_2gpu.json
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_1gpu.json 2> gpurun_out/r2i_bench_1gpu.err; echo "bench 1gpu exit $?"; cut -c1-200 gpurun_out/r2i_bench_1gpu.json
DEER_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_b224.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r2i_ncu_b224.log 2>&1; python scripts/summarize_launches.py gpurun_out/r2i_launches_b224.csv | grep "im2col\|vit_assemble\|launches"
