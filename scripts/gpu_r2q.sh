#!/bin/bash
# round 2, pass q (8 GPUs): BASELINE configs[4] -- exit_ratio x max_layer sweep on OpenFlamingo-3B, 224 episodes sharded over 8 GPUs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 scripts/sweep.py --steps 16 > gpurun_out/r2q_sweep_8gpu.jsonl 2> gpurun_out/r2q_sweep_8gpu.err; echo "sweep exit $?"; wc -l gpurun_out/r2q_sweep_8gpu.jsonl; cut -c1-160 gpurun_out/r2q_sweep_8gpu.jsonl | head -30; tail -3 gpurun_out/r2q_sweep_8gpu.err
