"""C5 sweep: exit_ratio x max_layer on OpenFlamingo-3B, device-timed action-steps/s vs mean exit depth.
One engine per max_layer (the model is built with min(ckpt_early_exit_layer, max_layer) layers, eval_calvin.py:529);
one full-depth delta table per engine feeds the reference's greedy threshold solver for every exit_ratio
(value_net.py:203-264).  Usage on the GPU box:  python scripts/sweep.py > gpurun_out/sweep.jsonl
Under torchrun (N ranks, one per GPU) every rank runs its own shard of B episodes, the delta tables are all-gathered for the
calibration like the reference does (value_net.py:195-201), times are the max over ranks and rank 0 prints the whole-job lines:
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep.py > gpurun_out/sweep_8gpu.jsonl"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=224)
    ap.add_argument("--seq", type=int, default=32)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--calib-steps", type=int, default=8)
    ap.add_argument("--max-layers", default="4,8,12,24")
    ap.add_argument("--exit-ratios", default="0.5,0.8,1.0,1.2,1.5,2.0")
    a = ap.parse_args()
    from deer_vla_b200 import config_3b
    from deer_vla_b200.engine import DeerEngine
    from deer_vla_b200.value_net import ExitController
    from deer_vla_b200.weights import synth_frames, synth_instruction_ids, synth_state_items
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = a.batch, a.seq
    frames = [(synth_frames(config_3b(), B, 1000 + 17 * (4 * rank + i), dev), synth_frames(config_3b(), B, 5000 + 17 * (4 * rank + i), dev)) for i in range(4)]
    for ml in [int(x) for x in a.max_layers.split(",")]:
        cfg = config_3b(ml)
        eng = DeerEngine(cfg, max_batch=B, max_slots=B, max_seq=S, device=local, precision="bf16")
        eng.load_state_dict(synth_state_items(cfg, 0))
        ids = synth_instruction_ids(cfg, B, S, seed=2000 + rank).to(dev)
        ctrl = ExitController(None, exit_id_list=eng.exit_list, steps_per_stage=1, exit_dist="exp", leq=True, max_layer=ml, engine=eng)
        eng.reset()
        cols = []
        for t in range(a.calib_steps):
            out = eng.step(frames[t % 4][0], frames[t % 4][1], ids, full_depth=True)
            cols.append(out["delta"].float().t().cpu())
        table = torch.cat(cols, dim=1)
        if world > 1:
            tl = [torch.zeros_like(table).to(dev) for _ in range(world)]
            dist.all_gather(tl, table.to(dev))
            table = torch.cat([x.cpu() for x in tl], dim=1)
        for er in [float(x) for x in a.exit_ratios.split(",")]:
            ctrl.set_threshold(None, None, None, er, cfg.llm_name, values=table)
            eng.reset()
            for t in range(a.warmup):
                eng.step(frames[t % 4][0], frames[t % 4][1], ids, want_delta=False)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            outs = []
            e0.record()
            for t in range(a.steps):
                outs.append(eng.step(frames[(a.warmup + t) % 4][0], frames[(a.warmup + t) % 4][1], ids, want_delta=False))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            hist = np.zeros(cfg.n_layers, dtype=np.int64)
            for o in outs:
                hist += np.bincount(o["exit_layer"].cpu().numpy(), minlength=cfg.n_layers)
            if world > 1:
                tt = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
                th = torch.from_numpy(hist).to(dev)
                dist.all_reduce(th)
                hist = th.cpu().numpy()
            exits = [e for e in cfg.exit_list if e <= cfg.ctrl_max_layer]
            if rank == 0:
                print(json.dumps({"max_layer": ml, "n_layers_built": cfg.n_layers, "exits": exits, "exit_ratio": er, "n_gpus": world,
                                  "action_steps_per_s": world * B * a.steps / (ms / 1e3), "ms_per_step": ms / a.steps,
                                  "mean_exit_depth_layers": float(sum((e + 1) * hist[e] for e in exits) / hist.sum()),
                                  "exit_histogram": {str(e): int(hist[e]) for e in exits},
                                  "thresholds": [ctrl.thresholds[e] for e in ctrl._real_exits()], "episodes_per_gpu": B, "seq_len": S}), flush=True)
        eng.close()
        del eng
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
