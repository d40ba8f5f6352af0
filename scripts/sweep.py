"""C5 sweep: exit_ratio x max_layer on OpenFlamingo-3B, device-timed action-steps/s vs mean exit depth.
One engine per max_layer (the model is built with min(ckpt_early_exit_layer, max_layer) layers, eval_calvin.py:529);
one delta table per engine -- the reference's calibration pass in its training-window layout (deer_step_generate, 12-frame
windows, value_net.py:300-399) -- feeds the reference's greedy threshold solver for every exit_ratio (value_net.py:203-264).
Like bench.py the job holds --episodes (224) episodes IN TOTAL, sharded over the ranks (eval_utils.py:523-527), and the timed
steps are episode timesteps >= 5 with carried LSTM state.  Usage on the GPU box:  python scripts/sweep.py > gpurun_out/sweep.jsonl
Under torchrun (N ranks, one per GPU) the delta tables are all-gathered for the calibration like the reference does
(value_net.py:195-201), times are the max over ranks and rank 0 prints the whole-job lines:
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
    ap.add_argument("--episodes", type=int, default=224, help="episodes of the whole job, sharded over the ranks")
    ap.add_argument("--batch", type=int, default=None, help="override: episodes per GPU")
    ap.add_argument("--seq", type=int, default=32)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--calib-windows", type=int, default=2)
    ap.add_argument("--max-layers", default="4,8,12,24")
    ap.add_argument("--exit-ratios", default="0.5,0.8,1.0,1.2,1.5,2.0")
    a = ap.parse_args()
    from deer_vla_b200 import config_3b
    from deer_vla_b200.engine import DeerEngine
    from deer_vla_b200.value_net import ExitController, generate_action_values_window
    from deer_vla_b200.sharding import all_gather_values
    from deer_vla_b200.weights import synth_frames, synth_instruction_ids, synth_state_items
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = (a.batch if a.batch is not None else a.episodes // world), a.seq
    NF, WIN = 6, 12
    frames = [(synth_frames(config_3b(), B, 1000 + 17 * (NF * rank + i), dev), synth_frames(config_3b(), B, 5000 + 17 * (NF * rank + i), dev)) for i in range(NF)]
    lead_in = max(0, 5 - a.warmup)
    for ml in [int(x) for x in a.max_layers.split(",")]:
        cfg = config_3b(ml)
        eng = DeerEngine(cfg, max_batch=B, max_slots=B, max_seq=S, device=local, precision="bf16")
        eng.load_state_dict(synth_state_items(cfg, 0))
        ids = synth_instruction_ids(cfg, B, S, seed=2000 + rank).to(dev)
        ctrl = ExitController(None, exit_id_list=eng.exit_list, steps_per_stage=1, exit_dist="exp", leq=True, max_layer=ml, engine=eng)
        gen = torch.Generator().manual_seed(4242 + rank)
        cols = []
        for w in range(max(a.calib_windows, -(-400 // (6 * B * world)))):
            order = [(w * 5 + t) % NF for t in range(WIN)]
            win = {"rgb": torch.stack([frames[i][0] for i in order], dim=1), "gripper": torch.stack([frames[i][1] for i in order], dim=1), "ids": ids}
            cols.append(generate_action_values_window(eng, [win], window_size=WIN, generator=gen))
            del win
        table = all_gather_values(torch.cat(cols, dim=1))
        for er in [float(x) for x in a.exit_ratios.split(",")]:
            ctrl.set_threshold(None, None, None, er, cfg.llm_name, values=table)
            eng.reset()
            for t in range(lead_in + a.warmup):
                eng.step(frames[t % NF][0], frames[t % NF][1], ids, want_delta=False)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            outs = []
            e0.record()
            for t in range(a.steps):
                outs.append(eng.step(frames[(lead_in + a.warmup + t) % NF][0], frames[(lead_in + a.warmup + t) % NF][1], ids, want_delta=False))
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
                                  "thresholds": [ctrl.thresholds[e] for e in ctrl._real_exits()], "episodes_per_gpu": B, "episodes_total": B * world,
                                  "calibration_samples": int(table.shape[1]), "seq_len": S}), flush=True)
        eng.close()
        del eng
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
