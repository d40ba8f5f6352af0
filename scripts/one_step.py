"""One DeeR-B action-step batch (B=224, S=32, golden thresholds) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/step_metrics.csv \
      python scripts/one_step.py
Prints the per-stage GEMM launch counts of that step (stage boundaries for scripts/ncu_stage_table.py)."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deer_vla_b200 import config_3b, config_9b
from deer_vla_b200.engine import DeerEngine
from deer_vla_b200.weights import synth_frames, synth_instruction_ids, synth_state_items

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=224)
ap.add_argument("--family", default="3b")
ap.add_argument("--out", default="gpurun_out/one_step.json")
a = ap.parse_args()
cfg = config_3b(12) if a.family == "3b" else config_9b(12)
B, S = a.batch, 32
eng = DeerEngine(cfg, max_batch=B, max_slots=B, max_seq=S, device=0, precision="bf16")
eng.load_state_dict(synth_state_items(cfg, 0))
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
if a.family == "3b":
    eng.set_thresholds(np.load(os.path.join(root, "tests", "golden", "full3b.npz"))["thresholds_B"].tolist())
else:
    eng.set_thresholds([-1e8, 0.02, 0.02, 0.02, 0.02, 1e8])
dev = torch.device("cuda", 0)
fr = [(synth_frames(cfg, B, 1000 + i, dev), synth_frames(cfg, B, 5000 + i, dev)) for i in range(3)]
ids = synth_instruction_ids(cfg, B, S, seed=2000).to(dev)
eng.reset()
for t in range(2):
    eng.step(fr[t][0], fr[t][1], ids, want_delta=False)
torch.cuda.synchronize()
l0 = eng.stats().kernel_launches_total
torch.cuda.profiler.start()
out = eng.step(fr[2][0], fr[2][1], ids, want_delta=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
l1 = eng.stats().kernel_launches_total
# same step again with the engine's own per-GEMM events to get the stage boundaries (identical launch sequence up to exits)
eng.set_profile(True)
eng.reset()
for t in range(3):
    eng.step(fr[t][0], fr[t][1], ids, want_delta=False)
p = eng.profile()
hist = np.bincount(out["exit_layer"].cpu().numpy(), minlength=cfg.n_layers).tolist()
json.dump({"launches_in_profiled_step": int(l1 - l0), "gemm_launches_by_stage": p["gemm_launches"], "gemm_ms_by_stage": p["gemm_ms"],
           "step_ms": p["step_ms"], "exit_histogram": hist, "batch": B, "family": a.family}, open(a.out, "w"))
print(open(a.out).read())
