"""Exit-head probe timing at small active sets (OpenFlamingo-3B): whole steps with the fused probe kernel vs the per-op path
(DEER_HEAD_FUSED=0), and the fused kernel's phase timeline (deer_head_debug: %globaltimer of CTA 0 at every grid barrier)."""
import ctypes as C
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deer_vla_b200 import _lib, config_3b
from deer_vla_b200.engine import DeerEngine
from deer_vla_b200.weights import synth_frames, synth_instruction_ids, synth_state_items

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
cfg = config_3b(12)
thr = np.load(os.path.join(root, "tests", "golden", "full3b.npz"))["thresholds_B"].tolist()
dev = torch.device("cuda", 0)
S = 32
sd = list(synth_state_items(cfg, 0))
res = {}
quick = "--quick" in sys.argv          # fused path only (cluster-size A/B via DEER_HEAD_CLUSTER)
for B in (28, 1):
    fr = [(synth_frames(cfg, B, 1000 + i, dev), synth_frames(cfg, B, 5000 + i, dev)) for i in range(4)]
    ids = synth_instruction_ids(cfg, B, S, seed=2000).to(dev)
    for fused in ((1,) if quick else (1, 0)):
        os.environ["DEER_HEAD_FUSED"] = str(fused)
        eng = DeerEngine(cfg, max_batch=B, max_slots=B, max_seq=S, device=0, precision="bf16")
        eng.load_state_dict(sd)
        eng.set_thresholds(thr)
        eng.reset()
        for t in range(6):
            eng.step(fr[t % 4][0], fr[t % 4][1], ids, want_delta=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for t in range(n):
            eng.step(fr[t % 4][0], fr[t % 4][1], ids, want_delta=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        # full-depth step (every exit probed: 7 head evaluations) for a clean per-probe figure
        for t in range(3):
            eng.step(fr[t % 4][0], fr[t % 4][1], ids, full_depth=True, want_delta=False)
        torch.cuda.synchronize()
        e0.record()
        for t in range(n):
            eng.step(fr[t % 4][0], fr[t % 4][1], ids, full_depth=True, want_delta=False)
        e1.record()
        torch.cuda.synchronize()
        ms_full = e0.elapsed_time(e1) / n
        res[f"B{B}_fused{fused}"] = {"ms_per_step_dynamic": round(ms, 4), "ms_per_step_full_depth": round(ms_full, 4),
                                      "launches_last_step": int(eng.stats().kernel_launches_last_step)}
        print(f"B={B} fused={fused}: dynamic {ms:.3f} ms/step, full depth {ms_full:.3f} ms/step, launches {eng.stats().kernel_launches_last_step}", flush=True)
        if fused:
            lib = _lib.load()
            lib.deer_head_debug(1, None, 0)
            eng.step(fr[0][0], fr[0][1], ids, full_depth=True, want_delta=False)
            buf = (C.c_longlong * 64)()
            lib.deer_head_debug(1, buf, 64)
            t = list(buf)[:16]
            names = ["prefetch+pool", "lstm0", "lstm1", "lstm2", "lstm3", "mlp1", "mlp2", "out"]
            line = []
            for k in range(8):
                work = t[2 * k + 1] - t[2 * k]
                wait = (t[2 * k + 2] - t[2 * k + 1]) if k < 7 else 0
                line.append(f"{names[k]} {work / 1e3:.1f}+{wait / 1e3:.1f}us")
            print(f"  head probe timeline (CTA 0; work + barrier wait): {', '.join(line)}; total {(t[15] - t[0]) / 1e3:.1f} us", flush=True)
            f = list(buf)[16:19]
            print(f"  lstm1 breakdown: stage+cluster sync {(f[0] - t[4]) / 1e3:.1f} us, gemm {(f[1] - f[0]) / 1e3:.1f} us, reduce+cell {(f[2] - f[1]) / 1e3:.1f} us "
                  f"(cluster size {os.environ.get('DEER_HEAD_CLUSTER', '8')})", flush=True)
            res[f"B{B}_head_timeline_ns"] = [int(x - t[0]) for x in t]
            lib.deer_head_debug(0, None, 0)
        eng.close()
    os.environ.pop("DEER_HEAD_FUSED", None)
print(json.dumps(res))
