import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
from deer_vla_b200 import DeerConfig
from deer_vla_b200.engine import DeerEngine
from oracle import deer_oracle as O
from oracle import pil_resize as R
from oracle.make_golden_preprocess import frames

def run(S, max_seq, host_control=False):
    ocfg = O.config_tiny(image_size=224)
    W = O.synth_weights(ocfg, 0)
    thr = [0.3, 0.25, 1e8]
    eng = DeerEngine(DeerConfig(**ocfg.to_dict()), max_batch=2, max_seq=max_seq, precision="fp32", host_control=host_control)
    eng.load_state_dict(W); eng.set_thresholds(thr)
    rgb = np.stack([frames(1, 200, 200), frames(2, 200, 200)]); grip = np.stack([frames(3, 84, 84), frames(4, 84, 84)])
    a = eng.preprocess(torch.from_numpy(rgb).cuda()); b = eng.preprocess(torch.from_numpy(grip).cuda())
    ids = torch.stack([O.synth_tokens(ocfg, 5, S), O.synth_tokens(ocfg, 6, S)])
    print(f"--- S={S} max_seq={max_seq} host_control={host_control}")
    for e in range(2):
        orc = O.DeerOracle(ocfg, W, thr); orc.reset()
        o = orc.step(a[e].cpu(), b[e].cpu(), ids[e])
        print("oracle ep", e, "exit", o.exit_layer, "deltas", {k: round(float(v), 5) for k, v in o.deltas.items()})
    for rep in range(2):
        eng.reset()
        d = eng.step(a, b, ids.cuda())
        torch.cuda.synchronize()
        print("dev ", d["exit_layer"].tolist(), d["delta"].tolist())
        eng.reset()
        h = eng.step(a.cpu().pin_memory(), b.cpu().pin_memory(), ids)
        print("host", h["exit_layer"].tolist(), h["delta"].tolist())
        eng.reset()
        u = eng.step_u8(torch.from_numpy(rgb).pin_memory(), torch.from_numpy(grip).pin_memory(), ids)
        print("u8  ", u["exit_layer"].tolist(), u["delta"].tolist())
    eng.close()

run(12, 16); run(16, 16); run(12, 16, True)
