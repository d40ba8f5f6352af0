"""Soak: thousands of consecutive dynamic-exit steps (fused head probe = thousands of cooperative launches on one barrier counter),
eager at 28 episodes and CUDA-graph at 1 episode; checks that nothing hangs, outputs stay finite and runs are reproducible."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deer_vla_b200 import config_3b
from deer_vla_b200.engine import DeerEngine
from deer_vla_b200.weights import synth_frames, synth_instruction_ids, synth_state_items

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
cfg = config_3b(12)
thr = np.load(os.path.join(root, "tests", "golden", "full3b.npz"))["thresholds_B"].tolist()
dev = torch.device("cuda", 0)
sd = list(synth_state_items(cfg, 0))
for B, graph, steps in ((28, False, 1500), (1, True, 3000)):
    fr = [(synth_frames(cfg, B, 1000 + i, dev), synth_frames(cfg, B, 5000 + i, dev)) for i in range(5)]
    ids = synth_instruction_ids(cfg, B, 32, seed=2000).to(dev)
    sums = []
    for rep in range(2):
        eng = DeerEngine(cfg, max_batch=B, max_slots=B, max_seq=32, device=0, precision="bf16", use_graph=graph)
        eng.load_state_dict(sd)
        eng.set_thresholds(thr)
        eng.reset()
        t0 = time.time()
        acc = torch.zeros(7, device=dev, dtype=torch.float64)
        hist = torch.zeros(cfg.n_layers, device=dev, dtype=torch.int64)
        for t in range(steps):
            o = eng.step(fr[t % 5][0], fr[t % 5][1], ids, want_delta=False)
            acc += o["action"].double().sum(0)
            hist += torch.bincount(o["exit_layer"].long(), minlength=cfg.n_layers)
            if t % 360 == 359:
                eng.reset()                                  # a new subtask (eval_utils.py:646)
        torch.cuda.synchronize()
        dt = time.time() - t0
        assert torch.isfinite(acc).all()
        sums.append((acc.cpu().numpy().copy(), hist.cpu().numpy().copy()))
        print(f"B={B} graph={graph} rep {rep}: {steps} steps in {dt:.1f} s ({1e3 * dt / steps:.2f} ms/step wall), exits {hist.cpu().tolist()}", flush=True)
        eng.close()
    assert np.array_equal(sums[0][0], sums[1][0]) and np.array_equal(sums[0][1], sums[1][1]), "runs differ"
    print(f"B={B} graph={graph}: two runs bit-identical over {steps} steps", flush=True)
