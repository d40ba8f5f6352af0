"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the committed golden
vectors generated from the unmodified reference (oracle/make_golden.py).

Tolerances (stated per test):
  * fp32 validation path: every tap <= 2e-4 abs, exit indices identical (bit-exact decisions).
  * bf16 product path: the north-star tolerance is "1e-3 bf16"; a bf16 stack of 24 ViT + 12 LLM
    layers cannot meet 1e-3 absolute on tanh outputs against an fp32 oracle, so "bf16 tolerance" is
    measured, not assumed: tests/golden/yardstick3b.npz holds the oracle run with bf16-rounded
    contraction operands (Y = 1.2e-2..1.8e-2 on poses vs the fp32 oracle at 3B size), and the engine
    must stay within 1.5*Y (+2e-3) of the fp32 oracle (test_full3b_bf16_within_rounding_yardstick;
    2.5*Y at 9B size).  The multi-step 3B comparisons against the reference's golden rollouts (asserted
    exact-match set, re-seeding after a rounding-level flip, fp32 path bit-exact on every decision) live
    in tests/test_gpu_round2.py.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import deer_oracle as O  # noqa: E402  (the checker)


def _engine(ocfg, W, **kw):
    from deer_vla_b200 import DeerConfig
    from deer_vla_b200.engine import DeerEngine
    eng = DeerEngine(DeerConfig(**ocfg.to_dict()), **kw)
    eng.load_state_dict(W)
    return eng


def _inputs(ocfg, episodes, t, S):
    ids = torch.stack([O.synth_tokens(ocfg, e, S) for e in episodes])
    rgb = torch.stack([O.synth_image(ocfg, e, t, 0) for e in episodes])
    grip = torch.stack([O.synth_image(ocfg, e, t, 1) for e in episodes])
    return rgb, grip, ids


@pytest.fixture(scope="module")
def tiny():
    ocfg = O.config_tiny()
    return ocfg, O.synth_weights(ocfg, 0)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,gemm,tol", [("fp32", "simt", 2e-4), ("bf16", "simt", 4e-2), ("bf16", "tcgen05", 4e-2)])
def test_tiny_stage_taps_vs_oracle(tiny, prec, gemm, tol):
    """Every stage of one step (ViT tokens, media, each layer's hidden state, pooled feature, pose,
    gripper, delta at every exit) against the oracle, full depth."""
    ocfg, W = tiny
    S = 12
    eng = _engine(ocfg, W, max_batch=2, max_seq=S, precision=prec, gemm_impl=gemm, debug_taps=True)
    eng.reset()
    rgb, grip, ids = _inputs(ocfg, [0, 1], 0, S)
    out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda(), full_depth=True)
    torch.cuda.synchronize()
    vit = eng.get_tap("vit")
    media = eng.get_tap("media")
    for e in range(2):
        orc = O.DeerOracle(ocfg, W)
        o = orc.step(rgb[e], grip[e], ids[e], full_depth=True, taps=True)
        scale = lambda t: max(1.0, float(t.abs().max()))
        assert np.abs(vit[e] - o.taps["vit_rgb"].numpy()).max() <= tol * scale(o.taps["vit_rgb"]), "vit rgb"
        assert np.abs(vit[2 + e] - o.taps["vit_gripper"].numpy()).max() <= tol * scale(o.taps["vit_gripper"]), "vit gripper"
        assert np.abs(media[e] - o.taps["media"].numpy()).max() <= tol * scale(o.taps["media"]), "media"
        for l in range(ocfg.ctrl_max_layer + 1):
            h = eng.get_tap("hidden", l)[e]
            ref = o.taps[f"hidden_{l}"]
            assert np.abs(h - ref.numpy()).max() <= tol * scale(ref), f"hidden {l}"
        assert np.abs(out["pose"][e].cpu().numpy() - o.pose.numpy()).max() <= tol
        assert abs(float(out["gripper"][e]) - float(o.gripper)) <= tol
        d = out["delta"][e].cpu().numpy()
        for k, ex in enumerate([x for x in ocfg.exit_list if x <= ocfg.ctrl_max_layer]):
            assert abs(d[k] - o.deltas[ex]) <= 2 * tol, (k, d[k], o.deltas[ex])
        assert int(out["exit_layer"][e]) == ocfg.ctrl_max_layer
        h, c = eng.get_state(e)
        assert np.abs(h - o.taps["h"].numpy()).max() <= tol and np.abs(c - o.taps["c"].numpy()).max() <= 2 * tol
    eng.close()


def _run_golden_episode(eng, ocfg, episode, T, S, slot=0):
    eng.reset([slot])
    res = []
    for t in range(T):
        rgb, grip, ids = _inputs(ocfg, [episode], t, S)
        out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda(), slots=torch.tensor([slot], dtype=torch.int32, device="cuda"),
                       timesteps=torch.tensor([t], dtype=torch.int32, device="cuda"))
        res.append({k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()})
    return res


def test_tiny_golden_dynamic_exit_fp32(tiny, gold_tiny):
    """Multi-step rollouts with dynamic exit and carried LSTM state vs the golden outputs of the
    UNMODIFIED reference MPTFlamingo + ExitController (tests/golden/tiny_dynamic.npz)."""
    ocfg, W = tiny
    g = gold_tiny
    S, T = int(g["S"]), int(g["T"])
    eng = _engine(ocfg, W, max_batch=1, max_slots=2, max_seq=S, precision="fp32")
    eng.set_thresholds(g["thresholds"].tolist())
    for e in g["episodes"].tolist():
        res = _run_golden_episode(eng, ocfg, e, T, S)
        for t in range(T):
            assert int(res[t]["exit_layer"][0]) == int(g[f"ep{e}_exit_layer"][t]), (e, t)
            assert np.abs(res[t]["pose"][0] - g[f"ep{e}_pose"][t]).max() <= 2e-4
            assert np.abs(res[t]["action"][0] - g[f"ep{e}_action"][t]).max() <= 2e-4
            gd = g[f"ep{e}_deltas"][t]
            m = ~np.isnan(gd)
            assert np.array_equal(np.isnan(res[t]["delta"][0]), ~m)            # same probes evaluated
            assert np.abs(res[t]["delta"][0][m] - gd[m]).max() <= 2e-4
        h, c = eng.get_state(0)
        assert np.abs(h - g[f"ep{e}_h_last"]).max() <= 3e-4 and np.abs(c - g[f"ep{e}_c_last"]).max() <= 1e-3
    eng.close()


def test_tiny_golden_batched_equals_single(tiny, gold_tiny):
    """All golden episodes advanced CONCURRENTLY in one batch (each exits at its own layer, survivors
    are compacted): every episode must get exactly the decision and action it gets alone."""
    ocfg, W = tiny
    g = gold_tiny
    S, T = int(g["S"]), int(g["T"])
    eps = g["episodes"].tolist()
    for host_control in (False, True):
        eng = _engine(ocfg, W, max_batch=len(eps), max_seq=S, precision="fp32", host_control=host_control)
        eng.set_thresholds(g["thresholds"].tolist())
        eng.reset()
        for t in range(T):
            rgb, grip, ids = _inputs(ocfg, eps, t, S)
            out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
            ex = out["exit_layer"].cpu().numpy()
            pose = out["pose"].cpu().numpy()
            for i, e in enumerate(eps):
                assert int(ex[i]) == int(g[f"ep{e}_exit_layer"][t]), (host_control, e, t, ex)
                assert np.abs(pose[i] - g[f"ep{e}_pose"][t]).max() <= 2e-4
        eng.close()


def test_tiny_golden_cuda_graph_conditional_exit(tiny, gold_tiny):
    """use_graph=1: the step is captured once into a CUDA graph whose post-exit segments are nested conditional IF nodes
    driven by exit_gate_kernel (cudaGraphSetConditional).  Step 0 runs eagerly, step 1 captures, later steps replay; every
    step must reproduce the reference's golden exits / poses, for a batch whose episodes leave at different exits."""
    ocfg, W = tiny
    g = gold_tiny
    S, T = int(g["S"]), int(g["T"])
    eps = g["episodes"].tolist()
    for prec, tol in (("fp32", 2e-4),):
        eng = _engine(ocfg, W, max_batch=len(eps), max_seq=S, precision=prec, use_graph=True)
        eng.set_thresholds(g["thresholds"].tolist())
        eng.reset()
        for t in range(T):
            rgb, grip, ids = _inputs(ocfg, eps, t, S)
            out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
            ex = out["exit_layer"].cpu().numpy()
            pose = out["pose"].cpu().numpy()
            for i, e in enumerate(eps):
                assert int(ex[i]) == int(g[f"ep{e}_exit_layer"][t]), (t, e, ex)
                assert np.abs(pose[i] - g[f"ep{e}_pose"][t]).max() <= tol
        eng.close()
    # single episode: when it exits at layer 1 every later layer sits inside a conditional node that must not run
    eng = _engine(ocfg, W, max_batch=1, max_seq=S, precision="fp32", use_graph=True)
    eng.set_thresholds(g["thresholds"].tolist())
    for e in eps[:2]:
        res = _run_golden_episode(eng, ocfg, e, T, S)
        for t in range(T):
            assert int(res[t]["exit_layer"][0]) == int(g[f"ep{e}_exit_layer"][t]), (e, t)
            assert np.abs(res[t]["pose"][0] - g[f"ep{e}_pose"][t]).max() <= 2e-4
    eng.close()


def test_tiny_steps_per_stage_and_max_layer(tiny, gold_tiny):
    """value_net.py:285-286 (steps_per_stage=2 reuses the previous exit id on odd steps) and the
    max_layer=4 truncation (exit list [1,3,4], forced exit at 3; eval_calvin.py:529, value_net.py:173)."""
    ocfg, W = tiny
    g = gold_tiny
    S, T = int(g["S"]), int(g["T"])
    c2 = O.config_tiny(steps_per_stage=2)
    eng = _engine(c2, W, max_batch=1, max_seq=S, precision="fp32")
    eng.set_thresholds(g["thresholds"].tolist())
    res = _run_golden_episode(eng, c2, 7, T, S)
    for t in range(T):
        assert int(res[t]["exit_layer"][0]) == int(g["sps2_ep7_exit_layer"][t]), t
        assert np.abs(res[t]["pose"][0] - g["sps2_ep7_pose"][t]).max() <= 2e-4
    eng.close()
    c3 = O.config_tiny(max_layer=4)
    assert c3.exit_list == [1, 3, 4] and c3.ctrl_max_layer == 3 and c3.n_thresholds == 2
    eng = _engine(c3, W, max_batch=1, max_seq=S, precision="fp32")
    assert eng.exit_list == [1, 3, 4]
    eng.set_thresholds(g["ml4_thresholds"].tolist())
    res = _run_golden_episode(eng, c3, 9, T, S)
    for t in range(T):
        assert int(res[t]["exit_layer"][0]) == int(g["ml4_ep9_exit_layer"][t]), t
        assert np.abs(res[t]["pose"][0] - g["ml4_ep9_pose"][t]).max() <= 2e-4
    eng.close()


def test_tiny_static_exit_and_padding(tiny):
    """exit_id mode (flamingo_mpt.py:396-405) and right-padded batches: an episode padded to a longer S
    must equal the un-padded batch-1 run."""
    ocfg, W = tiny
    S = 12
    eng = _engine(ocfg, W, max_batch=2, max_seq=16, precision="fp32")
    eng.reset()
    rgb, grip, ids = _inputs(ocfg, [0], 0, S)
    for exit_id in (0, 2, 5):
        orc = O.DeerOracle(ocfg, W)
        media, _ = O.encode_vision(ocfg, W, rgb, grip)
        x = W["lang_encoder.transformer.wte.weight"][ids]
        bias = O.mpt_attn_bias(ocfg, torch.ones(1, S, dtype=torch.int64))
        for l in range(exit_id + 1):
            x = O.flamingo_layer_forward(ocfg, W, l, x, media, ids == ocfg.media_token_id, bias)
        pose = O.head_forward(ocfg, W, x, None)[0]
        eng.reset()
        out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda(), exit_id=exit_id)
        assert int(out["exit_layer"][0]) == exit_id
        assert (out["pose"][0].cpu() - pose[0]).abs().max().item() <= 2e-4
    # padding: S=12 tokens padded with 4 pad ids, batched with a full-length episode
    eng.set_thresholds([0.0, 0.0, 1e8])
    ids16 = torch.cat([ids[0], torch.full((4,), 511, dtype=torch.int64)])[None]
    mask16 = torch.cat([torch.ones(12), torch.zeros(4)]).to(torch.uint8)[None]
    other = O.synth_tokens(ocfg, 5, 16)[None]
    rgb2, grip2, _ = _inputs(ocfg, [0, 5], 0, S)
    eng.reset()
    out = eng.step(rgb2.cuda(), grip2.cuda(), torch.cat([ids16, other]).cuda(),
                   mask=torch.cat([mask16, torch.ones(1, 16, dtype=torch.uint8)]).cuda())
    orc = O.DeerOracle(ocfg, W, [0.0, 0.0, 1e8])
    o = orc.step(rgb2[0], grip2[0], ids[0])
    assert int(out["exit_layer"][0]) == o.exit_layer
    assert (out["pose"][0].cpu() - o.pose).abs().max().item() <= 2e-4
    eng.close()


def test_errors_mirror_reference_asserts(tiny):
    from deer_vla_b200._lib import DeerError
    ocfg, W = tiny
    eng = _engine(ocfg, W, max_batch=1, max_seq=12, precision="fp32")
    rgb, grip, ids = _inputs(ocfg, [0], 0, 12)
    with pytest.raises(DeerError):           # "Please set thresholds before calling forward" (value_net.py:279)
        eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
    with pytest.raises(DeerError):           # len(thresholds) == real_num_exit (value_net.py:179)
        eng.set_thresholds([0.1])
    with pytest.raises(DeerError):           # 0 <= exit_id < n_layers (flamingo_mpt.py:449)
        eng.step(rgb.cuda(), grip.cuda(), ids.cuda(), exit_id=17)
    eng.close()


def test_tiny_9b_geometry_vs_oracle():
    """OpenFlamingo-9B structure (MPT-7B block: no q/k LayerNorm, norm_1/norm_2/ffn.* names; x-attn only every 4th layer;
    exits [1,3,...,12] with max_layer=12 forcing the exit at 11) at reduced dims, fp32 path vs the oracle."""
    ocfg = O.config_tiny(d_model=256, n_heads=2, qk_ln=False, cross_attn_every=4, llm_name="mpt_9b",
                         ckpt_early_exit_layer=15, max_layer=12)
    assert ocfg.exit_list == [1, 3, 5, 7, 9, 11, 12] and ocfg.ctrl_max_layer == 11 and ocfg.n_thresholds == 6
    W = O.synth_weights(ocfg, 2)
    S = 10
    eng = _engine(ocfg, W, max_batch=2, max_seq=S, precision="fp32", debug_taps=True)
    assert eng.exit_list == ocfg.exit_list
    thr = [-1e8, 0.2, 0.2, 0.2, 0.2, 1e8]            # p_0 = 0 for mpt_9b: the first exit never fires (value_net.py:235-236)
    eng.set_thresholds(thr)
    eng.reset()
    for t in range(2):
        rgb, grip, ids = _inputs(ocfg, [0, 1], t, S)
        out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
        for e in range(2):
            if t == 0:
                orc = O.DeerOracle(ocfg, W, thr)
                o = orc.step(rgb[e], grip[e], ids[e], taps=True)
                for l in range(o.exit_layer + 1):
                    assert np.abs(eng.get_tap("hidden", l)[e] - o.taps[f"hidden_{l}"].numpy()).max() <= 3e-4, l
                assert int(out["exit_layer"][e]) == o.exit_layer and o.exit_layer >= 3
                assert np.abs(out["pose"][e].cpu().numpy() - o.pose.numpy()).max() <= 2e-4
    eng.close()


# ------------------------------------------------------------------------------------------------
# OpenFlamingo-3B size
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full3b():
    ocfg = O.config_3b(12)
    return ocfg, O.synth_weights(ocfg, 0)


# (the 2-episode x 3-step bf16-vs-golden tests of round 1 are superseded by tests/test_gpu_round2.py: 8 episodes x 10 steps,
#  asserted exact-match set, re-seeding after a rounding-level flip, fp32 path bit-exact on every decision)


def test_full3b_bf16_within_rounding_yardstick(full3b):
    """'bf16 tolerance' made concrete: the oracle with bf16-rounded contraction operands (fp32
    accumulate, everything else fp32) differs from the fp32 oracle by Y on these inputs; the engine's bf16 path must be
    within 1.5*Y (+2e-3) of the fp32 oracle -- measured 1.0 .. 1.4*Y -- and within 2.5*Y of the bf16-operand oracle (two
    independent roundings of the same computation differ by up to ~2*Y), on poses and on the gate's deltas."""
    ocfg, W = full3b
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "yardstick3b.npz"))
    S, T = int(g["S"]), int(g["T"])
    eng = _engine(ocfg, W, max_batch=2, max_seq=S, precision="bf16", gemm_impl="tcgen05")
    eng.set_thresholds([1e8] * ocfg.n_thresholds)
    eng.reset()
    for t in range(T):
        rgb, grip, ids = _inputs(ocfg, [0, 1], t, S)
        out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda(), full_depth=True)
        pose, delta = out["pose"].cpu().numpy(), out["delta"].cpu().numpy()
        for e in range(2):
            for name, got in (("pose", pose[e]), ("deltas", delta[e])):
                f32, b16 = g[f"fp32_ep{e}_{name}"][t], g[f"bf16ops_ep{e}_{name}"][t]
                yard = float(np.abs(b16 - f32).max())
                err32, err16 = float(np.abs(got - f32).max()), float(np.abs(got - b16).max())
                print(f"\n[yardstick] ep{e} t{t} {name}: Y={yard:.3e} |engine-fp32|={err32:.3e} |engine-bf16ops|={err16:.3e}")
                assert err32 <= 1.5 * yard + 2e-3 and err16 <= 2.5 * yard + 2e-3, (e, t, name, yard, err32, err16)
    eng.close()


def test_full3b_batch_invariance(full3b, gold_full):
    """Size-independent property at the benchmark's shape: an episode's result is BIT-IDENTICAL whether it
    runs alone or inside a batch of 24 with compaction after every exit (row-independent kernels)."""
    ocfg, W = full3b
    S = int(gold_full["S"])
    thr = gold_full["thresholds_B"].tolist()
    eng = _engine(ocfg, W, max_batch=24, max_seq=S, precision="bf16")
    eng.set_thresholds(thr)
    eps = list(range(24))
    eng.reset()
    rgb, grip, ids = _inputs(ocfg, eps, 0, S)
    big = eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
    big = {k: v.cpu() for k, v in big.items()}
    assert len(set(big["exit_layer"].tolist())) > 1, "want a mix of exit depths in the batch"
    for i in (0, 7, 23):
        eng.reset()
        one = eng.step(rgb[i:i + 1].cuda(), grip[i:i + 1].cuda(), ids[i:i + 1].cuda())
        assert int(one["exit_layer"][0]) == int(big["exit_layer"][i])
        assert torch.equal(one["pose"][0].cpu(), big["pose"][i]), i
    eng.close()


def test_full3b_launch_overlap_is_bit_identical(full3b, gold_full):
    """Programmatic dependent launch only moves kernel START times: three consecutive 3B steps of 8 episodes (carried LSTM
    state, compaction after every exit) give bit-identical poses, grippers, deltas and exit layers with it on and off."""
    from deer_vla_b200.engine import DeerEngine
    ocfg, W = full3b
    S = int(gold_full["S"])
    thr = gold_full["thresholds_B"].tolist()
    eng = _engine(ocfg, W, max_batch=8, max_seq=S, precision="bf16")
    eng.set_thresholds(thr)
    eps = list(range(8))
    runs = {}
    prev = DeerEngine.set_launch_overlap(True)
    try:
        for mode in (True, False):
            DeerEngine.set_launch_overlap(mode)
            eng.reset()
            outs = []
            for t in range(3):
                rgb, grip, ids = _inputs(ocfg, eps, t, S)
                o = eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
                outs.append({k: v.cpu().clone() for k, v in o.items()})
            runs[mode] = outs
    finally:
        DeerEngine.set_launch_overlap(prev)
    for a, b in zip(runs[True], runs[False]):
        for k in ("pose", "gripper", "exit_layer", "delta"):
            if a.get(k) is not None:      # deltas of exits an episode never reached are NaN in both runs
                assert torch.equal(torch.nan_to_num(a[k].float(), nan=-7.0), torch.nan_to_num(b[k].float(), nan=-7.0)), k
    eng.close()


def test_full3b_snake_row_order_is_bit_identical(full3b, gold_full):
    """LayerNorm rows and ViT attention items are walked from the END by default, so that they start on what the producing GEMM left
    in L2 (deer_set_row_order, include/deer_engine.h).  The kernels are row-independent: three consecutive 3B steps of 8 episodes
    (carried LSTM state, compaction after every exit) give bit-identical poses, grippers, deltas and exit layers in both orders."""
    from deer_vla_b200.engine import DeerEngine
    ocfg, W = full3b
    S = int(gold_full["S"])
    thr = gold_full["thresholds_B"].tolist()
    eng = _engine(ocfg, W, max_batch=8, max_seq=S, precision="bf16")
    eng.set_thresholds(thr)
    eps = list(range(8))
    runs = {}
    prev = DeerEngine.set_row_order(True)
    try:
        for mode in (True, False):
            DeerEngine.set_row_order(mode)
            eng.reset()
            outs = []
            for t in range(3):
                rgb, grip, ids = _inputs(ocfg, eps, t, S)
                o = eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
                outs.append({k: v.cpu().clone() for k, v in o.items()})
            runs[mode] = outs
    finally:
        DeerEngine.set_row_order(prev)
    for a, b in zip(runs[True], runs[False]):
        for k in ("pose", "gripper", "exit_layer", "delta"):
            if a.get(k) is not None:
                assert torch.equal(torch.nan_to_num(a[k].float(), nan=-7.0), torch.nan_to_num(b[k].float(), nan=-7.0)), k
    eng.close()


# ------------------------------------------------------------------------------------------------
# OpenFlamingo-9B size (BASELINE configs[3])
# ------------------------------------------------------------------------------------------------
def test_full9b_bf16_within_rounding_yardstick():
    """OpenFlamingo-9B / MPT-7B at FULL size (d_model 4096, 32 heads x 128, MLP 16384, x-attn at layers 3, 7, 11; 13 layers built
    for max_layer=12) on the tcgen05 product path: two consecutive steps of two episodes (carried LSTM state), full depth, so
    every exit's delta and the final pose are compared and no threshold decision can flip.  Bar = the measured rounding
    yardstick of THIS model size (tests/golden/yardstick9b.npz, `python -m oracle.make_golden --yardstick-only --family 9b`):
    the oracle with bf16-rounded contraction operands differs from the fp32 oracle by Y (2.5e-2 .. 3.9e-2 on poses here, twice
    the 3B figure -- the stack is twice as wide); the engine must be within 2.5*Y of both (measured <= 2.1*Y).  The MPT-7B block itself is cross-checked
    against transformers.models.mpt (tests/golden/xcheck.json); the reference's own modeling_gpt_9b.py cannot be driven here
    (13 un-vendored sibling modules), so this checks the engine against the restatement only."""
    ocfg = O.config_9b(12)
    W = O.synth_weights(ocfg, 0)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "yardstick9b.npz"))
    S, T = int(g["S"]), int(g["T"])
    eng = _engine(ocfg, W, max_batch=2, max_seq=S, precision="bf16", gemm_impl="tcgen05")
    eng.reset()
    exits = [x for x in ocfg.exit_list if x <= ocfg.ctrl_max_layer]
    for t in range(T):
        rgb, grip, ids = _inputs(ocfg, [0, 1], t, S)
        out = eng.step(rgb.cuda(), grip.cuda(), ids.cuda(), full_depth=True)
        pose, delta, ex = out["pose"].cpu().numpy(), out["delta"].cpu().numpy(), out["exit_layer"].cpu().tolist()
        assert ex == [exits[-1]] * 2
        for e in range(2):
            for name, got in (("pose", pose[e]), ("deltas", delta[e])):
                f32, b16 = g[f"fp32_ep{e}_{name}"][t], g[f"bf16ops_ep{e}_{name}"][t]
                yard = float(np.abs(b16 - f32).max())
                err32, err16 = float(np.abs(got - f32).max()), float(np.abs(got - b16).max())
                print(f"\n[yardstick 9b] ep{e} t{t} {name}: Y={yard:.3e} |engine-fp32|={err32:.3e} |engine-bf16ops|={err16:.3e}")
                assert err32 <= 2.5 * yard + 2e-3 and err16 <= 2.5 * yard + 2e-3, (e, t, name, yard, err32, err16)   # measured <= 2.1*Y
    eng.close()


# ------------------------------------------------------------------------------------------------
# threshold calibration in the training layout
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 1.5e-2)])
def test_generate_mode_vs_reference_golden(prec, tol):
    """deer_step_generate over 11 frames of 2 windows == the unmodified reference's generate_action_values table
    (golden made by oracle/make_golden.py::golden_generate with the layer indices the reference drew)."""
    from deer_vla_b200.value_net import generate_action_values_window
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "generate_tiny.npz"))
    ocfg = O.config_tiny()
    W = O.synth_weights(ocfg, 0)
    win, S = int(g["window"]), int(g["S"])
    eps = [int(e) for e in g["episodes"]]
    eng = _engine(ocfg, W, max_batch=len(eps), max_seq=S, precision=prec)
    rgb = torch.stack([torch.stack([O.synth_image(ocfg, e, t, 0) for t in range(win)]) for e in eps]).cuda()
    grip = torch.stack([torch.stack([O.synth_image(ocfg, e, t, 1) for t in range(win)]) for e in eps]).cuda()
    ids = torch.stack([O.synth_tokens(ocfg, e, S) for e in eps]).cuda()
    li = torch.from_numpy(g["layer_idx"]).long()
    table = generate_action_values_window(eng, [{"rgb": rgb, "gripper": grip, "ids": ids}], window_size=win, layer_indices=[li])
    want = torch.from_numpy(g["deltas"])
    assert table.shape == want.shape
    err = float((table - want).abs().max())
    print(f"\n[generate {prec}] max |delta - reference| = {err:.3e} (table max {float(want.max()):.3e})")
    assert err <= tol, err
    # the dynamic-exit path is untouched by a calibration pass: same thresholds -> same step as a fresh engine
    eng.set_thresholds([float(x) for x in want.mean(1)[: ocfg.n_thresholds - 1]] + [1e8])
    eng.reset()
    a = eng.step(rgb[:, 0], grip[:, 0], ids)
    eng2 = _engine(ocfg, W, max_batch=len(eps), max_seq=S, precision=prec)
    eng2.set_thresholds(eng.thresholds)
    eng2.reset()
    b = eng2.step(rgb[:, 0], grip[:, 0], ids)
    assert torch.equal(a["exit_layer"], b["exit_layer"]) and torch.equal(a["action"], b["action"])
    eng.close(); eng2.close()
