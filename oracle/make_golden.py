"""Generate the committed golden fixtures under ``tests/golden/`` (BUILD CONTAINER ONLY:
needs the read-only reference checkout at /root/reference; test infrastructure).

    python oracle/make_golden.py [--full]   # --full also writes the 3B-size fixtures (minutes)

What it does
  1. tiny config: drives the UNMODIFIED reference ``MPTFlamingo.forward`` +
     ``ExitController`` (robot_flamingo/models/flamingo_mpt.py:308, value_net.py:277) through
     several episodes x timesteps with dynamic exit, asserts the oracle restatement
     (oracle/deer_oracle.py) reproduces it (exit ids identical, tensors <= 2e-5), and stores
     the reference outputs (+ per-layer hidden states of one step) as ``tiny_dynamic.npz``.
  2. threshold solver: reference ``ExitController.set_threshold(values=...)``
     (value_net.py:185-272) vs ``solve_thresholds`` -> ``thresholds.npz``.
  3. un-vendored pieces cross-checked against in-image third-party code (evidence, not a
     pin): ViT vs ``transformers`` CLIPVisionModel, MPT-7B block + ALiBi vs
     ``transformers.models.mpt`` -> printed, and stored as ``xcheck.json``.
  4. --full: OpenFlamingo-3B size (DeeR-B max_layer=12 and DeeR-S max_layer=4): reference vs
     oracle on 2 episodes x 3 steps, a delta table from a batched full-depth oracle sweep,
     thresholds for exit_ratio 1.0 / 0.8 -> ``full3b_*.npz``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import deer_oracle as O          # noqa: E402
from oracle import ref_shims                 # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run_reference_episode(cfg, model, ctrl, episode, T, S, capture_hidden_at=None, ensemble=False):
    """ModelWrapper.reset + T x ModelWrapper.step (eval_utils.py:252-277, 279-490) on the
    reference modules; returns per-step dicts."""
    model.clear_all_exit_memory()
    ctrl.value_net.hidden_state = None
    ctrl.value_net.history_memory = []
    ids = O.synth_tokens(cfg, episode, S)
    outs = []
    # capture deltas by wrapping the value net's forward
    deltas_log = []
    orig = ctrl.value_net.forward

    def wrapped(feats, i=None, **kw):
        d = orig(feats, i, **kw)
        deltas_log.append((i, float(d.flatten()[0])))
        return d
    ctrl.value_net.forward = wrapped
    try:
        for t in range(T):
            rgb = O.synth_image(cfg, episode, t, 0)
            grip = O.synth_image(cfg, episode, t, 1)
            ctrl.set_timestep(t)
            model.set_all_exit_window_size(1)
            ctrl.value_net.window_size = 1
            deltas_log.clear()
            with torch.no_grad():
                out = model(vision_x=rgb[None, None, None], lang_x=ids[None],
                            attention_mask=torch.ones(1, S, dtype=torch.long),
                            vision_gripper=grip[None, None, None], state_tensor=None, return_feature=True,
                            deterministic=True, exit_id=None, dynamic_early_exit=True, exit_controller=ctrl)
            # eval_utils.py:458,464
            action = torch.concat((out.logits[0], out.logits[1] > 0.5), dim=2).squeeze(0)[-1]
            action[-1] = (action[-1] - 0.5) * 2
            if ensemble:                      # use_action_ensemble (eval_utils.py:460-464)
                ens = ctrl.value_net.get_ensemble_action()
                ctrl.value_net.reset_actions()
                action = torch.concat((ens[0], ens[1] > 0.5), dim=2).squeeze(0)[-1]
                action[-1] = (action[-1] - 0.5) * 2
            rec = dict(pose=out.logits[0].flatten().clone(), gripper=out.logits[1].flatten().clone(),
                       feature=out.logits[2].flatten().clone(),
                       action=action.clone(), exit_layer=int(out.exit_layer), deltas=dict(deltas_log),
                       h=model.extra_exit.hidden_state[0][:, 0].clone(), c=model.extra_exit.hidden_state[1][:, 0].clone())
            if capture_hidden_at is not None and t == capture_hidden_at:
                rec["hidden"] = [h[0].clone() for h in out.hidden_states]
                rec["media"] = model.lang_encoder._get_decoder_layers()[0].vis_x[0, 0].clone()
            outs.append(rec)
    finally:
        ctrl.value_net.forward = orig
    return outs


def run_oracle_episode(cfg, W, thresholds, episode, T, S, taps_at=None):
    orc = O.DeerOracle(cfg, W, thresholds)
    orc.reset()
    ids = O.synth_tokens(cfg, episode, S)
    outs = []
    for t in range(T):
        orc.set_timestep(t)
        o = orc.step(O.synth_image(cfg, episode, t, 0), O.synth_image(cfg, episode, t, 1), ids,
                     taps=(taps_at is not None and t == taps_at))
        outs.append(o)
    return outs, orc


def batched_delta_table(cfg, W, episodes, T, S):
    """Full-depth delta table (n_thresholds, len(episodes)*T) with the oracle's batched
    functions (all episodes advance together; the last exit's LSTM state is committed)."""
    B = len(episodes)
    ids = torch.stack([O.synth_tokens(cfg, e, S) for e in episodes])
    mask = torch.ones(B, S, dtype=torch.int64)
    state = None
    exits = [e for e in cfg.exit_list if e <= cfg.ctrl_max_layer]
    table = torch.zeros(len(exits), B * T)
    with torch.no_grad():
        for t in range(T):
            rgb = torch.stack([O.synth_image(cfg, e, t, 0) for e in episodes])
            grip = torch.stack([O.synth_image(cfg, e, t, 1) for e in episodes])
            media, _ = O.encode_vision(cfg, W, rgb, grip)
            x = W["lang_encoder.transformer.wte.weight"][ids]
            bias = O.mpt_attn_bias(cfg, mask)
            ml = ids == cfg.media_token_id
            prev_x, prev_action, last = None, None, None
            for l in range(exits[-1] + 1):
                prev_x = x
                x = O.flamingo_layer_forward(cfg, W, l, x, media, ml, bias)
                if l in exits:
                    if l - cfg.exit_interval < 0:
                        prev_action = O.head_forward(cfg, W, prev_x, state)[0]
                    last = O.head_forward(cfg, W, x, state)
                    table[exits.index(l), t * B:(t + 1) * B] = O.get_delta(last[0], prev_action, cfg.threshold_type)
                    prev_action = last[0]
            state = last[3]
    return table


def compare(ref_steps, orc_steps, tol, tag):
    worst = 0.0
    for t, (r, o) in enumerate(zip(ref_steps, orc_steps)):
        assert r["exit_layer"] == o.exit_layer, (tag, t, r["exit_layer"], o.exit_layer)
        assert sorted(r["deltas"]) == sorted(o.deltas), (tag, t, r["deltas"], o.deltas)
        for k in r["deltas"]:
            worst = max(worst, abs(r["deltas"][k] - o.deltas[k]))
        worst = max(worst, float((r["pose"] - o.pose).abs().max()), float((r["gripper"] - o.gripper).abs().max()))
        assert torch.equal(r["action"][-1:], o.action[-1:]), (tag, t)
    assert worst <= tol, (tag, worst)
    return worst


def pack_steps(steps, exits):
    n = len(steps)
    d = np.full((n, len(exits)), np.nan, np.float32)
    for t, r in enumerate(steps):
        for k, v in r["deltas"].items():
            d[t, exits.index(k)] = v
    return dict(pose=np.stack([r["pose"].numpy() for r in steps]),
                gripper=np.stack([r["gripper"].numpy() for r in steps]),
                action=np.stack([r["action"].numpy() for r in steps]),
                exit_layer=np.array([r["exit_layer"] for r in steps], np.int32),
                deltas=d,
                # LSTM state after the LAST step only (it integrates every earlier commit)
                h_last=steps[-1]["h"].numpy(), c_last=steps[-1]["c"].numpy())


# ---------------------------------------------------------------------------------------
def golden_tiny():
    cfg = O.config_tiny()
    W = O.synth_weights(cfg, 0)
    S, T, EPIS = 12, 6, [0, 1, 2, 3]
    table = batched_delta_table(cfg, W, list(range(100, 116)), 4, S)
    thr = O.solve_thresholds(table, 1.0)
    model, ctrl, _ = ref_shims.build_reference_model(cfg, W, thr)
    exits = [e for e in cfg.exit_list if e <= cfg.ctrl_max_layer]
    out = {"thresholds": np.array(thr, np.float32), "S": S, "T": T, "episodes": np.array(EPIS),
           "delta_table": table.numpy()}
    worst = 0.0
    hist = {}
    for e in EPIS:
        ref = run_reference_episode(cfg, model, ctrl, e, T, S, capture_hidden_at=1 if e == 0 else None)
        orc, _ = run_oracle_episode(cfg, W, thr, e, T, S, taps_at=1 if e == 0 else None)
        worst = max(worst, compare(ref, orc, 2e-5, f"tiny ep{e}"))
        for k, v in pack_steps(ref, exits).items():
            out[f"ep{e}_{k}"] = v
        for r in ref:
            hist[r["exit_layer"]] = hist.get(r["exit_layer"], 0) + 1
        if e == 0:
            for i, h in enumerate(ref[1]["hidden"]):
                d = float((h - orc[1].taps[f"hidden_{i}"]).abs().max())
                assert d < 2e-5, (i, d)
                out[f"ep0_t1_hidden_{i}"] = h.numpy()
            assert float((ref[1]["media"] - orc[1].taps["media"]).abs().max()) < 2e-5
            out["ep0_t1_media"] = ref[1]["media"].numpy()
            out["ep0_t1_vit_rgb"] = orc[1].taps["vit_rgb"].numpy()
    print(f"[tiny] reference vs oracle: exit ids identical on {len(EPIS) * T} steps, worst |diff| {worst:.2e}; exit hist {hist}")
    # steps_per_stage = 2 variant (value_net.py:285-286)
    cfg2 = O.config_tiny(steps_per_stage=2)
    model2, ctrl2, _ = ref_shims.build_reference_model(cfg2, W, thr)
    ref = run_reference_episode(cfg2, model2, ctrl2, 7, T, S)
    orc, _ = run_oracle_episode(cfg2, W, thr, 7, T, S)
    w2 = compare(ref, orc, 2e-5, "tiny sps2")
    for k, v in pack_steps(ref, exits).items():
        out[f"sps2_ep7_{k}"] = v
    # max_layer = 4 variant (DeeR-S style truncation: exit list [1,3,4], forced exit at 3)
    cfg3 = O.config_tiny(max_layer=4)
    table3 = table[:cfg3.n_thresholds]
    thr3 = O.solve_thresholds(table3, 0.8)
    model3, ctrl3, _ = ref_shims.build_reference_model(cfg3, W, thr3)
    assert model3.get_all_exit_idx() == cfg3.exit_list and ctrl3.max_layer == cfg3.ctrl_max_layer
    ref = run_reference_episode(cfg3, model3, ctrl3, 9, T, S)
    orc, _ = run_oracle_episode(cfg3, W, thr3, 9, T, S)
    w3 = compare(ref, orc, 2e-5, "tiny ml4")
    exits3 = [e for e in cfg3.exit_list if e <= cfg3.ctrl_max_layer]
    for k, v in pack_steps(ref, exits3).items():
        out[f"ml4_ep9_{k}"] = v
    out["ml4_thresholds"] = np.array(thr3, np.float32)
    print(f"[tiny] steps_per_stage=2 worst {w2:.2e}; max_layer=4 worst {w3:.2e} exits {[r['exit_layer'] for r in ref]}")
    np.savez_compressed(os.path.join(GOLD, "tiny_dynamic.npz"), **out)


def golden_thresholds():
    ref_shims.install_shims()
    from robot_flamingo.models.value_net import ExitController
    g = torch.Generator().manual_seed(5)
    out = {}
    for name, n_exit, llm in (("3b", 6, "mpt_dolly_3b"), ("3b_s", 2, "mpt_dolly_3b"), ("9b", 6, "mpt_9b")):
        values = torch.rand(n_exit, 512, generator=g) ** 2
        out[f"{name}_values"] = values.numpy()
        for ratio in (0.5, 0.8, 1.0, 1.5, 2.0):
            for dist in ("exp", "gauss", "gamma"):
                exit_list = list(range(1, 2 * n_exit, 2))
                ctrl = ExitController(None, exit_id_list=exit_list, steps_per_stage=1, exit_dist=dist, leq=True,
                                      max_layer=exit_list[-1] + 1)
                args = types.SimpleNamespace(rank=1)
                ctrl.set_threshold(args, None, None, ratio, llm, values.clone())
                ref_T = [float(ctrl.thresholds[e]) for e in exit_list]
                mine = O.solve_thresholds(values.clone(), ratio, dist, llm)
                assert np.allclose(ref_T, mine, rtol=0, atol=0), (name, ratio, dist, ref_T, mine)
                out[f"{name}_{dist}_{ratio}"] = np.array(ref_T, np.float64)
    np.savez_compressed(os.path.join(GOLD, "thresholds.npz"), **out)
    print(f"[thresholds] reference ExitController.set_threshold == solve_thresholds on {len(out) - 3} cases (exact)")


def xcheck_third_party():
    """Evidence for the un-vendored restatements (not a pin): compare with transformers' CLIP and MPT."""
    res = {}
    from transformers import CLIPVisionConfig, CLIPVisionModel
    cfg = O.config_tiny(image_size=56, vit_width=128, vit_layers=3, vit_heads=2, vit_mlp=512)
    W = O.synth_weights(cfg, 3)
    hc = CLIPVisionConfig(hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
                          image_size=56, patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    hf = CLIPVisionModel(hc).eval()
    sd = {}
    p = "vision_encoder.visual."
    sd["vision_model.embeddings.class_embedding"] = W[p + "class_embedding"]
    sd["vision_model.embeddings.patch_embedding.weight"] = W[p + "conv1.weight"]
    sd["vision_model.embeddings.position_embedding.weight"] = W[p + "positional_embedding"]
    sd["vision_model.pre_layrnorm.weight"] = W[p + "ln_pre.weight"]
    sd["vision_model.pre_layrnorm.bias"] = W[p + "ln_pre.bias"]
    for i in range(3):
        q = f"{p}transformer.resblocks.{i}."
        h = f"vision_model.encoder.layers.{i}."
        wq, wk, wv = W[q + "attn.in_proj_weight"].chunk(3, 0)
        bq, bk, bv = W[q + "attn.in_proj_bias"].chunk(3, 0)
        sd.update({h + "self_attn.q_proj.weight": wq, h + "self_attn.k_proj.weight": wk, h + "self_attn.v_proj.weight": wv,
                   h + "self_attn.q_proj.bias": bq, h + "self_attn.k_proj.bias": bk, h + "self_attn.v_proj.bias": bv,
                   h + "self_attn.out_proj.weight": W[q + "attn.out_proj.weight"], h + "self_attn.out_proj.bias": W[q + "attn.out_proj.bias"],
                   h + "layer_norm1.weight": W[q + "ln_1.weight"], h + "layer_norm1.bias": W[q + "ln_1.bias"],
                   h + "layer_norm2.weight": W[q + "ln_2.weight"], h + "layer_norm2.bias": W[q + "ln_2.bias"],
                   h + "mlp.fc1.weight": W[q + "mlp.c_fc.weight"], h + "mlp.fc1.bias": W[q + "mlp.c_fc.bias"],
                   h + "mlp.fc2.weight": W[q + "mlp.c_proj.weight"], h + "mlp.fc2.bias": W[q + "mlp.c_proj.bias"]})
    miss = hf.load_state_dict(sd, strict=False)
    assert all("post_layernorm" in k or "position_ids" in k for k in miss.missing_keys), miss
    img = torch.stack([O.synth_image(cfg, 0, 0, 0), O.synth_image(cfg, 1, 0, 1)])
    with torch.no_grad():
        ref = hf(pixel_values=img).last_hidden_state[:, 1:]
        mine = O.vit_forward(cfg, W, img)
    res["vit_vs_hf_clip_maxabs"] = float((ref - mine).abs().max())
    # restated nn.Module ViT (nn.MultiheadAttention based) vs functional oracle
    vis = ref_shims.OpenClipVisualRestated(56, 14, 128, 3, 2, 512)
    vis.load_state_dict({k[len(p):]: v for k, v in W.items() if k.startswith(p)}, strict=False)
    with torch.no_grad():
        res["vit_vs_nnMHA_maxabs"] = float((vis.eval()(img)[1] - mine).abs().max())
    # MPT-7B style block vs transformers MptBlock
    from transformers.models.mpt.configuration_mpt import MptConfig
    from transformers.models.mpt.modeling_mpt import MptBlock, build_mpt_alibi_tensor
    c7 = O.config_tiny(d_model=256, n_heads=2, qk_ln=False, cross_attn_every=4, llm_name="mpt_9b")
    W7 = O.synth_weights(c7, 4)
    mc = MptConfig(d_model=256, n_heads=2, n_layers=1, expansion_ratio=4, max_seq_len=64, vocab_size=512, no_bias=True,
                   layer_norm_epsilon=1e-5)
    blk = MptBlock(mc, layer_idx=0).eval()
    pp = "lang_encoder.transformer.blocks.0.decoder_layer."
    m = blk.load_state_dict({"norm_1.weight": W7[pp + "norm_1.weight"], "attn.Wqkv.weight": W7[pp + "attn.Wqkv.weight"],
                             "attn.out_proj.weight": W7[pp + "attn.out_proj.weight"], "norm_2.weight": W7[pp + "norm_2.weight"],
                             "ffn.up_proj.weight": W7[pp + "ffn.up_proj.weight"], "ffn.down_proj.weight": W7[pp + "ffn.down_proj.weight"]},
                            strict=False)
    S = 10
    x = torch.randn(2, S, 256, generator=torch.Generator().manual_seed(1)) * 0.5
    alibi = build_mpt_alibi_tensor(2, S, 8)
    causal = ~torch.ones(S, S, dtype=torch.bool).tril()
    with torch.no_grad():
        ref = blk(x, position_bias=alibi, attention_mask=causal.view(1, 1, S, S))[0]
        mine = O.mpt_block_forward(c7, W7, 0, x, O.mpt_attn_bias(c7, torch.ones(2, S, dtype=torch.int64)))
    res["mpt7b_block_vs_hf_maxabs"] = float((ref - mine).abs().max())
    res["alibi_slopes_vs_hf_maxabs"] = float((build_mpt_alibi_tensor(16, 32, 8)[:, 0, :] -
                                              O.mpt_attn_bias(O.config_3b(), torch.ones(1, 32, dtype=torch.int64))[0, :, 0, :]).abs().max())
    res["alibi32_slopes_vs_hf_maxabs"] = float((build_mpt_alibi_tensor(32, 32, 8)[:, 0, :] -
                                                O.mpt_attn_bias(O.config_9b(), torch.ones(1, 32, dtype=torch.int64))[0, :, 0, :]).abs().max())
    print("[xcheck]", json.dumps(res))
    for k, v in res.items():
        assert v < 1e-4, (k, v)
    with open(os.path.join(GOLD, "xcheck.json"), "w") as f:
        json.dump(res, f, indent=1)


def golden_full(S=32):
    """OpenFlamingo-3B sized fixtures: DeeR-B (max_layer=12) and DeeR-S (max_layer=4)."""
    torch.set_num_threads(os.cpu_count())
    cfgB = O.config_3b(12)
    t0 = time.time()
    W = O.synth_weights(cfgB, 0)
    print(f"[full] weights in {time.time() - t0:.1f}s, {sum(v.numel() for v in W.values()) / 1e9:.3f} G params")
    t0 = time.time()
    episodes = list(range(100, 112))
    table = batched_delta_table(cfgB, W, episodes, 4, S)
    print(f"[full] delta table {tuple(table.shape)} in {time.time() - t0:.1f}s; mean per exit {table.mean(1).tolist()}")
    thrB = O.solve_thresholds(table, 1.0)
    out = {"delta_table": table.numpy(), "thresholds_B": np.array(thrB, np.float64), "S": S}
    model, ctrl, _ = ref_shims.build_reference_model(cfgB, W, thrB)
    exits = [e for e in cfgB.exit_list if e <= cfgB.ctrl_max_layer]
    T = 3
    worst = 0.0
    for e in (0, 1):
        t0 = time.time()
        ref = run_reference_episode(cfgB, model, ctrl, e, T, S, capture_hidden_at=0 if e == 0 else None)
        tref = time.time() - t0
        orc, _ = run_oracle_episode(cfgB, W, thrB, e, T, S, taps_at=0 if e == 0 else None)
        worst = max(worst, compare(ref, orc, 5e-5, f"3B-B ep{e}"))
        for k, v in pack_steps(ref, exits).items():
            out[f"B_ep{e}_{k}"] = v
        if e == 0:
            # a few slices of large intermediates (full tensors would be MBs)
            out["B_ep0_t0_media_slice"] = ref[0]["media"][:8, :64].numpy()
            out["B_ep0_t0_hidden_last_slice"] = ref[0]["hidden"][-1][:8, :64].numpy()
            assert float((ref[0]["media"] - orc[0].taps["media"]).abs().max()) < 5e-5
        print(f"[full] DeeR-B ep{e}: reference {tref / T:.2f} s/step, exits {[r['exit_layer'] for r in ref]}")
    del model, ctrl
    cfgS = O.config_3b(4)
    WS = {k: v for k, v in W.items() if k in {n for n, _, _ in O.weight_specs(cfgS)}}
    thrS = O.solve_thresholds(table[:cfgS.n_thresholds], 0.8)
    out["thresholds_S"] = np.array(thrS, np.float64)
    model, ctrl, _ = ref_shims.build_reference_model(cfgS, WS, thrS)
    exitsS = [e for e in cfgS.exit_list if e <= cfgS.ctrl_max_layer]
    for e in (0, 1):
        ref = run_reference_episode(cfgS, model, ctrl, e, T, S)
        orc, _ = run_oracle_episode(cfgS, WS, thrS, e, T, S)
        worst = max(worst, compare(ref, orc, 5e-5, f"3B-S ep{e}"))
        for k, v in pack_steps(ref, exitsS).items():
            out[f"S_ep{e}_{k}"] = v
        print(f"[full] DeeR-S ep{e}: exits {[r['exit_layer'] for r in ref]}")
    print(f"[full] reference vs oracle worst |diff| {worst:.2e}")
    np.savez_compressed(os.path.join(GOLD, "full3b.npz"), **out)


def golden_full_long(S=32, T=10, n_ep=8):
    """Round-2 fixture ``full3b_long.npz``: OpenFlamingo-3B size, ``n_ep`` episodes x ``T`` consecutive steps (carried LSTM
    state) for DeeR-B (max_layer=12, exit_ratio 1.0) and DeeR-S (max_layer=4, exit_ratio 0.8), produced by the UNMODIFIED
    reference modules and reproduced by the restatement (exit ids identical, tensors <= 5e-5).  Thresholds are calibrated
    on a delta table that covers the SAME timestep range (12 other episodes x T steps), so the exits spread over the list.
    Per decision the fixture also records the margin min_k |delta_k - T_k| over the probes the reference evaluated."""
    torch.set_num_threads(os.cpu_count())
    cfgB = O.config_3b(12)
    W = O.synth_weights(cfgB, 0)
    t0 = time.time()
    table = batched_delta_table(cfgB, W, list(range(100, 112)), T, S)
    print(f"[long] delta table {tuple(table.shape)} in {time.time() - t0:.1f}s; mean per exit {table.mean(1).tolist()}", flush=True)
    out = {"delta_table": table.numpy(), "S": S, "T": T, "episodes": np.arange(n_ep)}
    worst = 0.0
    for tag, cfg, ratio in (("B", cfgB, 1.0), ("S", O.config_3b(4), 0.8)):
        Wc = W if tag == "B" else {k: v for k, v in W.items() if k in {n for n, _, _ in O.weight_specs(cfg)}}
        thr = O.solve_thresholds(table[:cfg.n_thresholds], ratio)
        out[f"thresholds_{tag}"] = np.array(thr, np.float64)
        model, ctrl, _ = ref_shims.build_reference_model(cfg, Wc, thr)
        exits = [e for e in cfg.exit_list if e <= cfg.ctrl_max_layer]
        hist, margins = {}, []
        for e in range(n_ep):
            t0 = time.time()
            ref = run_reference_episode(cfg, model, ctrl, e, T, S)
            orc, _ = run_oracle_episode(cfg, Wc, thr, e, T, S)
            worst = max(worst, compare(ref, orc, 5e-5, f"3B-{tag} ep{e}"))
            for k, v in pack_steps(ref, exits).items():
                out[f"{tag}_ep{e}_{k}"] = v
            mg = []
            for r in ref:
                m = [abs(v - thr[exits.index(k)]) for k, v in r["deltas"].items() if thr[exits.index(k)] < 1e7]
                mg.append(min(m) if m else np.inf)
                hist[r["exit_layer"]] = hist.get(r["exit_layer"], 0) + 1
            out[f"{tag}_ep{e}_margin"] = np.array(mg, np.float64)
            margins += mg
            print(f"[long] DeeR-{tag} ep{e}: exits {[r['exit_layer'] for r in ref]} min margin {min(mg):.2e} ({time.time() - t0:.0f}s)", flush=True)
        print(f"[long] DeeR-{tag}: thresholds {thr}; exit hist {dict(sorted(hist.items()))}; "
              f"decisions with margin > 0.02: {sum(m > 0.02 for m in margins)} / {len(margins)}; smallest margin {min(margins):.2e}", flush=True)
        del model, ctrl
    print(f"[long] reference vs oracle worst |diff| {worst:.2e}")
    np.savez_compressed(os.path.join(GOLD, "full3b_long.npz"), **out)


HEAD_VARIANTS = {
    # tag: (dropout_mode, mlp_layernorm, lstm_layernorm, hidden layers, {canonical Sequential index -> the reference's})
    "oldest": ("wo_last", False, False, 3, {1: 0, 5: 2, 9: 4, 13: 6}),                     # action_head.py:149-158
    "wolast_ln": ("wo_last", True, True, 3, {1: 1, 2: 2, 5: 5, 6: 6, 9: 9, 10: 10, 13: 12}),   # action_head.py:118-133
    "last2": ("last", True, False, 2, {1: 0, 2: 1, 5: 3, 6: 4, 9: 7}),                     # action_head.py:86-113
}


def golden_variants(S=12, T=6):
    """``variants_tiny.npz``: (1) the head variants of older DeeR checkpoints (eval_calvin.py:464-476) -- the UNMODIFIED reference
    built with that dropout_mode / layernorm / hidden-layer setting, fed the synthetic weights under ITS OWN state-dict keys, run
    with dynamic exit; the restatement reproduces it.  The reference's head key names are stored so the product's key
    canonicalisation (checkpoint.canonical_head_keys) can be checked against them.  (2) ``use_action_ensemble``
    (eval_utils.py:457-464, value_net.py:92-95): the action the reference's ModelWrapper would emit, plus ``.logits[2]``."""
    import re
    out = {"S": S, "T": T}
    for tag, (mode, mlp_ln, lstm_ln, nh, idx) in HEAD_VARIANTS.items():
        cfg = O.config_tiny(mlp_dims=(1024, 512, 256)[:nh], mlp_layernorm=mlp_ln, lstm_layernorm=lstm_ln)
        W = O.synth_weights_variant(cfg, 0)
        Wref = {}
        for k, v in W.items():
            m = re.match(r"^(extra_exit\.(?:actions|gripper)\.mlp\.)(\d+)(\..+)$", k)
            Wref[f"{m.group(1)}{idx[int(m.group(2))]}{m.group(3)}" if m else k] = v
        table = batched_delta_table(cfg, W, list(range(100, 108)), 4, S)
        thr = O.solve_thresholds(table, 1.0)
        model, ctrl, res = ref_shims.build_reference_model(cfg, Wref, thr, dropout_mode=mode)
        exits = [e for e in cfg.exit_list if e <= cfg.ctrl_max_layer]
        out[f"{tag}_thresholds"] = np.array(thr, np.float32)
        out[f"{tag}_ref_head_keys"] = np.array(sorted(k for k in model.state_dict() if k.startswith("extra_exit.")))
        hist = {}
        for e in (0, 1):
            ref = run_reference_episode(cfg, model, ctrl, e, T, S)
            orc, _ = run_oracle_episode(cfg, W, thr, e, T, S)
            w = compare(ref, orc, 2e-5, f"variant {tag} ep{e}")
            for k, v in pack_steps(ref, exits).items():
                out[f"{tag}_ep{e}_{k}"] = v
            for r in ref:
                hist[r["exit_layer"]] = hist.get(r["exit_layer"], 0) + 1
        print(f"[variants] {tag}: reference == oracle (worst {w:.2e}); exits {hist}")
        del model, ctrl
    cfg = O.config_tiny()
    W = O.synth_weights(cfg, 0)
    g = np.load(os.path.join(GOLD, "tiny_dynamic.npz"))
    thr = g["thresholds"].tolist()
    model, ctrl, _ = ref_shims.build_reference_model(cfg, W, thr)
    n_two = 0
    for e in (0, 1, 2, 3):
        ref = run_reference_episode(cfg, model, ctrl, e, T, S, ensemble=True)
        orc, _ = run_oracle_episode(cfg, W, thr, e, T, S)
        for t, (r, o) in enumerate(zip(ref, orc)):
            assert r["exit_layer"] == o.exit_layer == int(g[f"ep{e}_exit_layer"][t]), (e, t)
            assert float((r["action"] - o.ensemble_action).abs().max()) <= 2e-5, (e, t)
            assert float((r["feature"] - o.feature).abs().max()) <= 2e-5, (e, t)
            n_two += int(len(r["deltas"]) > 1)
        out[f"ens_ep{e}_action"] = np.stack([r["action"].numpy() for r in ref])
        out[f"ens_ep{e}_feature"] = np.stack([r["feature"].numpy() for r in ref])
        out[f"ens_ep{e}_exit_layer"] = np.array([r["exit_layer"] for r in ref], np.int32)
    print(f"[variants] action ensemble: reference == oracle on {4 * T} steps ({n_two} of them average two probes)")
    np.savez_compressed(os.path.join(GOLD, "variants_tiny.npz"), **out)


def golden_generate(bs=2, win=12, S=16):
    """Calibration pass in the training layout: the unmodified reference (MPTFlamingo.forward with return_in_feat /
    only_extra_exit as value_net.py:370-381 calls it, then ActionValueNet.forward(mode='generate')) vs the oracle's
    restatement, tiny config -> generate_tiny.npz (random layer indices as drawn by the reference, delta table)."""
    cfg = O.config_tiny()
    W = O.synth_weights(cfg, 0)
    model, ctrl, _ = ref_shims.build_reference_model(cfg, W, [1e8] * cfg.n_thresholds)
    model.window_size = win
    model.set_all_exit_window_size(win)
    ctrl.value_net.window_size = win
    eps = list(range(40, 40 + bs))
    rgb = torch.stack([torch.stack([O.synth_image(cfg, e, t, 0) for t in range(win)]) for e in eps])
    grip = torch.stack([torch.stack([O.synth_image(cfg, e, t, 1) for t in range(win)]) for e in eps])
    ids = torch.stack([O.synth_tokens(cfg, e, S) for e in eps])
    images = rgb.unsqueeze(2).unsqueeze(2).flatten(0, 1)
    gripper = grip.unsqueeze(2).unsqueeze(2).flatten(0, 1)
    input_ids = ids.unsqueeze(1).repeat(1, win, 1).flatten(0, 1)
    torch.manual_seed(123)
    with torch.no_grad():
        final_output, _, _, rand_layer_feat, rand_idx = model(
            vision_x=images, lang_x=input_ids, attention_mask=torch.ones_like(input_ids), vision_gripper=gripper,
            state_tensor=None, with_gripper_logits=True, return_in_feat=True, only_extra_exit=True)
        sim = ctrl.value_net(final_output.hidden_states, mode="generate", rand_layer_feat=rand_layer_feat)
    hidden = O.window_hidden_states(cfg, W, rgb, grip, ids)
    mine = O.generate_deltas(cfg, W, hidden, rand_idx, win)
    hd = max(float((a - b).abs().max()) for a, b in zip(final_output.hidden_states, hidden))
    dd = float((sim - mine).abs().max())
    print(f"[generate] reference vs oracle: hidden {hd:.2e}, delta table {tuple(sim.shape)} {dd:.2e}; layer idx {rand_idx.tolist()}")
    assert hd < 5e-5 and dd < 5e-6
    np.savez_compressed(os.path.join(GOLD, "generate_tiny.npz"), layer_idx=rand_idx.numpy().astype(np.int32), deltas=sim.numpy(),
                        episodes=np.array(eps), window=win, S=S)


def golden_yardstick(S=32, T=2, family="3b"):
    """Rounding yardstick at OpenFlamingo-3B / -9B size: the oracle in fp32 and with bf16-rounded
    contraction operands on the same inputs, full depth (every exit probed), 2 episodes x T steps -> yardstick3b.npz /
    yardstick9b.npz (python -m oracle.make_golden --yardstick-only [--family 9b])."""
    torch.set_num_threads(os.cpu_count())
    cfg = O.config_3b(12) if family == "3b" else O.config_9b(12)
    W = O.synth_weights(cfg, 0)
    thr = [1e8] * cfg.n_thresholds
    out = {"S": S, "T": T}
    for mode, tag in ((None, "fp32"), (torch.bfloat16, "bf16ops")):
        with O.operand_rounding(mode):
            for e in (0, 1):
                orc = O.DeerOracle(cfg, W, thr)
                orc.reset()
                ids = O.synth_tokens(cfg, e, S)
                poses, grips, deltas = [], [], []
                for t in range(T):
                    o = orc.step(O.synth_image(cfg, e, t, 0), O.synth_image(cfg, e, t, 1), ids, full_depth=True)
                    poses.append(o.pose.flatten().numpy()); grips.append(float(o.gripper.flatten()[0]))
                    deltas.append(np.array([float(o.deltas[k]) for k in sorted(o.deltas)], np.float32))
                out[f"{tag}_ep{e}_pose"] = np.stack(poses)
                out[f"{tag}_ep{e}_gripper"] = np.array(grips, np.float32)
                out[f"{tag}_ep{e}_deltas"] = np.stack(deltas)
    for e in (0, 1):
        print(f"[yardstick] ep{e}: |bf16ops - fp32| pose {np.abs(out[f'bf16ops_ep{e}_pose'] - out[f'fp32_ep{e}_pose']).max():.3e} "
              f"deltas {np.abs(out[f'bf16ops_ep{e}_deltas'] - out[f'fp32_ep{e}_deltas']).max():.3e}")
    np.savez_compressed(os.path.join(GOLD, f"yardstick{family}.npz"), **out)


def golden_yardstick_long(S=32):
    """``yardstick3b_long.npz``: the rounding yardstick over the SAME rollouts as full3b_long.npz (8 episodes x 10 carried steps,
    DeeR-B and DeeR-S, dynamic exit): the oracle with bf16-rounded contraction operands vs the reference goldens.  Where the
    rounded oracle takes the reference's exit, |pose - golden pose| is one yardstick sample; where rounding flips the exit the
    carried state is re-seeded from an fp32 oracle run beside it (== the goldens) -- exactly what the GPU test does with the
    engine.  The test's pose bar is 1.5 x the largest sample."""
    torch.set_num_threads(os.cpu_count())
    g = np.load(os.path.join(GOLD, "full3b_long.npz"))
    T, eps = int(g["T"]), g["episodes"].tolist()
    W = O.synth_weights(O.config_3b(12), 0)
    out = {}
    for tag, ml in (("B", 12), ("S", 4)):
        cfg = O.config_3b(ml)
        Wc = {k: v for k, v in W.items() if k in {n for n, _, _ in O.weight_specs(cfg)}}
        thr = g[f"thresholds_{tag}"].tolist()
        errs, match = np.full((len(eps), T), np.nan, np.float32), np.zeros((len(eps), T), np.int8)
        for i, e in enumerate(eps):
            t0 = time.time()
            f32, b16 = O.DeerOracle(cfg, Wc, thr), O.DeerOracle(cfg, Wc, thr)
            f32.reset(); b16.reset()
            ids = O.synth_tokens(cfg, e, S)
            for t in range(T):
                rgb, grip = O.synth_image(cfg, e, t, 0), O.synth_image(cfg, e, t, 1)
                f32.set_timestep(t); b16.set_timestep(t)
                a = f32.step(rgb, grip, ids)
                with O.operand_rounding(torch.bfloat16):
                    b = b16.step(rgb, grip, ids)
                assert a.exit_layer == int(g[f"{tag}_ep{e}_exit_layer"][t])
                if b.exit_layer == a.exit_layer:
                    match[i, t] = 1
                    errs[i, t] = float((b.pose - torch.from_numpy(g[f"{tag}_ep{e}_pose"][t])).abs().max())
                else:
                    b16.state = (f32.state[0].clone(), f32.state[1].clone())
                    b16.cur_exit_id = f32.cur_exit_id
            print(f"[yardstick-long] DeeR-{tag} ep{e}: matched {int(match[i].sum())}/{T}, max pose err {np.nanmax(errs[i]):.3e} ({time.time() - t0:.0f}s)", flush=True)
        out[f"{tag}_pose_err"], out[f"{tag}_match"] = errs, match
        print(f"[yardstick-long] DeeR-{tag}: Y_pose = {np.nanmax(errs):.3e} (mean {np.nanmean(errs):.3e}); rounded oracle keeps the reference's exit on "
              f"{int(match.sum())}/{match.size} decisions", flush=True)
    np.savez_compressed(os.path.join(GOLD, "yardstick3b_long.npz"), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--yardstick-only", action="store_true")
    ap.add_argument("--generate-only", action="store_true")
    ap.add_argument("--long-only", action="store_true", help="only the 8-episode x 10-step 3B fixture (full3b_long.npz)")
    ap.add_argument("--variants-only", action="store_true", help="only the head-variant / action-ensemble fixture (variants_tiny.npz)")
    ap.add_argument("--family", default="3b", choices=["3b", "9b"], help="size of the --yardstick-only fixture")
    ap.add_argument("--yardstick-long-only", action="store_true", help="only yardstick3b_long.npz (bf16-operand oracle over the full3b_long rollouts)")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if a.yardstick_long_only:
        golden_yardstick_long()
        sys.exit(0)
    if a.yardstick_only:
        golden_yardstick(family=a.family)
        sys.exit(0)
    if a.generate_only:
        golden_generate()
        sys.exit(0)
    if a.long_only:
        golden_full_long()
        sys.exit(0)
    if a.variants_only:
        golden_variants()
        sys.exit(0)
    golden_thresholds()
    xcheck_third_party()
    golden_tiny()
    golden_variants()
    golden_generate()
    if a.full:
        golden_full()
        golden_full_long()
        golden_yardstick_long()
        golden_yardstick()
        golden_yardstick(family="9b")
