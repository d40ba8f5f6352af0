"""CPU tests of the product's host-side logic (no GPU, no CUDA calls)."""
import numpy as np
import pytest
import torch

from oracle import deer_oracle as O


def test_product_config_agrees_with_oracle_bookkeeping():
    from deer_vla_b200 import config_3b, config_9b, config_tiny
    for pc, oc in ((config_3b(12), O.config_3b(12)), (config_3b(4), O.config_3b(4)), (config_9b(12), O.config_9b(12)),
                   (config_9b(8), O.config_9b(8)), (config_tiny(), O.config_tiny()), (config_3b(24), O.config_3b(24))):
        assert pc.to_dict() == oc.to_dict()
        assert (pc.n_layers, pc.exit_list, pc.ctrl_max_layer, pc.n_thresholds) == (oc.n_layers, oc.exit_list, oc.ctrl_max_layer, oc.n_thresholds)


def test_synthetic_checkpoint_generators_agree():
    from deer_vla_b200 import DeerConfig
    from deer_vla_b200.weights import state_dict_spec, synth_state_dict
    oc = O.config_tiny()
    W = O.synth_weights(oc, 3)
    P = synth_state_dict(DeerConfig(**oc.to_dict()), 3)
    assert list(W) == list(P) and all(torch.equal(W[k], P[k]) for k in W)
    spec9 = {n for n, _, _ in state_dict_spec(DeerConfig(**O.config_9b(12).to_dict()))}
    assert "lang_encoder.transformer.blocks.3.gated_cross_attn_layer.attn.to_q.weight" in spec9
    assert "lang_encoder.transformer.blocks.2.gated_cross_attn_layer.attn.to_q.weight" not in spec9
    assert "lang_encoder.transformer.blocks.0.decoder_layer.norm_1.weight" in spec9


def test_exit_controller_threshold_solver_matches_reference(gold_thr):
    """The product's vectorised solver vs thresholds produced by the reference's ExitController.set_threshold."""
    from deer_vla_b200.value_net import ExitController, solve_thresholds
    g = gold_thr
    for name, llm in (("3b", "mpt_dolly_3b"), ("3b_s", "mpt_dolly_3b"), ("9b", "mpt_9b")):
        values = torch.from_numpy(g[f"{name}_values"])
        for ratio in (0.5, 0.8, 1.0, 1.5, 2.0):
            for dist in ("exp", "gauss", "gamma"):
                T = solve_thresholds(values, ratio, dist, llm)
                assert np.array_equal(T.double().numpy(), g[f"{name}_{dist}_{ratio}"]), (name, ratio, dist)
    n_exit = g["3b_values"].shape[0]
    exits = list(range(1, 2 * n_exit, 2))
    ctrl = ExitController(None, exit_id_list=exits, steps_per_stage=1, exit_dist="exp", leq=True, max_layer=exits[-1] + 1)
    ctrl.set_threshold(None, None, None, 1.0, "mpt_dolly_3b", values=torch.from_numpy(g["3b_values"]))
    assert [ctrl.thresholds[e] for e in exits] == [float(x) for x in g["3b_exp_1.0"]]
    with pytest.raises(AssertionError):
        ctrl._set_threshold_value([0.1, 0.2])            # len(thresholds) == real_num_exit (value_net.py:179)
    ctrl.set_timestep(5)
    assert ctrl.cur_step == 5 and ctrl.module is ctrl


def test_exit_controller_max_layer_semantics():
    from deer_vla_b200.value_net import ExitController
    c = ExitController(None, exit_id_list=[1, 3, 4], max_layer=4)        # DeeR-S
    assert c.max_layer == 3 and c._real_exits() == [1, 3]
    c = ExitController(None, exit_id_list=[1, 3, 5, 7, 9, 11], max_layer=12)
    assert c.max_layer == 11 and len(c._real_exits()) == 6
    c = ExitController(None, exit_id_list=[1, 3, 5, 7, 9, 11], max_layer=24)
    assert c.max_layer == 11


def test_sharding_block_split():
    from deer_vla_b200.sharding import owner_of, shard_sequences
    seqs = list(range(224))
    parts = [shard_sequences(seqs, r, 8) for r in range(8)]
    assert all(len(p) == 28 for p in parts) and sum(parts, []) == seqs
    assert owner_of(27, 224, 8) == 0 and owner_of(28, 224, 8) == 1 and owner_of(223, 224, 8) == 7
    with pytest.raises(AssertionError):
        shard_sequences(list(range(10)), 0, 4)            # NUM_SEQUENCES % device_num == 0 (eval_utils.py:523)


class _Tok:
    eos_token = "<eos>"
    padding_side = "left"

    def __call__(self, sample, max_length, padding, truncation, return_tensors):
        ids = [[7] + [ord(c) % 50 + 10 for c in s][: max_length - 1] for s in sample]
        L = max(len(i) for i in ids)
        mask = [[1] * len(i) + [0] * (L - len(i)) for i in ids]
        ids = [i + [0] * (L - len(i)) for i in ids]
        return {"input_ids": torch.tensor(ids), "attention_mask": torch.tensor(mask)}


def test_text_and_action_processing():
    from deer_vla_b200.eval_utils import clip_image_transform, postprocess_action, preprocess_text_calvin
    tok = _Tok()
    ids, mask = preprocess_text_calvin(["  push the block  ", "lift"], tok)
    assert tok.padding_side == "right" and ids.shape == mask.shape and ids.shape[1] <= 32
    a = postprocess_action(torch.tensor([[[0.1, -0.2, 0.3, 0.0, 0.5, -0.9]]]), torch.tensor([[[0.7]]]), multi_execution=2)
    assert a.dtype == np.float16 and a.shape == (2, 7) and a[0, 6] == 1.0
    a = postprocess_action(torch.zeros(1, 1, 6), torch.tensor([[[0.5]]]))        # gripper > 0.5 is strict (eval_utils.py:458)
    assert a[0, 6] == -1.0
    img = (np.random.RandomState(0).rand(200, 200, 3) * 255).astype(np.uint8)
    t = clip_image_transform(224)(img)
    assert t.shape == (3, 224, 224) and t.dtype == torch.float32
    g = clip_image_transform(224)((np.random.RandomState(1).rand(84, 84, 3) * 255).astype(np.uint8))
    assert g.shape == (3, 224, 224)


def test_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deer_vla_b200 import config_tiny
    from deer_vla_b200._lib import DeerError
    from deer_vla_b200.engine import DeerEngine
    with pytest.raises(DeerError, match="no CUDA device"):
        DeerEngine(config_tiny(), max_batch=1)


def test_checkpoint_name_and_hyperparameter_readout():
    """eval_calvin.py:356-489: configuration smuggled through the checkpoint file name + dict."""
    from deer_vla_b200.checkpoint import config_from_checkpoint, parse_checkpoint_name, readout_args
    name = "DeeR_task_D_D-exit-strategy-post_4+1-exit-mpt_dolly_3b-layer_11_multie_intv=2_extrae_nodth_reg_aug_10_4_traj_cons_ws_12_mpt_dolly_3b_4.pth"
    f = parse_checkpoint_name(name)
    assert f["llm_name"] == "mpt_dolly_3b" and f["rgb_pad"] == 10 and f["gripper_pad"] == 4 and f["traj_cons"] is True
    assert f["multi_step_action"] == 1 and not f["use_state"] and not f["sep_resampler"] and not f["dif_ws"]
    assert parse_checkpoint_name("x_mpt_9b_y.pth")["llm_name"] == "mpt_9b"
    assert parse_checkpoint_name("a_3_step_b_mpt_3b.pth")["multi_step_action"] == 3
    f2 = parse_checkpoint_name("m_difws_8_16_n.pth")
    assert f2["dif_ws"] and f2["window_size"] == 16
    ck = dict(head_type="deterministic", early_exit_layer=11, multi_exit=True, exit_interval=2, mlp_layernorm=True,
              lstm_layernorm=True, mlp_num_hidden_layers=2, lstm_num_layers=4, pooling="max", dropout_mode="layerwise")
    a = readout_args(ck, "mpt_dolly_3b", None)
    assert a["max_layer"] == 12 and a["exit_interval"] == 2
    old = readout_args({}, "mpt_9b", None)                  # old checkpoints: defaults; early_exit_layer -1 -> 31
    assert old["early_exit_layer"] == 31 and old["mlp_num_hidden_layers"] == 3 and old["dropout_mode"] == "wo_last"
    cfg = config_from_checkpoint(ck, name, max_layer=12)
    assert cfg.exit_list == [1, 3, 5, 7, 9, 11] and cfg.n_layers == 12
    assert config_from_checkpoint(ck, name, max_layer=4).exit_list == [1, 3, 4]
    old_cfg = config_from_checkpoint({}, name)                # an old-style head: no LayerNorms, 3 hidden layers (eval_calvin.py:464-476)
    assert old_cfg.mlp_dims == (1024, 512, 256) and not old_cfg.mlp_layernorm and not old_cfg.lstm_layernorm
    with pytest.raises(NotImplementedError):
        config_from_checkpoint(dict(ck, head_type="gaussian"), name)
    with pytest.raises(NotImplementedError):
        config_from_checkpoint(ck, name.replace("traj_cons", "state"))


def test_assemble_state_dict_from_published_pieces():
    """HF MPT + open_clip + OpenFlamingo + DeeR pieces with their native key names -> the model's state dict (f3)."""
    import torch
    from deer_vla_b200 import config_tiny
    from deer_vla_b200.checkpoint import assemble_state_dict
    from deer_vla_b200.weights import synth_state_dict
    cfg = config_tiny()
    full = synth_state_dict(cfg, 3)
    hf, clip, of, deer = {}, {"text_projection": torch.zeros(2), "logit_scale": torch.zeros(())}, {}, {}
    for k, v in full.items():
        if k.startswith("vision_encoder."):
            clip["visual." + k[len("vision_encoder."):]] = v
        elif ".decoder_layer." in k:
            i = k.split(".")[3]
            hf[f"transformer.blocks.{i}." + k.split(".decoder_layer.")[1]] = v
        elif k == "lang_encoder.transformer.wte.weight":
            hf["transformer.wte.weight"] = torch.cat([v, torch.zeros(24, v.shape[1])])         # HF vocab is larger (50432 vs 50280)
            of[k] = v + 1.0                                                                     # OpenFlamingo trains the new tokens' rows
        elif ".gated_cross_attn_layer." in k:
            i = k.split(".")[3]
            of[f"lang_encoder.gated_cross_attn_layers.{i}." + k.split(".gated_cross_attn_layer.")[1]] = v
        elif k.startswith("perceiver."):
            of[k] = v * 0.5
            deer["module." + k] = v                                                             # DeeR fine-tunes the resampler: later source wins
        else:
            deer["module." + k] = v
    hf["transformer.blocks.99.ln_1.weight"] = torch.zeros(4)                                    # deeper blocks are deleted by the model
    hf["transformer.ln_f.weight"] = torch.zeros(4)
    merged, missing = assemble_state_dict(cfg, hf_lm=hf, clip=clip, open_flamingo=of, deer={"model_state_dict": deer, "exit_interval": 2})
    assert missing == []
    assert set(merged) == set(full)
    for k, v in full.items():
        want = v + 1.0 if k == "lang_encoder.transformer.wte.weight" else v
        assert torch.equal(merged[k], want), k
    _, missing = assemble_state_dict(cfg, hf_lm=hf, clip=clip)
    assert any(m.startswith("perceiver.") for m in missing) and any(m.startswith("extra_exit.") for m in missing)


def test_gelu_epilogue_polynomial_constants():
    """The constants of the one-MUFU exact GELU that ships in csrc/common.cuh (gelu_erf_fast), evaluated here in float32 with the
    same Horner order, reproduce 0.5 x (1 + erf(x / sqrt 2)) -- nn.GELU() of the reference's FeedForward (helpers.py:15-22) and
    MPT MLP -- to <= 1e-6 absolute on [-8, 8] and keep relative accuracy in the negative tail.  (The GPU test
    test_gemm_epilogue_gelu_is_exact_erf_gelu checks the kernel itself.)"""
    import math
    import os
    import re
    import numpy as np
    src = open(os.path.join(os.path.dirname(__file__), "..", "deer_vla_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float gelu_erf_fast(float x)"):]
    body = body[:body.index("return")]
    first = float(re.search(r"float p = ([-+0-9.e]+)f;", body).group(1))
    rest = [float(x) for x in re.findall(r"p = fmaf\(p, a, ([-+0-9.e]+)f\);", body)]
    clamp = float(re.search(r"fminf\(fabsf\(x\), ([0-9.]+)f\)", body).group(1))
    assert len(rest) == 7 and rest[-1] == -1.0 and clamp == 6.0
    x = np.linspace(-8, 8, 200001).astype(np.float32)
    a = np.minimum(np.abs(x), np.float32(clamp))
    p = np.full_like(a, np.float32(first))
    for c in rest:
        p = (p * a + np.float32(c)).astype(np.float32)
    xt = x * np.exp2(p).astype(np.float32)
    got = np.where(x >= 0, x - xt, xt).astype(np.float64)
    xd = x.astype(np.float64)
    ref = np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in xd])
    assert np.abs(got - ref).max() <= 1e-6
    tail = (xd < -3) & (xd > -6)
    assert (np.abs(got - ref)[tail] / np.abs(ref[tail])).max() <= 2e-4


def test_bench_clock_sampler_keeps_only_samples_inside_the_timed_region(tmp_path, monkeypatch):
    """bench.py's ClockSampler streams `nvidia-smi -lms 50` and keeps the rows whose nvidia-smi timestamp falls inside
    [mark_begin, mark_end] (the driver rejects a line without clocks sampled DURING the timed region).  Driven here with a fake
    nvidia-smi: rows before the region carry a different clock and must not leak into the median."""
    import importlib
    import os
    import stat
    import sys
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("""#!/usr/bin/env python3
import sys, time, datetime
loop = "-lms" in sys.argv
t0 = time.time()
while True:
    ts = datetime.datetime.now().strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    clk = 1965 if time.time() - t0 < 0.35 else 1400
    print(f"{ts}, {clk}, 1965, 990.0, Not Active, Not Active, Not Active, {'Not Active' if clk == 1965 else 'Active'}", flush=True)
    if not loop:
        break
    time.sleep(0.05)
""")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    bench = importlib.import_module("bench")
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.6)                      # "warm-up": the fake reports 1965 MHz for the first 0.35 s
    s.mark_begin()
    time.sleep(0.5)
    s.mark_end()
    out = s.stop()
    assert out["samples"] >= 5 and out["sm_mhz"] == 1400.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["power_w"] == 990.0


def test_head_key_canonicalisation_vs_reference_layouts():
    """checkpoint.canonical_head_keys against the state-dict keys the UNMODIFIED reference builds for each head variant
    (stored in tests/golden/variants_tiny.npz by oracle/make_golden.py::golden_variants; action_head.py:82-161,183-269)."""
    import os
    import numpy as np
    from deer_vla_b200 import config_tiny
    from deer_vla_b200.checkpoint import canonical_head_keys
    from deer_vla_b200.weights import state_dict_spec
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "variants_tiny.npz"))
    variants = {"oldest": dict(dropout_mode="wo_last", mlp_layernorm=False, lstm_layernorm=False, mlp_num_hidden_layers=3, exit_dropout=0.0),
                "wolast_ln": dict(dropout_mode="wo_last", mlp_layernorm=True, lstm_layernorm=True, mlp_num_hidden_layers=3, exit_dropout=0.0),
                "last2": dict(dropout_mode="last", mlp_layernorm=True, lstm_layernorm=False, mlp_num_hidden_layers=2, exit_dropout=0.0)}
    for tag, a in variants.items():
        ref_keys = ["module." + str(k) for k in g[f"{tag}_ref_head_keys"]]
        canon = canonical_head_keys({k: 0 for k in ref_keys}, a)
        nh = 3 if a["dropout_mode"] == "wo_last" else a["mlp_num_hidden_layers"]
        cfg = config_tiny(mlp_dims=(1024, 512, 256)[:nh], mlp_layernorm=a["mlp_layernorm"], lstm_layernorm=a["lstm_layernorm"])
        want = {"module." + n for n, _, _ in state_dict_spec(cfg) if n.startswith("extra_exit.")}
        assert set(canon) == want, (tag, sorted(set(canon) ^ want)[:6])


def test_every_environment_switch_of_the_library_is_documented():
    """INTEGRATION.md section 4 lists the library's environment switches; a getenv added to csrc/ without a row there fails here."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, "deer_vla_b200", "csrc", "*.cu*")):
        names |= set(re.findall(r'getenv\("([A-Z0-9_]+)"\)', open(f).read()))
    assert names, "no getenv found: the scan is broken"
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = sorted(n for n in names if n not in doc)
    assert not missing, f"undocumented environment switches: {missing}"
