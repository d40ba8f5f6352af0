"""The reference-facing Python surface (MPTFlamingo / ExitController / ModelWrapper mirrors) on the GPU,
checked against the oracle; host-buffer path (deer_step_host) == device path; profiling hooks."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import deer_oracle as O  # noqa: E402


class _Tok:
    eos_token = "<eos>"
    padding_side = "left"

    def __init__(self, cfg):
        self.cfg = cfg

    def __call__(self, sample, max_length, padding, truncation, return_tensors):
        ids = []
        for s in sample:
            body = [1 + (ord(c) * 7) % 400 for c in s if c not in "<>|"][: max_length - 3]
            ids.append([self.cfg.media_token_id] + body + [self.cfg.eoc_token_id, 0])
        L = max(len(i) for i in ids)
        mask = [[1] * len(i) + [0] * (L - len(i)) for i in ids]
        ids = [i + [511] * (L - len(i)) for i in ids]
        return {"input_ids": torch.tensor(ids), "attention_mask": torch.tensor(mask)}


def _build(prec="fp32", max_batch=1):
    from deer_vla_b200 import DeerConfig
    from deer_vla_b200.engine import DeerEngine
    from deer_vla_b200.flamingo_mpt import MPTFlamingo
    from deer_vla_b200.value_net import ExitController
    ocfg = O.config_tiny()
    W = O.synth_weights(ocfg, 0)
    cfg = DeerConfig(**ocfg.to_dict())
    eng = DeerEngine(cfg, max_batch=max_batch, max_seq=32, precision=prec)
    model = MPTFlamingo(eng, cfg)
    model.load_state_dict({"model_state_dict": {"module." + k: v for k, v in W.items()}})    # DeeR ckpt schema (train_utils.py:27-60,631-638)
    model.finalize()
    ctrl = ExitController(None, exit_id_list=model.get_all_exit_idx(), steps_per_stage=1, leq=True, max_layer=ocfg.max_layer, engine=eng)
    return ocfg, W, eng, model, ctrl


def test_model_wrapper_rollout_matches_oracle():
    from deer_vla_b200.eval_utils import ModelWrapper, clip_image_transform, preprocess_text_calvin
    ocfg, W, eng, model, ctrl = _build()
    thr = [0.3, 0.25, 1e8]
    ctrl._set_threshold_value(thr)
    assert model.get_all_exit_idx() == ocfg.exit_list and ctrl.max_layer == ocfg.ctrl_max_layer
    tok, proc = _Tok(ocfg), clip_image_transform(ocfg.image_size)
    wrap = ModelWrapper(model, tok, proc, exit_controller=ctrl, early_exit=True, multi_execution=2)
    rs = np.random.RandomState(0)
    goal = "push the red block left"
    ids, mask = preprocess_text_calvin([goal], tok)
    orc = O.DeerOracle(ocfg, W, thr)
    wrap.reset()
    orc.reset()
    for t in range(4):
        obs = {"rgb_obs": {"rgb_static": (rs.rand(200, 200, 3) * 255).astype(np.uint8),
                           "rgb_gripper": (rs.rand(84, 84, 3) * 255).astype(np.uint8)}}
        ctrl.set_timestep(t)
        a = wrap.step(obs, goal)
        o = orc.step(proc(obs["rgb_obs"]["rgb_static"]), proc(obs["rgb_obs"]["rgb_gripper"]), ids[0], mask[0])
        assert a.shape == (2, 7) and a.dtype == np.float16
        assert wrap.current_exit_layer == o.exit_layer
        assert np.abs(a[0].astype(np.float32) - o.action.numpy()).max() <= 2e-3          # fp16 output rounding
    # a new subtask: reset clears the LSTM state (eval_utils.py:252-277)
    wrap.reset()
    h, c = eng.get_state(0)
    assert not h.any() and not c.any()
    eng.close()


def test_batched_wrapper_equals_single_env():
    from deer_vla_b200.eval_utils import BatchedModelWrapper, clip_image_transform, preprocess_text_calvin
    ocfg, W, eng, model, ctrl = _build(max_batch=3)
    thr = [0.3, 0.25, 1e8]
    ctrl._set_threshold_value(thr)
    tok, proc = _Tok(ocfg), clip_image_transform(ocfg.image_size)
    bw = BatchedModelWrapper(model, tok, proc, exit_controller=ctrl)
    goals = ["open the drawer", "lift the pink block from the table", "turn on the led"]       # ragged lengths -> right padding
    rs = np.random.RandomState(1)
    oracles = [O.DeerOracle(ocfg, W, thr) for _ in goals]
    bw.reset()
    for t in range(3):
        obs = [{"rgb_obs": {"rgb_static": (rs.rand(200, 200, 3) * 255).astype(np.uint8),
                            "rgb_gripper": (rs.rand(84, 84, 3) * 255).astype(np.uint8)}} for _ in goals]
        acts = bw.step(obs, goals)
        assert acts.shape == (3, 1, 7)
        for i, g in enumerate(goals):
            ids, mask = preprocess_text_calvin([g], tok)            # batch-1 tokenisation: no pads, like the reference
            o = oracles[i].step(proc(obs[i]["rgb_obs"]["rgb_static"]), proc(obs[i]["rgb_obs"]["rgb_gripper"]), ids[0], mask[0])
            assert bw.current_exit_layers[i] == o.exit_layer, (t, i)
            assert np.abs(acts[i, 0].astype(np.float32) - o.action.numpy()).max() <= 2e-3
    eng.close()


def test_host_path_equals_device_path_and_profile():
    ocfg, W, eng, model, ctrl = _build(prec="bf16", max_batch=2)
    eng.set_thresholds([0.3, 0.25, 1e8])
    S = 12
    ids = torch.stack([O.synth_tokens(ocfg, e, S) for e in range(2)])
    rgb = torch.stack([O.synth_image(ocfg, e, 0, 0) for e in range(2)])
    grip = torch.stack([O.synth_image(ocfg, e, 0, 1) for e in range(2)])
    eng.reset()
    d = eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
    d = {k: v.cpu().clone() for k, v in d.items()}
    eng.reset()
    h = eng.step(rgb.pin_memory(), grip.pin_memory(), ids.pin_memory())          # deer_step_host: H2D + run + D2H + sync inside
    assert torch.equal(d["pose"], h["pose"]) and torch.equal(d["exit_layer"], h["exit_layer"]) and torch.equal(d["action"], h["action"])
    eng.set_profile(True)
    eng.reset()
    eng.step(rgb.cuda(), grip.cuda(), ids.cuda())
    p = eng.profile()
    assert p["step_ms"] > 0 and p["gemm_launches"]["vit"] > 0 and p["gemm_launches"]["perceiver"] > 0 and p["gemm_launches"]["llm"] > 0
    assert sum(p["gemm_ms"].values()) <= p["step_ms"] * 1.01
    assert len(p["active_after"]) >= 1 and p["active_after"][-1] == 0
    st = eng.stats()
    assert st.kernel_launches_last_step > 50 and st.gemm_launches_last_step > 20
    eng.close()


def test_calls_on_different_streams_are_ordered():
    """deer_step on the caller's stream, deer_reset, deer_step_host on the engine's stream and a side stream, back to back
    WITHOUT host synchronisation in between: the ctx orders them itself (shared workspace / episode state)."""
    ocfg, W, eng, model, ctrl = _build(max_batch=2)
    eng.set_thresholds([0.3, 0.25, 1e8])
    rgb = torch.stack([O.synth_image(ocfg, e, 0, 0) for e in (0, 1)])
    grip = torch.stack([O.synth_image(ocfg, e, 0, 1) for e in (0, 1)])
    ids = torch.stack([O.synth_tokens(ocfg, e, 16) for e in (0, 1)])
    drgb, dgrip, dids = rgb.cuda(), grip.cuda(), ids.cuda()
    hrgb, hgrip, hids = rgb.pin_memory(), grip.pin_memory(), ids.pin_memory()
    torch.cuda.synchronize()
    eng.reset()
    want = eng.step(hrgb, hgrip, hids)
    want = {k: v.clone() for k, v in want.items()}
    side = torch.cuda.Stream()
    for rep in range(6):
        eng.reset()
        d = eng.step(drgb, dgrip, dids)                       # async on the current stream
        eng.reset()
        h = eng.step(hrgb, hgrip, hids)                       # engine's own stream, while d may still be running
        h = {k: v.clone() for k, v in h.items()}
        eng.reset()
        with torch.cuda.stream(side):
            s2 = eng.step(drgb, dgrip, dids)                  # a third stream
        torch.cuda.synchronize()
        for got in (d, h, s2):
            assert torch.equal(got["exit_layer"].cpu(), want["exit_layer"]), rep
            assert torch.equal(got["action"].cpu(), want["action"]), rep
    eng.close()


def test_batched_wrapper_device_transform_equals_host_transform():
    """Raw uint8 frames transformed on the GPU (deer_step_host_u8) vs the Pillow transform on the host: the transform is
    bit-exact, so actions and exits are identical."""
    pytest.importorskip("PIL.Image")
    from deer_vla_b200.eval_utils import BatchedModelWrapper, clip_image_transform
    ocfg, W, eng, model, ctrl = _build(max_batch=3)
    ctrl._set_threshold_value([0.3, 0.25, 1e8])
    tok, proc = _Tok(ocfg), clip_image_transform(ocfg.image_size)
    host = BatchedModelWrapper(model, tok, proc, exit_controller=ctrl)
    dev = BatchedModelWrapper(model, tok, None, exit_controller=ctrl, device_transform=True)
    goals = ["open the drawer", "lift the pink block from the table", "turn on the led"]
    rs = np.random.RandomState(4)
    obs_seq = [[{"rgb_obs": {"rgb_static": (rs.rand(200, 200, 3) * 255).astype(np.uint8),
                             "rgb_gripper": (rs.rand(84, 84, 3) * 255).astype(np.uint8)}} for _ in goals] for _ in range(3)]
    res = []
    for w in (host, dev):
        w.reset()
        res.append([(w.step(obs, goals).copy(), list(w.current_exit_layers)) for obs in obs_seq])
    for (a0, e0), (a1, e1) in zip(*res):
        assert e0 == e1 and np.array_equal(a0, a1)
    eng.close()
