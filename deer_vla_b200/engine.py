"""Python handle on the C-ABI engine (``include/deer_engine.h``).

PyTorch is used only as plumbing: device buffers, streams, pinned host memory.  All arithmetic of
the hot path runs inside ``libdeer_engine.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import DeerCConfig, DeerError, DeerProfile, DeerStats
from .config import DeerConfig


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeerEngine:
    """One engine per GPU.  Owns weights (packed bf16 arena), workspaces, the per-slot LSTM state and
    the exit thresholds.  ``step`` runs one action-step for a batch of concurrent episodes, each with
    the decision it would have had alone (the reference is batch-1: value_net.py:293)."""

    def __init__(self, cfg: DeerConfig, max_batch: int = 1, max_slots: Optional[int] = None, max_seq: int = 32,
                 device: int = 0, precision: str = "bf16", gemm_impl: str = "tcgen05", debug_taps: bool = False,
                 host_control: bool = False, use_graph: bool = False, action_ensemble: bool = False):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.max_batch = int(max_batch)
        self.max_slots = int(max_slots if max_slots is not None else max_batch)
        self.max_seq = int(max_seq)
        if len(cfg.mlp_dims) not in (2, 3):
            raise ValueError("mlp_num_hidden_layers must be 2 (the trained DeeR head, train_calvin_post_strategy.py:343-344) "
                             "or 3 (older checkpoints, eval_calvin.py:473)")
        cc = DeerCConfig()
        cc.abi_version = _lib.ABI_VERSION
        for k in ("image_size", "patch", "vit_width", "vit_layers", "vit_heads", "vit_mlp", "perc_depth", "perc_heads",
                  "perc_dim_head", "num_latents", "perc_ff_mult", "d_model", "n_heads", "mlp_ratio", "vocab",
                  "alibi_bias_max", "cross_attn_every", "xattn_heads", "xattn_dim_head", "xattn_ff_mult",
                  "media_token_id", "exit_interval", "max_layer", "lstm_hidden", "lstm_layers", "steps_per_stage"):
            setattr(cc, k, int(getattr(cfg, k)))
        cc.qk_ln = int(bool(cfg.qk_ln))
        cc.n_layers = cfg.n_layers
        cc.is_mpt_9b = int("mpt_9b" in cfg.llm_name)
        cc.mlp_dim0, cc.mlp_dim1 = int(cfg.mlp_dims[0]), int(cfg.mlp_dims[1])
        cc.mlp_dim2 = int(cfg.mlp_dims[2]) if len(cfg.mlp_dims) == 3 else 0
        cc.mlp_layernorm, cc.lstm_layernorm = int(bool(cfg.mlp_layernorm)), int(bool(cfg.lstm_layernorm))
        cc.action_ensemble = int(bool(action_ensemble))
        cc.threshold_type = _lib.THR_TYPES[cfg.threshold_type]
        cc.max_batch, cc.max_slots, cc.max_seq = self.max_batch, self.max_slots, self.max_seq
        cc.precision = {"bf16": _lib.PREC_BF16, "fp32": _lib.PREC_FP32}[precision]
        cc.gemm_impl = {"tcgen05": _lib.GEMM_TCGEN05, "simt": _lib.GEMM_SIMT}[gemm_impl]
        cc.debug_taps = int(debug_taps)
        cc.host_control = int(host_control)
        cc.use_graph = int(use_graph)
        self.precision = precision
        self.action_ensemble = bool(action_ensemble)
        self.debug_taps = bool(debug_taps)
        self._h = C.c_void_p()
        rc = self.lib.deer_create(C.byref(cc), device, C.byref(self._h))
        if rc != 0:
            raise DeerError(rc, self.lib.deer_last_error(None).decode())
        # the engine must agree with the host-side restatement of the reference's exit bookkeeping
        assert self.exit_list == cfg.exit_list, (self.exit_list, cfg.exit_list)
        assert self.lib.deer_num_thresholds(self._h) == cfg.n_thresholds
        assert self.lib.deer_ctrl_max_layer(self._h) == cfg.ctrl_max_layer
        self.thresholds: Optional[List[float]] = None
        self._pinned: Dict[str, torch.Tensor] = {}

    # ---- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.deer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise DeerError(rc, self.lib.deer_last_error(self._h).decode())
        return rc

    # ---- introspection ----------------------------------------------------------------------
    @property
    def exit_list(self) -> List[int]:
        """model.get_all_exit_idx() (flamingo_mpt.py:268-270)."""
        return [self.lib.deer_exit_layer(self._h, i) for i in range(self.lib.deer_num_exits(self._h))]

    @property
    def required_weights(self) -> List[str]:
        n = self.lib.deer_num_required_weights(self._h)
        return [self.lib.deer_required_weight_name(self._h, i).decode() for i in range(n)]

    def stats(self) -> DeerStats:
        s = DeerStats()
        self._check(self.lib.deer_get_stats(self._h, C.byref(s)))
        return s

    @staticmethod
    def set_launch_overlap(enable: bool) -> bool:
        """Process-wide switch of programmatic dependent launch (see include/deer_engine.h); returns the previous setting."""
        return bool(_lib.load().deer_set_launch_overlap(int(enable)))

    @staticmethod
    def set_row_order(snake: bool) -> bool:
        """Process-wide switch of the "snake" row order of LayerNorm / ViT attention (see include/deer_engine.h); returns the previous setting."""
        return bool(_lib.load().deer_set_row_order(int(snake)))

    def set_profile(self, enable: bool):
        self._check(self.lib.deer_set_profile(self._h, int(enable)))

    def profile(self) -> dict:
        """Per-stage GEMM device time / exact algorithmic FLOPs of the last step (synchronises)."""
        p = DeerProfile()
        self._check(self.lib.deer_get_profile(self._h, C.byref(p)))
        stages = ("vit", "perceiver", "llm", "head")
        return {"gemm_ms": {s: p.gemm_ms[i] for i, s in enumerate(stages)},
                "gemm_flops": {s: p.gemm_flops[i] for i, s in enumerate(stages)},
                "gemm_bytes": {s: p.gemm_bytes[i] for i, s in enumerate(stages)},
                "gemm_launches": {s: int(p.gemm_launches[i]) for i, s in enumerate(stages)},
                "step_ms": p.step_ms, "active_after": [int(p.active_after[i]) for i in range(p.n_probes)]}

    # ---- weights ----------------------------------------------------------------------------
    def load_state_dict(self, sd: Mapping[str, "torch.Tensor | np.ndarray"], strict: bool = True):
        """Accepts the reference's state-dict keys (optionally 'module.'-prefixed, train_utils.py:631-638);
        tensors not on the hot path are skipped like the reference's strict=False loads (eval_calvin.py:543,577)."""
        used = 0
        for name, t in (sd.items() if hasattr(sd, "items") else sd):
            if isinstance(t, torch.Tensor):
                a = t.detach().to(torch.float32).cpu().contiguous().numpy()
            else:
                a = np.ascontiguousarray(t, dtype=np.float32)
            shape = (C.c_int64 * max(1, a.ndim))(*a.shape)
            rc = self._check(self.lib.deer_load_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim))
            used += (rc == 0)
        if strict:
            self.finalize()
        return used

    def finalize(self):
        self._check(self.lib.deer_finalize_weights(self._h))

    # ---- exit controller state ------------------------------------------------------------------
    def set_thresholds(self, thresholds: Sequence[float]):
        """ExitController._set_threshold_value (value_net.py:177-182)."""
        arr = (C.c_float * len(thresholds))(*[float(x) for x in thresholds])
        self._check(self.lib.deer_set_thresholds(self._h, arr, len(thresholds)))
        self.thresholds = [float(x) for x in thresholds]

    def reset(self, slots: Optional[Iterable[int]] = None):
        """ModelWrapper.reset -> clear_all_exit_memory (eval_utils.py:252-277)."""
        if slots is None:
            self._check(self.lib.deer_reset(self._h, None, -1))
        else:
            sl = list(slots)
            arr = (C.c_int32 * len(sl))(*sl)
            self._check(self.lib.deer_reset(self._h, arr, len(sl)))

    # ---- the hot path -----------------------------------------------------------------------------
    def _pin(self, key: str, shape, dtype) -> torch.Tensor:
        t = self._pinned.get(key)
        numel = int(np.prod(shape))
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(numel, dtype=dtype, pin_memory=True)
            self._pinned[key] = t
        return t[:numel].view(*shape)

    def step(self, rgb: torch.Tensor, gripper: torch.Tensor, ids: torch.Tensor, mask: Optional[torch.Tensor] = None,
             slots: Optional[torch.Tensor] = None, timesteps: Optional[torch.Tensor] = None, exit_id: Optional[int] = None,
             full_depth: bool = False, want_delta: bool = True) -> Dict[str, torch.Tensor]:
        """rgb/gripper (n,3,H,W) f32 normalised, ids (n,S) int64, mask (n,S).  CUDA tensors -> device path
        (outputs are CUDA tensors, no sync); CPU tensors -> ``deer_step_host`` (pinned staging, H2D + D2H inside)."""
        n, S = int(ids.shape[0]), int(ids.shape[1])
        on_dev = rgb.is_cuda
        dev = self.device if on_dev else torch.device("cpu")
        if mask is None:
            mask = torch.ones(n, S, dtype=torch.uint8, device=dev)
        if slots is None:
            slots = torch.arange(n, dtype=torch.int32, device=dev)
        rgb = rgb.reshape(n, 3, self.cfg.image_size, self.cfg.image_size).to(torch.float32).contiguous()
        gripper = gripper.reshape(n, 3, self.cfg.image_size, self.cfg.image_size).to(torch.float32).contiguous()
        ids = ids.to(torch.int64).contiguous()
        mask = mask.to(torch.uint8).contiguous()
        slots = slots.to(torch.int32).contiguous()
        if timesteps is not None:
            timesteps = timesteps.to(torch.int32).contiguous()
        flags = _lib.FLAG_FULL_DEPTH if full_depth else 0
        sx = -1 if exit_id is None else int(exit_id if exit_id >= 0 else exit_id + self.cfg.n_layers)
        nthr = max(1, self.cfg.n_thresholds)
        if on_dev:
            for t in (gripper, ids, mask, slots):
                assert t.is_cuda, "all inputs must live on the same side"
            out = {"pose": torch.empty(n, 6, dtype=torch.float32, device=dev), "gripper": torch.empty(n, dtype=torch.float32, device=dev),
                   "action": torch.empty(n, 7, dtype=torch.float32, device=dev), "exit_layer": torch.empty(n, dtype=torch.int32, device=dev),
                   "delta": torch.empty(n, nthr, dtype=torch.float32, device=dev) if want_delta else None}
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.deer_step(self._h, n, _ptr(slots), _ptr(rgb), _ptr(gripper), _ptr(ids), _ptr(mask), S, _ptr(timesteps),
                                           sx, flags, _ptr(out["pose"]), _ptr(out["gripper"]), _ptr(out["action"]),
                                           _ptr(out["exit_layer"]), _ptr(out["delta"]), C.c_void_p(stream)))
            return out
        out = {"pose": self._pin("pose", (n, 6), torch.float32), "gripper": self._pin("grip", (n,), torch.float32),
               "action": self._pin("action", (n, 7), torch.float32), "exit_layer": self._pin("exit", (n,), torch.int32),
               "delta": self._pin("delta", (n, nthr), torch.float32) if want_delta else None}
        self._check(self.lib.deer_step_host(self._h, n, _ptr(slots), _ptr(rgb), _ptr(gripper), _ptr(ids), _ptr(mask), S, _ptr(timesteps),
                                            sx, flags, _ptr(out["pose"]), _ptr(out["gripper"]), _ptr(out["action"]),
                                            _ptr(out["exit_layer"]), _ptr(out["delta"])))
        return out

    def step_generate(self, rgb: torch.Tensor, gripper: torch.Tensor, ids: torch.Tensor, commit_layer: torch.Tensor,
                      mask: Optional[torch.Tensor] = None, slots: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One frame of the calibration pass in the training layout (``deer_step_generate``): CUDA tensors in, returns the
        delta columns of this frame, (n, len(exit_list)), on the device.  ``commit_layer`` (n,) int: the exit layer whose
        candidate LSTM state is carried to the next frame (flamingo_mpt.py:484-492)."""
        n, S = int(ids.shape[0]), int(ids.shape[1])
        dev = self.device
        assert rgb.is_cuda and gripper.is_cuda
        if mask is None:
            mask = torch.ones(n, S, dtype=torch.uint8, device=dev)
        if slots is None:
            slots = torch.arange(n, dtype=torch.int32, device=dev)
        rgb = rgb.reshape(n, 3, self.cfg.image_size, self.cfg.image_size).to(torch.float32).contiguous()
        gripper = gripper.reshape(n, 3, self.cfg.image_size, self.cfg.image_size).to(torch.float32).contiguous()
        ids = ids.to(dev, torch.int64).contiguous()
        mask = mask.to(dev, torch.uint8).contiguous()
        slots = slots.to(dev, torch.int32).contiguous()
        commit_layer = commit_layer.to(dev, torch.int32).contiguous()
        delta = torch.empty(n, len(self.exit_list), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.deer_step_generate(self._h, n, _ptr(slots), _ptr(rgb), _ptr(gripper), _ptr(ids), _ptr(mask), S,
                                                _ptr(commit_layer), _ptr(delta), C.c_void_p(stream)))
        return delta

    def preprocess(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """On-device open_clip image transform: CUDA uint8 [n,H,W,3] -> fp32 [n,3,S,S] (bit-exact Pillow bicubic + normalise)."""
        assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.ndim == 4 and frames_u8.shape[-1] == 3
        f = frames_u8.contiguous()
        n, H, W, _ = f.shape
        out = torch.empty(n, 3, self.cfg.image_size, self.cfg.image_size, dtype=torch.float32, device=f.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.deer_preprocess_u8(self._h, _ptr(f), n, H, W, _ptr(out), C.c_void_p(stream)))
        return out

    def step_u8(self, rgb_u8: torch.Tensor, gripper_u8: torch.Tensor, ids: torch.Tensor, mask: Optional[torch.Tensor] = None,
                slots: Optional[torch.Tensor] = None, timesteps: Optional[torch.Tensor] = None, exit_id: Optional[int] = None,
                full_depth: bool = False) -> Dict[str, torch.Tensor]:
        """Raw camera frames in (uint8 [n,H,W,3], HOST tensors; pinned for full speed): H2D of the uint8 frames, the image
        transform, the step and the D2H of the results all happen inside ``deer_step_host_u8``."""
        n, S = int(ids.shape[0]), int(ids.shape[1])
        assert not rgb_u8.is_cuda and rgb_u8.dtype == torch.uint8 and gripper_u8.dtype == torch.uint8
        rgb_u8, gripper_u8 = rgb_u8.contiguous(), gripper_u8.contiguous()
        if mask is None:
            mask = torch.ones(n, S, dtype=torch.uint8)
        if slots is None:
            slots = torch.arange(n, dtype=torch.int32)
        ids, mask, slots = ids.to(torch.int64).contiguous(), mask.to(torch.uint8).contiguous(), slots.to(torch.int32).contiguous()
        if timesteps is not None:
            timesteps = timesteps.to(torch.int32).contiguous()
        flags = _lib.FLAG_FULL_DEPTH if full_depth else 0
        sx = -1 if exit_id is None else int(exit_id if exit_id >= 0 else exit_id + self.cfg.n_layers)
        nthr = max(1, self.cfg.n_thresholds)
        out = {"pose": self._pin("pose", (n, 6), torch.float32), "gripper": self._pin("grip", (n,), torch.float32),
               "action": self._pin("action", (n, 7), torch.float32), "exit_layer": self._pin("exit", (n,), torch.int32),
               "delta": self._pin("delta", (n, nthr), torch.float32)}
        self._check(self.lib.deer_step_host_u8(self._h, n, _ptr(slots), _ptr(rgb_u8), rgb_u8.shape[1], rgb_u8.shape[2], _ptr(gripper_u8),
                                               gripper_u8.shape[1], gripper_u8.shape[2], _ptr(ids), _ptr(mask), S, _ptr(timesteps), sx, flags,
                                               _ptr(out["pose"]), _ptr(out["gripper"]), _ptr(out["action"]), _ptr(out["exit_layer"]),
                                               _ptr(out["delta"])))
        return out

    # ---- parity taps ------------------------------------------------------------------------------
    def get_tap(self, name: str, layer: int = 0) -> np.ndarray:
        cap = {"vit": 2 * self.max_batch * (self.cfg.image_size // self.cfg.patch) ** 2 * self.cfg.vit_width,
               "media": self.max_batch * 2 * self.cfg.num_latents * self.cfg.vit_width,
               "hidden": self.max_batch * self.max_seq * self.cfg.d_model,
               "pooled": self.max_batch * self.cfg.d_model}[name]
        buf = np.empty(cap, dtype=np.float32)
        shape = (C.c_int64 * 4)()
        self._check(self.lib.deer_get_tap(self._h, name.encode(), layer, buf.ctypes.data_as(C.c_void_p), cap, shape))
        dims = [int(s) for s in shape if s > 0]
        n = int(np.prod(dims))
        return buf[:n].reshape(dims).copy()

    def set_state(self, slot: int, h: np.ndarray, c: np.ndarray):
        """Overwrite one slot's carried LSTM state -- `model.extra_exit.hidden_state = (h, c)` (action_head.py:474,554-558)."""
        shp = (self.cfg.lstm_layers, self.cfg.lstm_hidden)
        h = np.ascontiguousarray(h, dtype=np.float32).reshape(shp)
        c = np.ascontiguousarray(c, dtype=np.float32).reshape(shp)
        self._check(self.lib.deer_set_state(self._h, slot, h.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)))

    def features(self, n: int, device: bool = True) -> torch.Tensor:
        """`.logits[2]` of the last step: (n, d_model) max-pooled token feature each episode's exit head consumed
        (action_head.py:519-520,608-609), as a CUDA tensor (stream-ordered copy) or a host tensor."""
        if device:
            out = torch.empty(n, self.cfg.d_model, dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.deer_get_features(self._h, n, _ptr(out), 0, C.c_void_p(stream)))
        else:
            out = torch.empty(n, self.cfg.d_model, dtype=torch.float32)
            self._check(self.lib.deer_get_features(self._h, n, _ptr(out), 1, None))
        return out

    def get_state(self, slot: int):
        shp = (self.cfg.lstm_layers, self.cfg.lstm_hidden)
        h = np.empty(shp, dtype=np.float32)
        c = np.empty(shp, dtype=np.float32)
        self._check(self.lib.deer_get_state(self._h, slot, h.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)))
        return h, c


# ---- single-operator access (parity tests of individual kernels) --------------------------------
def op_gemm(A: torch.Tensor, W: torch.Tensor, bias=None, act: int = 0, scale=None, resid=None, out_f32: bool = False,
            impl: str = "tcgen05", m_dev: Optional[torch.Tensor] = None, m_mul: int = 1) -> torch.Tensor:
    """C = epi(A @ W^T) through ``deer_op_gemm``.  A [M,K], W [N,K] both bf16 (or both fp32 for the SIMT path)."""
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    dtype = 1 if A.dtype == torch.float32 else 0
    odt = torch.float32 if (out_f32 or dtype == 1) else torch.bfloat16
    if resid is not None:
        Cm = resid.clone()
        rp = _ptr(Cm)
    else:
        Cm = torch.zeros(M, N, dtype=odt, device=A.device)
        rp = None
    stream = torch.cuda.current_stream(A.device).cuda_stream
    rc = lib.deer_op_gemm({"tcgen05": 0, "simt": 1, "tcgen05_1cta": 2, "tcgen05_splitk": 3, "tcgen05_1cta_m128": 4}[impl], dtype, _ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(Cm), Cm.stride(0),
                          M, N, K, _ptr(bias), act, _ptr(scale), rp, int(out_f32 or resid is not None), _ptr(m_dev), m_mul,
                          C.c_void_p(stream))
    if rc != 0:
        raise DeerError(rc, lib.deer_last_error(None).decode())
    return Cm


def op_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, impl: str = "tc") -> torch.Tensor:
    """q [B,Sq,H,hd], k/v [B,Sk,H,hd] (contiguous) -> [B,Sq,H,hd] through ``deer_op_attention`` (no mask)."""
    lib = _lib.load()
    B, Sq, H, hd = q.shape
    Sk = k.shape[1]
    dtype = 1 if q.dtype == torch.float32 else 0
    kv = torch.stack([k, v], dim=2).contiguous()            # [B,Sk,2,H,hd]: K and V share a row pitch like in the fused qkv buffers
    o = torch.empty_like(q)
    stream = torch.cuda.current_stream(q.device).cuda_stream
    kp = kv.data_ptr()
    vp = kp + H * hd * kv.element_size()
    rc = lib.deer_op_attention({"tc": 0, "simt": 1, "flash": 3}[impl], dtype, _ptr(q), H * hd, C.c_void_p(kp), C.c_void_p(vp), 2 * H * hd,
                               _ptr(o), H * hd, B, H, Sq, Sk, hd, float(scale), C.c_void_p(stream))
    if rc != 0:
        raise DeerError(rc, lib.deer_last_error(None).decode())
    return o


def op_attention_masked(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, impl: str = "auto", causal: bool = False,
                        alibi_slopes: Optional[torch.Tensor] = None, kv_len: Optional[torch.Tensor] = None,
                        qmode: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The LLM-side attentions with their masks through ``deer_op_attention_masked``; shapes as in ``op_attention``.
    impl "simt" (reference), "mma" (mma.sync kernel) or "auto" (the engine's choice)."""
    lib = _lib.load()
    B, Sq, H, hd = q.shape
    Sk = k.shape[1]
    dtype = 1 if q.dtype == torch.float32 else 0
    kv = torch.stack([k, v], dim=2).contiguous()
    o = torch.zeros_like(q)
    stream = torch.cuda.current_stream(q.device).cuda_stream
    kp = kv.data_ptr()
    vp = kp + H * hd * kv.element_size()
    opt = [None if t is None else t.contiguous() for t in (alibi_slopes, kv_len, qmode, n_dev)]
    rc = lib.deer_op_attention_masked({"simt": 1, "auto": 2, "mma": 4}[impl], dtype, _ptr(q), H * hd, C.c_void_p(kp), C.c_void_p(vp),
                                      2 * H * hd, _ptr(o), H * hd, B, H, Sq, Sk, hd, float(scale), int(causal),
                                      _ptr(opt[0]), _ptr(opt[1]), _ptr(opt[2]), _ptr(opt[3]), C.c_void_p(stream))
    if rc != 0:
        raise DeerError(rc, lib.deer_last_error(None).decode())
    return o
