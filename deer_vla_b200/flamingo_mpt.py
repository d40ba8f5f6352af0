"""``MPTFlamingo`` call surface (robot_flamingo/models/flamingo_mpt.py) on top of the CUDA engine.

Same constructor-time attributes the reference's harness reads (eval_utils.py:192-247,300-324,456,537,666;
eval_calvin.py:583) and the same ``forward`` keywords used on the evaluation path
(flamingo_mpt.py:308-333).  One instance serves ``batch`` concurrent episodes (slot i = episode i); with
batch 1 it is a drop-in for the reference's one-episode-per-model use.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import torch

from .config import DeerConfig
from .engine import DeerEngine


class _HeadState:
    """Stands in for ``model.extra_exit``: the harness only touches ``hidden_state`` / ``history_memory``
    (flamingo_mpt.py:292-306; eval_utils.py:276-277).  The real (h,c) live in the engine, per slot."""

    def __init__(self, owner):
        self._owner = owner
        self.history_memory: list = []
        self.window_size = 1

    @property
    def hidden_state(self):
        return self._owner._state_marker

    @hidden_state.setter
    def hidden_state(self, v):
        if v is None:
            self._owner.engine.reset()
            self._owner._state_marker = None
        else:
            # (h, c) each (layers, B, hidden) like nn.LSTM's state (action_head.py:554-558): written into slots 0..B-1
            h, c = v
            h, c = torch.as_tensor(h).detach().float().cpu().numpy(), torch.as_tensor(c).detach().float().cpu().numpy()
            for b in range(h.shape[1]):
                self._owner.engine.set_state(b, h[:, b], c[:, b])
            self._owner._state_marker = "engine"


class MPTFlamingo:
    def __init__(self, engine: DeerEngine, cfg: DeerConfig, window_size: int = 12, use_gripper: bool = True,
                 fusion_mode: str = "post", pad_length: int = -1, replan: int = -1, refresh: int = -1,
                 multi_step_action: int = 1, tcp_rel: bool = False, use_state: bool = False):
        assert fusion_mode == "post" and use_gripper and not use_state, \
            "the engine implements the DeeR evaluation configuration (eval_calvin.py:173-177; SURVEY App. A-12)"
        self.engine, self.cfg = engine, cfg
        self.window_size, self.pad_length, self.replan, self.refresh = window_size, pad_length, replan, refresh
        if self.replan != -1:                              # flamingo_mpt.py:102-104
            self.replan = min(int(replan * self.window_size), 180)
        self.use_gripper, self.use_state, self.fusion_mode = use_gripper, use_state, fusion_mode
        self.decoder_type, self.head_type = "lstm", "deterministic"
        self.sep_lm_head, self.act_step, self.tcp_rel = True, multi_step_action, tcp_rel
        self.use_hist, self.use_diff, self.sep_resampler = False, False, False
        self.early_exit_layer = cfg.early_exit_layer
        self.llm_inference_time = -1.0
        self.lang_encoder = SimpleNamespace(config=SimpleNamespace(n_layers=cfg.n_layers, d_model=cfg.d_model))
        self.lm_exits = {i: None for i in range(cfg.exit_interval - 1, cfg.early_exit_layer, cfg.exit_interval)}
        self._state_marker = None
        self.extra_exit = _HeadState(self)
        self.lm_head = _HeadState(self)
        self.training = False

    # DDP / nn.Module shims used by the harness (eval_calvin.py:565-568)
    @property
    def module(self):
        return self

    def eval(self):
        return self

    def get_all_exit_idx(self) -> List[int]:           # flamingo_mpt.py:268-270
        return list(self.lm_exits.keys()) + [self.lang_encoder.config.n_layers - 1]

    def get_exit_num(self) -> int:
        return len(self.get_all_exit_idx())

    def set_all_exit_window_size(self, new_window_size: int) -> int:   # flamingo_mpt.py:275-290
        old = self.extra_exit.window_size
        if new_window_size != 1:
            raise NotImplementedError("the engine runs the single-step inference mode (window_size=1, eval_utils.py:309)")
        self.extra_exit.window_size = self.lm_head.window_size = new_window_size
        return old

    def clear_all_exit_memory(self):                   # flamingo_mpt.py:292-306
        self.engine.reset()
        self._state_marker = None
        self.extra_exit.history_memory = []
        self.lm_head.history_memory = []

    def load_state_dict(self, sd, strict: bool = False):
        sd = sd.get("model_state_dict", sd) if isinstance(sd, dict) else sd
        return self.engine.load_state_dict(sd, strict=False)

    def finalize(self):
        self.engine.finalize()

    @torch.no_grad()
    def forward(self, vision_x: torch.Tensor, lang_x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, labels=None,
                use_cached_vision_x: bool = False, clear_conditioned_layers: bool = True, past_key_values=None, use_cache: bool = False,
                vision_gripper: Optional[torch.Tensor] = None, state_tensor=None, return_feature: bool = False, policy_mask=None,
                act=None, deterministic: bool = False, with_gripper_logits: bool = False, exit_id: Optional[int] = None,
                dynamic_early_exit: bool = False, exit_controller=None, slots: Optional[torch.Tensor] = None, **unused):
        assert vision_x is not None and vision_gripper is not None and not use_cached_vision_x
        assert vision_x.ndim == 6 and vision_x.shape[1] == 1 and vision_x.shape[2] == 1, \
            "vision_x should be of shape (b, T_img=1, F=1, C, H, W)"     # flamingo_mpt.py:570-572
        assert exit_controller is None or exit_id is None, "Only one exit indicator can be sepcified!"   # mosaic_gpt_3b.py:300
        B = vision_x.shape[0]
        timesteps = None
        if dynamic_early_exit and exit_controller is not None:
            assert exit_controller.thresholds is not None, "Please set thresholds before calling forward"  # value_net.py:279
            if exit_controller.engine is not self.engine:
                exit_controller.attach(self.engine)
            timesteps = torch.full((B,), int(getattr(exit_controller, "cur_step", 0)), dtype=torch.int32, device=vision_x.device)
            static_exit = None
        else:
            static_exit = (self.cfg.n_layers - 1) if exit_id is None else exit_id      # static LLM forward -> last layer's head
        out = self.engine.step(vision_x.reshape(B, *vision_x.shape[3:]), vision_gripper.reshape(B, *vision_gripper.shape[3:]),
                               lang_x, attention_mask, slots=slots, timesteps=timesteps, exit_id=static_exit)
        self._state_marker = "engine"
        exit_layer = out["exit_layer"]
        # .logits[2]: the pooled feature the exit head consumed (return_feature=True, flamingo_mpt.py:458-460)
        feat = self.engine.features(B, device=vision_x.is_cuda) if return_feature else None
        # .hidden_states (opt-in: engine built with debug_taps): tuple of (B, S, d) post-block states of layers 0..exit
        # (mosaic_gpt_3b.py:424-427); rows of episodes that left earlier keep the state they left with
        hidden = None
        if self.engine.debug_taps:
            last = int(exit_layer.max())
            hidden = tuple(torch.from_numpy(self.engine.get_tap("hidden", l)) for l in range(last + 1))
        return SimpleNamespace(
            logits=(out["pose"].view(B, 1, 6), out["gripper"].view(B, 1, 1), feat),
            exit_layer=int(exit_layer[0]) if B == 1 else exit_layer,
            exit_layers=exit_layer, action=out["action"], delta=out["delta"], hidden_states=hidden, past_key_values=None)

    __call__ = forward
