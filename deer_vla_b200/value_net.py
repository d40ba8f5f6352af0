"""Host-side mirror of ``robot_flamingo/models/value_net.py`` for the engine.

``ExitController`` keeps the reference's constructor, attributes and methods
(``set_threshold`` / ``_set_threshold_value`` / ``set_timestep`` / ``thresholds`` dict /
``max_layer``), but its ``forward`` is not a Python probe: the consistency gate
(value_net.py:105-133, 277-297) runs inside the engine's ``exit_gate_kernel`` for every active
episode at once.  Attaching an engine pushes the thresholds to the device.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


def exit_probabilities(n_exit: int, exit_ratio: float, exit_dist: str = "exp", model_name: str = "mpt_dolly_3b") -> torch.Tensor:
    """p_k of value_net.py:216-238 (float32 torch arithmetic, like the reference, so floor(N*p_k) agrees)."""
    if exit_dist == "exp":
        probs = exit_ratio ** torch.arange(1, n_exit + 1)
    elif exit_dist == "gauss":
        probs = torch.tensor([math.exp(-(i - exit_ratio) ** 2 / (2 * 1.0 ** 2)) for i in range(n_exit)])
    elif exit_dist == "gamma":
        import scipy.stats
        x = torch.arange(1, n_exit + 1, dtype=torch.float32)
        probs = torch.tensor([scipy.stats.gamma.pdf(val, exit_ratio, scale=2.0) for val in x], dtype=torch.float32)
    else:
        raise ValueError("Unsupported exit distribution")
    if "mpt_9b" in model_name:
        probs[0] = 0          # only enable exits from at least the 4th layer for the deep model (value_net.py:235-236)
    probs /= probs.sum()
    return probs


def solve_thresholds(values: torch.Tensor, exit_ratio: float, exit_dist: str = "exp", model_name: str = "mpt_dolly_3b",
                     leq: bool = True) -> torch.Tensor:
    """The greedy solver of ``ExitController.set_threshold`` (value_net.py:203-264): at exit k, the
    floor(N*p_k) smallest-delta samples that have not left at an earlier exit leave; T_k is the delta of
    the last of them.  Vectorised (argsort + masked cumulative count) instead of the reference's python scan."""
    vals = values.detach().to(torch.float32).cpu()
    n_stage, n_sample = vals.shape
    order = torch.argsort(vals, dim=1, descending=not leq, stable=True).numpy()
    v = vals.numpy()
    filtered = np.zeros(n_sample, dtype=bool)
    T = torch.full((n_stage,), -1e8 if leq else 1e8)
    probs = exit_probabilities(n_stage, exit_ratio, exit_dist, model_name)
    for k in range(n_stage - 1):
        out_n = math.floor(n_sample * probs[k])
        alive_sorted = ~filtered[order[k]]
        count = np.cumsum(alive_sorted)
        hit = np.nonzero(alive_sorted & (count == out_n))[0]
        if out_n >= 1 and hit.size:
            T[k] = float(v[k, order[k][hit[0]]])
        tk = np.float32(T[k])
        filtered |= (v[k] <= tk) if leq else (v[k] >= tk)
    T[n_stage - 1] = 1e8 if leq else -1e8
    return T


class ExitController:
    """Same construction as the reference (value_net.py:163-173).  ``value_net`` may be None: the
    action-consistency metric lives in the engine."""

    def __init__(self, value_net=None, exit_id_list: Sequence[int] = (), steps_per_stage: int = 1, exit_dist: str = "exp",
                 leq: bool = True, max_layer: int = 12, engine=None):
        assert leq, "the engine implements the reference's evaluation setting leq=True (eval_calvin.py:584)"
        self.value_net = value_net
        self.thresholds: Optional[Dict[int, float]] = None
        self.leq = leq
        self.exit_id_list = list(exit_id_list)
        self.num_exit = len(self.exit_id_list)
        self.steps_per_stage = steps_per_stage
        self.exit_dist = exit_dist
        self.max_layer = min(max_layer - 1, self.exit_id_list[-1])
        self.cur_step = 0
        self.engine = None
        if engine is not None:
            self.attach(engine)

    @property
    def module(self):          # the reference wraps the controller in DDP and reaches through .module (eval_calvin.py:590)
        return self

    def eval(self):
        return self

    def attach(self, engine):
        """Bind to an engine.  The gate runs inside the engine with the ENGINE's exit bookkeeping, so a controller built with
        different options would silently be ignored: refuse instead (value_net.py:163-173,285-286)."""
        ecfg = engine.cfg
        if list(engine.exit_list) != list(self.exit_id_list):
            raise ValueError(f"exit_id_list {self.exit_id_list} != the engine's exit list {engine.exit_list} (flamingo_mpt.py:268-270)")
        if ecfg.ctrl_max_layer != self.max_layer:
            raise ValueError(f"controller max_layer {self.max_layer} != engine {ecfg.ctrl_max_layer} (value_net.py:173)")
        if int(ecfg.steps_per_stage) != int(self.steps_per_stage):
            raise ValueError(f"controller steps_per_stage {self.steps_per_stage} != engine {ecfg.steps_per_stage} (value_net.py:285)")
        self.engine = engine
        if self.thresholds is not None:
            engine.set_thresholds([self.thresholds[e] for e in self._real_exits()])

    def _real_exits(self) -> List[int]:
        return [x for x in self.exit_id_list if x <= self.max_layer]

    def _set_threshold_value(self, thresholds: Sequence[float]):
        real = self._real_exits()
        assert len(thresholds) == len(real)                              # value_net.py:179
        self.thresholds = {real[i]: float(thresholds[i]) for i in range(len(real))}
        if self.engine is not None:
            self.engine.set_thresholds([self.thresholds[e] for e in real])

    def set_threshold(self, args, model, dataloader, exit_ratio, model_name, values=None):
        """values: (n_real_exit, n_sample) delta table.  With values=None the table is produced by the engine in
        full-depth mode over ``dataloader`` (an iterable of step-input dicts) -- the GPU replacement of
        ``generate_action_values`` (value_net.py:300-399)."""
        if values is None:
            values = generate_action_values(self.engine if self.engine is not None else model, dataloader)
            # every rank calibrates on its own shard; the solver sees all of them (value_net.py:192-201)
            from .sharding import all_gather_values
            values = all_gather_values(torch.as_tensor(values))
        real = self._real_exits()
        values = torch.as_tensor(values)[: len(real)]
        T = solve_thresholds(values, exit_ratio, self.exit_dist, model_name, self.leq)
        self.thresholds = {real[i]: float(T[i]) for i in range(len(real))}
        if self.engine is not None:
            self.engine.set_thresholds([self.thresholds[e] for e in real])
        return values

    def set_timestep(self, t: int):
        self.cur_step = t


@torch.no_grad()
def generate_action_values(engine, batches) -> torch.Tensor:
    """Delta table (n_exit, n_sample) from full-depth engine steps.  ``batches`` yields dicts with the keyword
    arguments of ``DeerEngine.step`` (rgb, gripper, ids, [mask, slots]); consecutive dicts are consecutive
    timesteps of the same episodes (the LSTM state is carried), a dict with ``reset=True`` starts new episodes."""
    cols = []
    for b in batches:
        b = dict(b)
        if b.pop("reset", False):
            engine.reset()
        out = engine.step(full_depth=True, **b)
        cols.append(out["delta"].detach().float().cpu().t())
    return torch.cat(cols, dim=1)


@torch.no_grad()
def generate_action_values_window(engine, windows, window_size: int = 12, layer_indices=None, generator=None) -> torch.Tensor:
    """``generate_action_values`` in the reference's own layout (value_net.py:300-399): ``windows`` yields dicts with
    ``rgb``/``gripper`` (bs, window, 3, H, W) CUDA tensors and ``ids`` (bs, S) [``mask`` (bs, S)] -- what the CALVIN loader
    hands over (value_net.py:331-346).  Per window the exit layer whose features stand in for each past frame is drawn
    uniformly from the exit list (flamingo_mpt.py:486-487) unless ``layer_indices`` supplies an iterable of (bs, window)
    tensors.  Returns the (n_exit, n_windows * window_size//2) delta table ``ExitController.set_threshold(values=...)`` takes
    (value_net.py:134-160: seq_id = window//2-1 .. window-2, bs-major order)."""
    exits = torch.tensor(engine.exit_list, dtype=torch.int64)
    li_iter = iter(layer_indices) if layer_indices is not None else None
    cols = []
    for w in windows:
        rgb, grip, ids = w["rgb"], w["gripper"], w["ids"]
        bs = rgb.shape[0]
        assert rgb.shape[1] == window_size and bs <= engine.max_batch
        if li_iter is not None:
            li = next(li_iter).to(torch.int64).cpu()
        else:
            li = exits[torch.randint(0, len(exits), (bs, window_size), generator=generator)]
        engine.reset(range(bs))
        per_t = []
        for t in range(window_size - 1):
            d = engine.step_generate(rgb[:, t], grip[:, t], ids, li[:, t], mask=w.get("mask"))
            if t >= window_size // 2 - 1:
                per_t.append(d)
        # (T/2, bs, n_exit) -> (n_exit, bs, T/2) -> (n_exit, bs*T/2)
        cols.append(torch.stack(per_t).permute(2, 1, 0).reshape(len(exits), -1).float().cpu())
    return torch.cat(cols, dim=1)
