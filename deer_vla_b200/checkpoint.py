"""DeeR checkpoint handling (robot_flamingo/eval/eval_calvin.py:356-489, robot_flamingo/train/train_utils.py:27-60,631-638).

A DeeR ``.pth`` is a dict: ``model_state_dict`` (trainable tensors only, keys prefixed ``module.``), head hyper-parameters
(``head_type, early_exit_layer, multi_exit, exit_interval, mlp_layernorm, lstm_layernorm, mlp_num_hidden_layers,
lstm_num_layers, pooling`` ...), optionally the cached delta table ``values`` (eval_calvin.py:601-611), optimizer state.
More configuration is smuggled through the FILE NAME and re-parsed at evaluation time.  This module restates that
bookkeeping for the engine and refuses configurations the engine does not implement instead of silently diverging.
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, List, Mapping, Optional, Tuple

from .config import DeerConfig, config_3b, config_9b

LLM_NAMES = ["mpt_3b", "mpt_4b", "mpt_9b", "mpt_dolly_3b", "mpt_base_4b"]          # eval_calvin.py:417
_N_LAYERS = {"mpt_3b": 24, "mpt_dolly_3b": 24, "mpt_9b": 32}                        # eval_calvin.py:481-487


def parse_checkpoint_name(path: str, default_llm: str = "mpt_dolly_3b") -> Dict[str, Any]:
    """Flags derived from the checkpoint file name exactly as eval_calvin.py:356-421 does (substring tests included)."""
    name = path
    out: Dict[str, Any] = {}
    out["use_state"] = "state" in name
    out["real_data"] = "real" in name
    m = re.search(r"aug_(\d+)_(\d+)", name)
    out["rgb_pad"], out["gripper_pad"] = (int(m.group(1)), int(m.group(2))) if m else (-1, -1)
    out["sep_resampler"] = "sep" in name
    out["sep_lm_head"] = "lm_head" in name
    out["residual"] = "res_" in name
    out["tcp_rel"] = "tcp" in name
    parts = name.split("_")
    out["multi_step_action"] = int(parts[parts.index("step") - 1]) if "step" in parts else 1
    out["text_aug"] = "text_aug" in name
    out["traj_cons"] = "traj_cons" in name
    if "difws" in name:
        ix = parts.index("difws")
        out["dif_ws"], out["min_window_size"], out["max_window_size"] = True, int(parts[ix + 1]), int(parts[ix + 2])
        out["window_size"] = out["max_window_size"]
    else:
        out["dif_ws"] = False
    if "latent" in name:
        out["global_latent"] = int(parts[parts.index("latent") + 1])
    out["no_image_patch"] = "no_image_patch" in name
    if "gpt" in name:
        out["decoder_type"], out["hidden_size"] = "gpt", int(parts[parts.index("gpt") + 1])
    out["llm_name"] = default_llm
    for n in LLM_NAMES:                           # first match in this order wins, like the reference's loop
        if n in name:
            out["llm_name"] = n
            break
    return out


def readout_args(ckpt: Mapping[str, Any], llm_name: str, max_layer: Optional[int] = None) -> Dict[str, Any]:
    """Head hyper-parameters with the reference's defaults for old checkpoints (eval_calvin.py:455-489)."""
    defaults = dict(head_type="deterministic", early_exit_layer=-1, multi_exit=False, exit_interval=1, exit_dropout=0.0,
                    lstm_dropout=0.0, dropout_mode="wo_last", mlp_layernorm=False, lstm_layernorm=False,
                    mlp_num_hidden_layers=3, lstm_num_layers=4, pooling="max")
    a = {k: ckpt.get(k, v) for k, v in defaults.items()}
    if "layernorm" in ckpt:                       # compatibility with old code
        a["mlp_layernorm"] = ckpt["layernorm"]
    if a["early_exit_layer"] < 0:
        if llm_name not in _N_LAYERS:
            raise NotImplementedError(llm_name)
        a["early_exit_layer"] += _N_LAYERS[llm_name]
    a["max_layer"] = a["early_exit_layer"] + 1 if max_layer in (None, -1) else max_layer
    return a


def config_from_checkpoint(ckpt: Mapping[str, Any], path: str = "", max_layer: Optional[int] = None,
                           steps_per_stage: int = 1, threshold_type: str = "L2") -> DeerConfig:
    """DeerConfig for a DeeR checkpoint; raises for configurations outside the engine's scope."""
    flags = parse_checkpoint_name(os.path.basename(path)) if path else {"llm_name": "mpt_dolly_3b"}
    a = readout_args(ckpt, flags["llm_name"], max_layer)
    unsupported = []
    if a["head_type"] != "deterministic":
        unsupported.append(f"head_type={a['head_type']}")
    n_hidden = head_num_hidden_layers(a)
    if n_hidden not in (2, 3):
        unsupported.append(f"mlp_num_hidden_layers={n_hidden} (the engine builds 2 or 3 hidden layers)")
    if a["dropout_mode"] not in ("layerwise", "last", "wo_last"):
        unsupported.append(f"dropout_mode={a['dropout_mode']}")
    if a["pooling"] != "max":
        unsupported.append(f"pooling={a['pooling']}")
    for k in ("use_state", "sep_resampler", "residual", "tcp_rel", "no_image_patch"):
        if flags.get(k):
            unsupported.append(k)
    if flags.get("multi_step_action", 1) != 1 or flags.get("decoder_type", "lstm") != "lstm":
        unsupported.append("multi_step_action / gpt decoder")
    if unsupported:
        raise NotImplementedError("checkpoint needs features outside the engine's scope: " + "; ".join(unsupported))
    build = config_9b if "mpt_9b" in flags["llm_name"] else config_3b
    return build(a["max_layer"], ckpt_early_exit_layer=a["early_exit_layer"], exit_interval=a["exit_interval"],
                 lstm_layers=a["lstm_num_layers"], steps_per_stage=steps_per_stage, threshold_type=threshold_type,
                 llm_name=flags["llm_name"], mlp_dims=tuple(MLP_HIDDEN_DIMS[:n_hidden]),
                 mlp_layernorm=bool(a["mlp_layernorm"]), lstm_layernorm=bool(a["lstm_layernorm"]))


MLP_HIDDEN_DIMS = (1024, 512, 256)                 # action_head.py:87,188


def head_num_hidden_layers(a: Mapping[str, Any]) -> int:
    """Hidden layers MLPTanhHead / MLPSigmoidHead actually build: ``mlp_num_hidden_layers`` for dropout_mode 'last' /
    'layerwise', always 3 for the older hard-coded Sequentials ('wo_last'; action_head.py:116-158)."""
    return int(a["mlp_num_hidden_layers"]) if a["dropout_mode"] in ("last", "layerwise") else 3


def head_key_map(a: Mapping[str, Any]) -> Dict[str, str]:
    """``nn.Sequential`` index of every parametrised module of the MLP heads for this checkpoint's head variant -> the
    index the engine uses (the 'layerwise' layout: Linear 4j+1, LayerNorm 4j+2, output Linear 4n+1).  The reference builds a
    different Sequential per (dropout_mode, layernorm, dropout > 0) (action_head.py:82-161, 183-269), so the same tensors
    sit under different keys in checkpoints of different vintages."""
    n, ln = head_num_hidden_layers(a), bool(a["mlp_layernorm"])
    mode = a["dropout_mode"]
    if mode == "layerwise":
        lin = [4 * j + 1 for j in range(n)]; norm = [4 * j + 2 for j in range(n)]; out = 4 * n + 1
    elif mode == "last":                          # [Linear, LN|Identity, ReLU] * n, Dropout after the last hidden layer
        lin = [3 * j for j in range(n)]; norm = [3 * j + 1 for j in range(n)]; out = 3 * n + 1
    elif ln:                                      # 'wo_last' with LayerNorm: Dropout, Linear, LN, ReLU, ... (3 hidden), Linear
        lin = [1, 5, 9]; norm = [2, 6, 10]; out = 12
    elif float(a.get("exit_dropout", 0.0)) > 0:   # Dropout, Linear, ReLU, ... : Linear at 1, 4, 7; output at 9
        lin = [1, 4, 7]; norm = []; out = 9
    else:                                         # the oldest layout: Linear, ReLU, ... : 0, 2, 4; output at 6
        lin = [0, 2, 4]; norm = []; out = 6
    m = {str(lin[j]): str(4 * j + 1) for j in range(n)}
    if ln:
        m.update({str(norm[j]): str(4 * j + 2) for j in range(n)})
    m[str(out)] = str(4 * n + 1)
    return m


_HEAD_KEY = re.compile(r"^(module\.)?((?:extra_exit|lm_head)\.(?:actions|gripper)\.mlp\.)(\d+)(\.(?:weight|bias))$")


def canonical_head_keys(sd: Mapping[str, Any], a: Mapping[str, Any]) -> Dict[str, Any]:
    """Rename the MLP-head tensors of a checkpoint to the engine's canonical 'layerwise' indices (``head_key_map``);
    everything else passes through."""
    km = head_key_map(a)
    out: Dict[str, Any] = {}
    for k, v in sd.items():
        m = _HEAD_KEY.match(k)
        if m and m.group(3) in km:
            k = f"{m.group(1) or ''}{m.group(2)}{km[m.group(3)]}{m.group(4)}"
        out[k] = v
    return out


def state_dict_of(ckpt: Mapping[str, Any]) -> Mapping[str, Any]:
    """``model_state_dict`` if present (DeeR .pth), else the mapping itself (OpenFlamingo .pt / plain state dict)."""
    return ckpt.get("model_state_dict", ckpt)


# ---------------------------------------------------------------------------------------------------
# assembling the model's state dict from its published pieces
# ---------------------------------------------------------------------------------------------------
_XATTN_ALIAS = re.compile(r"^lang_encoder\.gated_cross_attn_layers\.(\d+)\.(.+)$")
_HF_BLOCK = re.compile(r"^transformer\.blocks\.(\d+)\.(.+)$")


def _strip_module(k: str) -> str:
    return k[len("module."):] if k.startswith("module.") else k


def assemble_state_dict(cfg: DeerConfig, hf_lm: Optional[Mapping[str, Any]] = None, clip: Optional[Mapping[str, Any]] = None,
                        open_flamingo: Optional[Mapping[str, Any]] = None, deer: Optional[Mapping[str, Any]] = None,
                        ) -> Tuple[Dict[str, Any], List[str]]:
    """Merge the pieces a DeeR model is built from, in the reference's load order, into ONE dict keyed like
    ``MPTFlamingo.state_dict()`` (what ``DeerEngine.load_state_dict`` / ``deer_load_weight`` consume):

    1. ``clip``: an open_clip checkpoint / ``state_dict()`` -- ``visual.*`` becomes ``vision_encoder.*`` (factory.py:109-114;
       MPTFlamingo keeps only ``vision_encoder.visual``), the text tower is dropped;
    2. ``hf_lm``: the HF MPT causal-LM state dict -- ``transformer.blocks.i.*`` becomes
       ``lang_encoder.transformer.blocks.i.decoder_layer.*`` (FlamingoLayer wrapping, flamingo_lm.py:145-176), blocks
       ``>= cfg.n_layers`` are dropped (flamingo_mpt.py:198-246 deletes them), ``transformer.wte.weight`` is cut to
       ``cfg.vocab`` rows (``resize_token_embeddings(len(tokenizer))``, factory.py:159; growing is refused because HF
       initialises new rows randomly);
    3. ``open_flamingo``: the OpenFlamingo ``checkpoint.pt`` (perceiver, gated x-attn layers, trained embeddings), loaded with
       ``strict=False`` (eval_calvin.py:541-543) -- ``lang_encoder.gated_cross_attn_layers.i.*`` is the same module as
       ``lang_encoder.transformer.blocks.i.gated_cross_attn_layer.*`` and is stored under the latter name;
    4. ``deer``: a DeeR ``.pth`` (or its ``model_state_dict``), ``module.`` stripped, ``strict=False`` (eval_calvin.py:577).

    Later sources override earlier ones.  Returns (state dict restricted to the tensors the engine needs, missing names)."""
    from .weights import state_dict_spec
    out: Dict[str, Any] = {}
    if clip is not None:
        for k, v in state_dict_of(clip).items():
            k = _strip_module(k)
            if k.startswith("visual."):
                out["vision_encoder." + k[len("visual."):]] = v
            elif k.startswith("vision_encoder."):
                out[k] = v
    if hf_lm is not None:
        for k, v in state_dict_of(hf_lm).items():
            k = _strip_module(k)
            m = _HF_BLOCK.match(k)
            if m:
                if int(m.group(1)) < cfg.n_layers:
                    out[f"lang_encoder.transformer.blocks.{m.group(1)}.decoder_layer.{m.group(2)}"] = v
            elif k == "transformer.wte.weight":
                if v.shape[0] < cfg.vocab:
                    raise ValueError(f"embedding has {v.shape[0]} rows < vocab {cfg.vocab}: new rows would be random (factory.py:159)")
                out["lang_encoder.transformer.wte.weight"] = v[: cfg.vocab]
            elif k.startswith("transformer."):
                out["lang_encoder." + k] = v
    for src in (open_flamingo, deer):
        if src is None:
            continue
        for k, v in state_dict_of(src).items():
            k = _strip_module(k)
            m = _XATTN_ALIAS.match(k)
            if m:
                k = f"lang_encoder.transformer.blocks.{m.group(1)}.gated_cross_attn_layer.{m.group(2)}"
            out[k] = v
    need = [n for n, _, _ in state_dict_spec(cfg)]
    missing = [n for n in need if n not in out]
    return {n: out[n] for n in need if n in out}, missing
