"""``create_model_and_transforms`` (robot_flamingo/models/factory.py:53-269) for the engine.

Keeps the reference's argument names that matter on the evaluation path and returns the same triple
``(model, image_processor, tokenizer)``.  ``llm_name`` selects the backbone geometry the way ``mpt_dict``
does (factory.py:13-26); weights come from ``state_dict`` / ``checkpoint`` (reference state-dict keys) or,
when absent (no checkpoints exist offline), from the seeded synthetic generator.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .config import DeerConfig, config_3b, config_9b
from .engine import DeerEngine
from .eval_utils import clip_image_transform
from .flamingo_mpt import MPTFlamingo
from .weights import synth_state_dict

# llm_name -> (config builder, cross_attn_every_n_layers)   mirrors mpt_dict (factory.py:13-26)
mpt_dict = {
    "mpt_dolly_3b": {"config": config_3b, "cross_attn_every_n_layers": 1},
    "mpt_3b": {"config": config_3b, "cross_attn_every_n_layers": 1},
    "mpt_9b": {"config": config_9b, "cross_attn_every_n_layers": 4},
}


def create_model_and_transforms(clip_vision_encoder_path: str = "ViT-L-14", clip_vision_encoder_pretrained: str = "openai",
                                lang_encoder_path: Optional[str] = None, tokenizer_path: Optional[str] = None,
                                cross_attn_every_n_layers: Optional[int] = None, window_size: int = 12, use_gripper: bool = True,
                                fusion_mode: str = "post", pad_length: int = -1, llm_name: str = "mpt_dolly_3b", replan: int = -1,
                                refresh: int = -1, multi_step_action: int = 1, early_exit_layer: int = 11, exit_interval: int = 2,
                                max_layer: int = 12, steps_per_stage: int = 1, threshold_type: str = "L2",
                                state_dict=None, checkpoint: Optional[str] = None, seed: int = 0,
                                max_batch: int = 1, max_slots: Optional[int] = None, device: int = 0, precision: str = "bf16",
                                host_control: bool = False, use_graph: bool = False, use_action_ensemble: bool = False,
                                debug_taps: bool = False, dropout_mode: str = "layerwise", mlp_layernorm: bool = True,
                                lstm_layernorm: bool = True, mlp_num_hidden_layers: int = 2, lstm_num_layers: int = 4,
                                exit_dropout: float = 0.0, base_state_dict=None, cfg: Optional[DeerConfig] = None, **unused):
    """``checkpoint``: a DeeR ``.pth`` (train_utils.py:27-60).  Like eval_calvin.py:356-421,455-489,529 the configuration is read
    from the file NAME (llm, flags) and from the hyper-parameters stored beside ``model_state_dict`` (with the reference's
    old-checkpoint defaults); the head keys are renamed to the engine's canonical layout.  The DeeR file holds only the
    trainable tensors, so the frozen ones (CLIP ViT, MPT blocks, embeddings: OpenFlamingo / HF files) come from
    ``base_state_dict`` (model keys; ``checkpoint.assemble_state_dict`` builds it) -- or, offline, from the seeded generator."""
    assert clip_vision_encoder_path == "ViT-L-14", "the engine implements the ViT-L/14 tower (factory.py:109-114)"
    head_args = dict(dropout_mode=dropout_mode, mlp_layernorm=mlp_layernorm, mlp_num_hidden_layers=mlp_num_hidden_layers,
                     exit_dropout=exit_dropout)
    ck = None
    if checkpoint is not None:
        from .checkpoint import config_from_checkpoint, parse_checkpoint_name, readout_args
        ck = torch.load(checkpoint, map_location="cpu", weights_only=False)
        if cfg is None:
            cfg = config_from_checkpoint(ck, checkpoint, max_layer=max_layer, steps_per_stage=steps_per_stage, threshold_type=threshold_type)
        head_args = readout_args(ck, parse_checkpoint_name(os.path.basename(checkpoint))["llm_name"], max_layer)
    if cfg is None:
        from .checkpoint import MLP_HIDDEN_DIMS, head_num_hidden_layers
        entry = mpt_dict[llm_name]
        cfg = entry["config"](max_layer, ckpt_early_exit_layer=early_exit_layer, exit_interval=exit_interval,
                              steps_per_stage=steps_per_stage, threshold_type=threshold_type, llm_name=llm_name,
                              lstm_layers=lstm_num_layers, mlp_dims=tuple(MLP_HIDDEN_DIMS[:head_num_hidden_layers(head_args)]),
                              mlp_layernorm=mlp_layernorm, lstm_layernorm=lstm_layernorm)
    if cross_attn_every_n_layers is not None:
        cfg.cross_attn_every = cross_attn_every_n_layers
    engine = DeerEngine(cfg, max_batch=max_batch, max_slots=max_slots, device=device, precision=precision, host_control=host_control,
                        use_graph=use_graph, action_ensemble=use_action_ensemble, debug_taps=debug_taps)
    if ck is not None:
        from .checkpoint import canonical_head_keys, state_dict_of
        base = dict(base_state_dict) if base_state_dict is not None else synth_state_dict(cfg, seed)
        base.update({(k[len("module."):] if k.startswith("module.") else k): v
                     for k, v in canonical_head_keys(state_dict_of(ck), head_args).items()})     # strict=False (eval_calvin.py:577)
        state_dict = base
    if state_dict is None:
        state_dict = synth_state_dict(cfg, seed)
    engine.load_state_dict(state_dict)
    model = MPTFlamingo(engine, cfg, window_size=window_size, use_gripper=use_gripper, fusion_mode=fusion_mode,
                        pad_length=pad_length, replan=replan, refresh=refresh, multi_step_action=multi_step_action)
    tokenizer = None
    if tokenizer_path is not None and os.path.exists(tokenizer_path):
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(tokenizer_path, local_files_only=True)
        tokenizer.add_special_tokens({"additional_special_tokens": ["<|endofchunk|>", "<image>"]})    # factory.py:120-126
        if tokenizer.pad_token is None:
            tokenizer.add_special_tokens({"pad_token": "<PAD>"})
    return model, clip_image_transform(cfg.image_size), tokenizer
