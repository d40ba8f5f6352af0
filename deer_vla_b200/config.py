"""Model / exit configuration of the DeeR engine (host side).

Field names follow the reference: ``max_layer`` / ``exit_interval`` / ``early_exit_layer`` are the
flags of robot_flamingo/eval/eval_calvin.py:36-345 and the derived quantities reproduce
eval_calvin.py:529, robot_flamingo/models/flamingo_mpt.py:198,248-250,268-270 and
robot_flamingo/models/value_net.py:173,178.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import List, Tuple


@dataclass
class DeerConfig:
    # vision tower: open_clip ViT-L-14 (robot_flamingo/models/factory.py:109-114)
    image_size: int = 224
    patch: int = 14
    vit_width: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    # perceiver resampler (open_flamingo/src/helpers.py:68-81)
    perc_depth: int = 6
    perc_heads: int = 8
    perc_dim_head: int = 64
    num_latents: int = 64
    perc_ff_mult: int = 4
    # language model
    d_model: int = 2048
    n_heads: int = 16
    mlp_ratio: int = 4
    vocab: int = 50280
    qk_ln: bool = True
    alibi_bias_max: int = 8
    cross_attn_every: int = 1
    xattn_heads: int = 8
    xattn_dim_head: int = 64
    xattn_ff_mult: int = 4
    llm_name: str = "mpt_dolly_3b"
    media_token_id: int = 50278
    eoc_token_id: int = 50277
    # exits
    ckpt_early_exit_layer: int = 11
    exit_interval: int = 2
    max_layer: int = 12
    # action head
    lstm_hidden: int = 1024
    lstm_layers: int = 4
    mlp_dims: Tuple[int, ...] = (1024, 512)
    threshold_type: str = "L2"
    steps_per_stage: int = 1
    # head variants of older checkpoints (eval_calvin.py:464-476): mlp_dims (1024, 512, 256), no LayerNorms
    mlp_layernorm: bool = True
    lstm_layernorm: bool = True

    @property
    def early_exit_layer(self) -> int:          # eval_calvin.py:529
        return min(self.ckpt_early_exit_layer, self.max_layer)

    @property
    def n_layers(self) -> int:                  # flamingo_mpt.py:198
        return self.early_exit_layer + 1

    @property
    def exit_list(self) -> List[int]:           # flamingo_mpt.py:248-250,268-270
        e = list(range(self.exit_interval - 1, self.early_exit_layer, self.exit_interval))
        return e + [self.n_layers - 1]

    @property
    def ctrl_max_layer(self) -> int:            # value_net.py:173
        return min(self.max_layer - 1, self.exit_list[-1])

    @property
    def n_thresholds(self) -> int:              # value_net.py:178
        return len([x for x in self.exit_list if x <= self.ctrl_max_layer])

    def has_xattn(self, layer: int) -> bool:    # flamingo_lm.py:174-176
        return (layer + 1) % self.cross_attn_every == 0

    def to_dict(self):
        return asdict(self)


def config_3b(max_layer: int = 12, **kw) -> DeerConfig:
    """OpenFlamingo-3B (MPT-1B-dolly, x-attn every layer; factory.py:14-19). max_layer 12 = DeeR-B, 4 = DeeR-S (README.md:124)."""
    return DeerConfig(max_layer=max_layer, **kw)


def config_9b(max_layer: int = 12, **kw) -> DeerConfig:
    """OpenFlamingo-9B (MPT-7B, x-attn every 4th layer; factory.py:20-25, README.md:311)."""
    base = dict(d_model=4096, n_heads=32, qk_ln=False, cross_attn_every=4, llm_name="mpt_9b",
                ckpt_early_exit_layer=15, max_layer=max_layer)
    base.update(kw)
    return DeerConfig(**base)


def config_tiny(max_layer: int = 6, **kw) -> DeerConfig:
    """Structure-preserving reduced dims used by the parity tests."""
    base = dict(image_size=56, patch=14, vit_width=128, vit_layers=2, vit_heads=2, vit_mlp=512,
                d_model=256, n_heads=2, vocab=512, media_token_id=510, eoc_token_id=509,
                ckpt_early_exit_layer=5, exit_interval=2, max_layer=max_layer)
    base.update(kw)
    return DeerConfig(**base)
