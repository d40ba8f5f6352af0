"""Checkpoint schema of the hot path and a seeded synthetic state-dict generator.

``state_dict_spec`` lists every tensor the engine consumes under the REFERENCE's state-dict keys
(what ``MPTFlamingo.state_dict()`` / the DeeR ``model_state_dict`` hold: train_utils.py:27-60,
631-638; eval_calvin.py:541-577), so a real checkpoint can be passed to
``DeerEngine.load_state_dict`` unchanged.  No real checkpoints exist in the build environment, so
``synth_state_dict`` draws seeded synthetic values (one generator per tensor, seeded by
(seed, crc32(name)), so the tensors of a truncated model equal those of the full one): Linear/Embedding N(0,0.02), LayerNorm gamma 1+0.1N / beta 0.1N, LSTM
U(+-1/sqrt(H)), Perceiver latents N(0,1), non-zero tanh gates (helpers.py:255,258 initialise the
gates to 0, which would null the cross-attention path).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Tuple

import torch

from .config import DeerConfig

Spec = Tuple[str, Tuple[int, ...], str]


def state_dict_spec(cfg: DeerConfig) -> List[Spec]:
    s: List[Spec] = []
    W = cfg.vit_width
    L = (cfg.image_size // cfg.patch) ** 2 + 1
    v = "vision_encoder.visual."
    s += [(v + "conv1.weight", (W, 3, cfg.patch, cfg.patch), "w"), (v + "class_embedding", (W,), "w"),
          (v + "positional_embedding", (L, W), "w"), (v + "ln_pre.weight", (W,), "g"), (v + "ln_pre.bias", (W,), "be")]
    for i in range(cfg.vit_layers):
        p = f"{v}transformer.resblocks.{i}."
        s += [(p + "ln_1.weight", (W,), "g"), (p + "ln_1.bias", (W,), "be"),
              (p + "attn.in_proj_weight", (3 * W, W), "w"), (p + "attn.in_proj_bias", (3 * W,), "b"),
              (p + "attn.out_proj.weight", (W, W), "w"), (p + "attn.out_proj.bias", (W,), "b"),
              (p + "ln_2.weight", (W,), "g"), (p + "ln_2.bias", (W,), "be"),
              (p + "mlp.c_fc.weight", (cfg.vit_mlp, W), "w"), (p + "mlp.c_fc.bias", (cfg.vit_mlp,), "b"),
              (p + "mlp.c_proj.weight", (W, cfg.vit_mlp), "w"), (p + "mlp.c_proj.bias", (W,), "b")]
    inner = cfg.perc_heads * cfg.perc_dim_head
    s += [("perceiver.latents", (cfg.num_latents, W), "n1")]
    for i in range(cfg.perc_depth):
        p = f"perceiver.layers.{i}."
        s += [(p + "0.norm_media.weight", (W,), "g"), (p + "0.norm_media.bias", (W,), "be"),
              (p + "0.norm_latents.weight", (W,), "g"), (p + "0.norm_latents.bias", (W,), "be"),
              (p + "0.to_q.weight", (inner, W), "w"), (p + "0.to_kv.weight", (2 * inner, W), "w"),
              (p + "0.to_out.weight", (W, inner), "w"), (p + "1.0.weight", (W,), "g"), (p + "1.0.bias", (W,), "be"),
              (p + "1.1.weight", (cfg.perc_ff_mult * W, W), "w"), (p + "1.3.weight", (W, cfg.perc_ff_mult * W), "w")]
    s += [("perceiver.norm.weight", (W,), "g"), ("perceiver.norm.bias", (W,), "be")]
    D = cfg.d_model
    s += [("lang_encoder.transformer.wte.weight", (cfg.vocab, D), "w")]
    xin = cfg.xattn_heads * cfg.xattn_dim_head
    for i in range(cfg.n_layers):
        if cfg.has_xattn(i):
            p = f"lang_encoder.transformer.blocks.{i}.gated_cross_attn_layer."
            s += [(p + "attn.norm.weight", (D,), "g"), (p + "attn.norm.bias", (D,), "be"),
                  (p + "attn.to_q.weight", (xin, D), "w"), (p + "attn.to_kv.weight", (2 * xin, W), "w"),
                  (p + "attn.to_out.weight", (D, xin), "w"), (p + "attn_gate", (1,), "gate"),
                  (p + "ff.0.weight", (D,), "g"), (p + "ff.0.bias", (D,), "be"),
                  (p + "ff.1.weight", (cfg.xattn_ff_mult * D, D), "w"),
                  (p + "ff.3.weight", (D, cfg.xattn_ff_mult * D), "w"), (p + "ff_gate", (1,), "gate")]
        p = f"lang_encoder.transformer.blocks.{i}.decoder_layer."
        if cfg.qk_ln:
            s += [(p + "ln_1.weight", (D,), "g"), (p + "attn.Wqkv.weight", (3 * D, D), "w"),
                  (p + "attn.q_ln.weight", (D,), "g"), (p + "attn.k_ln.weight", (D,), "g"),
                  (p + "attn.out_proj.weight", (D, D), "w"), (p + "ln_2.weight", (D,), "g"),
                  (p + "mlp.mlp_up.weight", (cfg.mlp_ratio * D, D), "w"), (p + "mlp.mlp_down.weight", (D, cfg.mlp_ratio * D), "w")]
        else:
            s += [(p + "norm_1.weight", (D,), "g"), (p + "attn.Wqkv.weight", (3 * D, D), "w"),
                  (p + "attn.out_proj.weight", (D, D), "w"), (p + "norm_2.weight", (D,), "g"),
                  (p + "ffn.up_proj.weight", (cfg.mlp_ratio * D, D), "w"), (p + "ffn.down_proj.weight", (D, cfg.mlp_ratio * D), "w")]
    H = cfg.lstm_hidden
    for l in range(cfg.lstm_layers):
        ins = D if l == 0 else H
        if cfg.lstm_layernorm:                     # LayerNormLSTM (action_head.py:15-40)
            p = f"extra_exit.rnn.layers.{3 * l}."
            s += [(p + "weight_ih_l0", (4 * H, ins), "lstm"), (p + "weight_hh_l0", (4 * H, H), "lstm"),
                  (p + "bias_ih_l0", (4 * H,), "lstm"), (p + "bias_hh_l0", (4 * H,), "lstm"),
                  (f"extra_exit.rnn.layers.{3 * l + 1}.weight", (H,), "g"), (f"extra_exit.rnn.layers.{3 * l + 1}.bias", (H,), "be")]
        else:                                      # stacked nn.LSTM (action_head.py:71-79)
            p = "extra_exit.rnn."
            s += [(p + f"weight_ih_l{l}", (4 * H, ins), "lstm"), (p + f"weight_hh_l{l}", (4 * H, H), "lstm"),
                  (p + f"bias_ih_l{l}", (4 * H,), "lstm"), (p + f"bias_hh_l{l}", (4 * H,), "lstm")]
    for head, out in (("actions", 6), ("gripper", 1)):
        cur = H
        for j, dim in enumerate(cfg.mlp_dims):     # canonical 'layerwise' Sequential indices (checkpoint.canonical_head_keys)
            p = f"extra_exit.{head}.mlp."
            s += [(p + f"{4 * j + 1}.weight", (dim, cur), "w"), (p + f"{4 * j + 1}.bias", (dim,), "b")]
            if cfg.mlp_layernorm:
                s += [(p + f"{4 * j + 2}.weight", (dim,), "g"), (p + f"{4 * j + 2}.bias", (dim,), "be")]
            cur = dim
        k = 4 * len(cfg.mlp_dims) + 1
        s += [(f"extra_exit.{head}.mlp.{k}.weight", (out, cur), "wout"), (f"extra_exit.{head}.mlp.{k}.bias", (out,), "b")]
    return s


_GATES = [0.5, -0.4, 0.7, 0.35, -0.6, 0.45]


def synth_state_dict(cfg: DeerConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    return dict(synth_state_items(cfg, seed))


def synth_state_items(cfg: DeerConfig, seed: int = 0):
    """Generator form of ``synth_state_dict``: one (name, tensor) at a time, so loading a replica never holds the whole
    fp32 state dict (6 GB for the 3B model) in host memory -- N ranks on one box load concurrently."""
    gi = 0
    for idx, (name, shape, kind) in enumerate(state_dict_spec(cfg)):
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
        if kind == "w":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "wout":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "b":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "g":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "be":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "n1":
            t = torch.randn(shape, generator=g)
        elif kind == "lstm":
            t = (torch.rand(shape, generator=g) * 2 - 1) * (1.0 / math.sqrt(cfg.lstm_hidden))
        elif kind == "gate":
            t = torch.tensor([_GATES[gi % len(_GATES)]])
            gi += 1
        else:
            raise ValueError(kind)
        yield name, t.float().contiguous()


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def synth_frames(cfg: DeerConfig, n: int, seed: int, device="cuda") -> torch.Tensor:
    """(n,3,H,W) f32: uint8 U[0,255] frames, CLIP-normalised like open_clip's image transform (data.py:898-902)."""
    g = torch.Generator(device=device).manual_seed(seed)
    u8 = torch.randint(0, 256, (n, 3, cfg.image_size, cfg.image_size), generator=g, device=device, dtype=torch.uint8)
    mean = torch.tensor(CLIP_MEAN, device=device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, device=device).view(1, 3, 1, 1)
    return ((u8.float() / 255.0 - mean) / std).contiguous()


def synth_instruction_ids(cfg: DeerConfig, n: int, S: int = 32, seed: int = 2000) -> torch.Tensor:
    """(n,S) int64: [<image>] + S-3 random ids + [<|endofchunk|>, eos=0]  (preprocess_text_calvin, data.py:905-919)."""
    g = torch.Generator().manual_seed(seed)
    hi = min(cfg.vocab, cfg.eoc_token_id) - 1
    body = torch.randint(1, hi, (n, S - 3), generator=g, dtype=torch.int64)
    return torch.cat([torch.full((n, 1), cfg.media_token_id, dtype=torch.int64), body,
                      torch.full((n, 1), cfg.eoc_token_id, dtype=torch.int64), torch.zeros(n, 1, dtype=torch.int64)], dim=1)
