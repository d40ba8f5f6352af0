"""CPU fp32 ORACLE for the DeeR-VLA per-timestep multi-exit forward.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``deer_vla_b200``) never imports anything from
``oracle/`` and fails loudly when its CUDA library is missing.

It is a plain-PyTorch (CPU, fp32, batch-1-per-episode) restatement of the reference
algorithm; every function cites the reference ``file:line`` it follows (paths relative
to the reference checkout).

Pinning status (see ``oracle/make_golden.py`` and ``DESIGN.md``):
  * vendored parts (Perceiver, gated x-attn, FlamingoLayer order, MosaicGPT exit loop,
    action head, ActionValueNet, ExitController, threshold solver) are PINNED: the
    restatement is checked against the reference's own modules executed in the build
    container, and golden vectors from those runs are committed under ``tests/golden``.
  * un-vendored third-party arithmetic (open_clip 2.20.0 ViT-L/14 visual tower; HF-hub
    ``mosaicml/mpt-1b-redpajama-200b-dolly`` GPTBlock / attention; ``mosaicml/mpt-7b``
    MPTBlock) is restated from the published algorithm and cross-checked against
    ``transformers.models.clip`` / ``transformers.models.mpt``; the reference holds no
    golden vectors for them, so for those two boundaries: **parity unpinned**.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# Rounding yardstick: with OPERAND_DTYPE = torch.bfloat16 both operands of every contraction
# (linear, conv patch-embed, QK^T, PV) are rounded to bf16 and accumulated in fp32 -- what ANY bf16 tensor-core forward
# does -- while everything else stays fp32.  |this - fp32| is the size of "bf16 tolerance" for a given model and input;
# the tests require the engine's bf16 path to sit inside a small multiple of it.  None = the plain fp32 reference.
OPERAND_DTYPE = None


class operand_rounding:
    """``with operand_rounding(torch.bfloat16): ...`` -- context manager for OPERAND_DTYPE."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global OPERAND_DTYPE
        self.prev, OPERAND_DTYPE = OPERAND_DTYPE, self.dtype

    def __exit__(self, *exc):
        global OPERAND_DTYPE
        OPERAND_DTYPE = self.prev


def _r(x: Tensor) -> Tensor:
    return x if OPERAND_DTYPE is None else x.to(OPERAND_DTYPE).to(torch.float32)


def _linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return F.linear(_r(x), _r(w), b)


def _mm(a: Tensor, b: Tensor) -> Tensor:
    return torch.matmul(_r(a), _r(b))


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class DeerConfig:
    # vision tower: open_clip "ViT-L-14", pretrained="openai" (robot_flamingo/models/factory.py:109-114)
    image_size: int = 224
    patch: int = 14
    vit_width: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    # perceiver resampler defaults (open_flamingo/src/helpers.py:68-81)
    perc_depth: int = 6
    perc_heads: int = 8
    perc_dim_head: int = 64
    num_latents: int = 64
    perc_ff_mult: int = 4
    # language model: MPT-1B (3B flamingo) / MPT-7B (9B flamingo)
    d_model: int = 2048
    n_heads: int = 16
    mlp_ratio: int = 4
    vocab: int = 50280
    qk_ln: bool = True                # MPT-1B attn_qk_ln=true; MPT-7B false
    alibi_bias_max: int = 8
    cross_attn_every: int = 1         # factory.py:14-25 (3B: 1, 9B: 4)
    xattn_heads: int = 8              # helpers.py:242-243
    xattn_dim_head: int = 64
    xattn_ff_mult: int = 4
    llm_name: str = "mpt_dolly_3b"
    media_token_id: int = 50278
    eoc_token_id: int = 50277
    # exits (flamingo_mpt.py:191-259, eval_calvin.py:481-489,529)
    ckpt_early_exit_layer: int = 11
    exit_interval: int = 2
    max_layer: int = 12
    # action head (action_head.py:408-491; trained config SURVEY §8a-14)
    lstm_hidden: int = 1024
    lstm_layers: int = 4
    mlp_dims: Tuple[int, ...] = (1024, 512)
    threshold_type: str = "L2"
    steps_per_stage: int = 1
    # head variants of older checkpoints (eval_calvin.py:464-476): mlp_dims (1024, 512, 256), no LayerNorms
    mlp_layernorm: bool = True
    lstm_layernorm: bool = True

    # ---- derived quantities -------------------------------------------------------
    @property
    def early_exit_layer(self) -> int:
        # eval_calvin.py:529  early_exit_layer=min(args.early_exit_layer, args.max_layer)
        return min(self.ckpt_early_exit_layer, self.max_layer)

    @property
    def n_layers(self) -> int:
        # flamingo_mpt.py:198  _delete_decoder_layers(range(early_exit_layer+1, n_layers))
        return self.early_exit_layer + 1

    @property
    def exit_list(self) -> List[int]:
        # flamingo_mpt.py:248-250,268-270
        e = list(range(self.exit_interval - 1, self.early_exit_layer, self.exit_interval))
        return e + [self.n_layers - 1]

    @property
    def ctrl_max_layer(self) -> int:
        # value_net.py:173
        return min(self.max_layer - 1, self.exit_list[-1])

    @property
    def n_thresholds(self) -> int:
        # value_net.py:178
        return len([x for x in self.exit_list if x <= self.ctrl_max_layer])

    @property
    def n_tokens_vit(self) -> int:
        return (self.image_size // self.patch) ** 2 + 1

    def has_xattn(self, layer: int) -> bool:
        # flamingo_lm.py:174-176
        return (layer + 1) % self.cross_attn_every == 0

    def to_dict(self):
        return asdict(self)


def config_3b(max_layer: int = 12, **kw) -> DeerConfig:
    return DeerConfig(max_layer=max_layer, **kw)


def config_9b(max_layer: int = 12, **kw) -> DeerConfig:
    # factory.py:20-25; README.md:311 (early_exit_layer=15); modeling_gpt_9b.py
    base = dict(d_model=4096, n_heads=32, qk_ln=False, cross_attn_every=4, vocab=50280,
                llm_name="mpt_9b", ckpt_early_exit_layer=15, max_layer=max_layer)
    base.update(kw)
    return DeerConfig(**base)


def config_tiny(max_layer: int = 6, **kw) -> DeerConfig:
    """Reduced dims (same structure) so that full intermediate tensors fit in small fixtures."""
    base = dict(image_size=56, patch=14, vit_width=128, vit_layers=2, vit_heads=2, vit_mlp=512,
                d_model=256, n_heads=2, vocab=512,
                media_token_id=510, eoc_token_id=509, ckpt_early_exit_layer=5, exit_interval=2,
                max_layer=max_layer)
    base.update(kw)
    return DeerConfig(**base)


# --------------------------------------------------------------------------------------
# deterministic synthetic weights (keys == reference state_dict keys of MPTFlamingo)
# --------------------------------------------------------------------------------------
def weight_specs(cfg: DeerConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered (name, shape, init) list.  Names are the reference's state-dict keys.

    init kinds: 'w' N(0,0.02) linear/embedding; 'b' 0.02*N bias; 'g' LayerNorm gamma
    1+0.1*N; 'be' LayerNorm beta 0.1*N; 'n1' N(0,1); 'lstm' U(+-1/sqrt(H)); 'gate' fixed
    non-zero x-attn gates (helpers.py:255,258 init them to 0 which would null the path).
    """
    s: List[Tuple[str, Tuple[int, ...], str]] = []
    W, L = cfg.vit_width, cfg.n_tokens_vit
    v = "vision_encoder.visual."
    s += [(v + "conv1.weight", (W, 3, cfg.patch, cfg.patch), "w"),
          (v + "class_embedding", (W,), "w"),
          (v + "positional_embedding", (L, W), "w"),
          (v + "ln_pre.weight", (W,), "g"), (v + "ln_pre.bias", (W,), "be")]
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
              (p + "0.to_out.weight", (W, inner), "w"),
              (p + "1.0.weight", (W,), "g"), (p + "1.0.bias", (W,), "be"),
              (p + "1.1.weight", (cfg.perc_ff_mult * W, W), "w"),
              (p + "1.3.weight", (W, cfg.perc_ff_mult * W), "w")]
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
        if cfg.qk_ln:   # MPT-1B GPTBlock naming
            s += [(p + "ln_1.weight", (D,), "g"), (p + "attn.Wqkv.weight", (3 * D, D), "w"),
                  (p + "attn.q_ln.weight", (D,), "g"), (p + "attn.k_ln.weight", (D,), "g"),
                  (p + "attn.out_proj.weight", (D, D), "w"), (p + "ln_2.weight", (D,), "g"),
                  (p + "mlp.mlp_up.weight", (cfg.mlp_ratio * D, D), "w"),
                  (p + "mlp.mlp_down.weight", (D, cfg.mlp_ratio * D), "w")]
        else:           # MPT-7B MPTBlock naming
            s += [(p + "norm_1.weight", (D,), "g"), (p + "attn.Wqkv.weight", (3 * D, D), "w"),
                  (p + "attn.out_proj.weight", (D, D), "w"), (p + "norm_2.weight", (D,), "g"),
                  (p + "ffn.up_proj.weight", (cfg.mlp_ratio * D, D), "w"),
                  (p + "ffn.down_proj.weight", (D, cfg.mlp_ratio * D), "w")]
    H = cfg.lstm_hidden
    for l in range(cfg.lstm_layers):
        ins = D if l == 0 else H
        if cfg.lstm_layernorm:
            p = f"extra_exit.rnn.layers.{3 * l}."      # action_head.py:23-39: [LSTM, LN, Dropout] * n
            s += [(p + "weight_ih_l0", (4 * H, ins), "lstm"), (p + "weight_hh_l0", (4 * H, H), "lstm"),
                  (p + "bias_ih_l0", (4 * H,), "lstm"), (p + "bias_hh_l0", (4 * H,), "lstm"),
                  (f"extra_exit.rnn.layers.{3 * l + 1}.weight", (H,), "g"),
                  (f"extra_exit.rnn.layers.{3 * l + 1}.bias", (H,), "be")]
        else:                                          # action_head.py:71-79: one stacked nn.LSTM
            p = "extra_exit.rnn."
            s += [(p + f"weight_ih_l{l}", (4 * H, ins), "lstm"), (p + f"weight_hh_l{l}", (4 * H, H), "lstm"),
                  (p + f"bias_ih_l{l}", (4 * H,), "lstm"), (p + f"bias_hh_l{l}", (4 * H,), "lstm")]
    for head, out in (("actions", 6), ("gripper", 1)):
        cur = H
        for j, dim in enumerate(cfg.mlp_dims):     # action_head.py:96-108 ('layerwise' Sequential indices 1,2 / 5,6 / 9,10)
            p = f"extra_exit.{head}.mlp."
            s += [(p + f"{4 * j + 1}.weight", (dim, cur), "w"), (p + f"{4 * j + 1}.bias", (dim,), "b")]
            if cfg.mlp_layernorm:
                s += [(p + f"{4 * j + 2}.weight", (dim,), "g"), (p + f"{4 * j + 2}.bias", (dim,), "be")]
            cur = dim
        k = 4 * len(cfg.mlp_dims) + 1
        s += [(f"extra_exit.{head}.mlp.{k}.weight", (out, cur), "wout"),
              (f"extra_exit.{head}.mlp.{k}.bias", (out,), "b")]
    return s


def synth_weights(cfg: DeerConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded synthetic weights, generated in ``weight_specs`` order (there are no real
    checkpoints in the build environment).  Each tensor uses its own generator seeded by
    (seed, crc32(name)): a model truncated by max_layer gets the same tensors as the full one."""
    out: Dict[str, Tensor] = {}
    gate_vals = [0.5, -0.4, 0.7, 0.35, -0.6, 0.45]
    gi = 0
    for idx, (name, shape, kind) in enumerate(weight_specs(cfg)):
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
            k = 1.0 / math.sqrt(cfg.lstm_hidden)
            t = (torch.rand(shape, generator=g) * 2 - 1) * k
        elif kind == "gate":
            t = torch.tensor([gate_vals[gi % len(gate_vals)]])
            gi += 1
        else:
            raise ValueError(kind)
        out[name] = t.float().contiguous()
    return out


def synth_weights_variant(cfg: DeerConfig, seed: int = 0) -> Dict[str, Tensor]:
    """``synth_weights`` for the head variants WITHOUT LayerNorms (old checkpoints): with N(0, 0.02) / U(+-1/32) weights and
    no normalisation the un-trained head is nearly constant (deltas ~1e-6, every exit decision would sit inside fp32 noise),
    so the LSTM weights are scaled x6 and the MLP weights x3 where the corresponding LayerNorm is absent (deltas ~0.1)."""
    W = synth_weights(cfg, seed)
    for k in W:
        if not cfg.lstm_layernorm and k.startswith("extra_exit.rnn.weight"):
            W[k] = W[k] * 6.0
        if not cfg.mlp_layernorm and ".mlp." in k and k.endswith("weight"):
            W[k] = W[k] * 3.0
    return W


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def synth_image(cfg: DeerConfig, episode: int, t: int, cam: int) -> Tensor:
    """uint8 U[0,255] frame, CLIP-normalised -> f32 (3,H,W)."""
    g = torch.Generator().manual_seed(1000 + episode * 400 + t * 2 + cam)
    u8 = torch.randint(0, 256, (3, cfg.image_size, cfg.image_size), generator=g, dtype=torch.uint8)
    x = u8.float() / 255.0
    mean = torch.tensor(CLIP_MEAN).view(3, 1, 1)
    std = torch.tensor(CLIP_STD).view(3, 1, 1)
    return ((x - mean) / std).contiguous()


def synth_tokens(cfg: DeerConfig, episode: int, S: int = 32) -> Tensor:
    """[<image>] + (S-3) random ids + [<|endofchunk|>, eos=0]  (data.py:905-919)."""
    g = torch.Generator().manual_seed(2000 + episode)
    hi = min(cfg.vocab, cfg.eoc_token_id) - 1
    body = torch.randint(1, hi, (S - 3,), generator=g, dtype=torch.int64)
    return torch.cat([torch.tensor([cfg.media_token_id]), body,
                      torch.tensor([cfg.eoc_token_id, 0])]).to(torch.int64)


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def _ln(x: Tensor, w: Tensor, b: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def vit_forward(cfg: DeerConfig, W: Dict[str, Tensor], images: Tensor) -> Tensor:
    """open_clip 2.20.0 ``VisionTransformer.forward`` with ``output_tokens=True``
    [UNVENDORED; call sites robot_flamingo/models/factory.py:109-114, flamingo_mpt.py:580].
    Returns the patch tokens ``x[:, 1:]`` of the last residual block: no ``ln_post``, no
    ``proj``, CLS dropped.  conv1 has no bias; blocks are pre-LN with
    nn.MultiheadAttention and a QuickGELU MLP (the 'openai' pretrained tag forces
    quick_gelu); LayerNorm eps 1e-5.

    images: (N,3,H,W) f32 -> (N, n_patches, width)
    """
    p = "vision_encoder.visual."
    N = images.shape[0]
    x = F.conv2d(_r(images), _r(W[p + "conv1.weight"]), None, stride=cfg.patch)           # (N,width,g,g)
    x = x.reshape(N, cfg.vit_width, -1).permute(0, 2, 1)                          # (N,g*g,width)
    cls = W[p + "class_embedding"].view(1, 1, -1).expand(N, 1, -1)
    x = torch.cat([cls, x], dim=1) + W[p + "positional_embedding"]
    x = _ln(x, W[p + "ln_pre.weight"], W[p + "ln_pre.bias"])
    H, hd = cfg.vit_heads, cfg.vit_width // cfg.vit_heads
    for i in range(cfg.vit_layers):
        q_ = f"{p}transformer.resblocks.{i}."
        a = _ln(x, W[q_ + "ln_1.weight"], W[q_ + "ln_1.bias"])
        qkv = _linear(a, W[q_ + "attn.in_proj_weight"], W[q_ + "attn.in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)
        q = q.view(N, -1, H, hd).transpose(1, 2) * (hd ** -0.5)    # nn.MultiheadAttention scales q
        k = k.view(N, -1, H, hd).transpose(1, 2)
        v = v.view(N, -1, H, hd).transpose(1, 2)
        att = torch.softmax(_mm(q, k.transpose(-1, -2)), dim=-1)
        o = _mm(att, v).transpose(1, 2).reshape(N, -1, cfg.vit_width)
        x = x + _linear(o, W[q_ + "attn.out_proj.weight"], W[q_ + "attn.out_proj.bias"])
        m = _ln(x, W[q_ + "ln_2.weight"], W[q_ + "ln_2.bias"])
        h = _linear(m, W[q_ + "mlp.c_fc.weight"], W[q_ + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)                                             # QuickGELU
        x = x + _linear(h, W[q_ + "mlp.c_proj.weight"], W[q_ + "mlp.c_proj.bias"])
    return x[:, 1:]


def perceiver_forward(cfg: DeerConfig, W: Dict[str, Tensor], feats: Tensor, prefix: str = "perceiver.") -> Tensor:
    """``PerceiverResampler.forward`` (open_flamingo/src/helpers.py:107-132) with
    ``PerceiverAttention.forward`` (39-65) and ``FeedForward`` (15-22); no frame / media-time
    embeddings (constructed with defaults, flamingo_mpt.py:102).

    feats: (N, v, D) patch tokens of one image each (T=F=1) -> (N, num_latents, D)
    """
    N = feats.shape[0]
    h, dh = cfg.perc_heads, cfg.perc_dim_head
    lat = W[prefix + "latents"].unsqueeze(0).expand(N, -1, -1)
    for i in range(cfg.perc_depth):
        p = f"{prefix}layers.{i}."
        xm = _ln(feats, W[p + "0.norm_media.weight"], W[p + "0.norm_media.bias"])
        ln_lat = _ln(lat, W[p + "0.norm_latents.weight"], W[p + "0.norm_latents.bias"])
        q = _linear(ln_lat, W[p + "0.to_q.weight"])
        kv_in = torch.cat([xm, ln_lat], dim=-2)                                     # helpers.py:53
        k, v = _linear(kv_in, W[p + "0.to_kv.weight"]).chunk(2, dim=-1)
        q = q.view(N, -1, h, dh).transpose(1, 2) * (dh ** -0.5)
        k = k.view(N, -1, h, dh).transpose(1, 2)
        v = v.view(N, -1, h, dh).transpose(1, 2)
        sim = _mm(q, k.transpose(-1, -2))
        sim = sim - sim.amax(dim=-1, keepdim=True)                                   # helpers.py:60
        att = sim.softmax(dim=-1)
        o = _mm(att, v).transpose(1, 2).reshape(N, -1, h * dh)
        lat = _linear(o, W[p + "0.to_out.weight"]) + lat                           # helpers.py:130
        f = _ln(lat, W[p + "1.0.weight"], W[p + "1.0.bias"])
        f = _linear(F.gelu(_linear(f, W[p + "1.1.weight"])), W[p + "1.3.weight"])
        lat = f + lat                                                                # helpers.py:131
    return _ln(lat, W[prefix + "norm.weight"], W[prefix + "norm.bias"])


def encode_vision(cfg: DeerConfig, W: Dict[str, Tensor], rgb: Tensor, gripper: Tensor) -> Tuple[Tensor, Dict[str, Tensor]]:
    """``_encode_multi_vision_post_fusion`` (robot_flamingo/models/flamingo_mpt.py:609-668):
    ViT on rgb and gripper, the shared perceiver on each, concat along the latent dim.
    rgb/gripper: (B,3,H,W) -> media (B, 2*num_latents, D)."""
    tok_rgb = vit_forward(cfg, W, rgb)
    tok_grip = vit_forward(cfg, W, gripper)
    lat_rgb = perceiver_forward(cfg, W, tok_rgb)
    lat_grip = perceiver_forward(cfg, W, tok_grip)
    media = torch.cat([lat_rgb, lat_grip], dim=1)                                    # flamingo_mpt.py:661
    return media, {"vit_rgb": tok_rgb, "vit_gripper": tok_grip, "media": media}


def gated_xattn_forward(cfg: DeerConfig, W: Dict[str, Tensor], layer: int, x: Tensor, media: Tensor,
                        media_locations: Tensor) -> Tensor:
    """``GatedCrossAttentionBlock.forward`` (helpers.py:260-279) +
    ``MaskedCrossAttention.forward`` (160-233), ``only_attend_immediate_media=True``,
    ``use_cached_media=False``, T_img = 1.

    x (B,S,d), media (B,n,Dv), media_locations (B,S) bool."""
    p = f"lang_encoder.transformer.blocks.{layer}.gated_cross_attn_layer."
    B, S, _ = x.shape
    h, dh = cfg.xattn_heads, cfg.xattn_dim_head
    xn = _ln(x, W[p + "attn.norm.weight"], W[p + "attn.norm.bias"])
    q = _linear(xn, W[p + "attn.to_q.weight"])
    k, v = _linear(media, W[p + "attn.to_kv.weight"]).chunk(2, dim=-1)
    q = q.view(B, S, h, dh).transpose(1, 2) * (dh ** -0.5)
    k = k.view(B, -1, h, dh).transpose(1, 2)
    v = v.view(B, -1, h, dh).transpose(1, 2)
    sim = _mm(q, k.transpose(-1, -2))                                                    # (B,h,S,n)
    text_time = media_locations.cumsum(dim=-1)                                       # helpers.py:208
    media_time = torch.ones(media.shape[1], dtype=text_time.dtype)                   # T_img=1 -> all latents time 1
    mask = text_time.view(B, 1, S, 1) == media_time.view(1, 1, 1, -1)                # helpers.py:214-218
    sim = sim.masked_fill(~mask, -torch.finfo(sim.dtype).max)
    sim = sim - sim.amax(dim=-1, keepdim=True)
    att = sim.softmax(dim=-1)
    att = att.masked_fill((text_time == 0).view(B, 1, S, 1), 0.0)                    # helpers.py:223-229
    o = _mm(att, v).transpose(1, 2).reshape(B, S, h * dh)
    x = _linear(o, W[p + "attn.to_out.weight"]) * torch.tanh(W[p + "attn_gate"]) + x
    f = _ln(x, W[p + "ff.0.weight"], W[p + "ff.0.bias"])
    f = _linear(F.gelu(_linear(f, W[p + "ff.1.weight"])), W[p + "ff.3.weight"])
    return f * torch.tanh(W[p + "ff_gate"]) + x


def alibi_slopes(n_heads: int, alibi_bias_max: int = 8) -> Tensor:
    """MPT ``attention.alibi_bias`` slopes [UNVENDORED]: m = (1..H) * bias_max / H,
    slope = 2^-m (H a power of two for both MPT-1B (16) and MPT-7B (32); for other H the
    7B ``gen_slopes`` interleaving applies).  Cross-checked against
    ``transformers.models.mpt.modeling_mpt.build_mpt_alibi_tensor``."""
    hp = 2 ** math.ceil(math.log2(n_heads))
    m = torch.arange(1, hp + 1, dtype=torch.float32) * (alibi_bias_max / hp)
    slopes = 1.0 / torch.pow(2.0, m)
    if hp != n_heads:
        slopes = torch.cat([slopes[1::2], slopes[::2]])[:n_heads]
    return slopes


def mpt_attn_bias(cfg: DeerConfig, attention_mask: Tensor, dtype=torch.float32) -> Tensor:
    """``MosaicGPT._attn_bias`` (mosaic_gpt_3b.py:158-219): key-position-only ALiBi bias
    (1,H,1,S) = slope_h * (j - (S-1)); padded keys get finfo.min (215-217)."""
    B, S = attention_mask.shape
    pos = torch.arange(1 - S, 1, dtype=dtype).view(1, 1, 1, S)
    bias = pos * alibi_slopes(cfg.n_heads, cfg.alibi_bias_max).to(dtype).view(1, -1, 1, 1)
    return bias.expand(B, -1, -1, -1).masked_fill(~attention_mask.bool().view(B, 1, 1, S), torch.finfo(dtype).min)


def mpt_block_forward(cfg: DeerConfig, W: Dict[str, Tensor], layer: int, x: Tensor, attn_bias: Tensor) -> Tensor:
    """MPT-1B ``GPTBlock.forward`` / MPT-7B ``MPTBlock.forward`` [UNVENDORED; call sites
    mosaic_gpt_3b.py:413-417, modeling_gpt_9b.py:465-468]: pre-LN, bias-free (no_bias),
    ``attn_impl='torch'`` scaled_multihead_dot_product_attention: scores*1/sqrt(hd) +
    attn_bias, causal masked_fill(finfo.min), softmax, exact-erf GELU MLP.  MPT-1B applies
    LayerNorm over the full d_model to q and k before the head split (attn_qk_ln)."""
    p = f"lang_encoder.transformer.blocks.{layer}.decoder_layer."
    B, S, D = x.shape
    H, hd = cfg.n_heads, D // cfg.n_heads
    n1, n2 = ("ln_1", "ln_2") if cfg.qk_ln else ("norm_1", "norm_2")
    up, down = ("mlp.mlp_up", "mlp.mlp_down") if cfg.qk_ln else ("ffn.up_proj", "ffn.down_proj")
    a = _ln(x, W[p + n1 + ".weight"], None)
    q, k, v = _linear(a, W[p + "attn.Wqkv.weight"]).chunk(3, dim=2)
    if cfg.qk_ln:
        q = _ln(q, W[p + "attn.q_ln.weight"], None)
        k = _ln(k, W[p + "attn.k_ln.weight"], None)
    q = q.view(B, S, H, hd).transpose(1, 2)
    k = k.view(B, S, H, hd).transpose(1, 2)
    v = v.view(B, S, H, hd).transpose(1, 2)
    min_val = torch.finfo(x.dtype).min
    w = _mm(q, k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    w = w + attn_bias
    causal = ~torch.ones(S, S, dtype=torch.bool).tril()
    w = w.masked_fill(causal.view(1, 1, S, S), min_val)
    w = torch.softmax(w, dim=-1)
    o = _mm(w, v).transpose(1, 2).reshape(B, S, D)
    x = x + _linear(o, W[p + "attn.out_proj.weight"])
    m = _ln(x, W[p + n2 + ".weight"], None)
    n = _linear(F.gelu(_linear(m, W[p + up + ".weight"])), W[p + down + ".weight"])
    return x + n


def flamingo_layer_forward(cfg, W, layer, x, media, media_locations, attn_bias) -> Tensor:
    """``FlamingoLayer.forward`` (open_flamingo/src/flamingo_lm.py:46-83): gated x-attn
    (if this layer has one) THEN the decoder block."""
    if cfg.has_xattn(layer):
        x = gated_xattn_forward(cfg, W, layer, x, media, media_locations)
    return mpt_block_forward(cfg, W, layer, x, attn_bias)


# --------------------------------------------------------------------------------------
# action head
# --------------------------------------------------------------------------------------
def lstm_cell(x: Tensor, h: Tensor, c: Tensor, w_ih, w_hh, b_ih, b_hh) -> Tuple[Tensor, Tensor]:
    """One step of torch.nn.LSTM (gate order i,f,g,o)."""
    gates = _linear(x, w_ih, b_ih) + _linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def head_forward(cfg: DeerConfig, W: Dict[str, Tensor], feat: Tensor, state: Optional[Tuple[Tensor, Tensor]],
                 valid_len: Optional[Tensor] = None, prefix: str = "extra_exit."):
    """``DeterministicDecoder.forward`` single-step path (action_head.py:499-611, 548-560):
    max-pool over the S token positions (519-520) -> LayerNormLSTM (42-64: the stored h is
    the raw LSTM h, the next layer's input is LayerNorm(h)) -> MLPTanhHead / MLPSigmoidHead
    (82-161 / 183-269, layerwise mode: [Dropout, Linear, LN, ReLU]*k, Linear, Tanh|Sigmoid).

    feat (B,S,d); state = (h,c) each (layers,B,H) or None; valid_len (B,) optional: pool only
    the first valid_len positions (the batch-1 reference never sees pads).
    Returns pose (B,6), gripper (B,1), pooled (B,d), new_state."""
    B, S, D = feat.shape
    if valid_len is not None:
        pos = torch.arange(S).view(1, S, 1)
        feat = feat.masked_fill(pos >= valid_len.view(B, 1, 1), float("-inf"))
    pooled = feat.amax(dim=1)
    Hh = cfg.lstm_hidden
    if state is None:
        h0 = torch.zeros(cfg.lstm_layers, B, Hh)
        c0 = torch.zeros(cfg.lstm_layers, B, Hh)
    else:
        h0, c0 = state
    x = pooled
    hs, cs = [], []
    for l in range(cfg.lstm_layers):
        if cfg.lstm_layernorm:
            p = f"{prefix}rnn.layers.{3 * l}."
            h2, c2 = lstm_cell(x, h0[l], c0[l], W[p + "weight_ih_l0"], W[p + "weight_hh_l0"],
                               W[p + "bias_ih_l0"], W[p + "bias_hh_l0"])
            x = _ln(h2, W[f"{prefix}rnn.layers.{3 * l + 1}.weight"], W[f"{prefix}rnn.layers.{3 * l + 1}.bias"])
        else:                                      # stacked nn.LSTM: layer l+1 consumes h of layer l (inter-layer dropout is identity in eval)
            p = f"{prefix}rnn."
            h2, c2 = lstm_cell(x, h0[l], c0[l], W[p + f"weight_ih_l{l}"], W[p + f"weight_hh_l{l}"],
                               W[p + f"bias_ih_l{l}"], W[p + f"bias_hh_l{l}"])
            x = h2
        hs.append(h2)
        cs.append(c2)
    outs = []
    for head in ("actions", "gripper"):
        y = x
        for j in range(len(cfg.mlp_dims)):
            p = f"{prefix}{head}.mlp."
            y = _linear(y, W[p + f"{4 * j + 1}.weight"], W[p + f"{4 * j + 1}.bias"])
            if cfg.mlp_layernorm:
                y = _ln(y, W[p + f"{4 * j + 2}.weight"], W[p + f"{4 * j + 2}.bias"])
            y = torch.relu(y)
        k = 4 * len(cfg.mlp_dims) + 1
        y = _linear(y, W[f"{prefix}{head}.mlp.{k}.weight"], W[f"{prefix}{head}.mlp.{k}.bias"])
        outs.append(torch.tanh(y) if head == "actions" else torch.sigmoid(y))
    return outs[0], outs[1], pooled, (torch.stack(hs), torch.stack(cs))


def get_delta(a1: Tensor, a2: Tensor, threshold_type: str = "L2") -> Tensor:
    """``ActionValueNet.forward.get_delta`` (value_net.py:105-117)."""
    d = (a1 - a2).abs()
    if threshold_type == "mean":
        return d.mean(-1)
    if threshold_type == "L2":
        return d.pow(2).mean(-1).pow(0.5)
    if threshold_type == "max":
        return d.max(-1)[0]
    if threshold_type == "cosine":
        n1 = F.normalize(a1, p=2.0, dim=-1, eps=1e-5)
        n2 = F.normalize(a2, p=2.0, dim=-1, eps=1e-5)
        return 1 - (n1 * n2).sum(-1)
    raise NotImplementedError(threshold_type)


# --------------------------------------------------------------------------------------
# threshold solver
# --------------------------------------------------------------------------------------
def exit_probs(n_exit: int, exit_ratio: float, exit_dist: str = "exp", llm_name: str = "mpt_dolly_3b") -> Tensor:
    """value_net.py:216-238."""
    if exit_dist == "exp":
        probs = exit_ratio ** torch.arange(1, n_exit + 1)
        probs = probs.to(torch.float32) if not probs.is_floating_point() else probs
    elif exit_dist == "gauss":
        probs = torch.tensor([math.exp(-(i - exit_ratio) ** 2 / 2.0) for i in range(n_exit)])
    elif exit_dist == "gamma":
        import scipy.stats
        xs = torch.arange(1, n_exit + 1, dtype=torch.float32)
        probs = torch.tensor([scipy.stats.gamma.pdf(float(v), exit_ratio, scale=2.0) for v in xs], dtype=torch.float32)
    else:
        raise ValueError("Unsupported exit distribution")
    if "mpt_9b" in llm_name:
        probs[0] = 0
    return probs / probs.sum()


def solve_thresholds(values: Tensor, exit_ratio: float, exit_dist: str = "exp",
                     llm_name: str = "mpt_dolly_3b") -> List[float]:
    """``ExitController.set_threshold`` with ``values`` given, ``leq=True``
    (value_net.py:203-264).  values: (n_exit, n_sample) delta table."""
    n_stage, n_sample = values.shape
    _, sorted_idx = values.sort(dim=1, descending=False)
    filtered = torch.zeros(n_sample)
    T = torch.full((n_stage,), -1e8)
    probs = exit_probs(n_stage, exit_ratio, exit_dist, llm_name)
    for k in range(n_stage - 1):
        count = 0
        out_n = math.floor(n_sample * probs[k])
        for i in range(n_sample):
            ori = sorted_idx[k][i]
            if filtered[ori] == 0:
                count += 1
                if count == out_n:
                    T[k] = values[k][ori]
                    break
        filtered.add_(values[k].le(T[k]).type_as(filtered))
    T[n_stage - 1] = 1e8
    return [float(t) for t in T]


# --------------------------------------------------------------------------------------
# threshold calibration in the training layout ("generate" mode)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def window_hidden_states(cfg: DeerConfig, W: Dict[str, Tensor], rgb: Tensor, gripper: Tensor, ids: Tensor,
                         mask: Optional[Tensor] = None) -> List[Tensor]:
    """``MPTFlamingo.forward`` in the training layout (flamingo_mpt.py:308-461 without an exit controller): every frame
    of every window runs all ``n_layers`` layers.  rgb/gripper (bs, win, 3, H, W); ids (bs, S) repeated over the window
    (value_net.py:337-346).  Returns the post-block hidden states, a list of n_layers tensors (bs*win, S, d)."""
    bs, win = rgb.shape[:2]
    S = ids.shape[1]
    ids_f = ids.unsqueeze(1).repeat(1, win, 1).flatten(0, 1)
    mask_f = torch.ones_like(ids_f) if mask is None else mask.unsqueeze(1).repeat(1, win, 1).flatten(0, 1)
    media, _ = encode_vision(cfg, W, rgb.flatten(0, 1), gripper.flatten(0, 1))
    media_locations = ids_f == cfg.media_token_id
    x = W["lang_encoder.transformer.wte.weight"][ids_f]
    attn_bias = mpt_attn_bias(cfg, mask_f)
    hidden = []
    for l in range(cfg.n_layers):
        x = flamingo_layer_forward(cfg, W, l, x, media, media_locations, attn_bias)
        hidden.append(x)
    return hidden


@torch.no_grad()
def generate_deltas(cfg: DeerConfig, W: Dict[str, Tensor], hidden: List[Tensor], layer_idx: Tensor, window: int = 12) -> Tensor:
    """``ActionValueNet.forward(mode='generate')`` (value_net.py:134-160) on top of the random-layer features of
    flamingo_mpt.py:484-492.  ``layer_idx`` (bs, window): the exit layer whose features stand in for each PAST frame
    (the reference draws it with torch.randint over get_all_exit_idx()).  For seq_id in window//2-1 .. window-2 and every
    exit i in [0]+exit_list the head runs as a fresh LSTM over [rand feats of frames < seq_id] + [layer-i feats of frame
    seq_id] (window path of DeterministicDecoder.forward, action_head.py:596-602, last_action) and delta is taken between
    consecutive exits.  Returns (n_exit, bs * window//2) in the reference's (bs-major, seq-minor) order."""
    S, d = hidden[0].shape[1:]
    feats = [h.reshape(-1, window, S, d) for h in hidden]                     # (bs, win, S, d) per layer
    bs = feats[0].shape[0]
    all_f = torch.stack(feats, dim=2)                                          # (bs, win, n_layers, S, d)
    gi = layer_idx.reshape(bs, window, 1, 1, 1).expand(-1, -1, 1, S, d)
    rand_feat = torch.gather(all_f, 2, gi).squeeze(2)                          # (bs, win, S, d)  flamingo_mpt.py:489-490
    exits = [0] + list(cfg.exit_list)
    per_seq = []
    for seq_id in range(window // 2 - 1, window - 1):
        acts = []
        for i in exits:
            seq = torch.cat([rand_feat[:, :seq_id], feats[i][:, seq_id:seq_id + 1]], dim=1)      # (bs, seq_id+1, S, d)
            state = None
            for t in range(seq_id + 1):                                        # nn.LSTM over the sequence from a zero state
                pose, _, _, state = head_forward(cfg, W, seq[:, t], state)
            acts.append(pose)                                                  # last_action: the last frame's output
        per_seq.append(torch.stack(acts))                                      # (n_exit+1, bs, 6)
    a = torch.stack(per_seq).permute(1, 2, 0, 3)                               # (n_exit+1, bs, win/2, 6)
    return get_delta(a[:-1], a[1:], cfg.threshold_type).flatten(1, 2)


# --------------------------------------------------------------------------------------
# the per-timestep forward with dynamic exit
# --------------------------------------------------------------------------------------
@dataclass
class StepOutput:
    pose: Tensor                     # (6,)
    gripper: Tensor                  # (1,) sigmoid output
    action: Tensor                   # (7,) post-processed (eval_utils.py:458-464)
    exit_layer: int
    deltas: Dict[int, float] = field(default_factory=dict)    # exit index -> delta evaluated
    taps: Dict[str, Tensor] = field(default_factory=dict)
    ensemble_action: Optional[Tensor] = None                   # (7,) use_action_ensemble variant (eval_utils.py:457-464)
    feature: Optional[Tensor] = None                           # (d,) pooled feature = .logits[2] (action_head.py:608-609)


class DeerOracle:
    """One episode's view of ``ModelWrapper.step`` -> ``MPTFlamingo.forward`` with
    ``dynamic_early_exit=True`` (eval_utils.py:279-490, flamingo_mpt.py:308-461,
    mosaic_gpt_3b.py:274-449, value_net.py:97-133,277-297).  Holds the carried LSTM state of
    ONE episode, exactly like the reference's module attributes do."""

    def __init__(self, cfg: DeerConfig, weights: Dict[str, Tensor], thresholds: Optional[List[float]] = None):
        self.cfg = cfg
        self.W = weights
        self.thresholds: Optional[Dict[int, float]] = None
        if thresholds is not None:
            self.set_thresholds(thresholds)
        self.state: Optional[Tuple[Tensor, Tensor]] = None
        self.cur_exit_id: Optional[int] = None
        self.timestep = 0

    # value_net.py:177-182
    def set_thresholds(self, thresholds: List[float]):
        assert len(thresholds) == self.cfg.n_thresholds, (len(thresholds), self.cfg.n_thresholds)
        self.thresholds = {self.cfg.exit_list[i]: float(thresholds[i]) for i in range(self.cfg.n_thresholds)}

    # eval_utils.py:252-277, flamingo_mpt.py:292-306
    def reset(self):
        self.state = None
        self.cur_exit_id = None
        self.timestep = 0

    def set_timestep(self, t: int):      # value_net.py:274-275
        self.timestep = t

    @torch.no_grad()
    def step(self, rgb: Tensor, gripper: Tensor, ids: Tensor, mask: Optional[Tensor] = None,
             full_depth: bool = False, taps: bool = False) -> StepOutput:
        """rgb, gripper (3,H,W) f32 normalised; ids (S,) int64; mask (S,) {0,1}.
        ``full_depth=True`` ignores thresholds, runs every layer and records delta at every
        exit (calibration helper) but still exits/commits at the last exit."""
        cfg, W = self.cfg, self.W
        S = ids.shape[0]
        if mask is None:
            mask = torch.ones(S, dtype=torch.int64)
        media, vt = encode_vision(cfg, W, rgb.unsqueeze(0), gripper.unsqueeze(0))
        out_taps: Dict[str, Tensor] = {}
        if taps:
            out_taps.update({k: v[0].clone() for k, v in vt.items()})
        media_locations = (ids == cfg.media_token_id).view(1, S)                     # flamingo_lm.py:211
        x = W["lang_encoder.transformer.wte.weight"][ids].unsqueeze(0)               # mosaic_gpt_3b.py:341-343
        attn_bias = mpt_attn_bias(cfg, mask.view(1, S))
        valid_len = mask.sum().view(1)
        hidden: List[Tensor] = []
        deltas: Dict[int, float] = {}
        prev_action: Optional[Tensor] = None
        exit_layer = cfg.n_layers - 1
        last_probe = None
        action_list: List[Tuple[Tensor, Tensor]] = []        # value_net.py:129 (reset per step in ensemble mode, eval_utils.py:461)
        stage_boundary = (self.timestep % cfg.steps_per_stage == 0)
        for b_idx in range(cfg.n_layers):                                            # mosaic_gpt_3b.py:397
            x = flamingo_layer_forward(cfg, W, b_idx, x, media, media_locations, attn_bias)
            hidden.append(x)                                                         # :424-427 post-block state
            if taps:
                out_taps[f"hidden_{b_idx}"] = x[0].clone()
            # ---- ExitController.forward (value_net.py:277-297)
            if b_idx not in cfg.exit_list:
                continue
            if not stage_boundary and not full_depth:                                # value_net.py:285-286
                if b_idx >= self.cur_exit_id:
                    exit_layer = b_idx
                    break
                continue
            # ---- ActionValueNet.forward infer (value_net.py:120-133)
            assert b_idx > 0
            if b_idx - cfg.exit_interval < 0 or prev_action is None:
                prev_action = head_forward(cfg, W, hidden[b_idx - 1], self.state, valid_len)[0]
            probe = head_forward(cfg, W, hidden[b_idx], self.state, valid_len)
            delta = float(get_delta(probe[0], prev_action, cfg.threshold_type)[0])
            deltas[b_idx] = delta
            prev_action = probe[0]
            last_probe = probe
            action_list.append((probe[0], probe[1]))
            if full_depth:
                if b_idx >= cfg.ctrl_max_layer:
                    exit_layer = b_idx
                    break
                continue
            thr = self.thresholds.get(b_idx) if self.thresholds is not None else None
            leq = (thr is not None) and (delta <= thr)
            if leq or b_idx >= cfg.ctrl_max_layer:                                   # value_net.py:293
                self.cur_exit_id = b_idx
                exit_layer = b_idx
                break
        # ---- final head evaluation on the exit layer, committing the LSTM state
        #      (flamingo_mpt.py:444-461; action_head.py:554-556)
        pose, grip, pooled, new_state = head_forward(cfg, W, hidden[exit_layer], self.state, valid_len)
        self.state = new_state
        g = (grip[0] > 0.5).float()
        action = torch.cat([pose[0], (g - 0.5) * 2])                                 # eval_utils.py:458,464
        if taps:
            out_taps["pooled"] = pooled[0].clone()
            out_taps["h"] = new_state[0][:, 0].clone()
            out_taps["c"] = new_state[1][:, 0].clone()
        ens = None
        if action_list:                                                              # value_net.py:92-95, eval_utils.py:460-464
            acts, grips = zip(*action_list[-2:])
            ea, eg = torch.stack(acts, dim=0).mean(0)[0], torch.stack(grips, dim=0).mean(0)[0]
            ens = torch.cat([ea, ((eg > 0.5).float() - 0.5) * 2])
        return StepOutput(pose=pose[0], gripper=grip[0], action=action, exit_layer=exit_layer,
                          deltas=deltas, taps=out_taps, ensemble_action=ens, feature=pooled[0].clone())
