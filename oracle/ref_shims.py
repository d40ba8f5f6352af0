"""BUILD-CONTAINER-ONLY helper (test infrastructure): import the *unmodified* reference
modules from ``/root/reference`` so the oracle restatement can be validated against them
and golden vectors can be generated (``oracle/make_golden.py``).

``/root/reference`` does not exist on the GPU box: nothing in ``tests -m gpu``, ``smoke()``
or ``bench.py`` imports this file.

What is shimmed (packages absent from the image, none of them on the arithmetic path):
``einops_exts.rearrange_many`` (3-line einops wrapper), ``zarr``, ``thop.profile``,
``fvcore.nn.FlopCountAnalysis``, ``open_clip`` (empty), and
``robot_flamingo.models.trajectory_gpt2.get_gpt_model`` (GPT head, unused by DeeR).

What is RESTATED here because it is third-party code that the reference downloads at run
time and is not vendored under /root/reference:
  * HF-hub ``mosaicml/mpt-1b-redpajama-200b-dolly`` remote code (``attention.py``,
    ``gpt_blocks.py``, ``configuration_mosaic_gpt.py``, ``param_init_fns.py``,
    ``low_precision_layernorm.py``) -- the relative imports of ``mosaic_gpt_3b.py:18-23``.
    With these five sibling modules provided, the reference's own patched
    ``MosaicGPT.forward`` (mosaic_gpt_3b.py:274-449) is executed unmodified.
  * open_clip 2.20.0 ``VisionTransformer`` (ViT-L-14, quick_gelu, output_tokens=True).
"""
from __future__ import annotations

import importlib.util
import math
import sys
import types
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference"


# --------------------------------------------------------------------------------------
# shims for absent python packages
# --------------------------------------------------------------------------------------
def install_shims():
    if "einops_exts" not in sys.modules:
        from einops import rearrange
        m = types.ModuleType("einops_exts")
        m.rearrange_many = lambda tensors, pattern, **kw: tuple(rearrange(t, pattern, **kw) for t in tensors)
        sys.modules["einops_exts"] = m
    if "zarr" not in sys.modules:
        m = types.ModuleType("zarr")
        m.Array = type("Array", (), {})
        m.Group = type("Group", (), {})
        sys.modules["zarr"] = m
    if "thop" not in sys.modules:
        m = types.ModuleType("thop")
        m.profile = lambda *a, **k: (0, 0)
        sys.modules["thop"] = m
    if "fvcore" not in sys.modules:
        m = types.ModuleType("fvcore")
        n = types.ModuleType("fvcore.nn")
        n.FlopCountAnalysis = type("FlopCountAnalysis", (), {"__init__": lambda s, *a, **k: None, "total": lambda s: 0})
        m.nn = n
        sys.modules["fvcore"] = m
        sys.modules["fvcore.nn"] = n
    if "open_clip" not in sys.modules:
        sys.modules["open_clip"] = types.ModuleType("open_clip")
    for p in (REF, REF + "/open_flamingo"):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "robot_flamingo.models.trajectory_gpt2" not in sys.modules:
        m = types.ModuleType("robot_flamingo.models.trajectory_gpt2")
        m.get_gpt_model = lambda *a, **k: None
        sys.modules["robot_flamingo.models.trajectory_gpt2"] = m


# --------------------------------------------------------------------------------------
# restated MPT-1B remote code (HF hub mosaicml/mpt-1b-redpajama-200b-dolly)
# --------------------------------------------------------------------------------------
def _build_mpt1b_package():
    """Creates the package ``mpt1b_remote`` holding restated sibling modules and loads the
    reference's ``mosaic_gpt_3b.py`` into it as ``mpt1b_remote.mosaic_gpt``."""
    if "mpt1b_remote.mosaic_gpt" in sys.modules:
        return sys.modules["mpt1b_remote.mosaic_gpt"]
    from transformers import PretrainedConfig
    from einops import rearrange

    pkg = types.ModuleType("mpt1b_remote")
    pkg.__path__ = []
    sys.modules["mpt1b_remote"] = pkg

    # ---- configuration_mosaic_gpt.py
    cfgm = types.ModuleType("mpt1b_remote.configuration_mosaic_gpt")

    class MosaicGPTConfig(PretrainedConfig):
        model_type = "mosaic_gpt"

        def __init__(self, d_model=2048, n_heads=16, n_layers=24, mlp_ratio=4, max_seq_len=2048,
                     vocab_size=50368, attn_pdrop=0.0, resid_pdrop=0.0, emb_pdrop=0.0,
                     attn_impl="torch", attn_qk_ln=False, attn_clip_qkv=None, softmax_scale=None,
                     prefix_lm=False, attn_uses_sequence_id=False, alibi=False, alibi_bias_max=8,
                     init_device="cpu", logit_scale=None, no_bias=False, verbose=0,
                     param_init_fn="kaiming_normal_", init_div_is_residual=True, init_std=0.02,
                     emb_init_std=None, emb_init_uniform_lim=None, init_gain=0, fan_mode="fan_in",
                     init_nonlinearity="relu", embedding_fraction=1.0, low_precision_layernorm=True,
                     use_cache=False, **kwargs):
            self.d_model, self.n_heads, self.n_layers, self.mlp_ratio = d_model, n_heads, n_layers, mlp_ratio
            self.max_seq_len, self.vocab_size = max_seq_len, vocab_size
            self.attn_pdrop, self.resid_pdrop, self.emb_pdrop = attn_pdrop, resid_pdrop, emb_pdrop
            self.attn_impl, self.attn_qk_ln, self.attn_clip_qkv = attn_impl, attn_qk_ln, attn_clip_qkv
            self.softmax_scale, self.prefix_lm = softmax_scale, prefix_lm
            self.attn_uses_sequence_id, self.alibi, self.alibi_bias_max = attn_uses_sequence_id, alibi, alibi_bias_max
            self.init_device, self.logit_scale, self.no_bias, self.verbose = init_device, logit_scale, no_bias, verbose
            self.param_init_fn, self.init_div_is_residual, self.init_std = param_init_fn, init_div_is_residual, init_std
            self.emb_init_std, self.emb_init_uniform_lim, self.init_gain = emb_init_std, emb_init_uniform_lim, init_gain
            self.fan_mode, self.init_nonlinearity = fan_mode, init_nonlinearity
            self.embedding_fraction, self.low_precision_layernorm = embedding_fraction, low_precision_layernorm
            self.use_cache = use_cache
            super().__init__(**kwargs)

    cfgm.MosaicGPTConfig = MosaicGPTConfig
    sys.modules[cfgm.__name__] = cfgm

    # ---- low_precision_layernorm.py
    lpm = types.ModuleType("mpt1b_remote.low_precision_layernorm")

    class LPLayerNorm(nn.LayerNorm):
        """On CPU fp32 (no autocast) the low-precision LayerNorm is a plain LayerNorm."""

        def forward(self, x):
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)

    lpm.LPLayerNorm = LPLayerNorm
    sys.modules[lpm.__name__] = lpm

    # ---- param_init_fns.py
    pim = types.ModuleType("mpt1b_remote.param_init_fns")
    pim.MODEL_INIT_REGISTRY = {"kaiming_normal_": lambda module, **kw: None}
    sys.modules[pim.__name__] = pim

    # ---- attention.py
    atm = types.ModuleType("mpt1b_remote.attention")

    def alibi_bias(n_heads, seq_len, full=False, alibi_bias_max=8, device=None, dtype=None):
        ab = torch.arange(1 - seq_len, 1, dtype=dtype, device=device).view(1, 1, 1, seq_len)
        if full:
            ab = ab - torch.arange(1 - seq_len, 1, dtype=dtype, device=device).view(1, 1, seq_len, 1)
            ab = ab.abs().mul(-1)
        m = torch.arange(1, n_heads + 1, dtype=dtype, device=device)
        m = m.mul(alibi_bias_max / n_heads)
        return ab * (1.0 / 2 ** m.view(1, n_heads, 1, 1))

    def attn_bias_shape(attn_impl, n_heads, seq_len, alibi, prefix_lm, causal, use_sequence_id):
        if attn_impl == "flash":
            return None
        if attn_impl in ("torch", "triton"):
            if alibi:
                if (prefix_lm or not causal) or use_sequence_id:
                    return (1, n_heads, seq_len, seq_len)
                return (1, n_heads, 1, seq_len)
            if prefix_lm or use_sequence_id:
                return (1, 1, seq_len, seq_len)
            return None
        raise ValueError(attn_impl)

    def attn_bias(attn_impl, attn_bias, n_heads, seq_len, causal=False, alibi=False, alibi_bias_max=8):
        if attn_impl == "flash":
            return None
        if attn_impl in ("torch", "triton"):
            if alibi:
                device, dtype = attn_bias.device, attn_bias.dtype
                attn_bias = attn_bias.add(alibi_bias(n_heads, seq_len, full=not causal,
                                                     alibi_bias_max=alibi_bias_max, device=device, dtype=dtype))
            return attn_bias
        raise ValueError(attn_impl)

    def scaled_multihead_dot_product_attention(query, key, value, n_heads, softmax_scale=None, attn_bias=None,
                                               key_padding_mask=None, is_causal=False, dropout_p=0.0,
                                               training=False, needs_weights=False):
        q = rearrange(query, "b s (h d) -> b h s d", h=n_heads)
        k = rearrange(key, "b s (h d) -> b h d s", h=n_heads)
        v = rearrange(value, "b s (h d) -> b h s d", h=n_heads)
        min_val = torch.finfo(q.dtype).min
        b, _, s_q, d = q.shape
        s_k = k.size(-1)
        if softmax_scale is None:
            softmax_scale = 1 / math.sqrt(d)
        attn_weight = q.matmul(k) * softmax_scale
        if attn_bias is not None:
            attn_weight = attn_weight + attn_bias
        if key_padding_mask is not None:
            attn_weight = attn_weight.masked_fill(~key_padding_mask.view((b, 1, 1, s_k)), min_val)
        if is_causal:
            s = max(s_q, s_k)
            causal_mask = attn_weight.new_ones(s, s, dtype=torch.float16).tril().to(torch.bool)
            causal_mask = ~causal_mask
            causal_mask = causal_mask[-s_q:, -s_k:]
            attn_weight = attn_weight.masked_fill(causal_mask.view(1, 1, s_q, s_k), min_val)
        attn_weight = torch.softmax(attn_weight, dim=-1)
        out = attn_weight.matmul(v)
        out = rearrange(out, "b h s d -> b s (h d)")
        return out, (attn_weight if needs_weights else None)

    class MultiheadAttention(nn.Module):
        def __init__(self, d_model, n_heads, attn_impl="torch", attn_clip_qkv=None, attn_qk_ln=False,
                     softmax_scale=None, attn_pdrop=0.0, low_precision_layernorm=False, device=None):
            super().__init__()
            self.clip_qkv, self.attn_qk_ln = attn_clip_qkv, attn_qk_ln
            self.d_model, self.n_heads = d_model, n_heads
            self.softmax_scale = softmax_scale if softmax_scale is not None else 1 / math.sqrt(d_model / n_heads)
            self.attn_dropout_p = attn_pdrop
            self.Wqkv = nn.Linear(d_model, 3 * d_model, device=device)
            if attn_qk_ln:
                lnc = LPLayerNorm if low_precision_layernorm else nn.LayerNorm
                self.q_ln = lnc(d_model, device=device)
                self.k_ln = lnc(d_model, device=device)
            self.attn_fn = scaled_multihead_dot_product_attention
            self.out_proj = nn.Linear(d_model, d_model, device=device)

        def forward(self, x, past_key_value=None, attn_bias=None, attention_mask=None, is_causal=True,
                    needs_weights=False):
            qkv = self.Wqkv(x)
            if self.clip_qkv:
                qkv.clamp_(min=-self.clip_qkv, max=self.clip_qkv)
            query, key, value = qkv.chunk(3, dim=2)
            key_padding_mask = attention_mask
            if self.attn_qk_ln:
                dtype = query.dtype
                query = self.q_ln(query).to(dtype)
                key = self.k_ln(key).to(dtype)
            if past_key_value is not None:
                if len(past_key_value) != 0:
                    key = torch.cat([past_key_value[0], key], dim=1)
                    value = torch.cat([past_key_value[1], value], dim=1)
                past_key_value = (key, value)
            if attn_bias is not None:
                attn_bias = attn_bias[:, :, -query.size(1):, -key.size(1):]
            context, attn_weights = self.attn_fn(query, key, value, self.n_heads, softmax_scale=self.softmax_scale,
                                                 attn_bias=attn_bias, key_padding_mask=key_padding_mask,
                                                 is_causal=is_causal, dropout_p=self.attn_dropout_p,
                                                 training=self.training, needs_weights=needs_weights)
            return self.out_proj(context), attn_weights, past_key_value

    atm.alibi_bias, atm.attn_bias_shape, atm.attn_bias = alibi_bias, attn_bias_shape, attn_bias
    atm.scaled_multihead_dot_product_attention = scaled_multihead_dot_product_attention
    atm.MultiheadAttention = MultiheadAttention
    sys.modules[atm.__name__] = atm

    # ---- gpt_blocks.py
    gbm = types.ModuleType("mpt1b_remote.gpt_blocks")

    class GPTMLP(nn.Module):
        def __init__(self, d_model, mlp_ratio, device=None):
            super().__init__()
            self.mlp_up = nn.Linear(d_model, mlp_ratio * d_model, device=device)
            self.mlp_act = nn.GELU(approximate="none")
            self.mlp_down = nn.Linear(mlp_ratio * d_model, d_model, device=device)

        def forward(self, x):
            return self.mlp_down(self.mlp_act(self.mlp_up(x)))

    class GPTBlock(nn.Module):
        def __init__(self, attn_impl, d_model, n_heads, mlp_ratio, attn_clip_qkv=None, attn_qk_ln=False,
                     softmax_scale=None, attn_pdrop=0.0, alibi=False, resid_pdrop=0.0,
                     low_precision_layernorm=False, device=None, **kwargs):
            super().__init__()
            lnc = LPLayerNorm if low_precision_layernorm else nn.LayerNorm
            self.ln_1 = lnc(d_model, device=device)
            self.attn = MultiheadAttention(attn_impl=attn_impl, attn_clip_qkv=attn_clip_qkv, attn_qk_ln=attn_qk_ln,
                                           softmax_scale=softmax_scale, attn_pdrop=attn_pdrop, d_model=d_model,
                                           n_heads=n_heads, low_precision_layernorm=low_precision_layernorm,
                                           device=device)
            self.ln_2 = lnc(d_model, device=device)
            self.mlp = GPTMLP(d_model=d_model, mlp_ratio=mlp_ratio, device=device)
            self.resid_attn_dropout = nn.Dropout(resid_pdrop)
            self.resid_mlp_dropout = nn.Dropout(resid_pdrop)

        def forward(self, x, past_key_value=None, attn_bias=None, attention_mask=None, is_causal=True):
            a = self.ln_1(x)
            b, _, past_key_value = self.attn(a, past_key_value=past_key_value, attn_bias=attn_bias,
                                             attention_mask=attention_mask, is_causal=is_causal)
            x = x + self.resid_attn_dropout(b)
            m = self.ln_2(x)
            n = self.mlp(m)
            x = x + self.resid_mlp_dropout(n)
            return x, past_key_value

    gbm.GPTMLP, gbm.GPTBlock = GPTMLP, GPTBlock
    sys.modules[gbm.__name__] = gbm

    # ---- the reference's own patched top-level model, loaded unmodified
    spec = importlib.util.spec_from_file_location("mpt1b_remote.mosaic_gpt", REF + "/mosaic_gpt_3b.py",
                                                  submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "mpt1b_remote"
    sys.modules["mpt1b_remote.mosaic_gpt"] = mod
    spec.loader.exec_module(mod)
    return mod


# --------------------------------------------------------------------------------------
# restated open_clip 2.20.0 VisionTransformer (parameter names == open_clip's)
# --------------------------------------------------------------------------------------
class _QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class _ResBlock(nn.Module):
    def __init__(self, width, heads, mlp):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.MultiheadAttention(width, heads)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(width, mlp)), ("gelu", _QuickGELU()),
                                              ("c_proj", nn.Linear(mlp, width))]))

    def forward(self, x):
        a = self.ln_1(x)
        x = x + self.attn(a, a, a, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads, mlp):
        super().__init__()
        self.resblocks = nn.ModuleList([_ResBlock(width, heads, mlp) for _ in range(layers)])

    def forward(self, x):
        for r in self.resblocks:
            x = r(x)
        return x


class OpenClipVisualRestated(nn.Module):
    """open_clip 2.20.0 ``VisionTransformer`` with ``output_tokens=True`` (factory.py:109-114)."""

    def __init__(self, image_size, patch, width, layers, heads, mlp, output_dim=768):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, patch, patch, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(width))
        self.positional_embedding = nn.Parameter(torch.zeros((image_size // patch) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads, mlp)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(torch.zeros(width, output_dim))
        self.output_tokens = True

    def forward(self, x):
        x = self.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        x = torch.cat([self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        pooled, tokens = x[:, 0], x[:, 1:]
        pooled = self.ln_post(pooled) @ self.proj
        return pooled, tokens


class _ClipShell(nn.Module):
    def __init__(self, visual):
        super().__init__()
        self.visual = visual


# --------------------------------------------------------------------------------------
# build the reference's real MPTFlamingo + ExitController on CPU
# --------------------------------------------------------------------------------------
def build_reference_model(cfg, weights, thresholds=None, dropout_mode="layerwise"):
    """Returns (model: reference MPTFlamingo, exit_controller: reference ExitController).
    ``cfg`` is an ``oracle.deer_oracle.DeerConfig`` with qk_ln=True (MPT-1B path).  ``dropout_mode`` / ``cfg.mlp_layernorm``
    / ``cfg.lstm_layernorm`` / ``len(cfg.mlp_dims)`` select the head variant the reference builds (action_head.py:66-161);
    ``weights`` must then carry the reference's OWN key names for that variant."""
    install_shims()
    mg = _build_mpt1b_package()
    import torch.distributed as dist
    if not dist.is_initialized():
        import os
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        dist.init_process_group("gloo", rank=0, world_size=1)
    from open_flamingo.src.flamingo_lm import FlamingoLMMixin
    from open_flamingo.src.utils import extend_instance
    from robot_flamingo.models.flamingo_mpt import MPTFlamingo
    from robot_flamingo.models.value_net import ActionValueNet, ExitController

    assert cfg.qk_ln, "only the MPT-1B (3B flamingo) top-level file is driven through the reference"
    total_layers = max(cfg.ckpt_early_exit_layer + 1, cfg.n_layers)
    mcfg = mg.MosaicGPTConfig(d_model=cfg.d_model, n_heads=cfg.n_heads, n_layers=total_layers,
                              mlp_ratio=cfg.mlp_ratio, max_seq_len=2048, vocab_size=cfg.vocab,
                              attn_impl="torch", attn_qk_ln=True, alibi=True, alibi_bias_max=cfg.alibi_bias_max,
                              no_bias=True, low_precision_layernorm=True, init_device="cpu")
    lang = mg.MosaicGPT(mcfg)
    extend_instance(lang, FlamingoLMMixin)
    lang.set_decoder_layers_attr_name("transformer.blocks")          # open_flamingo factory.py:139-140
    visual = OpenClipVisualRestated(cfg.image_size, cfg.patch, cfg.vit_width, cfg.vit_layers, cfg.vit_heads,
                                    cfg.vit_mlp)
    model = MPTFlamingo(_ClipShell(visual), lang, cfg.eoc_token_id, cfg.media_token_id, vis_dim=cfg.vit_width,
                        cross_attn_every_n_layers=cfg.cross_attn_every, window_size=12, use_gripper=True,
                        fusion_mode="post", llm=cfg.llm_name, pooling="max", decoder_type="lstm",
                        head_type="deterministic", return_feature=True,
                        early_exit_layer=cfg.early_exit_layer, multi_exit=False, exit_interval=cfg.exit_interval,
                        dropout_mode=dropout_mode, mlp_layernorm=cfg.mlp_layernorm, lstm_layernorm=cfg.lstm_layernorm,
                        mlp_num_hidden_layers=len(cfg.mlp_dims), lstm_num_layers=cfg.lstm_layers)
    sd = model.state_dict()
    # a checkpoint trained with more layers than max_layer keeps is loaded strict=False
    # (eval_calvin.py:576-577, "disabled due to max layer constrain")
    extra = [k for k in weights if k not in sd]
    assert all(".blocks." in k for k in extra), extra[:5]
    res = model.load_state_dict({k: v for k, v in weights.items() if k in sd}, strict=False)
    assert not [k for k in res.missing_keys if k.startswith("extra_exit.")], res.missing_keys[:8]   # head keys of this variant
    model.eval()
    value_net = ActionValueNet(exit_list=model.get_all_exit_idx(), exit_head=model.extra_exit,
                               interval=cfg.exit_interval, window_size=12, threshold_type=cfg.threshold_type)
    ctrl = ExitController(value_net, exit_id_list=model.get_all_exit_idx(), steps_per_stage=cfg.steps_per_stage,
                          leq=True, exit_dist="exp", max_layer=cfg.max_layer)
    ctrl.eval()
    if thresholds is not None:
        ctrl._set_threshold_value(list(thresholds))
    return model, ctrl, res
