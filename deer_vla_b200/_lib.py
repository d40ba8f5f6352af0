"""ctypes binding of the C ABI declared in ``include/deer_engine.h``.

The library is the product: if ``libdeer_engine.so`` is missing the import fails loudly --
there is no CPU / PyTorch fallback path anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEER_LIB") or os.path.join(_HERE, "libdeer_engine.so")   # DEER_LIB: an instrumented build (scripts/)

ABI_VERSION = 2
PREC_BF16, PREC_FP32 = 0, 1
GEMM_TCGEN05, GEMM_SIMT = 0, 1
THR_TYPES = {"L2": 0, "mean": 1, "max": 2, "cosine": 3}
FLAG_FULL_DEPTH = 1


class DeerCConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version",
        "image_size", "patch", "vit_width", "vit_layers", "vit_heads", "vit_mlp",
        "perc_depth", "perc_heads", "perc_dim_head", "num_latents", "perc_ff_mult",
        "d_model", "n_heads", "mlp_ratio", "vocab", "qk_ln", "alibi_bias_max", "cross_attn_every",
        "xattn_heads", "xattn_dim_head", "xattn_ff_mult", "media_token_id",
        "n_layers", "exit_interval", "max_layer", "is_mpt_9b",
        "lstm_hidden", "lstm_layers", "mlp_dim0", "mlp_dim1", "threshold_type", "steps_per_stage",
        "max_batch", "max_slots", "max_seq", "precision", "gemm_impl", "debug_taps", "host_control", "use_graph",
        "mlp_dim2", "mlp_layernorm", "lstm_layernorm", "action_ensemble")]


class DeerStats(C.Structure):
    _fields_ = [("kernel_launches_total", C.c_int64), ("kernel_launches_last_step", C.c_int64),
                ("gemm_launches_last_step", C.c_int64), ("gemm_flops_last_step", C.c_double),
                ("steps", C.c_int64), ("weight_bytes", C.c_int64), ("workspace_bytes", C.c_int64)]


class DeerProfile(C.Structure):
    _fields_ = [("gemm_ms", C.c_double * 4), ("gemm_flops", C.c_double * 4), ("gemm_launches", C.c_int64 * 4),
                ("step_ms", C.c_double), ("n_probes", C.c_int32), ("active_after", C.c_int32 * 16),
                ("gemm_bytes", C.c_double * 4)]


# every symbol include/deer_engine.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "deer_abi_version": (C.c_int, []),
    "deer_build_info": (C.c_char_p, []),
    "deer_source_hash": (C.c_char_p, []),
    "deer_last_error": (C.c_char_p, [_P]),
    "deer_create": (C.c_int, [C.POINTER(DeerCConfig), C.c_int, C.POINTER(_P)]),
    "deer_destroy": (None, [_P]),
    "deer_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "deer_finalize_weights": (C.c_int, [_P]),
    "deer_num_required_weights": (C.c_int, [_P]),
    "deer_required_weight_name": (C.c_char_p, [_P, C.c_int]),
    "deer_set_thresholds": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "deer_num_exits": (C.c_int, [_P]),
    "deer_exit_layer": (C.c_int, [_P, C.c_int]),
    "deer_num_thresholds": (C.c_int, [_P]),
    "deer_ctrl_max_layer": (C.c_int, [_P]),
    "deer_reset": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int]),
    "deer_step": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    "deer_step_host": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_uint32, _P, _P, _P, _P, _P]),
    "deer_step_generate": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P]),
    "deer_preprocess_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "deer_step_host_u8": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_uint32,
                                    _P, _P, _P, _P, _P]),
    "deer_get_state": (C.c_int, [_P, C.c_int, _P, _P]),
    "deer_set_state": (C.c_int, [_P, C.c_int, _P, _P]),
    "deer_get_features": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    "deer_get_tap": (C.c_int, [_P, C.c_char_p, C.c_int, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "deer_get_stats": (C.c_int, [_P, C.POINTER(DeerStats)]),
    "deer_set_launch_overlap": (C.c_int, [C.c_int]),
    "deer_set_row_order": (C.c_int, [C.c_int]),
    "deer_set_profile": (C.c_int, [_P, C.c_int]),
    "deer_get_profile": (C.c_int, [_P, C.POINTER(DeerProfile)]),
    "deer_op_gemm": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                               _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P]),
    "deer_attn_debug": (C.c_int, [C.c_int, _P, C.c_int]),
    "deer_head_debug": (C.c_int, [C.c_int, _P, C.c_int]),
    "deer_op_pv_probe": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "deer_op_attention": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "deer_op_attention_masked": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, _P, _P, _P]),
}

_lib = None


def _tree_source_hash():
    """sha256 over csrc/ + include/deer_engine.h (same recipe as __graft_entry__.source_hash); None when the sources are
    not beside the library (an installed binary)."""
    import hashlib
    csrc = os.path.join(_HERE, "csrc")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "deer_engine.h")
    if not os.path.isdir(csrc) or not os.path.exists(hdr):
        return None
    h = hashlib.sha256()
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))) + [hdr]
    for p in files:
        h.update(os.path.basename(p).encode() + b"\0")
        h.update(open(p, "rb").read())
    return h.hexdigest()[:32]


def load():
    """Load (once) and return the ctypes library handle.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C deer_vla_b200/csrc`). The DeeR engine has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.deer_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.deer_abi_version()} != binding {ABI_VERSION}")
    want = _tree_source_hash()
    have = lib.deer_source_hash().decode()
    if want is not None and have != want and not os.environ.get("DEER_LIB"):
        raise ImportError(f"{LIB_PATH} was built from sources {have}, the tree holds {want}: stale library, rebuild it "
                          "(python -c 'import __graft_entry__ as g; g.build()')")
    _lib = lib
    return lib


class DeerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"deer_engine error {code}: {msg}")
        self.code = code
