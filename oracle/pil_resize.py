"""CPU restatement of the image transform the reference applies before the hot path (test infrastructure).

``preprocess_image`` (robot_flamingo/data/data.py:898-902) calls open_clip's ``image_transform(224, is_train=False)``
[UNVENDORED: open-clip-torch 2.20.0] = torchvision ``Resize(224, BICUBIC)`` on a PIL image -> ``CenterCrop(224)`` ->
``ToTensor`` -> ``Normalize(CLIP mean, std)``.  The resize is Pillow's two-pass (horizontal, then vertical) 8-bit
convolution resampler with 22-bit fixed-point coefficients and a uint8 round-trip between the passes
(Pillow ``src/libImaging/Resample.c``: ``bicubic_filter``, ``precompute_coeffs``, ``normalize_coeffs_8bpc``,
``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``).  Pillow IS present in the image, so this restatement is pinned
bit-exactly against the real library (tests/test_preprocess.py) and then serves as the oracle of the CUDA kernel.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)


def bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """bounds[out,2] (xmin, count) and fixed-point kk[out,ksize] (int32), as Pillow computes them for box=(0,in_size)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img uint8 [H,W,C] -> uint8 [out_h,out_w,C], bit-exact with PIL.Image.resize((out_w,out_h), BICUBIC)."""
    H, W, C = img.shape
    src = img.astype(np.int64)
    if out_w != W:
        bx, kx, _ = precompute_coeffs(W, out_w)
        tmp = np.empty((H, out_w, C), dtype=np.uint8)
        for xx in range(out_w):
            xmin, n = bx[xx]
            acc = np.full((H, C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for x in range(n):
                acc += src[:, xmin + x, :] * int(kx[xx, x])
            tmp[:, xx, :] = _clip8(acc)
        src = tmp.astype(np.int64)
    else:
        tmp = img
    if out_h != H:
        by, ky, _ = precompute_coeffs(H, out_h)
        out = np.empty((out_h, src.shape[1], C), dtype=np.uint8)
        for yy in range(out_h):
            ymin, n = by[yy]
            acc = np.full((src.shape[1], C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for y in range(n):
                acc += src[ymin + y, :, :] * int(ky[yy, y])
            out[yy] = _clip8(acc)
        return out
    return tmp.astype(np.uint8)


def resized_dims(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision ``Resize(size:int)``: the smaller edge becomes ``size``, the other ``int(size * long / short)``."""
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


def clip_transform(img: np.ndarray, size: int = 224) -> np.ndarray:
    """uint8 [H,W,3] -> float32 [3,size,size]: Resize(size, BICUBIC) -> CenterCrop(size) -> /255 -> Normalize."""
    h, w, _ = img.shape
    nh, nw = resized_dims(h, w, size)
    r = resize_bicubic_u8(img, nh, nw)
    top, left = int(round((nh - size) / 2.0)), int(round((nw - size) / 2.0))
    r = r[top:top + size, left:left + size]
    x = r.astype(np.float32) / np.float32(255.0)
    x = (x - CLIP_MEAN) / CLIP_STD
    return np.ascontiguousarray(x.transpose(2, 0, 1))
