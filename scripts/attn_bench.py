"""ViT attention micro-benchmark through deer_op_attention (fused qkv layout like the engine's v_qkv buffer)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deer_vla_b200 import _lib
lib = _lib.load()
B, H, S, W = 448, 16, 257, 1024
qkv = torch.randn(B * S, 3 * W, device="cuda").bfloat16()
out = torch.empty(B * S, W, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
p = qkv.data_ptr()
for name, impl in (("tcgen05", 0), ("mma.sync flash", 3)):
    def run():
        rc = lib.deer_op_attention(impl, 0, C.c_void_p(p), 3 * W, C.c_void_p(p + 2 * W), C.c_void_p(p + 4 * W), 3 * W, C.c_void_p(out.data_ptr()), W,
                                   B, H, S, S, 64, 0.125, C.c_void_p(st))
        assert rc == 0
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:16s} {ms:7.3f} ms  {4 * B * H * S * S * 64 / ms / 1e9:7.1f} TFLOP/s")

import numpy as np
lib.deer_attn_debug(1, None, 0)
rc = lib.deer_op_attention(0, 0, C.c_void_p(p), 3 * W, C.c_void_p(p + 2 * W), C.c_void_p(p + 4 * W), 3 * W, C.c_void_p(out.data_ptr()), W, B, H, S, S, 64, 0.125, C.c_void_p(st))
buf = np.zeros(512 * 8, dtype=np.int64)
lib.deer_attn_debug(1, buf.ctypes.data_as(C.c_void_p), buf.size)
d = buf.reshape(512, 8)[:148, :8]
tiles = 2 * (B * H) / 148
m = d.mean(0) / tiles
if os.environ.get("DEER_ATTN_V3") or os.environ.get("DEER_ATTN_V4"):
    print("cycles per tile (softmax warp 2): wait_S %.0f | S loads issued->landed %.0f, max+exchange %.0f | epilogue: wait_O %.0f, rest %.0f | exp+P-store %.0f | total %.0f" % (m[0], m[6], m[1], m[5], m[2], m[3], d.sum(1).mean() / tiles))
else:   # v5: warp 2 belongs to group 0 and sees every other tile
    m = m * 2
    print("v5 cycles per tile of ONE group (warp 2): wait_S %.0f | softmax (8 chunks) %.0f | wait_O %.0f | output %.0f | total %.0f  (two groups overlap: SM time per tile ~ total/2)" % (m[0], m[1], m[2], m[3], 2 * d.sum(1).mean() / tiles))
