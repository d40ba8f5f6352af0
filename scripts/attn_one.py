"""One ViT-shaped attention launch set (for ncu): python scripts/attn_one.py [B]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deer_vla_b200 import _lib
lib = _lib.load()
B, H, S, W = (int(sys.argv[1]) if len(sys.argv) > 1 else 448), 16, 257, 1024
qkv = torch.randn(B * S, 3 * W, device="cuda").bfloat16()
out = torch.empty(B * S, W, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
p = qkv.data_ptr()
for _ in range(3):
    assert lib.deer_op_attention(0, 0, C.c_void_p(p), 3 * W, C.c_void_p(p + 2 * W), C.c_void_p(p + 4 * W), 3 * W, C.c_void_p(out.data_ptr()), W,
                                 B, H, S, S, 64, 0.125, C.c_void_p(st)) == 0
torch.cuda.synchronize()
