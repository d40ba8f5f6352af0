"""Bring-up probe: which (LBO, SBO, k-step) encoding makes the MN-major V operand + TMEM A operand produce P @ V."""
import ctypes as C, sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deer_vla_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
for K in (64, 128, 256):
    P = torch.randn(128, K, device="cuda").bfloat16()
    V = torch.randn(K, 64, device="cuda").bfloat16()
    ref = P.float() @ V.float()
    for lbo, sbo, kstep in itertools.product((16, 1024, 2048), (1024, 128, 2048), (2048, 1024, 256)):
        O = torch.zeros(128, 64, device="cuda")
        rc = lib.deer_op_pv_probe(C.c_void_p(P.data_ptr()), C.c_void_p(V.data_ptr()), C.c_void_p(O.data_ptr()), K, lbo, sbo, kstep, None)
        torch.cuda.synchronize()
        err = (O - ref).abs().max().item()
        tag = "OK " if err < 0.05 else "   "
        print(f"{tag} K={K} lbo={lbo} sbo={sbo} kstep={kstep} maxerr={err:.4f}", flush=True)
