"""Per-k-block timeline of the 1-CTA tcgen05 GEMM kernel (instrumented build: -DDEER_GEMM_PROF, scripts/libdeer_engine_prof.so).
CTA 0's first tile: clock64 of the producer after its wait on the stage's `empty` barrier [0] and after issuing the two TMA loads [1];
of the MMA thread after its wait on `full` [2] and after issuing the four MMAs + commit [3].  Usage (GPU box):
  nvcc ... -DDEER_GEMM_PROF engine.cu -o scripts/libdeer_engine_prof.so ; DEER_LIB=scripts/libdeer_engine_prof.so python scripts/gemm_kblock_prof.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deer_vla_b200 import _lib

lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
raw.deer_gemm_debug.argtypes = [C.c_void_p, C.c_int]
SHAPES = [("llm.down+res M=32", 32, 2048, 8192, True, 2), ("llm.up M=32", 32, 8192, 2048, False, 2), ("vit.fc2+res M=514", 514, 1024, 4096, True, 2),
          ("vit.fc2+res M=514 (128-row tiles)", 514, 1024, 4096, True, 4), ("llm.down+res M=896 (128x128)", 896, 2048, 8192, True, 2),
          ("one tile 128x64, K=8192 (1 CTA alone)", 128, 64, 8192, False, 4)]
for name, M, N, K, resid, impl in SHAPES:
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    Cm = torch.zeros(M, N, device="cuda", dtype=torch.float32 if resid else torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        rc = lib.deer_op_gemm(impl, 0, C.c_void_p(A.data_ptr()), K, C.c_void_p(W.data_ptr()), K, C.c_void_p(Cm.data_ptr()), N, M, N, K,
                              None, 0, None, C.c_void_p(Cm.data_ptr()) if resid else None, int(resid), None, 1, C.c_void_p(st))
        assert rc == 0
    buf = np.zeros(4 * 256, dtype=np.int64)
    raw.deer_gemm_debug(buf.ctypes.data_as(C.c_void_p), buf.size)
    nkb = min(256, (K + 63) // 64)
    t = buf.reshape(256, 4)[:nkb]
    t0 = t[0, 0]
    full = t[:, 2] - t0
    d_full = np.diff(full)
    print(f"== {name}: {nkb} k-blocks; MMA thread sees `full` every {np.median(d_full):.0f} clk (median; mean {d_full.mean():.0f}; first arrival {full[0]} clk after the first TMA issue)")
    print("   producer: wait-on-empty -> TMA issued  median %.0f clk;  producer lead over the MMA thread (k-blocks in flight) median %.1f" %
          (np.median(t[:, 1] - t[:, 0]), np.median([(t[:, 0] <= x).sum() - i for i, x in enumerate(t[:, 2])])))
    print("   MMA thread: `full` seen -> 4 MMAs + commit issued  median %.0f clk; commit(kb) -> producer passes empty(kb + stages): see lead" % np.median(t[:, 3] - t[:, 2]))
    print("   first 12 `full` intervals:", d_full[:12].tolist(), " last 6:", d_full[-6:].tolist())
