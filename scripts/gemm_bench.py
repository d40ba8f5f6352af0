"""GEMM micro-benchmark through the C ABI (deer_op_gemm): the step's real shapes, CUDA-event timed."""
import ctypes as C
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deer_vla_b200 import _lib

lib = _lib.load()
SHAPES = [  # name, M, N, K, act, resid(out_f32)
    ("vit.qkv", 115136, 3072, 1024, 0, False), ("vit.out+res", 115136, 1024, 1024, 0, True),
    ("vit.fc1+qgelu", 115136, 4096, 1024, 2, False), ("vit.fc2+res", 115136, 1024, 4096, 0, True),
    ("perc.kv", 143360, 1024, 1024, 0, False), ("perc.ff1+gelu", 28672, 4096, 1024, 1, False),
    ("llm.qkv", 7168, 6144, 2048, 0, False), ("llm.out+res", 7168, 2048, 2048, 0, True),
    ("llm.up+gelu", 7168, 8192, 2048, 1, False), ("llm.down+res", 7168, 2048, 8192, 0, True),
    ("xattn.q", 7168, 512, 2048, 0, False), ("head.lstm0", 224, 4096, 3072, 0, True),
]
SETS = {
    "b28": [("vit.qkv", 14392, 3072, 1024, 0, False), ("vit.out+res", 14392, 1024, 1024, 0, True), ("vit.fc1+qgelu", 14392, 4096, 1024, 2, False),
            ("vit.fc2+res", 14392, 1024, 4096, 0, True), ("perc.kv", 17920, 1024, 1024, 0, False), ("perc.q", 3584, 512, 1024, 0, False),
            ("perc.ff1+gelu", 3584, 4096, 1024, 1, False), ("perc.ff2+res", 3584, 1024, 4096, 0, True),
            ("llm.qkv", 896, 6144, 2048, 0, False), ("llm.out+res", 896, 2048, 2048, 0, True), ("llm.up+gelu", 896, 8192, 2048, 1, False),
            ("llm.down+res", 896, 2048, 8192, 0, True), ("xattn.q", 896, 512, 2048, 0, False), ("xattn.kv", 3584, 1024, 1024, 0, False),
            ("xattn.out+res", 896, 2048, 512, 0, True), ("llm.qkv@448", 448, 6144, 2048, 0, False), ("llm.down+res@448", 448, 2048, 8192, 0, True)],
    "b1": [("vit.qkv", 514, 3072, 1024, 0, False), ("vit.out+res", 514, 1024, 1024, 0, True), ("vit.fc1+qgelu", 514, 4096, 1024, 2, False),
           ("vit.fc2+res", 514, 1024, 4096, 0, True), ("perc.kv", 640, 1024, 1024, 0, False), ("perc.ff1+gelu", 128, 4096, 1024, 1, False),
           ("llm.qkv", 32, 6144, 2048, 0, False), ("llm.out+res", 32, 2048, 2048, 0, True), ("llm.up+gelu", 32, 8192, 2048, 1, False),
           ("llm.down+res", 32, 2048, 8192, 0, True), ("xattn.q", 32, 512, 2048, 0, False), ("xattn.kv", 128, 1024, 1024, 0, False)],
}
if len(sys.argv) > 3 and sys.argv[3] in SETS:
    SHAPES = SETS[sys.argv[3]]
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
only = sys.argv[2] if len(sys.argv) > 2 else None
if only and only != "all":
    SHAPES = [x for x in SHAPES if x[0] == only]
res = {}
print("impl", impl, "(0 = auto: 2-CTA when eligible, 2 = force 1-CTA, 3 = 1-CTA with split-K for few-row launches)")
for name, M, N, K, act, resid in SHAPES:
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda") * 0.1
    Cm = torch.zeros(M, N, device="cuda", dtype=torch.float32 if resid else torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    def run():
        rc = lib.deer_op_gemm(impl, 0, C.c_void_p(A.data_ptr()), K, C.c_void_p(W.data_ptr()), K, C.c_void_p(Cm.data_ptr()), N, M, N, K,
                              C.c_void_p(bias.data_ptr()), act, None, C.c_void_p(Cm.data_ptr()) if resid else None, int(resid), None, 1, C.c_void_p(st))
        assert rc == 0, lib.deer_last_error(None)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if M > 20000 else 50
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    res[name] = {"ms": round(ms, 4), "tflops": round(tf, 1)}
    print(f"{name:16s} M={M:6d} N={N:5d} K={K:5d}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s", flush=True)
    del A, W, Cm
print(json.dumps(res))
