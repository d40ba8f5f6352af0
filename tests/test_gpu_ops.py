"""Single-kernel parity through the C ABI (deer_op_gemm / deer_op_attention) against torch fp32."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_gemm(A, W, bias, act, scale, resid):
    v = A.float() @ W.float().t()
    if bias is not None:
        v = v + bias
    if act == 1:
        v = torch.nn.functional.gelu(v)
    elif act == 2:
        v = v * torch.sigmoid(1.702 * v)
    elif act == 3:
        v = torch.relu(v)
    if scale is not None:
        v = v * scale
    if resid is not None:
        v = v + resid
    return v


SHAPES = [(128, 256, 64), (128, 64, 128), (200, 512, 1024), (1000, 3072, 1024), (1028, 1024, 4096),
          (32, 6144, 2048), (7, 4096, 3072), (300, 200, 136), (5000, 4096, 1024), (1, 6, 512), (38000, 1024, 640),
          (9000, 1000, 520), (256 * 74, 256, 64), (19999, 2048, 2048)]     # the last rows take the 2-CTA (cta_group::2) kernel


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_plain(impl, M, N, K):
    from deer_vla_b200.engine import op_gemm
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    out = op_gemm(A, W, out_f32=True, impl=impl)
    torch.cuda.synchronize()
    ref = _ref_gemm(A, W, None, 0, None, None)
    err = (out - ref).abs().max().item()
    tol = 2e-3 * math.sqrt(K / 64) * 0.5
    assert err <= tol, f"{impl} {M}x{N}x{K}: max abs err {err} > {tol}"


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("M", [777, 5003])          # 5003 rows -> 2-CTA kernel with the smem-staged coalesced epilogue
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_epilogues(impl, act, M):
    from deer_vla_b200.engine import op_gemm
    N, K = 1000, 512
    g = torch.Generator(device="cuda").manual_seed(act)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    scale = torch.tensor([0.4621], device="cuda")
    resid = torch.randn(M, N, device="cuda", generator=g)
    # bf16 output, bias + act
    out = op_gemm(A, W, bias=bias, act=act, impl=impl)
    ref = _ref_gemm(A, W, bias, act, None, None)
    assert (out.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    # fp32 residual stream update with the tanh(gate) scale
    out = op_gemm(A, W, bias=bias, act=act, scale=scale, resid=resid, impl=impl)
    ref = _ref_gemm(A, W, bias, act, scale, resid)
    assert (out - ref).abs().max().item() <= 2e-3


@pytest.mark.parametrize("M", [640, 9000])            # 1-CTA and 2-CTA tcgen05 kernels
def test_gemm_epilogue_gelu_is_exact_erf_gelu(M):
    """The tensor-core epilogue's GELU is nn.GELU() (erf form, the reference's FeedForward / MPT MLP activation) to fp32
    accuracy, not the tanh approximation: identity weights feed bf16-exact pre-activations on [-9, 9] (dense around 0 and in
    the negative tail, where 1 + erf cancels) through the epilogue with fp32 output; |err| <= 2e-6 against float64."""
    from deer_vla_b200.engine import op_gemm
    K = N = 256
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.cat([torch.linspace(-9, 9, M * N // 2, device="cuda"), torch.randn(M * N - M * N // 2, device="cuda", generator=g) * 2.0])
    x = x[torch.randperm(M * N, device="cuda", generator=g)].view(M, N).bfloat16()
    W = torch.eye(N, K, device="cuda").bfloat16()
    out = op_gemm(x.contiguous(), W, act=1, resid=torch.zeros(M, N, device="cuda"), impl="tcgen05")
    torch.cuda.synchronize()
    xd = x.double()
    ref = 0.5 * xd * (1.0 + torch.erf(xd / math.sqrt(2.0)))
    err = (out.double() - ref).abs()
    assert err.max().item() <= 2e-6, (err.max().item(), xd.flatten()[err.argmax()].item())
    tail = (xd < -3.0) & (xd > -6.0)                    # relative accuracy where the value itself is 4e-3 .. 6e-9 (beyond -6 it is clamped)
    rel = (err[tail] / ref[tail].abs().clamp_min(1e-12)).max().item()
    assert rel <= 1e-3, rel


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
def test_gemm_dynamic_m_and_strides(impl):
    from deer_vla_b200.engine import op_gemm
    g = torch.Generator(device="cuda").manual_seed(11)
    big = (torch.randn(640, 2048, device="cuda", generator=g) * 0.5).bfloat16()
    A = big[:, 1024:1536]                       # lda 2048, column offset: the head MLP reads one half of y1 like this
    W = (torch.randn(512, 512, device="cuda", generator=g) * 0.05).bfloat16()
    m_dev = torch.tensor([9], dtype=torch.int32, device="cuda")
    out = op_gemm(A, W, out_f32=True, impl=impl, m_dev=m_dev, m_mul=32)     # only 9*32 = 288 rows are live
    ref = _ref_gemm(A, W, None, 0, None, None)
    assert (out[:288] - ref[:288]).abs().max().item() <= 2e-3
    assert out[288:].abs().max().item() == 0.0                              # rows past the device-side count untouched


def test_gemm_fp32_simt():
    from deer_vla_b200.engine import op_gemm
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(130, 300, device="cuda", generator=g)
    W = torch.randn(70, 300, device="cuda", generator=g) * 0.1
    out = op_gemm(A, W, impl="simt")
    assert (out - A @ W.t()).abs().max().item() <= 1e-4


def _ref_attn(q, k, v, scale):
    s = torch.einsum("bihd,bjhd->bhij", q.float(), k.float()) * scale
    return torch.einsum("bhij,bjhd->bihd", s.softmax(-1), v.float())


@pytest.mark.parametrize("B,H,Sq,Sk", [(3, 4, 257, 257), (2, 8, 64, 320), (1, 2, 17, 17), (2, 16, 130, 70), (40, 16, 257, 257),
                                       (5, 3, 256, 272), (2, 2, 5, 9), (3, 2, 300, 200)])
@pytest.mark.parametrize("impl", ["tc", "flash", "simt"])
def test_attention_hd64(impl, B, H, Sq, Sk):
    from deer_vla_b200.engine import op_attention
    g = torch.Generator(device="cuda").manual_seed(B + H + Sq)
    q = torch.randn(B, Sq, H, 64, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Sk, H, 64, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Sk, H, 64, device="cuda", generator=g).bfloat16()
    out = op_attention(q, k, v, 0.125, impl=impl)
    torch.cuda.synchronize()
    ref = _ref_attn(q, k, v, 0.125)
    assert (out.float() - ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("kind", ["peaked", "ramp", "late_spike", "tail_spike"])
def test_attention_tc_large_logits(kind):
    """The chunk-pipelined softmax of the tcgen05 attention kernels takes probabilities relative to a running reference instead
    of the full-row max (v5: the reference moves, with an exact power-of-two rescale of what is already stored, only when a chunk
    max exceeds it by 2^60; v4: clamped to chunk max - 30k at stage k).  Rows whose late keys beat the early ones by far more
    than that must still come out as the exact softmax."""
    from deer_vla_b200.engine import op_attention
    B, H, S = 4, 4, 257
    g = torch.Generator(device="cuda").manual_seed(11)
    q = torch.randn(B, S, H, 64, device="cuda", generator=g)
    k = torch.randn(B, S, H, 64, device="cuda", generator=g)
    v = torch.randn(B, S, H, 64, device="cuda", generator=g).bfloat16()
    if kind == "peaked":                 # logits ~ N(0, 36^2): nearly one-hot rows, winners anywhere
        q, k = q * 6.0, k * 6.0
    else:
        u = torch.nn.functional.normalize(torch.randn(64, device="cuda", generator=g), dim=0)
        q = q * 0.3 + 16.0 * u           # every query has a large component along u ...
        pos = torch.arange(S, device="cuda", dtype=torch.float32)
        if kind == "ramp":               # ... and key j's component along u grows with j: logit(j) ~ 2*16*j/S*12 -> +100 nats end to end
            amp = 50.0 * pos / S
        elif kind == "late_spike":       # one key in the last chunk of each half towers over everything before it
            amp = torch.zeros(S, device="cuda")
            amp[120], amp[250] = 40.0, 45.0
        else:                            # the 257th key (handled on CUDA cores) is the winner
            amp = torch.zeros(S, device="cuda")
            amp[256] = 40.0
        k = k * 0.3 + amp.view(1, S, 1, 1) * u
    q, k = q.bfloat16(), k.bfloat16()
    out = op_attention(q, k, v, 0.125, impl="tc")
    torch.cuda.synchronize()
    ref = _ref_attn(q, k, v, 0.125)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 3e-2


def test_attention_fp32_simt():
    from deer_vla_b200.engine import op_attention
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(2, 33, 3, 128, device="cuda", generator=g)
    k = torch.randn(2, 50, 3, 128, device="cuda", generator=g)
    v = torch.randn(2, 50, 3, 128, device="cuda", generator=g)
    out = op_attention(q, k, v, 1 / math.sqrt(128), impl="simt")
    assert (out - _ref_attn(q, k, v, 1 / math.sqrt(128))).abs().max().item() <= 1e-4


def test_pv_probe_tmem_a_operand_and_mn_major_b():
    """Regression for the two tcgen05 encodings the attention kernel relies on: P read from tensor memory (two bf16 per
    32-bit column, lane = row) and V read in its natural [key][d] layout as an MN-major B operand (SBO 1024 B, 2048 B / 16 keys)."""
    import ctypes as C
    from deer_vla_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    for K in (16, 64, 256):
        P = torch.randn(128, K, device="cuda", generator=g).bfloat16()
        V = torch.randn(K, 64, device="cuda", generator=g).bfloat16()
        O = torch.zeros(128, 64, device="cuda")
        assert lib.deer_op_pv_probe(C.c_void_p(P.data_ptr()), C.c_void_p(V.data_ptr()), C.c_void_p(O.data_ptr()), K, 1024, 1024, 2048, None) == 0
        torch.cuda.synchronize()
        assert (O - P.float() @ V.float()).abs().max().item() <= 1e-3 * K ** 0.5


def _masked_ref(q, k, v, scale, causal, slopes, kv_len, qmode):
    """fp32 torch reference with the kernels' semantics (masked keys excluded; mode 0 -> zeros; mode 2 -> uniform over all keys)."""
    B, Sq, H, hd = q.shape
    Sk = k.shape[1]
    s = torch.einsum("bihd,bjhd->bhij", q.float(), k.float()) * scale
    j = torch.arange(Sk, device=q.device)
    i = torch.arange(Sq, device=q.device)
    if slopes is not None:
        s = s + slopes.view(1, H, 1, 1) * (j - (Sk - 1)).view(1, 1, 1, Sk).float()
    dead = torch.zeros(B, 1, Sq, Sk, dtype=torch.bool, device=q.device)
    if causal:
        dead |= (j.view(1, Sk) > (i.view(Sq, 1) + (Sk - Sq))).view(1, 1, Sq, Sk)
    if kv_len is not None:
        dead |= (j.view(1, 1, 1, Sk) >= kv_len.view(B, 1, 1, 1))
    s = s.masked_fill(dead, float("-inf"))
    if qmode is not None:
        s = torch.where((qmode == 2).view(B, 1, Sq, 1), torch.zeros_like(s), s)
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhij,bjhd->bihd", p, v.float())
    if qmode is not None:
        o = o * (qmode != 0).view(B, Sq, 1, 1).float()
    return o


@pytest.mark.parametrize("case", [
    dict(B=5, H=16, Sq=32, Sk=32, hd=128, causal=True, alibi=True, ragged=True, modes=False),     # MPT-1B self-attention
    dict(B=3, H=32, Sq=32, Sk=32, hd=128, causal=True, alibi=True, ragged=False, modes=False),    # MPT-7B self-attention
    dict(B=4, H=8, Sq=32, Sk=128, hd=64, causal=False, alibi=False, ragged=False, modes=True),    # gated x-attn, media classes
    dict(B=2, H=2, Sq=13, Sk=13, hd=128, causal=True, alibi=True, ragged=True, modes=False),      # short instruction
    dict(B=3, H=3, Sq=7, Sk=100, hd=64, causal=False, alibi=False, ragged=False, modes=True),     # ragged key count
    dict(B=2, H=4, Sq=17, Sk=128, hd=64, causal=False, alibi=False, ragged=False, modes=False),
    dict(B=6, H=2, Sq=32, Sk=32, hd=64, causal=True, alibi=True, ragged=True, modes=False),
])
def test_llm_attention_mma_vs_simt_and_torch(case):
    from deer_vla_b200.engine import op_attention_masked
    B, H, Sq, Sk, hd = case["B"], case["H"], case["Sq"], case["Sk"], case["hd"]
    g = torch.Generator(device="cuda").manual_seed(B * 131 + Sq)
    q = torch.randn(B, Sq, H, hd, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Sk, H, hd, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Sk, H, hd, device="cuda", generator=g).bfloat16()
    scale = hd ** -0.5
    slopes = (0.5 ** torch.arange(1, H + 1, device="cuda").float() * 4.0 / H).contiguous() if case["alibi"] else None
    kv_len = torch.randint(max(1, Sk // 2), Sk + 1, (B,), device="cuda", generator=g, dtype=torch.int32) if case["ragged"] else None
    qmode = torch.randint(0, 3, (B, Sq), device="cuda", generator=g, dtype=torch.int32).to(torch.uint8) if case["modes"] else None
    n_dev = torch.tensor([B - 1], device="cuda", dtype=torch.int32)                      # last episode inactive: its rows stay untouched
    ref = _masked_ref(q, k, v, scale, case["causal"], slopes, kv_len, qmode)
    outs = {}
    for impl in ("simt", "mma", "auto"):
        outs[impl] = op_attention_masked(q, k, v, scale, impl=impl, causal=case["causal"], alibi_slopes=slopes, kv_len=kv_len,
                                         qmode=qmode, n_dev=n_dev).float()
        assert torch.all(outs[impl][B - 1] == 0), impl                                     # dynamic batch honoured
        err = (outs[impl][: B - 1] - ref[: B - 1]).abs().max().item()
        assert err <= 2e-2, (impl, err)                                                    # bf16 output rounding of O(1) values
    assert (outs["mma"] - outs["simt"]).abs().max().item() <= 2e-2


@pytest.mark.parametrize("M,N,K", [(5003, 3072, 1024), (9000, 512, 2048), (14392, 4096, 512)])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_tma_store_epilogue_is_bit_identical(M, N, K, act):
    """2-CTA kernel, bf16 output without residual: the TMA-store epilogue (row-per-lane, 64-byte-swizzled tile, one
    cp.async.bulk.tensor per warp) against the transposing per-thread-store epilogue (DEER_GEMM_TMA_STORE=0) on the same inputs:
    BIT-identical (same bias / activation / scale / rounding), the ragged last row tile (M % 32 != 0) clipped by the tensor map,
    with and without bias."""
    import os
    from deer_vla_b200.engine import op_gemm
    g = torch.Generator(device="cuda").manual_seed(M + act)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    outs = {}
    for flag in ("1", "0"):
        os.environ["DEER_GEMM_TMA_STORE"] = flag
        try:
            outs[flag] = (op_gemm(A, W, bias=bias, act=act, impl="tcgen05").clone(), op_gemm(A, W, act=act, impl="tcgen05").clone())
            torch.cuda.synchronize()
        finally:
            os.environ.pop("DEER_GEMM_TMA_STORE", None)
    for a, b in zip(outs["1"], outs["0"]):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    ref = _ref_gemm(A, W, bias, act, None, None)
    assert (outs["1"][0].float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(32, 2048, 8192), (32, 2048, 2048), (1, 512, 2048), (128, 1024, 4096), (97, 520, 2000), (130, 256, 2048)])
@pytest.mark.parametrize("mode", ["bf16", "f32_resid", "bias_gelu"])
def test_gemm_split_k_for_few_row_launches(M, N, K, mode):
    """Few tiles on 148 SMs (1-4 episode steps): the OPT-IN split-K kernel (fp32 partial tiles in a workspace, the tile's last CTA
    reduces in split order and runs the ordinary epilogue) against the one-CTA-per-tile launch of the same 128x64 kernel.  The two differ
    only in the fp32 summation order over K (one TMEM accumulation vs `splits` partial sums), so they agree to fp32 rounding of the
    accumulator -- bar 2e-5 of the output scale before the bf16 store, one bf16 ulp after it -- and the split-K result is
    BIT-reproducible run to run (the reduction order does not depend on which CTA arrives last).  Ragged N / K and a second row
    tile (M = 130) included; the counters re-arm themselves (three launches in a row)."""
    from deer_vla_b200.engine import op_gemm
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    resid = torch.randn(M, N, device="cuda", generator=g)
    kw = {"bf16": dict(), "f32_resid": dict(bias=bias, resid=resid), "bias_gelu": dict(bias=bias, act=1)}[mode]
    runs = [op_gemm(A, W, impl="tcgen05_splitk", **kw).clone() for _ in range(3)]
    base = op_gemm(A, W, impl="tcgen05_1cta_m128", **kw).clone()
    torch.cuda.synchronize()
    for r in runs[1:]:
        assert torch.equal(r.view(torch.int32 if r.dtype == torch.float32 else torch.int16), runs[0].view(torch.int32 if r.dtype == torch.float32 else torch.int16))
    ref = _ref_gemm(A, W, kw.get("bias"), kw.get("act", 0), None, kw.get("resid"))
    scale = max(1.0, ref.abs().max().item())
    if runs[0].dtype == torch.float32:
        assert (runs[0] - base).abs().max().item() <= 2e-5 * scale
    else:
        assert (runs[0].float() - base.float()).abs().max().item() <= 2.0 ** -7 * scale      # one bf16 ulp at the output scale
    assert (runs[0].float() - ref).abs().max().item() <= 1e-2 * scale


@pytest.mark.parametrize("M,N,K", [(514, 1024, 4096), (514, 1024, 1024), (32, 2048, 8192), (1, 512, 2048), (97, 520, 1000), (896, 512, 2048), (130, 264, 72)])
@pytest.mark.parametrize("mode", ["bf16", "f32_resid", "bias_gelu"])
def test_gemm_half_height_tiles_are_bit_identical(M, N, K, mode):
    """64 x 64 tiles (tcgen05.mma M = 64, accumulator row r in TMEM lane (r % 16) + 32 * (r / 16)) for launches whose 128-row tiles
    leave most SMs idle (a 1-episode ViT: M = 514; the LLM at M = 32): K is accumulated in the same order as in the 128-row
    kernels, so the outputs are BIT-identical to them -- the batch-invariance of the backbone is kept.  Ragged M / N / K included."""
    from deer_vla_b200.engine import op_gemm
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    resid = torch.randn(M, N, device="cuda", generator=g)
    kw = {"bf16": dict(), "f32_resid": dict(bias=bias, resid=resid), "bias_gelu": dict(bias=bias, act=1)}[mode]
    half = op_gemm(A, W, impl="tcgen05_1cta", **kw).clone()          # the cost model picks 64 x 64 tiles for every shape above
    full = op_gemm(A, W, impl="tcgen05_1cta_m128", **kw).clone()
    torch.cuda.synchronize()
    it = torch.int32 if half.dtype == torch.float32 else torch.int16
    assert torch.equal(half.view(it), full.view(it))
    ref = _ref_gemm(A, W, kw.get("bias"), kw.get("act", 0), None, kw.get("resid"))
    assert (half.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
