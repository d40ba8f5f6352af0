// Attention kernels.
//   * attn_simt_kernel<T>: generic softmax(scale*QK^T + bias, masks)V, one CTA per (query row,
//     head, batch).  Handles every flavour on the path: ViT (no mask), Perceiver (no mask),
//     Flamingo gated x-attn (media-time mask classes, helpers.py:205-229) and MPT self-attention
//     (ALiBi key bias + causal + key padding, mosaic_gpt_3b.py:158-219).  Used for the tiny
//     S<=32 LLM attentions (negligible FLOPs) and for everything on the fp32 validation path.
//   * flash_attn_hd64_kernel: bf16 tensor-core (mma.sync m16n8k16) flash attention, head_dim 64,
//     no mask -- the ViT 257x257 and Perceiver 64x320 attentions, which are the only ones with
//     meaningful FLOPs (~4% of a step).
#pragma once
#include "common.cuh"

struct AttnParams {
  const void* Q; const void* K; const void* V; void* O;
  int ldq, ldk, ldv, ldo;        // row strides (elements)
  int Sq, Sk, H, hd;
  float scale;
  int causal;                    // key j visible to query i iff j <= i + (Sk - Sq)
  const float* alibi_slopes;     // [H] or nullptr: bias = slope_h * (j - (Sk-1))
  const int* kv_len;             // [B] number of valid (non-pad) keys, or nullptr
  const uint8_t* qmode;          // [B,Sq] x-attn class: 0 zero output, 1 attend all, 2 uniform; or nullptr
  int B_host;
  const int* n_dev;              // dynamic batch (active episodes) or nullptr
};

template <typename T>
__global__ void __launch_bounds__(128)
attn_simt_kernel(AttnParams p) {
  extern __shared__ float sm[];      // [hd] q, [Sk] scores, [8] red
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nb = p.n_dev ? min(*p.n_dev, p.B_host) : p.B_host;
  if (b >= nb) return;
  float* sq = sm;
  float* sc = sm + p.hd;
  float* red = sc + p.Sk;
  const T* Q = reinterpret_cast<const T*>(p.Q) + (static_cast<size_t>(b) * p.Sq + i) * p.ldq + h * p.hd;
  const T* K = reinterpret_cast<const T*>(p.K) + static_cast<size_t>(b) * p.Sk * p.ldk + h * p.hd;
  const T* V = reinterpret_cast<const T*>(p.V) + static_cast<size_t>(b) * p.Sk * p.ldv + h * p.hd;
  T* O = reinterpret_cast<T*>(p.O) + (static_cast<size_t>(b) * p.Sq + i) * p.ldo + h * p.hd;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int mode = 1;
  if (p.qmode) mode = p.qmode[b * p.Sq + i];
  if (mode == 0) {                                   // text token before any <image>: attention zeroed
    for (int d = tid; d < p.hd; d += 128) O[d] = from_f32<T>(0.0f);
    return;
  }
  for (int d = tid; d < p.hd; d += 128) sq[d] = to_f32<T>(Q[d]);
  __syncthreads();
  const int kvl = p.kv_len ? p.kv_len[b] : p.Sk;
  const int jmax = p.causal ? (i + (p.Sk - p.Sq)) : (p.Sk - 1);
  const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.0f;
  // scores: one warp per key, lanes across head_dim (coalesced)
  for (int j = warp; j < p.Sk; j += 4) {
    float acc = 0.0f;
    const T* kr = K + static_cast<size_t>(j) * p.ldk;
    for (int d = lane; d < p.hd; d += 32) acc += sq[d] * to_f32<T>(kr[d]);
    acc = warp_sum(acc);
    if (lane == 0) {
      float s = acc * p.scale + slope * static_cast<float>(j - (p.Sk - 1));
      if (mode == 2) s = 0.0f;                         // every key masked with the same value -> uniform
      else if (j > jmax || j >= kvl) s = -INFINITY;
      sc[j] = s;
    }
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = tid; j < p.Sk; j += 128) m = fmaxf(m, sc[j]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float l = 0.0f;
  for (int j = tid; j < p.Sk; j += 128) {
    const float e = (sc[j] == -INFINITY) ? 0.0f : expf(sc[j] - m);
    sc[j] = e;
    l += e;
  }
  l = warp_sum(l);
  __syncthreads();
  if (lane == 0) red[4 + warp] = l;
  __syncthreads();
  l = red[4] + red[5] + red[6] + red[7];
  const float inv = 1.0f / l;
  for (int d = tid; d < p.hd; d += 128) {
    float acc = 0.0f;
    for (int j = 0; j < p.Sk; ++j) acc += sc[j] * to_f32<T>(V[static_cast<size_t>(j) * p.ldv + d]);
    O[d] = from_f32<T>(acc * inv);
  }
}

// ---------------------------------------------------------------------------------------------
// attn_small_kernel<T>: the LLM-side attentions (gated x-attn S x 128 latents, MPT self-attention S x S,
// S <= 32).  One CTA (256 threads) per (batch, head): Q, K, V staged ONCE in shared memory as fp32
// (rows padded to hd+1 floats -> conflict-free), scores + softmax + PV entirely on chip.  Same masks /
// bias semantics as attn_simt_kernel.  The FLOPs are negligible; what matters is that K/V are read from
// HBM/L2 once per (b,h) instead of once per query row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
attn_small_kernel(AttnParams p) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int nb = p.n_dev ? min(*p.n_dev, p.B_host) : p.B_host;
  if (b >= nb) return;
  const int hd = p.hd, hp = hd + 1, Sq = p.Sq, Sk = p.Sk, skp = Sk + 1;
  float* sq = sm;                         // [Sq][hp]
  float* sk = sq + Sq * hp;               // [Sk][hp]
  float* sv = sk + Sk * hp;               // [Sk][hp]
  float* ss = sv + Sk * hp;               // [Sq][skp]
  const T* Q = reinterpret_cast<const T*>(p.Q) + static_cast<size_t>(b) * Sq * p.ldq + h * hd;
  const T* K = reinterpret_cast<const T*>(p.K) + static_cast<size_t>(b) * Sk * p.ldk + h * hd;
  const T* V = reinterpret_cast<const T*>(p.V) + static_cast<size_t>(b) * Sk * p.ldv + h * hd;
  T* O = reinterpret_cast<T*>(p.O) + static_cast<size_t>(b) * Sq * p.ldo + h * hd;
  const int tid = threadIdx.x;
  // stage Q, K, V as fp32 (16-byte global loads: 8 bf16 or 4 floats per thread per request)
  constexpr int VE = 16 / sizeof(T);
  const bool vec = (hd % VE == 0) && (p.ldq % VE == 0) && (p.ldk % VE == 0) && (p.ldv % VE == 0);
  if (vec) {
    const int hv = hd / VE;
    for (int idx = tid; idx < Sq * hv; idx += 256) {
      const int i = idx / hv, c = idx % hv;
      T tmp[VE];
      *reinterpret_cast<uint4*>(tmp) = *reinterpret_cast<const uint4*>(Q + static_cast<size_t>(i) * p.ldq + c * VE);
#pragma unroll
      for (int e = 0; e < VE; ++e) sq[i * hp + c * VE + e] = to_f32<T>(tmp[e]);
    }
    for (int idx = tid; idx < Sk * hv; idx += 256) {
      const int j = idx / hv, c = idx % hv;
      T tk[VE], tv[VE];
      *reinterpret_cast<uint4*>(tk) = *reinterpret_cast<const uint4*>(K + static_cast<size_t>(j) * p.ldk + c * VE);
      *reinterpret_cast<uint4*>(tv) = *reinterpret_cast<const uint4*>(V + static_cast<size_t>(j) * p.ldv + c * VE);
#pragma unroll
      for (int e = 0; e < VE; ++e) { sk[j * hp + c * VE + e] = to_f32<T>(tk[e]); sv[j * hp + c * VE + e] = to_f32<T>(tv[e]); }
    }
  } else {
    for (int idx = tid; idx < Sq * hd; idx += 256) { const int i = idx / hd, d = idx % hd; sq[i * hp + d] = to_f32<T>(Q[static_cast<size_t>(i) * p.ldq + d]); }
    for (int idx = tid; idx < Sk * hd; idx += 256) {
      const int j = idx / hd, d = idx % hd;
      sk[j * hp + d] = to_f32<T>(K[static_cast<size_t>(j) * p.ldk + d]);
      sv[j * hp + d] = to_f32<T>(V[static_cast<size_t>(j) * p.ldv + d]);
    }
  }
  __syncthreads();
  const int kvl = p.kv_len ? p.kv_len[b] : Sk;
  const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.0f;
  const bool tiled = (Sq <= 32) && (Sk <= 128) && (hd <= 128);
  if (tiled) {
    // register-tiled scores: thread (g = tid/32, t = tid%32) owns queries 4g..4g+3 x keys t, t+32, t+64, t+96
    // (q values are warp broadcasts, key rows t + 32r hit 32 distinct banks: 8 LDS per 16 FMA instead of 2 per 1)
    const int g = tid >> 5, t = tid & 31;
    const int kpt = (Sk + 31) >> 5;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[a][r] = 0.0f;
    for (int d = 0; d < hd; ++d) {
      float qv[4], kv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qv[a] = (4 * g + a < Sq) ? sq[(4 * g + a) * hp + d] : 0.0f;
#pragma unroll
      for (int r = 0; r < 4; ++r) kv[r] = (r < kpt && t + 32 * r < Sk) ? sk[(t + 32 * r) * hp + d] : 0.0f;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[a][r] = fmaf(qv[a], kv[r], acc[a][r]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = 4 * g + a;
      if (i < Sq) {
        const int mode = p.qmode ? p.qmode[b * Sq + i] : 1;
        const int jmax = p.causal ? (i + (Sk - Sq)) : (Sk - 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int j = t + 32 * r;
          if (j < Sk) {
            float sv_ = acc[a][r] * p.scale + slope * static_cast<float>(j - (Sk - 1));
            if (mode == 2) sv_ = 0.0f;
            else if (j > jmax || j >= kvl) sv_ = -INFINITY;
            ss[i * skp + j] = sv_;
          }
        }
      }
    }
  } else {
  for (int idx = tid; idx < Sq * Sk; idx += 256) {
    const int i = idx / Sk, j = idx % Sk;          // consecutive threads -> consecutive keys (distinct smem rows)
    const int mode = p.qmode ? p.qmode[b * Sq + i] : 1;
    const int jmax = p.causal ? (i + (Sk - Sq)) : (Sk - 1);
    float acc = 0.0f;
    const float* qr = sq + i * hp;
    const float* kr = sk + j * hp;
#pragma unroll 8
    for (int d = 0; d < hd; ++d) acc = fmaf(qr[d], kr[d], acc);
    float s = acc * p.scale + slope * static_cast<float>(j - (Sk - 1));
    if (mode == 2) s = 0.0f;
    else if (j > jmax || j >= kvl) s = -INFINITY;
    ss[i * skp + j] = s;
  }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < Sq; i += 8) {
    float m = -INFINITY;
    for (int j = lane; j < Sk; j += 32) m = fmaxf(m, ss[i * skp + j]);
    m = warp_max(m);
    float l = 0.0f;
    for (int j = lane; j < Sk; j += 32) {
      const float v = ss[i * skp + j];
      const float e = (v == -INFINITY) ? 0.0f : expf(v - m);
      ss[i * skp + j] = e;
      l += e;
    }
    l = warp_sum(l);
    const int mode = p.qmode ? p.qmode[b * Sq + i] : 1;
    const float inv = (mode == 0) ? 0.0f : 1.0f / l;     // mode 0: attention zeroed (helpers.py:223-229)
    for (int j = lane; j < Sk; j += 32) ss[i * skp + j] *= inv;
  }
  __syncthreads();
  if (tiled) {
    // register-tiled PV: thread owns queries 4g..4g+3 x output dims t, t+32, t+64, t+96
    const int g = tid >> 5, t = tid & 31;
    const int dpt = (hd + 31) >> 5;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[a][r] = 0.0f;
    for (int j = 0; j < Sk; ++j) {
      float pv[4], vv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) pv[a] = (4 * g + a < Sq) ? ss[(4 * g + a) * skp + j] : 0.0f;
#pragma unroll
      for (int r = 0; r < 4; ++r) vv[r] = (r < dpt && t + 32 * r < hd) ? sv[j * hp + t + 32 * r] : 0.0f;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[a][r] = fmaf(pv[a], vv[r], acc[a][r]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = 4 * g + a;
      if (i < Sq) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int d = t + 32 * r;
          if (d < hd) O[static_cast<size_t>(i) * p.ldo + d] = from_f32<T>(acc[a][r]);
        }
      }
    }
  } else {
  for (int idx = tid; idx < Sq * hd; idx += 256) {
    const int i = idx / hd, d = idx % hd;              // consecutive threads -> consecutive d (conflict-free V reads)
    float acc = 0.0f;
    const float* pr = ss + i * skp;
    for (int j = 0; j < Sk; ++j) acc = fmaf(pr[j], sv[j * hp + d], acc);
    O[static_cast<size_t>(i) * p.ldo + d] = from_f32<T>(acc);
  }
  }
}
static inline size_t attn_small_smem(int Sq, int Sk, int hd) { return (static_cast<size_t>(Sq + 2 * Sk) * (hd + 1) + static_cast<size_t>(Sq) * (Sk + 1)) * sizeof(float); }

// ---------------------------------------------------------------------------------------------
// flash attention, bf16, head_dim 64, no mask.  grid (ceil(Sq/64), H, B), 128 threads.
// Each warp owns 16 query rows; K/V streamed in 64-key tiles through double-buffered smem.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// [64 rows][64 bf16] tile, 16-byte chunks XOR-swizzled by (row & 7) -> conflict-free ldmatrix
__device__ __forceinline__ int swz(int row, int col) { return row * 64 + ((((col >> 3) ^ (row & 7)) << 3) | (col & 7)); }

__device__ __forceinline__ void flash_load_tile(bf16* s, const bf16* g, int ld, int row0, int nrows_total, int tid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = tid + it * 128;
    const int r = c >> 3, ch = c & 7;
    const bool ok = (row0 + r) < nrows_total;
    const bf16* src = g + static_cast<size_t>(ok ? (row0 + r) : 0) * ld + ch * 8;
    cp_async16(s + swz(r, ch * 8), src, ok);
  }
}

__global__ void __launch_bounds__(128)
flash_attn_hd64_kernel(const bf16* __restrict__ Qg, int ldq, const bf16* __restrict__ Kg, const bf16* __restrict__ Vg, int ldk,
                       bf16* __restrict__ Og, int ldo, int Sq, int Sk, float scale_log2) {
  __shared__ __align__(128) bf16 sQ[64 * 64];
  __shared__ __align__(128) bf16 sK[2][64 * 64];
  __shared__ __align__(128) bf16 sV[2][64 * 64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const bf16* Q = Qg + static_cast<size_t>(b) * Sq * ldq + h * 64;
  const bf16* K = Kg + static_cast<size_t>(b) * Sk * ldk + h * 64;
  const bf16* V = Vg + static_cast<size_t>(b) * Sk * ldk + h * 64;
  bf16* O = Og + static_cast<size_t>(b) * Sq * ldo + h * 64;

  pdl_wait();
  pdl_trigger();
  flash_load_tile(sQ, Q, ldq, q0, Sq, tid);
  flash_load_tile(sK[0], K, ldk, 0, Sk, tid);
  flash_load_tile(sV[0], V, ldk, 0, Sk, tid);
  cp_async_commit();
  const int ntiles = (Sk + 63) / 64;

  uint32_t qa[4][4];
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      flash_load_tile(sK[buf ^ 1], K, ldk, (t + 1) * 64, Sk, tid);
      flash_load_tile(sV[buf ^ 1], V, ldk, (t + 1) * 64, Sk, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = kk * 16 + (lane >> 4) * 8;
        ldsm_x4(qa[kk], sQ + swz(r, c));
      }
    }
    const int kbase = t * 64;
    const int kvalid = min(64, Sk - kbase);                 // valid keys in this tile
    const int np_valid = (kvalid + 15) >> 4;                // 16-key groups that contain a valid key
    const bool warp_active = (q0 + warp * 16) < Sq;         // warps whose 16 query rows are all padding skip the math
    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.0f; }
    if (warp_active) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np < np_valid) {
          uint32_t kb[4];
          const int r = np * 16 + (lane & 7) + (lane >> 4) * 8;
          const int c = kk * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(kb, sK[buf] + swz(r, c));
          mma_bf16_16816(s[2 * np], qa[kk], kb[0], kb[1]);
          mma_bf16_16816(s[2 * np + 1], qa[kk], kb[2], kb[3]);
        }
      }
    }
    }
    // online softmax (log2 domain)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int col = kbase + n * 8 + (lane & 3) * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool valid = (col + (e & 1)) < Sk;
        s[n][e] = valid ? s[n][e] * scale_log2 : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
      mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile has >= 1 valid key
    const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.0f, rs1 = 0.0f;
    uint32_t pa[4][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float p0 = exp2f(s[n][0] - mn0), p1 = exp2f(s[n][1] - mn0);
      const float p2 = exp2f(s[n][2] - mn1), p3 = exp2f(s[n][3] - mn1);
      rs0 += p0 + p1; rs1 += p2 + p3;
      const int kk = n >> 1;
      if ((n & 1) == 0) { pa[kk][0] = pack_bf16x2(p0, p1); pa[kk][1] = pack_bf16x2(p2, p3); }
      else              { pa[kk][2] = pack_bf16x2(p0, p1); pa[kk][3] = pack_bf16x2(p2, p3); }
    }
    l0 = l0 * a0 + rs0; l1 = l1 * a1 + rs1;
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
    // O += P V   (16-key groups past the last valid key contribute p = 0: skipped)
    if (warp_active) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk >= np_valid) break;
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t vb[4];
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(vb, sV[buf] + swz(r, c));
        mma_bf16_16816(o[2 * np], pa[kk], vb[0], vb[1]);
        mma_bf16_16816(o[2 * np + 1], pa[kk], vb[2], vb[3]);
      }
    }
    }
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = n * 8 + (lane & 3) * 2;
    if (r0 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(r0) * ldo + col) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
    if (r1 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(r1) * ldo + col) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
  }
}

// ---------------------------------------------------------------------------------------------
// attn_llm_mma_kernel<HD>: bf16 tensor-core (mma.sync m16n8k16) version of the LLM-side attentions -- gated x-attn
// (Sq<=32 text rows x 128 latents, head_dim 64) and MPT self-attention (S<=32, head_dim 128) -- with the full mask / bias
// semantics of attn_small_kernel (ALiBi key bias, causal, key padding, x-attn media classes, dynamic batch).
// One CTA of 2 warps per (episode, head); warp w owns query rows 16w..16w+15.  Q/K/V are staged ONCE in shared memory as
// bf16 (cp.async, XOR-swizzled rows; 36 KB for x-attn instead of 75 KB fp32 -> 6 CTAs/SM), all Sk<=128 scores of a row
// live in registers, so the softmax is single-pass (no online rescale).  These launches are HBM/L2-latency bound (88 MB of
// q/k/v per launch at B=224): the point is bytes and occupancy, not FLOPs.
// ---------------------------------------------------------------------------------------------
template <int HD> __device__ __forceinline__ int swz_hd(int row, int col) {
  return row * HD + (((((col >> 3) ^ (row & 7)) & 7) | ((col >> 3) & ~7)) << 3 | (col & 7));
}
template <int HD>
__device__ __forceinline__ void llm_load_rows(bf16* s, const bf16* g, int ld, int nrows_pad, int nrows, int tid) {
  constexpr int CH = HD / 8;                                  // 16-byte chunks per row
  for (int c = tid; c < nrows_pad * CH; c += 64) {
    const int r = c / CH, ch = c % CH;
    const bool ok = r < nrows;
    cp_async16(s + swz_hd<HD>(r, ch * 8), g + static_cast<size_t>(ok ? r : 0) * ld + ch * 8, ok);
  }
}

template <int HD>
__global__ void __launch_bounds__(64)
attn_llm_mma_kernel(AttnParams p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  pdl_wait();
  pdl_trigger();
  const int h = blockIdx.x, b = blockIdx.y;
  const int nb = p.n_dev ? min(*p.n_dev, p.B_host) : p.B_host;
  if (b >= nb) return;
  const int Sq = p.Sq, Sk = p.Sk;
  const int skp = (Sk + 15) & ~15;                            // keys padded to the 16-key MMA granule (zero rows)
  bf16* sQ = reinterpret_cast<bf16*>(smraw);                  // [32][HD]
  bf16* sK = sQ + 32 * HD;                                    // [skp][HD]
  bf16* sV = sK + skp * HD;                                   // [skp][HD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bf16* Q = reinterpret_cast<const bf16*>(p.Q) + static_cast<size_t>(b) * Sq * p.ldq + h * HD;
  const bf16* K = reinterpret_cast<const bf16*>(p.K) + static_cast<size_t>(b) * Sk * p.ldk + h * HD;
  const bf16* V = reinterpret_cast<const bf16*>(p.V) + static_cast<size_t>(b) * Sk * p.ldv + h * HD;
  bf16* O = reinterpret_cast<bf16*>(p.O) + static_cast<size_t>(b) * Sq * p.ldo + h * HD;
  llm_load_rows<HD>(sQ, Q, p.ldq, 32, Sq, tid);
  llm_load_rows<HD>(sK, K, p.ldk, skp, Sk, tid);
  cp_async_commit();
  llm_load_rows<HD>(sV, V, p.ldv, skp, Sk, tid);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  const bool active = warp * 16 < Sq;
  const int ngrp = skp >> 4;                                  // 16-key groups (<= 8)
  // ---- S = Q K^T: 16 rows x skp keys per warp ----
  float sc[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) { sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.0f; }
  if (active) {
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t qa[4];
      ldsm_x4(qa, sQ + swz_hd<HD>(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 16 + (lane >> 4) * 8));
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        if (np < ngrp) {
          uint32_t kb[4];
          ldsm_x4(kb, sK + swz_hd<HD>(np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 16 + ((lane >> 3) & 1) * 8));
          mma_bf16_16816(sc[2 * np], qa, kb[0], kb[1]);
          mma_bf16_16816(sc[2 * np + 1], qa, kb[2], kb[3]);
        }
      }
    }
  }
  // ---- scale, bias, masks, single-pass softmax (rows r0 = 16w + lane/4 and r1 = r0 + 8) ----
  const int r0 = warp * 16 + (lane >> 2), r1 = r0 + 8;
  const int kvl = p.kv_len ? p.kv_len[b] : Sk;
  const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.0f;
  const int mode0 = (p.qmode && r0 < Sq) ? p.qmode[b * Sq + r0] : 1;
  const int mode1 = (p.qmode && r1 < Sq) ? p.qmode[b * Sq + r1] : 1;
  const int jmax0 = p.causal ? (r0 + (Sk - Sq)) : (Sk - 1), jmax1 = p.causal ? (r1 + (Sk - Sq)) : (Sk - 1);
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = n * 8 + (lane & 3) * 2 + (e & 1);
      const int mode = (e < 2) ? mode0 : mode1;
      const int jmax = (e < 2) ? jmax0 : jmax1;
      float v = sc[n][e] * p.scale + slope * static_cast<float>(j - (Sk - 1));
      if (j >= Sk) v = -INFINITY;
      else if (mode == 2) v = 0.0f;
      else if (j > jmax || j >= kvl) v = -INFINITY;
      sc[n][e] = v;
    }
    mx0 = fmaxf(mx0, fmaxf(sc[n][0], sc[n][1]));
    mx1 = fmaxf(mx1, fmaxf(sc[n][2], sc[n][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.0f, l1 = 0.0f;
  uint32_t pa[8][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    const float p0 = (sc[n][0] == -INFINITY) ? 0.0f : __expf(sc[n][0] - mx0), p1 = (sc[n][1] == -INFINITY) ? 0.0f : __expf(sc[n][1] - mx0);
    const float p2 = (sc[n][2] == -INFINITY) ? 0.0f : __expf(sc[n][2] - mx1), p3 = (sc[n][3] == -INFINITY) ? 0.0f : __expf(sc[n][3] - mx1);
    l0 += p0 + p1; l1 += p2 + p3;
    const int kk = n >> 1;
    if ((n & 1) == 0) { pa[kk][0] = pack_bf16x2(p0, p1); pa[kk][1] = pack_bf16x2(p2, p3); }
    else              { pa[kk][2] = pack_bf16x2(p0, p1); pa[kk][3] = pack_bf16x2(p2, p3); }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = (mode0 == 0) ? 0.0f : 1.0f / l0, i1 = (mode1 == 0) ? 0.0f : 1.0f / l1;   // mode 0: attention zeroed (helpers.py:223-229)
  cp_async_wait<0>();
  __syncthreads();
  // ---- O = P V ----
  float o[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f; }
  if (active) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      if (kk < ngrp) {
#pragma unroll
        for (int np = 0; np < HD / 16; ++np) {
          uint32_t vb[4];
          ldsm_x4_t(vb, sV + swz_hd<HD>(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 16 + (lane >> 4) * 8));
          mma_bf16_16816(o[2 * np], pa[kk], vb[0], vb[1]);
          mma_bf16_16816(o[2 * np + 1], pa[kk], vb[2], vb[3]);
        }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    const int col = n * 8 + (lane & 3) * 2;
    if (r0 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(r0) * p.ldo + col) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
    if (r1 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(r1) * p.ldo + col) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
  }
}
static inline size_t attn_llm_mma_smem(int Sk, int hd) { return static_cast<size_t>(32 + 2 * ((Sk + 15) & ~15)) * hd * sizeof(bf16); }
