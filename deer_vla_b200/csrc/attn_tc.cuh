// tcgen05 attention building blocks.
//  * pv_probe_kernel: a probe used while bringing the kernel up (and kept as a regression test): O[128,64] = P[128,K] @ V[K,64]
//    with P written to TENSOR MEMORY by the threads (tcgen05.st, two bf16 per 32-bit column) and consumed as the A operand
//    (tcgen05.mma "TS" form), and V consumed straight from its natural [key][d] layout as an MN-major B operand.
#pragma once
#include "gemm.cuh"

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// generic smem descriptor: start address, LBO, SBO (bytes), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// idesc with selectable operand majors (bit 15: A MN-major, bit 16: B MN-major)
__host__ __device__ constexpr uint32_t make_idesc_bf16_major(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// P [128, K] bf16 row-major (gmem), V [K, 64] bf16 row-major via TMA (box 64 x K rows, SW128), O [128, 64] fp32.  K % 16 == 0, K <= 256.
__global__ void __launch_bounds__(128, 1)
pv_probe_kernel(const __grid_constant__ CUtensorMap tmV, const bf16* __restrict__ P, float* __restrict__ O, int K,
                uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t kstep_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bar_v = reinterpret_cast<uint64_t*>(smem + 256 * 128);
  uint64_t* bar_o = bar_v + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_v, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_p = tmem_base;            // P: K/2 columns
  const uint32_t tmem_o = tmem_base + 128;      // O: 64 columns
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_v, static_cast<uint32_t>(K) * 128u);
    tma_load_2d(smem, &tmV, bar_v, 0, 0);
  }
  // every thread owns one row of P: pack two bf16 per 32-bit TMEM column
  const int row = warp * 32 + lane;
  const uint32_t* prow = reinterpret_cast<const uint32_t*>(P + static_cast<size_t>(row) * K);
  for (int c = 0; c < K / 2; c += 8) {
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = prow[c + j];
    tmem_st_32x32b_x8(tmem_p + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c), r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_wait(bar_v, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16_major(128, 64, 0, 1);      // A K-major (TMEM), B MN-major
    const uint32_t v_addr = smem_u32(smem);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t bdesc = make_desc(v_addr + static_cast<uint32_t>(k) * kstep_bytes, lbo_bytes, sbo_bytes);
      umma_bf16_ts(tmem_o, tmem_p + static_cast<uint32_t>(k * 8), bdesc, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(bar_o);
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    float v[32];
    tmem_ld_32x32(tmem_o + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) O[static_cast<size_t>(row) * 64 + c * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// attn_tc_kernel: softmax(scale * Q K^T) V on the 5th-gen tensor cores, head_dim 64, no mask, Sk <= 272 keys.
// (the ViT-L/14 attention: 257 x 257 per (image, head); ~3 % of a step's FLOPs but 18 % of its time on mma.sync)
//
// Persistent CTAs, one work item = one (batch, head): K and V are TMA-loaded ONCE per item in their natural [key][d]
// layout (two 136-row boxes each, 128B swizzle) and every 128-row query tile of the item is processed against them.
//   S = Q K^T : tcgen05.mma  M=128, N=256 + N=16, K=64   (A = Q smem K-major, B = K smem K-major)      -> TMEM cols [0,272)
//   softmax   : 4 warps, one query row per thread (TMEM lane): pass 1 row max, pass 2 exp2 + row sum; P packed as bf16 pairs
//               and written back to TENSOR MEMORY (tcgen05.st)                                          -> TMEM cols [272,408)
//   O = P V   : tcgen05.mma  M=128, N=64, K=272          (A = P in TMEM, B = V smem MN-major)           -> TMEM cols [408,472)
//   epilogue  : O / rowsum -> bf16 -> global
// The MMA thread issues PV(n) and then immediately QK(n+1), so the next tile's scores are computed while the softmax warps
// drain O(n).  The 1-8 query rows that do not fill a 128-row tile (row 256 of a 257-token image) are done by a spare warp on
// CUDA cores against the K/V already in shared memory.
// warps: 0 TMA producer, 1 MMA issuer + TMEM owner, 2-9 softmax/epilogue, 10 leftover rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}

struct AttnTc {
  static constexpr int SKP = 272;                        // padded key count (17 x 16)
  static constexpr int KV_BYTES = SKP * 128;             // 34816 per K or V buffer
  static constexpr int HALF_ROWS = 136, HALF_BYTES = HALF_ROWS * 128;
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int OFF_K = 0, OFF_V = 2 * KV_BYTES, OFF_Q = 4 * KV_BYTES, OFF_P = OFF_Q + 2 * Q_BYTES;   // OFF_P: fp32 scratch of the side warp
  static constexpr int OFF_X = OFF_P + SKP * 4 + 64;         // partial row max / sum exchange between the two warps of a lane quarter
  static constexpr int OFF_BAR = OFF_X + 4 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int THREADS = 11 * 32;
  static constexpr uint32_t COL_S = 0, COL_P = 272, COL_O = 408;
};

__global__ void __launch_bounds__(AttnTc::THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const bf16* __restrict__ Qg, int ldq, bf16* __restrict__ Og, int ldo, int B, int H, int Sq, int Sk, float scale_log2) {
  using A = AttnTc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A::OFF_BAR);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]
  uint64_t* q_full = bars + 4;       // [2]
  uint64_t* q_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* o_full = bars + 10;
  uint64_t* o_empty = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = B * H;
  const int n_full = Sq / 128, left = Sq % 128;
  const int n_sub = n_full + (left > 8 ? 1 : 0);          // query tiles done on the tensor cores
  const int side_rows = (left > 8) ? 0 : left;            // rows done by the spare warp
  const int side_row0 = n_full * 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);      // MMA commit (last PV of the item) + spare warp
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int kb = 0; uint32_t kph = 0; int qb = 0; uint32_t qph = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int b = it / H, h = it % H;
        mbar_wait(&kv_empty[kb], kph ^ 1u);
        mbar_arrive_expect_tx(&kv_full[kb], 2 * A::KV_BYTES);
        uint8_t* sk = smem + A::OFF_K + kb * A::KV_BYTES;
        uint8_t* sv = smem + A::OFF_V + kb * A::KV_BYTES;
        tma_load_2d(sk, &tmK, &kv_full[kb], h * 64, b * Sk);
        tma_load_2d(sk + A::HALF_BYTES, &tmK, &kv_full[kb], h * 64, b * Sk + A::HALF_ROWS);
        tma_load_2d(sv, &tmV, &kv_full[kb], h * 64, b * Sk);
        tma_load_2d(sv + A::HALF_BYTES, &tmV, &kv_full[kb], h * 64, b * Sk + A::HALF_ROWS);
        if (++kb == 2) { kb = 0; kph ^= 1u; }
        for (int t = 0; t < n_sub; ++t) {
          mbar_wait(&q_empty[qb], qph ^ 1u);
          mbar_arrive_expect_tx(&q_full[qb], A::Q_BYTES);
          tma_load_2d(smem + A::OFF_Q + qb * A::Q_BYTES, &tmQ, &q_full[qb], h * 64, b * Sq + t * 128);
          if (++qb == 2) { qb = 0; qph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_s256 = make_idesc_bf16_major(128, 256, 0, 0);
      constexpr uint32_t idesc_s16 = make_idesc_bf16_major(128, 16, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(128, 64, 0, 1);     // B = V, MN-major
      int kb = 0; uint32_t kph = 0; int qb = 0; uint32_t qph = 0;
      uint32_t pph = 0, oeph = 0;
      long long n = 0;                                    // running query-tile counter
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        mbar_wait(&kv_full[kb], kph);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + A::OFF_K + kb * A::KV_BYTES);
        const uint32_t v_addr = smem_u32(smem + A::OFF_V + kb * A::KV_BYTES);
        for (int t = 0; t < n_sub; ++t, ++n) {
          mbar_wait(&q_full[qb], qph);
          tc_fence_after();
          const uint32_t q_addr = smem_u32(smem + A::OFF_Q + qb * A::Q_BYTES);
          // S = Q K^T   (the S columns were released when p_full of the previous tile was observed below)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = make_desc_sw128(q_addr) + static_cast<uint64_t>(2 * k);
            umma_bf16(tmem_base + A::COL_S, adesc, make_desc_sw128(k_addr) + static_cast<uint64_t>(2 * k), idesc_s256, k != 0 ? 1u : 0u);
            umma_bf16(tmem_base + A::COL_S + 256, adesc, make_desc_sw128(k_addr + 256 * 128) + static_cast<uint64_t>(2 * k), idesc_s16, k != 0 ? 1u : 0u);
          }
          umma_commit(s_full);
          umma_commit(&q_empty[qb]);
          if (++qb == 2) { qb = 0; qph ^= 1u; }
          // O = P V
          mbar_wait(p_full, pph); pph ^= 1u;
          if (n > 0) { mbar_wait(o_empty, oeph); oeph ^= 1u; }
          tc_fence_after();
#pragma unroll 1
          for (int k = 0; k < A::SKP / 16; ++k)
            umma_bf16_ts(tmem_base + A::COL_O, tmem_base + A::COL_P + static_cast<uint32_t>(8 * k),
                         make_desc(v_addr + static_cast<uint32_t>(k) * 2048u, 1024u, 1024u), idesc_pv, k != 0 ? 1u : 0u);
          umma_commit(o_full);
          if (t == n_sub - 1) umma_commit(&kv_empty[kb]);
        }
        if (n_sub == 0) mbar_arrive(&kv_empty[kb]);
        if (++kb == 2) { kb = 0; kph ^= 1u; }
      }
    }
  } else if (warp < 10) {
    // ===== softmax + epilogue: 8 warps.  A query row lives in one TMEM lane; the two warps that share a lane quarter split the
    //       row's key columns (half 0: keys [0,128), half 1: keys [128,272)) and exchange their partial max / sum through smem.
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rloc = lane_grp * 32 + lane;                                 // row inside the 128-row tile
    const uint32_t lane_off = static_cast<uint32_t>(lane_grp * 32) << 16;
    float* xmax = reinterpret_cast<float*>(smem + A::OFF_X);               // [2][128]
    float* xsum = xmax + 256;                                              // [2][128]
    const int c_begin = half * 4, c_end = c_begin + 4;                     // 32-key chunks of this half (the 16-key tail belongs to half 1)
    uint32_t sph = 0, oph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int b = it / H, h = it % H;
      for (int t = 0; t < n_sub; ++t) {
        mbar_wait(s_full, sph); sph ^= 1u;
        tc_fence_after();
        const uint32_t ts = tmem_base + lane_off + A::COL_S;
        // ---- pass 1: partial row max (TMEM loads software-pipelined: the next chunk is in flight while this one is reduced)
        float m = -INFINITY;
        float va[32], vb[32];
        tmem_ld_32x32(ts + static_cast<uint32_t>(c_begin * 32), va);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = c_begin + cc;
          float (&cur)[32] = (cc & 1) ? vb : va;
          float (&nxt)[32] = (cc & 1) ? va : vb;
          if (cc < 3) tmem_ld_32x32(ts + static_cast<uint32_t>((c + 1) * 32), nxt);
          if (c * 32 + 32 <= Sk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, cur[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (c * 32 + j < Sk) m = fmaxf(m, cur[j]);
          }
          if (cc < 3) tmem_ld_wait();
        }
        float vt[16];
        if (half == 1) {
          tmem_ld_32x32_x16(ts + 256u, vt);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) if (256 + j < Sk) m = fmaxf(m, vt[j]);
        }
        xmax[half * 128 + rloc] = m;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + lane_grp) : "memory");
        m = fmaxf(m, xmax[(half ^ 1) * 128 + rloc]);
        const float ms = m * scale_log2;
        // ---- pass 2: p = exp2(s*scale - m*scale), partial row sum, bf16 pairs -> TMEM
        float l = 0.0f;
        const uint32_t tp = tmem_base + lane_off + A::COL_P;
        tmem_ld_32x32(ts + static_cast<uint32_t>(c_begin * 32), va);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = c_begin + cc;
          float (&cur)[32] = (cc & 1) ? vb : va;
          float (&nxt)[32] = (cc & 1) ? va : vb;
          if (cc < 3) tmem_ld_32x32(ts + static_cast<uint32_t>((c + 1) * 32), nxt);
          uint32_t pk[16];
          if (c * 32 + 32 <= Sk) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = exp2f(fmaf(cur[j], scale_log2, -ms)), p1 = exp2f(fmaf(cur[j + 1], scale_log2, -ms));
              l += p0 + p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = (c * 32 + j < Sk) ? exp2f(fmaf(cur[j], scale_log2, -ms)) : 0.0f;
              const float p1 = (c * 32 + j + 1 < Sk) ? exp2f(fmaf(cur[j + 1], scale_log2, -ms)) : 0.0f;
              l += p0 + p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(tp + static_cast<uint32_t>(c * 16), pk);
          if (cc < 3) tmem_ld_wait();
        }
        if (half == 1) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float p0 = (256 + j < Sk) ? exp2f(fmaf(vt[j], scale_log2, -ms)) : 0.0f;
            const float p1 = (256 + j + 1 < Sk) ? exp2f(fmaf(vt[j + 1], scale_log2, -ms)) : 0.0f;
            l += p0 + p1;
            pk[j >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x8(tp + 128u, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        xsum[half * 128 + rloc] = l;
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- epilogue: this warp drains O columns [32*half, 32*half+32) of its rows
        mbar_wait(o_full, oph); oph ^= 1u;
        tc_fence_after();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + lane_grp) : "memory");      // partner's partial sum is visible
        const float inv = 1.0f / (l + xsum[(half ^ 1) * 128 + rloc]);
        const int row = t * 128 + rloc;
        bf16* orow = Og + (static_cast<size_t>(b) * Sq + row) * ldo + h * 64 + half * 32;
        float v[32];
        tmem_ld_32x32(tmem_base + lane_off + A::COL_O + static_cast<uint32_t>(half * 32), v);
        tmem_ld_wait();
        if (row < Sq) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(v[j] * inv, v[j + 1] * inv); o.y = pack_bf16x2(v[j + 2] * inv, v[j + 3] * inv);
            o.z = pack_bf16x2(v[j + 4] * inv, v[j + 5] * inv); o.w = pack_bf16x2(v[j + 6] * inv, v[j + 7] * inv);
            *reinterpret_cast<uint4*>(orow + j) = o;
          }
        }
        tc_fence_before();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + lane_grp) : "memory");      // both halves done with xmax/xsum before they are rewritten
        if (lane == 0) mbar_arrive(o_empty);
      }
    }
  } else {
    // ===== spare warp: the few query rows that do not fill a 128-row tile, on CUDA cores, against the smem K/V =====
    int kb = 0; uint32_t kph = 0;
    float* pbuf = reinterpret_cast<float*>(smem + A::OFF_P);
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int b = it / H, h = it % H;
      mbar_wait(&kv_full[kb], kph);
      const uint8_t* sk = smem + A::OFF_K + kb * A::KV_BYTES;
      const uint8_t* sv = smem + A::OFF_V + kb * A::KV_BYTES;
      for (int r = 0; r < side_rows; ++r) {
        const int row = side_row0 + r;
        const uint4* qp = reinterpret_cast<const uint4*>(Qg + (static_cast<size_t>(b) * Sq + row) * ldq + h * 64);
        uint4 q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = __ldg(qp + c);
        float sc[9];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const int j = lane + 32 * i;
          float acc = 0.0f;
          if (j < Sk) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 kk = *reinterpret_cast<const uint4*>(sk + j * 128 + ((c ^ (j & 7)) << 4));
              const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&q[c]);
              const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&kk);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 fa = __bfloat1622float2(a2[e]), fb = __bfloat1622float2(b2[e]);
                acc = fmaf(fa.x, fb.x, acc); acc = fmaf(fa.y, fb.y, acc);
              }
            }
            acc *= scale_log2;
            m = fmaxf(m, acc);
          } else acc = -INFINITY;
          sc[i] = acc;
        }
        m = warp_max(m);
        float l = 0.0f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const int j = lane + 32 * i;
          const float p = (j < Sk) ? exp2f(sc[i] - m) : 0.0f;
          l += p;
          if (j < A::SKP) pbuf[j] = p;
        }
        l = warp_sum(l);
        __syncwarp();
        float o0 = 0.0f, o1 = 0.0f;                        // lane owns d = 2*lane, 2*lane+1
        for (int j = 0; j < Sk; ++j) {
          const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(sv + j * 128 + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4);
          const float2 f = __bfloat1622float2(vv);
          const float p = pbuf[j];
          o0 = fmaf(p, f.x, o0); o1 = fmaf(p, f.y, o1);
        }
        const float inv = 1.0f / l;
        *reinterpret_cast<uint32_t*>(Og + (static_cast<size_t>(b) * Sq + row) * ldo + h * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
        __syncwarp();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&kv_empty[kb]);
      if (++kb == 2) { kb = 0; kph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
