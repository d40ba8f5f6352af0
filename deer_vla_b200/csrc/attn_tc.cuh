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
// attn_tc_kernel: softmax(scale * Q K^T) V on the 5th-gen tensor cores, head_dim 64, no mask, Sk <= 264 keys.
// (the ViT-L/14 attention: 257 x 257 per (image, head); ~3 % of a step's FLOPs but 18 % of its time on mma.sync)
//
// What bounds it: tensor-memory -> register reads run at only ~64 B/clk/SM, so the 128 x 256 fp32 score tile (128 KB) costs
// ~2000 cycles to read ONCE.  The kernel therefore reads S exactly once (each thread keeps its 128 scores in registers between
// the max and the exp), and overlaps that read with the tensor-core work of the neighbouring tiles:
//
// Persistent CTA per SM; one work item = one (batch, head).  K and V are TMA-loaded once per item in their natural [key][d]
// layout (two 136-row boxes each, 128B swizzle, double-buffered across items); each 128-row query tile n uses TMEM buffer
// u = n & 1 (256 columns each):
//   S = Q K^T : tcgen05.mma M=128, N=256, K=64 (A = Q smem, B = K smem, both K-major)                  -> cols [256u, 256u+256)
//   softmax   : 8 warps; a query row lives in one TMEM lane, the two warps sharing a lane quarter split its keys
//               ([0,128) / [128,256)), exchange their partial max / sum through smem, and write P (bf16 pairs) IN PLACE over
//               the score columns both have already consumed (tcgen05.st)                              -> cols [256u, 256u+128)
//   O = P V   : tcgen05.mma M=128, N=64, K<=256 (A = P in TMEM, B = V smem MN-major)                   -> cols [256u+128, 256u+192)
//   epilogue  : (O + tail terms) / rowsum -> bf16 -> global; DEFERRED into the next tile's iteration (between its score read and
//               its exp), so PV(n) and QK(n+1) run while the softmax warps are busy and nobody waits on the tensor core.
// Keys 256..Sk-1 (the 257th token of a 257-token image; at most 8) never enter the tensor core: their scores and their
// rank-1 contribution to O are computed per row on CUDA cores from the K/V rows already in shared memory -- that is what lets
// a tile fit 256 columns and two tiles fit the 512-column tensor memory.  The 1-8 query rows that do not fill a 128-row tile
// are done by a spare warp, also on CUDA cores.
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
  static constexpr int SKP = 272;                        // rows staged per K / V buffer (17 x 16)
  static constexpr int NMMA = 256;                       // keys that go through the tensor core
  static constexpr int MAX_TAIL = 8;                     // keys 256..263 at most on CUDA cores
  static constexpr int MAX_SK = NMMA + MAX_TAIL;
  static constexpr int KV_BYTES = SKP * 128;             // 34816
  static constexpr int HALF_ROWS = 136, HALF_BYTES = HALF_ROWS * 128;
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int OFF_K = 0, OFF_V = 2 * KV_BYTES, OFF_Q = 4 * KV_BYTES, OFF_P = OFF_Q + 2 * Q_BYTES;   // OFF_P: fp32 scratch of the spare warp
  static constexpr int OFF_XMAX = OFF_P + SKP * 4 + 64;          // [2 buf][4 stages][2 half][128]
  static constexpr int OFF_XSUM = OFF_XMAX + 2 * 4 * 2 * 128 * 4;  // [2 buf][2 half][128]
  static constexpr int OFF_XTAIL = OFF_XSUM + 2 * 2 * 128 * 4;   // [2 buf][16][128]
  static constexpr int OST_ROW = 80;                             // bytes per staged output row (64 B of bf16 + pad: conflict-free uint4 writes)
  static constexpr int OFF_OST = OFF_XTAIL + 2 * MAX_TAIL * 128 * 4;   // [8 warps][32 rows][OST_ROW]
  static constexpr int OFF_BAR = OFF_OST + 8 * 32 * OST_ROW;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int THREADS = 11 * 32;
  static constexpr int TMEM_COLS = 512;
};

// dot of row ra of one swizzled [row][64] bf16 tile with row rb of another (128B swizzle)
__device__ __forceinline__ float dot64_swz(const uint8_t* a_tile, int ra, const uint8_t* b_tile, int rb) {
  float acc = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 x = *reinterpret_cast<const uint4*>(a_tile + ra * 128 + ((c ^ (ra & 7)) << 4));
    const uint4 y = *reinterpret_cast<const uint4*>(b_tile + rb * 128 + ((c ^ (rb & 7)) << 4));
    const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __bfloat1622float2(x2[e]), fb = __bfloat1622float2(y2[e]);
      acc = fmaf(fa.x, fb.x, acc); acc = fmaf(fa.y, fb.y, acc);
    }
  }
  return acc;
}

template <int VER>      // 3 = whole-row softmax (read 128 scores, then exp), 4 = chunk-pipelined softmax (the default)
__global__ void __launch_bounds__(AttnTc::THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const bf16* __restrict__ Qg, int ldq, bf16* __restrict__ Og, int ldo, int B, int H, int Sq, int Sk, float scale_log2,
               long long* __restrict__ dbg) {
  using A = AttnTc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A::OFF_BAR);
  uint64_t* kv_full = bars;          // [2]  K/V of an item landed
  uint64_t* kv_empty = bars + 2;     // [2]  MMA commit + 8 softmax warps + spare warp are done with the item's K/V
  uint64_t* q_full = bars + 4;       // [2]
  uint64_t* q_empty = bars + 6;      // [2]  MMA commit (QK done) + 4 tail-scoring warps
  uint64_t* s_full = bars + 8;       // [2]  per TMEM buffer
  uint64_t* p_full = bars + 10;      // [2]
  uint64_t* o_full = bars + 12;      // [2]
  uint64_t* o_empty = bars + 14;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = B * H;
  const int n_full = Sq / 128, left = Sq % 128;
  const int n_sub = n_full + (left > 8 ? 1 : 0);          // query tiles on the tensor cores (host guarantees >= 1)
  const int side_rows = (left > 8) ? 0 : left;            // rows done by the spare warp
  const int side_row0 = n_full * 128;
  const int n_mma = min(Sk, A::NMMA);
  const int n_tail = Sk - n_mma;
  const int pv_steps = (n_mma + 15) / 16;
  const int my_items = (n_items > static_cast<int>(blockIdx.x)) ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  const int total = my_items * n_sub;                     // query tiles this CTA processes, in order

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 10);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 5);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, A::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int is = 0; is < my_items; ++is) {
        const int it = blockIdx.x + is * gridDim.x;
        const int b = it / H, h = it % H;
        const int kb = is & 1;
        mbar_wait(&kv_empty[kb], ((is >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[kb], 2 * A::KV_BYTES);
        uint8_t* sk = smem + A::OFF_K + kb * A::KV_BYTES;
        uint8_t* sv = smem + A::OFF_V + kb * A::KV_BYTES;
        tma_load_2d(sk, &tmK, &kv_full[kb], h * 64, b * Sk);
        tma_load_2d(sk + A::HALF_BYTES, &tmK, &kv_full[kb], h * 64, b * Sk + A::HALF_ROWS);
        tma_load_2d(sv, &tmV, &kv_full[kb], h * 64, b * Sk);
        tma_load_2d(sv + A::HALF_BYTES, &tmV, &kv_full[kb], h * 64, b * Sk + A::HALF_ROWS);
        for (int t = 0; t < n_sub; ++t) {
          const int n = is * n_sub + t, qb = n & 1;
          mbar_wait(&q_empty[qb], ((n >> 1) & 1) ^ 1u);
          mbar_arrive_expect_tx(&q_full[qb], A::Q_BYTES);
          tma_load_2d(smem + A::OFF_Q + qb * A::Q_BYTES, &tmQ, &q_full[qb], h * 64, b * Sq + t * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: QK runs one tile ahead of PV =====
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_major(128, A::NMMA, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(128, 64, 0, 1);     // B = V, MN-major
      auto issue_qk = [&](int n) {
        const int is = n / n_sub, t = n % n_sub, kb = is & 1, qb = n & 1, u = n & 1;
        if (t == 0) mbar_wait(&kv_full[kb], (is >> 1) & 1);
        mbar_wait(&q_full[qb], (n >> 1) & 1);
        if (n >= 2) mbar_wait(&o_empty[u], ((n >> 1) - 1) & 1);        // the buffer's previous tile has been drained
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + A::OFF_Q + qb * A::Q_BYTES);
        const uint32_t k_addr = smem_u32(smem + A::OFF_K + kb * A::KV_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + static_cast<uint32_t>(256 * u), make_desc_sw128(q_addr) + static_cast<uint64_t>(2 * k),
                    make_desc_sw128(k_addr) + static_cast<uint64_t>(2 * k), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[u]);
        umma_commit(&q_empty[qb]);
      };
      if (total > 0) issue_qk(0);
      for (int n = 0; n < total; ++n) {
        if (n + 1 < total) issue_qk(n + 1);
        const int is = n / n_sub, t = n % n_sub, kb = is & 1, u = n & 1;
        mbar_wait(&p_full[u], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + A::OFF_V + kb * A::KV_BYTES);
#pragma unroll 1
        for (int k = 0; k < pv_steps; ++k)
          umma_bf16_ts(tmem_base + static_cast<uint32_t>(256 * u + 128), tmem_base + static_cast<uint32_t>(256 * u + 8 * k),
                       make_desc(v_addr + static_cast<uint32_t>(k) * 2048u, 1024u, 1024u), idesc_pv, k != 0 ? 1u : 0u);
        umma_commit(&o_full[u]);
        if (t == n_sub - 1) umma_commit(&kv_empty[kb]);
      }
    }
  } else if (warp < 10) {
    // ===== softmax + (deferred) epilogue =====
    const int qd = warp & 3;                               // TMEM lane quarter
    const int hf = (warp - 2) >> 2;                        // which half of the keys / of the O columns
    const int rloc = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    float* xmax = reinterpret_cast<float*>(smem + A::OFF_XMAX);
    float* xsum = reinterpret_cast<float*>(smem + A::OFF_XSUM);
    float* xtail = reinterpret_cast<float*>(smem + A::OFF_XTAIL);
#ifdef DEER_ATTN_PROF
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool prof = (dbg != nullptr) && warp == 2 && lane == 0;
#define TICK(i) do { if (prof) { const long long _t = clock64(); tacc[i] += _t - t0; t0 = _t; } } while (0)
    long long t0 = prof ? clock64() : 0;
#else
#define TICK(i) do { } while (0)
#endif

    // epilogue of tile n (buffer u): (O + tail keys) / l -> global.  Both warps of a quarter have passed a bar.sync since the
    // partner published xsum / xtail for tile n.
    auto epilogue = [&](int n) {
      const int is = n / n_sub, t = n % n_sub, kb = is & 1, u = n & 1;
      const int it = blockIdx.x + is * gridDim.x;
      const int b = it / H, h = it % H;
      mbar_wait(&o_full[u], (n >> 1) & 1);
      tc_fence_after();
      TICK(5);
      const float inv = 1.0f / (xsum[(u * 2 + 0) * 128 + rloc] + xsum[(u * 2 + 1) * 128 + rloc]);
      const uint8_t* sv = smem + A::OFF_V + kb * A::KV_BYTES;
      uint8_t* ost = smem + A::OFF_OST + (warp - 2) * (32 * A::OST_ROW);
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {                     // 16 O columns at a time (keeps the register peak down)
        float v[16];
        tmem_ld_32x32_x16(tmem_base + lane_off + static_cast<uint32_t>(256 * u + 128 + hf * 32 + hh * 16), v);
        tmem_ld_wait();
        for (int j = 0; j < n_tail; ++j) {
          const int kr = A::NMMA + j;
          const float pj = xtail[(u * A::MAX_TAIL + j) * 128 + rloc];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {                 // two 16-byte chunks = 16 d; every lane reads the same address (broadcast)
            const uint4 y = *reinterpret_cast<const uint4*>(sv + kr * 128 + (((hf * 4 + hh * 2 + cc) ^ (kr & 7)) << 4));
            const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(y2[e]);
              v[cc * 8 + 2 * e] = fmaf(pj, f.x, v[cc * 8 + 2 * e]);
              v[cc * 8 + 2 * e + 1] = fmaf(pj, f.y, v[cc * 8 + 2 * e + 1]);
            }
          }
        }
        // this lane's row -> its slot of the warp's staging tile (a lane owns one ROW here; the stores below want rows contiguous)
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          uint4 o;
          o.x = pack_bf16x2(v[j] * inv, v[j + 1] * inv); o.y = pack_bf16x2(v[j + 2] * inv, v[j + 3] * inv);
          o.z = pack_bf16x2(v[j + 4] * inv, v[j + 5] * inv); o.w = pack_bf16x2(v[j + 6] * inv, v[j + 7] * inv);
          *reinterpret_cast<uint4*>(ost + lane * A::OST_ROW + hh * 32 + j * 2) = o;
        }
      }
      __syncwarp();
      // coalesced stores: 4 lanes cover the 64 contiguous bytes this warp owns of a row, 8 rows per instruction (a lane-per-row
      // store touched 32 different 128-byte lines per instruction and made the epilogue the longest phase of a tile)
      {
        const int ch = lane & 3, rsub = lane >> 2;
        bf16* obase = Og + (static_cast<size_t>(b) * Sq + t * 128 + qd * 32) * ldo + h * 64 + hf * 32 + ch * 8;
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          const int r = itr * 8 + rsub;
          const uint4 o = *reinterpret_cast<const uint4*>(ost + r * A::OST_ROW + ch * 16);
          if (t * 128 + qd * 32 + r < Sq) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(r) * ldo) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&o_empty[u]);
        if (t == n_sub - 1) mbar_arrive(&kv_empty[kb]);
      }
    };

    for (int n = 0; n <= total; ++n) {
      if (n == total) {                                    // drain: only the last tile's epilogue is left
        if (total > 0) {
          asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
          epilogue(total - 1);
        }
        break;
      }
      const int is = n / n_sub, kb = is & 1, qb = n & 1, u = n & 1;
      mbar_wait(&s_full[u], (n >> 1) & 1);
      tc_fence_after();
      TICK(0);
      if (VER == 4) {
        // ---- chunk-pipelined softmax.  The 128 scores of this thread arrive as four 32-column chunks; chunk k+1 is in flight
        // (tcgen05.ld is asynchronous until wait::ld) while chunk k goes through max -> exp2 -> pack, so the ~64 B/clk tensor-memory
        // read port and the MUFU pipe work at the same time instead of one after the other.
        // Exactness without the full-row max: probabilities are taken relative to a running reference r_k (scaled log2 domain),
        //   r_0 = max(stage-0 chunk of both halves, tail keys),   r_k = max(r_{k-1}, M_k - 30k)   (M_k = stage-k chunk max),
        // identical in the two warps that share a row (they exchange M_k through smem at every stage).  Stored p <= 2^(30k) <= 2^90;
        // if r_k moves, everything stored before is <= 2^(30(k-1)) against a new top term of exactly 2^(30k): the stale terms
        // carry <= 256 * 2^-30 = 2^-22 of the row's mass, i.e. the softmax is exact to fp32 rounding, with no second pass.
        const uint32_t ts = tmem_base + lane_off + static_cast<uint32_t>(256 * u + 128 * hf);
        const int c0 = 128 * hf;
        float sa[32], sb[32];
        uint32_t pk[64];
        tmem_ld_32x32(ts, sa);
        float m = -INFINITY;
        if (hf == 1) {
          const uint8_t* sq = smem + A::OFF_Q + qb * A::Q_BYTES;
          const uint8_t* sk = smem + A::OFF_K + kb * A::KV_BYTES;
#pragma unroll 1
          for (int j = 0; j < n_tail; ++j) {
            const float sj = dot64_swz(sq, rloc, sk, A::NMMA + j);
            xtail[(u * A::MAX_TAIL + j) * 128 + rloc] = sj;
            m = fmaxf(m, sj);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&q_empty[qb]);
        }
        tmem_ld_wait();
        tmem_ld_32x32(ts + 32u, sb);
        TICK(6);
        float* xm = xmax + u * (4 * 2 * 128);
        float ref = 0.0f, l = 0.0f;
#define STAGE_MAX(SX, K, MINIT)                                                                              \
        {                                                                                                    \
          if (c0 + 32 * ((K) + 1) > n_mma) {                                                                 \
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (c0 + 32 * (K) + j >= n_mma) SX[j] = -INFINITY; \
          }                                                                                                  \
          float cm = (MINIT);                                                                                \
          _Pragma("unroll") for (int j = 0; j < 32; j += 2) cm = fmaxf(cm, fmaxf(SX[j], SX[j + 1]));         \
          xm[((K) * 2 + hf) * 128 + rloc] = cm;                                                              \
          asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");                                         \
          cm = fmaxf(cm, xm[((K) * 2 + (hf ^ 1)) * 128 + rloc]) * scale_log2;                                \
          ref = ((K) == 0) ? cm : fmaxf(ref, cm - 30.0f * (K));                                              \
        }
#define STAGE_EXP(SX, K)                                                                                     \
        {                                                                                                    \
          _Pragma("unroll") for (int j = 0; j < 32; j += 2) {                                                \
            const float p0 = ex2_approx(fmaf(SX[j], scale_log2, -ref)), p1 = ex2_approx(fmaf(SX[j + 1], scale_log2, -ref)); \
            l += p0 + p1;                                                                                    \
            pk[(K) * 16 + (j >> 1)] = pack_bf16x2(p0, p1);                                                   \
          }                                                                                                  \
        }
        STAGE_MAX(sa, 0, m)
        STAGE_EXP(sa, 0)
        tmem_ld_wait();
        tmem_ld_32x32(ts + 64u, sa);
        STAGE_MAX(sb, 1, -INFINITY)
        TICK(1);
        if (n > 0) epilogue(n - 1);         // PV(n-1) has had two stages to finish; frees the other TMEM buffer for QK(n+1),
        TICK(2);                            // which still has two stages of this tile to complete in
        STAGE_EXP(sb, 1)
        tmem_ld_wait();
        tmem_ld_32x32(ts + 96u, sb);
        STAGE_MAX(sa, 2, -INFINITY)
        STAGE_EXP(sa, 2)
        tmem_ld_wait();
        STAGE_MAX(sb, 3, -INFINITY)         // after this stage's bar.sync BOTH warps of the quarter have read all their scores:
        STAGE_EXP(sb, 3)                    // P may now overwrite the score columns in place
#undef STAGE_MAX
#undef STAGE_EXP
        const uint32_t tp = tmem_base + lane_off + static_cast<uint32_t>(256 * u + 64 * hf);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          uint32_t q16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) q16[j] = pk[ci * 16 + j];
          tmem_st_32x32b_x16(tp + static_cast<uint32_t>(ci * 16), q16);
        }
        if (hf == 1) {
#pragma unroll 1
          for (int j = 0; j < n_tail; ++j) {
            const float pj = ex2_approx(fmaf(xtail[(u * A::MAX_TAIL + j) * 128 + rloc], scale_log2, -ref));
            l += pj;
            xtail[(u * A::MAX_TAIL + j) * 128 + rloc] = pj;
          }
        }
        xsum[(u * 2 + hf) * 128 + rloc] = l;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[u]);
        TICK(3);
        continue;
      }
      // ---- (VER 3) this thread's 128 scores, read ONCE
      const uint32_t ts = tmem_base + lane_off + static_cast<uint32_t>(256 * u + 128 * hf);
      float s0[32], s1[32], s2[32], s3[32];
      tmem_ld_32x32(ts, s0);
      tmem_ld_32x32(ts + 32u, s1);
      tmem_ld_32x32(ts + 64u, s2);
      tmem_ld_32x32(ts + 96u, s3);
      // tail keys (>= 256) on CUDA cores while the loads are in flight: this row's q (smem) . staged K rows
      float m = -INFINITY;
      if (hf == 1) {
        const uint8_t* sq = smem + A::OFF_Q + qb * A::Q_BYTES;
        const uint8_t* sk = smem + A::OFF_K + kb * A::KV_BYTES;
#pragma unroll 1
        for (int j = 0; j < n_tail; ++j) {                 // raw scores parked in smem (each thread reads back only its own slot)
          const float sj = dot64_swz(sq, rloc, sk, A::NMMA + j);
          xtail[(u * A::MAX_TAIL + j) * 128 + rloc] = sj;
          m = fmaxf(m, sj);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_empty[qb]);
      }
      tmem_ld_wait();
      TICK(6);
      const int c0 = 128 * hf;
      if (c0 + 128 > n_mma) {                              // keys past n_mma (padding / next image's rows): score -inf -> p = 0
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (c0 + j >= n_mma) s0[j] = -INFINITY;
          if (c0 + 32 + j >= n_mma) s1[j] = -INFINITY;
          if (c0 + 64 + j >= n_mma) s2[j] = -INFINITY;
          if (c0 + 96 + j >= n_mma) s3[j] = -INFINITY;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(fmaxf(m, s0[j]), fmaxf(fmaxf(s1[j], s2[j]), s3[j]));
      xmax[(u * 8 + hf) * 128 + rloc] = m;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");     // partner's max is visible AND both halves of S have been read
      m = fmaxf(m, xmax[(u * 8 + (hf ^ 1)) * 128 + rloc]);
      const float ms = m * scale_log2;
      TICK(1);
      // ---- the previous tile's epilogue goes here: its PV finished while we were reading scores
      if (n > 0) epilogue(n - 1);
      TICK(2);
      // ---- p = exp2(s*scale - m*scale), partial row sum; bf16 pairs written in place over the consumed score columns
      float l = 0.0f;
      const uint32_t tp = tmem_base + lane_off + static_cast<uint32_t>(256 * u + 64 * hf);
#define SOFTMAX_CHUNK(SX, CI)                                                                                          \
      {                                                                                                                 \
        uint32_t pk[16];                                                                                                \
        _Pragma("unroll") for (int j = 0; j < 32; j += 2) {                                                             \
          const float p0 = ex2_approx(fmaf(SX[j], scale_log2, -ms)), p1 = ex2_approx(fmaf(SX[j + 1], scale_log2, -ms)); \
          l += p0 + p1;                                                                                                 \
          pk[j >> 1] = pack_bf16x2(p0, p1);                                                                             \
        }                                                                                                               \
        tmem_st_32x32b_x16(tp + static_cast<uint32_t>((CI) * 16), pk);                                                  \
      }
      SOFTMAX_CHUNK(s0, 0)
      SOFTMAX_CHUNK(s1, 1)
      SOFTMAX_CHUNK(s2, 2)
      SOFTMAX_CHUNK(s3, 3)
#undef SOFTMAX_CHUNK
      if (hf == 1) {
#pragma unroll 1
        for (int j = 0; j < n_tail; ++j) {
          const float pj = ex2_approx(fmaf(xtail[(u * A::MAX_TAIL + j) * 128 + rloc], scale_log2, -ms));
          l += pj;
          xtail[(u * A::MAX_TAIL + j) * 128 + rloc] = pj;
        }
      }
      xsum[(u * 2 + hf) * 128 + rloc] = l;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[u]);
      TICK(3);
    }
    TICK(4);
#ifdef DEER_ATTN_PROF
    if (prof) for (int i = 0; i < 8; ++i) dbg[blockIdx.x * 8 + i] = tacc[i];
#endif
#undef TICK
  } else {
    // ===== spare warp: the few query rows that do not fill a 128-row tile, on CUDA cores, against the smem K/V =====
    float* pbuf = reinterpret_cast<float*>(smem + A::OFF_P);
    for (int is = 0; is < my_items; ++is) {
      const int it = blockIdx.x + is * gridDim.x;
      const int b = it / H, h = it % H, kb = is & 1;
      mbar_wait(&kv_full[kb], (is >> 1) & 1);
      const uint8_t* sK = smem + A::OFF_K + kb * A::KV_BYTES;
      const uint8_t* sV = smem + A::OFF_V + kb * A::KV_BYTES;
      for (int r = 0; r < side_rows; ++r) {
        const int row = side_row0 + r;
        const uint4* qp = reinterpret_cast<const uint4*>(Qg + (static_cast<size_t>(b) * Sq + row) * ldq + h * 64);
        uint4 q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = __ldg(qp + c);
        float m = -INFINITY;
#pragma unroll 1
        for (int j = lane; j < Sk; j += 32) {              // rolled on purpose: this rarely-run path must stay small in the I-cache
          float acc = 0.0f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 kk = *reinterpret_cast<const uint4*>(sK + j * 128 + ((c ^ (j & 7)) << 4));
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&q[c]);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&kk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fa = __bfloat1622float2(a2[e]), fb = __bfloat1622float2(b2[e]);
              acc = fmaf(fa.x, fb.x, acc); acc = fmaf(fa.y, fb.y, acc);
            }
          }
          acc *= scale_log2;
          pbuf[j] = acc;
          m = fmaxf(m, acc);
        }
        m = warp_max(m);
        float l = 0.0f;
#pragma unroll 1
        for (int j = lane; j < Sk; j += 32) {
          const float p = ex2_approx(pbuf[j] - m);
          pbuf[j] = p;
          l += p;
        }
        l = warp_sum(l);
        __syncwarp();
        float o0 = 0.0f, o1 = 0.0f;                        // lane owns d = 2*lane, 2*lane+1
        for (int j = 0; j < Sk; ++j) {
          const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(sV + j * 128 + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4);
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A::TMEM_COLS);
  }
}
