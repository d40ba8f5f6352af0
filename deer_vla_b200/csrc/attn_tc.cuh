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
// Shared pieces of the tensor-core ViT attention (attn_tc5.cuh): tensor-memory store / load shapes, the shared-memory
// budget of one (image, head) work item and the CUDA-core dot product for the tail keys.  (The round-1 schedules v3 / v4 --
// whole-row and two-warps-per-row softmax -- were measured against v5, profiles/r01_attn_ncu_and_tmem_probe.txt, and removed.)
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
