// GEMM kernels:  C[M,N] = epilogue(A[M,K] @ W[N,K]^T)
//   * gemm_tc_kernel<BN>  -- the product path: persistent, warp-specialised, TMA-fed
//     (cp.async.bulk.tensor, 128B swizzle) tcgen05.mma with the fp32 accumulator in TMEM
//     (double buffered so the epilogue of tile i overlaps the main loop of tile i+1).
//     A and W are both K-major bf16 (activations row-major, nn.Linear weights [out,in]).
//   * gemm_simt_kernel<T> -- CUDA-core reference used by the fp32 validation path and as
//     the in-library cross-check of the tensor-core kernel.
// Both share one epilogue: v = acc (+bias[n]); v = act(v); v *= *scale; v += resid[m,n];
// store as bf16 or fp32.  M can be clamped on the device (dynamic early exit: the number
// of still-active episodes lives in device memory).
#pragma once
#include "common.cuh"

struct GemmEpi {
  void* C;             // [M, ldc] bf16 (out_f32=0) or fp32 (out_f32=1); fp32 validation path stores T=float
  int ldc;
  int out_f32;
  const float* bias;   // [N] or nullptr
  int act;             // ACT_*
  const float* scale;  // device scalar (e.g. tanh(attn_gate)) or nullptr
  const float* resid;  // fp32 [M, ldr] or nullptr (may alias C)
  int ldr;
  int dbg;             // 0 normal; tuning only: 2 = skip the epilogue's math and stores in the 2-CTA kernel (mainloop-only timing),
                       // 4 = no L2 prefetch of the next tile's residual (A/B)
  int tma_store;       // 2-CTA kernel, bf16 output without residual: rows leave through shared memory + cp.async.bulk.tensor stores (tmC)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread SLEEPS in hardware until the phase flips or ~the hint (ns) expires, instead
// of spinning.  (Without the hint the single-thread TMA / MMA roles' spin loops issued as many instructions as the math
// warps sharing their scheduler -- visible as TRYWAIT/BRA/YIELD in the ncu instruction mix.)
#ifndef DEER_WAIT_HINT_NS
#define DEER_WAIT_HINT_NS 20000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(static_cast<uint32_t>(DEER_WAIT_HINT_NS))
      : "memory");
  return ok != 0;
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();            // >> seconds: each failed try sleeps up to ~20 us
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared -> global tile store (bulk async-group completion); coordinates (c0 = column, c1 = row); rows / columns outside the tensor
// map are clipped by the hardware
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// One lane of a CONVERGED warp.  The single-thread tcgen05 / TMA instructions run on the uniform datapath; issued from a branch the
// compiler cannot prove single-threaded (`if (lane == 0)`) every one of them is wrapped in an ELECT / BRA.U.ANY serialisation loop
// (~70 cycles per tcgen05.mma, measured: profiles/r02_mma_issue.txt), issued under elect.sync it is a plain predicated instruction.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// this warp's 32 TMEM lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t = lane t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
// start_address[0,14)=addr>>4, LBO[16,30)=1 (unused for swizzled K-major), SBO[32,46)=1024B>>4
// (stride between 8-row groups), version[46,48)=1, layout_type[61,64)=2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format[4,6)=1 (f32), a_format[7,10)=1 (bf16), b_format[10,13)=1 (bf16),
// a_major[15]=0, b_major[16]=0 (both K-major), n_dim[17,23)=N>>3, m_dim[24,29)=M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// shared epilogue: one output row, NV consecutive columns starting at col0
// ---------------------------------------------------------------------------------------------
// activation + scale with the (warp-uniform) activation switch hoisted OUT of the element loop
template <int NV, int ACT>
__device__ __forceinline__ void epi_act_scale_t(float sc, float (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = act_fast<ACT>(v[j]) * sc;
}
template <int NV>
__device__ __forceinline__ void epi_act_scale(int act, float sc, float (&v)[NV]) {
  if (act == ACT_NONE) epi_act_scale_t<NV, ACT_NONE>(sc, v);
  else if (act == ACT_GELU) epi_act_scale_t<NV, ACT_GELU>(sc, v);
  else if (act == ACT_QUICKGELU) epi_act_scale_t<NV, ACT_QUICKGELU>(sc, v);
  else epi_act_scale_t<NV, ACT_RELU>(sc, v);
}

template <int NV, bool FAST = false>
__device__ __forceinline__ void epi_store(const GemmEpi& e, float sc, int row, int col0, int N, float (&v)[NV]) {
  const bool full = (col0 + NV <= N);
  if (e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (full || col0 + j < N) v[j] += __ldg(e.bias + col0 + j);
  }
  if (FAST) epi_act_scale<NV>(e.act, sc, v);
  else {
    if (e.act != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = apply_act(v[j], e.act);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] *= sc;
  }
  if (e.resid != nullptr) {
    const float* rp = e.resid + static_cast<size_t>(row) * e.ldr + col0;
    if (full && (e.ldr & 3) == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        float4 r = *reinterpret_cast<const float4*>(rp + j);
        v[j] += r.x; v[j + 1] += r.y; v[j + 2] += r.z; v[j + 3] += r.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (col0 + j < N) v[j] += rp[j];
    }
  }
  if (e.out_f32) {
    float* cp = reinterpret_cast<float*>(e.C) + static_cast<size_t>(row) * e.ldc + col0;
    if (full && (e.ldc & 3) == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (col0 + j < N) cp[j] = v[j];
    }
  } else {
    bf16* cp = reinterpret_cast<bf16*>(e.C) + static_cast<size_t>(row) * e.ldc + col0;
    if (NV % 8 == 0 && full && (e.ldc & 7) == 0) {
#pragma unroll
      for (int j = 0; j + 8 <= NV; j += 8) {
        uint4 o;
        o.x = pack_bf16x2(v[j], v[j + 1]); o.y = pack_bf16x2(v[j + 2], v[j + 3]);
        o.z = pack_bf16x2(v[j + 4], v[j + 5]); o.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(cp + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (col0 + j < N) cp[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 GEMM
// ---------------------------------------------------------------------------------------------
// BM_ = 64: half-height tiles (tcgen05.mma M = 64: accumulator row r sits in TMEM lane (r % 16) + 32 * (r / 16), i.e. the first 16
// lanes of every lane quarter) for launches whose 128-row tiles would fill fewer than half of the SMs (a 1-episode ViT: M = 514
// rows x N = 1024 -> 80 tiles of 128 x 64 but 144 of 64 x 64).  Such launches are bound by what ONE SM can pull through its TMA
// pipeline per k-block (24 KB for a 128 x 64 tile, 16 KB for 64 x 64, and 12 stages instead of 8 in flight), not by the tensor
// pipe, which M = 64 runs at half rate.  K is accumulated in the same order: results are bit-identical to the 128-row tiles.
template <int BN, int BM_ = 128>
struct GemmTc {
  static constexpr int BM = BM_;
  static constexpr int BK = 64;                       // 64 bf16 = one 128B swizzle row
  static constexpr int UMMA_K = 16;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BM == 64) ? 12 : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = 2 * BN;            // two accumulator buffers (128 / 256 / 512 columns)
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // +1024: manual 1 KiB alignment
  static constexpr int EPI_WARPS = 8;                 // two warps per TMEM lane quarter, each owning half of the columns
  static constexpr int THREADS = 64 + 32 * EPI_WARPS; // warp0 TMA, warp1 MMA+TMEM alloc, warps2-9 epilogue
  static constexpr int ROWS_PER_QUARTER = BM / 4;     // accumulator rows held by one TMEM lane quarter (32, or 16 for BM = 64)
};

#ifdef DEER_GEMM_PROF
// instrumented build only (scripts/gemm_kblock_prof.sh): clock64 of CTA 0's producer / MMA threads per k-block of its first tile
__device__ long long g_gemm_prof[4 * 256];
#define GEMM_PROF(slot, kb) do { if (blockIdx.x == 0 && tile == 0 && (kb) < 256) g_gemm_prof[(kb) * 4 + (slot)] = clock64(); } while (0)
#else
#define GEMM_PROF(slot, kb) do { } while (0)
#endif

template <int BN, int BM_ = 128>
__global__ void __launch_bounds__(GemmTc<BN, BM_>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M_host, const int* __restrict__ m_dev, int m_mul, int N, int K, GemmEpi epi) {
  using G = GemmTc<BN, BM_>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G::STAGES * G::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G::STAGES;
  uint64_t* tfull_bar = empty_bar + G::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], G::EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, G::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                           // everything above overlapped the previous kernel's tail
  pdl_trigger();
  const int M = dyn_rows(M_host, m_dev, m_mul);
  const int m_blocks = (M + G::BM - 1) / G::BM;
  const int n_blocks = (N + BN - 1) / BN;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = (K + G::BK - 1) / G::BK;

  if (warp == 0) {
    // ===== TMA producer: converged warp, one elected lane issues =====
    {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;    // N fastest: the A tile stays hot in L2
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (elect_one_sync()) {
            GEMM_PROF(0, kb);
            uint8_t* sa = smem + s * G::STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[s], G::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full_bar[s], kb * G::BK, m_blk * G::BM);
            tma_load_2d(sa + G::A_BYTES, &tmB, &full_bar[s], kb * G::BK, n_blk * BN);
            GEMM_PROF(1, kb);
          }
          __syncwarp();
          if (++s == G::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp stays converged, one elected lane issues (see elect_one_sync) =====
    {
      constexpr uint32_t idesc = make_idesc_bf16(G::BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aph ^ 1u);          // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);                  // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * G::STAGE_BYTES);
          const uint64_t adesc = make_desc_sw128(a_addr);
          const uint64_t bdesc = make_desc_sw128(a_addr + G::A_BYTES);
          if (elect_one_sync()) {
            GEMM_PROF(2, kb);
#pragma unroll
            for (int k = 0; k < G::BK / G::UMMA_K; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128B swizzle row: +2 in the (addr>>4) field
              umma_bf16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                        (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);                 // frees the smem stage when these MMAs retire
            GEMM_PROF(3, kb);
          }
          __syncwarp();
          if (++s == G::STAGES) { s = 0; ph ^= 1u; }
        }
        if (elect_one_sync()) umma_commit(&tfull_bar[as]);   // accumulator complete -> epilogue
        __syncwarp();
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global =====
    const int lane_grp = warp & 3;                      // TMEM lane quarter this warp may access
    const int col_half = (warp - 2) >> 2;               // which half of the tile's columns this warp drains
    constexpr int CH = BN / 64;                         // 32-column chunks per half
    const float sc = (epi.scale != nullptr) ? __ldg(epi.scale) : 1.0f;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m_blk * G::BM + lane_grp * G::ROWS_PER_QUARTER + lane;
      const bool row_ok = (G::BM == 128 || lane < G::ROWS_PER_QUARTER) && row < M;   // BM = 64: lanes 16-31 of a quarter hold no row
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = col_half * CH; c < (col_half + 1) * CH; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && col0 < N) epi_store<32, true>(epi, sc, row, col0, N, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Split-K variant of the 128 x 64 kernel for the few-row GEMMs of a 1-4 episode step (M <= 128 token rows x N = 2048:
// 32 tiles on 148 SMs).  Those launches stream their weight matrix once and are bound by the bytes in flight, i.e. by
// how many SMs have a TMA pipeline open; grid = (tiles, splits) puts `splits` CTAs on each tile, each accumulating
// kb_per k-blocks in TMEM and writing its fp32 partial tile to a workspace.  The last CTA to arrive on the tile's
// counter (threadfence + atomic) sums the partials IN SPLIT ORDER -- the result does not depend on which CTA is last
// -- and runs the ordinary epilogue; it also re-arms the counter for the next launch.
// ws: [splits, M_host, N] fp32.  counters: one int per tile, zero between launches.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GemmTc<64>::THREADS, 1)
gemm_tc_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      int M_host, const int* __restrict__ m_dev, int m_mul, int N, int K, GemmEpi epi,
                      float* __restrict__ ws, int* __restrict__ counters, int kb_per) {
  constexpr int BN = 64;
  using G = GemmTc<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G::STAGES * G::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G::STAGES;
  uint64_t* tfull_bar = empty_bar + G::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  const int M = dyn_rows(M_host, m_dev, m_mul);
  const int n_blocks = (N + BN - 1) / BN;
  const int tile = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
  const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
  const int num_kb = (K + G::BK - 1) / G::BK;
  const int kb0 = split * kb_per;
  const int kb1 = min(num_kb, kb0 + kb_per);          // the host sizes `splits` so that every split owns >= 1 k-block
  const bool live = m_blk * G::BM < M;                // tiles past the device-side row count: all of their splits skip alike

  if (warp == 0) {
    if (live) {                                         // converged warp, one elected lane issues (see elect_one_sync)
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (elect_one_sync()) {
          uint8_t* sa = smem + s * G::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], G::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[s], kb * G::BK, m_blk * G::BM);
          tma_load_2d(sa + G::A_BYTES, &tmB, &full_bar[s], kb * G::BK, n_blk * BN);
        }
        __syncwarp();
        if (++s == G::STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (live) {
      constexpr uint32_t idesc = make_idesc_bf16(G::BM, BN);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * G::STAGE_BYTES);
        const uint64_t adesc = make_desc_sw128(a_addr);
        const uint64_t bdesc = make_desc_sw128(a_addr + G::A_BYTES);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < G::BK / G::UMMA_K; ++k)
            umma_bf16(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == G::STAGES) { s = 0; ph ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(tfull_bar);
      __syncwarp();
    }
  } else if (live) {
    // ===== epilogue warps: TMEM -> fp32 partial tile in the workspace; the tile's last CTA reduces and stores =====
    const int lane_grp = warp & 3;
    const int col_half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;                    // 0..255 among the epilogue threads
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int row = m_blk * G::BM + lane_grp * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    {
      float v[32];
      tmem_ld_32x32(taddr + static_cast<uint32_t>(col_half * 32), v);
      tmem_ld_wait();
      const int col0 = n_blk * BN + col_half * 32;
      if (row < M && col0 < N) {                       // N % 8 == 0 (host): an 8-column group is inside or outside as a whole
        float* wp = ws + (static_cast<size_t>(split) * M_host + row) * N + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (col0 + j < N) __stcg(reinterpret_cast<float4*>(wp + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
    }
    __threadfence();                                    // partial tile visible device-wide before the arrival below
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (et == 0) {
      __threadfence();
      const int old = atomicAdd(&counters[tile], 1);
      const int last = (old == splits - 1);
      if (last) counters[tile] = 0;                     // re-armed for the next launch (stream order)
      *last_flag = last;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (*last_flag) {
      __threadfence();
      const float sc = (epi.scale != nullptr) ? __ldg(epi.scale) : 1.0f;
      const int rows_here = min(G::BM, M - m_blk * G::BM);
      for (int idx = et; idx < rows_here * (BN / 8); idx += 256) {
        const int r = m_blk * G::BM + idx / (BN / 8);
        const int col0 = n_blk * BN + (idx % (BN / 8)) * 8;
        if (col0 >= N) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        const float* wp = ws + static_cast<size_t>(r) * N + col0;
        const size_t sstride = static_cast<size_t>(M_host) * N;
#pragma unroll 4
        for (int s = 0; s < splits; ++s) {
          const float4 p0 = __ldcg(reinterpret_cast<const float4*>(wp + s * sstride));
          const float4 p1 = __ldcg(reinterpret_cast<const float4*>(wp + s * sstride + 4));
          v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
          v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
        }
        epi_store<8, true>(epi, sc, r, col0, N, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------------------------
// 2-CTA tcgen05 GEMM (cta_group::2): a CTA pair (cluster of 2, same TPC) computes a 256 x 256 tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N-rows), so shared-memory
// fill + operand-read traffic per SM is 2/3 of the 1-CTA kernel's and the L2->SM traffic per FLOP halves
// for B.  The leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M=256,N=256,K=16); the
// accumulator rows 0-127 land in the leader's TMEM, rows 128-255 in the peer's.  Both CTAs' TMA loads
// complete_tx on the LEADER's full barrier; tcgen05.commit multicasts the "stage free" / "accumulator
// ready" arrivals to both CTAs.  Epilogue: TMEM -> registers -> per-warp smem transpose -> fully
// coalesced global loads/stores (row-contiguous), 8 epilogue warps per CTA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the consumer only needs this warp's TMEM reads to be done (tcgen05.wait::ld + fence::before_thread_sync);
  // a release at cluster scope would wait for the tile's global stores to drain (ERRBAR stall in the ncu source view)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* tm, uint32_t leader_bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {     // arrives on `bar` in BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// Coalesced epilogue for one 32x32 fp32 accumulator chunk held one-row-per-lane in v[].  The raw accumulators are
// transposed through this warp's 32x36-float smem buffer; bias / activation / scale / residual are then applied in the
// ROW-CONTIGUOUS layout (8 lanes cover the 32 columns of a row, a warp instruction covers 4 rows), so every global
// access is a full 128 B (fp32) or 64 B (bf16) row segment and the bias / residual loads are issued before the
// TMEM data is needed (nothing on the critical path waits on a dependent global load).
template <int ACT>
__device__ __forceinline__ float4 epi_math4(float4 t, float4 b, float sc) {
  t.x = act_fast<ACT>(t.x + b.x) * sc; t.y = act_fast<ACT>(t.y + b.y) * sc;
  t.z = act_fast<ACT>(t.z + b.z) * sc; t.w = act_fast<ACT>(t.w + b.w) * sc;
  return t;
}
// DEEP (instantiated for ACT_NONE only -- every residual GEMM of the model): all eight fp32 residual row segments of the chunk
// are requested at once, as soon as the accumulator registers are free (4 KB per warp in flight).  The shallow form requests four
// at a time and waits for each batch before its stores: two serialized HBM round trips per 32 x 32 chunk, which made the residual
// GEMMs latency-bound at 4.4 TB/s.  (Deep loads in every activation variant spill at this kernel's register budget.)
template <int ACT, bool DEEP>
__device__ __forceinline__ void epi_store_coalesced_t(const GemmEpi& e, float sc, int row0, int col0, int M, int N, float (&v)[32],
                                                      float* stage, int lane) {
  const bool fullc = (col0 + 32 <= N);
  const bool vec_ok = fullc && ((e.ldc & 7) == 0) && (e.resid == nullptr || (e.ldr & 3) == 0);
  const int sub = lane >> 3, c4 = (lane & 7) * 4;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool has_res = vec_ok && (e.resid != nullptr);
  if (vec_ok && e.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + c4));
#pragma unroll
  for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  if (DEEP && vec_ok) {
    float4 rs[8];                             // v[] is dead from here on
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row0 + it * 4 + sub;
      rs[it] = (has_res && row < M) ? *reinterpret_cast<const float4*>(e.resid + static_cast<size_t>(row) * e.ldr + col0 + c4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + sub;
      const int row = row0 + rr;
      float4 t = epi_math4<ACT>(*reinterpret_cast<const float4*>(stage + rr * 36 + c4), b4, sc);
      if (row < M) {
        t.x += rs[it].x; t.y += rs[it].y; t.z += rs[it].z; t.w += rs[it].w;
        if (e.out_f32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.C) + static_cast<size_t>(row) * e.ldc + col0 + c4) = t;
        } else {
          uint2 o; o.x = pack_bf16x2(t.x, t.y); o.y = pack_bf16x2(t.z, t.w);
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(e.C) + static_cast<size_t>(row) * e.ldc + col0 + c4) = o;
        }
      }
    }
    __syncwarp();
    return;
  }
  __syncwarp();
  if (vec_ok) {
    // v[] is dead from here on: the residual rows are fetched in two batches of four 128-byte row segments per 8 lanes
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 rs[4];
      if (has_res) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = row0 + (half * 4 + it) * 4 + sub;
          rs[it] = (row < M) ? *reinterpret_cast<const float4*>(e.resid + static_cast<size_t>(row) * e.ldr + col0 + c4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = (half * 4 + it) * 4 + sub;
        const int row = row0 + rr;
        float4 t = epi_math4<ACT>(*reinterpret_cast<const float4*>(stage + rr * 36 + c4), b4, sc);
        if (row < M) {
          if (has_res) { t.x += rs[it].x; t.y += rs[it].y; t.z += rs[it].z; t.w += rs[it].w; }
          if (e.out_f32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.C) + static_cast<size_t>(row) * e.ldc + col0 + c4) = t;
          } else {
            uint2 o; o.x = pack_bf16x2(t.x, t.y); o.y = pack_bf16x2(t.z, t.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(e.C) + static_cast<size_t>(row) * e.ldc + col0 + c4) = o;
          }
        }
      }
    }
  } else {
    const int col = col0 + lane;
    const float b = (e.bias != nullptr && col < N) ? __ldg(e.bias + col) : 0.0f;
    for (int rr = 0; rr < 32; ++rr) {
      const int row = row0 + rr;
      if (row < M && col < N) {
        float t = act_fast<ACT>(stage[rr * 36 + lane] + b) * sc;
        if (e.resid != nullptr) t += e.resid[static_cast<size_t>(row) * e.ldr + col];
        if (e.out_f32) reinterpret_cast<float*>(e.C)[static_cast<size_t>(row) * e.ldc + col] = t;
        else reinterpret_cast<bf16*>(e.C)[static_cast<size_t>(row) * e.ldc + col] = __float2bfloat16_rn(t);
      }
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void epi_store_coalesced(const GemmEpi& e, float sc, int row0, int col0, int M, int N, float (&v)[32],
                                                    float* stage, int lane) {
  if (e.act == ACT_NONE) {
    if (e.resid != nullptr) epi_store_coalesced_t<ACT_NONE, true>(e, sc, row0, col0, M, N, v, stage, lane);
    else epi_store_coalesced_t<ACT_NONE, false>(e, sc, row0, col0, M, N, v, stage, lane);
  } else if (e.act == ACT_GELU) epi_store_coalesced_t<ACT_GELU, false>(e, sc, row0, col0, M, N, v, stage, lane);
  else if (e.act == ACT_QUICKGELU) epi_store_coalesced_t<ACT_QUICKGELU, false>(e, sc, row0, col0, M, N, v, stage, lane);
  else epi_store_coalesced_t<ACT_RELU, false>(e, sc, row0, col0, M, N, v, stage, lane);
}

// bf16 output without residual, through shared memory and a TMA store: the thread that owns accumulator row r (TMEM lane) applies
// bias / activation / scale to its 32 columns, packs them to 64 bytes and writes them as the row of a [32 rows][64 B] tile in the
// 64-byte-swizzle layout the tensor map expects (16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3): the eight lanes of a
// store phase hit eight different bank groups), then ONE cp.async.bulk.tensor instruction per warp moves the tile.  Against the
// transposing epilogue this drops the fp32 staging round trip through shared memory (4.6 KB written + read per chunk -> 2 KB
// written) and all per-thread global stores.  `buf`: this warp's 2 KB buffer (512-byte aligned).
template <int ACT>
__device__ __forceinline__ void epi_tma_store_t(const GemmEpi& e, const CUtensorMap* tmC, float sc, int row0, int col0, float (&v)[32],
                                                uint8_t* buf, int lane) {
  uint32_t w[16];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j));
    const float4 t = epi_math4<ACT>(make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), b4, sc);
    w[j >> 1] = pack_bf16x2(t.x, t.y);
    w[(j >> 1) + 1] = pack_bf16x2(t.z, t.w);
  }
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(buf + lane * 64 + ((c ^ sw) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (elect_one_sync()) {                                // the same leader lane every time (same member mask): it owns the bulk groups
    tma_store_2d(tmC, buf, col0, row0);
    tma_store_commit();
  }
}
__device__ __forceinline__ void epi_tma_store(const GemmEpi& e, const CUtensorMap* tmC, float sc, int row0, int col0, float (&v)[32],
                                              uint8_t* buf, int lane) {
  if (e.act == ACT_NONE) epi_tma_store_t<ACT_NONE>(e, tmC, sc, row0, col0, v, buf, lane);
  else if (e.act == ACT_GELU) epi_tma_store_t<ACT_GELU>(e, tmC, sc, row0, col0, v, buf, lane);
  else if (e.act == ACT_QUICKGELU) epi_tma_store_t<ACT_QUICKGELU>(e, tmC, sc, row0, col0, v, buf, lane);
  else epi_tma_store_t<ACT_RELU>(e, tmC, sc, row0, col0, v, buf, lane);
}

struct GemmTc2 {
  static constexpr int BM = 256, BN = 256, BK = 64, UMMA_K = 16;
  static constexpr int A_BYTES = 128 * BK * 2;        // this CTA's 128 rows of A
  static constexpr int B_BYTES = 128 * BK * 2;        // this CTA's half (128 N-rows) of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 5;
  static constexpr int EPI_WARPS = 12;                // three warps per TMEM lane quarter; 448 threads -> 128 registers each
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * 36 * 4;
  static constexpr int TMEM_COLS = 512;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + BAR_BYTES + 1024;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GemmTc2::THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                int M_host, const int* __restrict__ m_dev, int m_mul, int N, int K, GemmEpi epi) {
  using G = GemmTc2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* epi_stage = reinterpret_cast<float*>(smem + G::STAGES * G::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G::STAGES * G::STAGE_BYTES + G::EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G::STAGES;
  uint64_t* tfull_bar = empty_bar + G::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (epi.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                      // only the leader's is used: its own arrive.expect_tx
      mbar_init(&empty_bar[s], 1);                     // one multicast tcgen05.commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * G::EPI_WARPS);     // leader's: epilogue warps of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, G::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                           // everything above overlapped the previous kernel's tail
  pdl_trigger();
  const int M = dyn_rows(M_host, m_dev, m_mul);
  const int m_blocks = (M + G::BM - 1) / G::BM;
  const int n_blocks = (N + G::BN - 1) / G::BN;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = (K + G::BK - 1) / G::BK;

  if (warp == 0) {
    // ===== TMA producer (one per CTA): own A rows + own half of B; bytes are accounted on the leader's barrier =====
    {                                                   // converged warp, one elected lane issues (see elect_one_sync)
      int s = 0;
      uint32_t ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (elect_one_sync()) {
            uint8_t* sa = smem + s * G::STAGE_BYTES;
            const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * G::STAGE_BYTES);
            tma_load_2d_2cta(sa, &tmA, lead_full, kb * G::BK, m_blk * G::BM + static_cast<int>(rank) * 128);
            tma_load_2d_2cta(sa + G::A_BYTES, &tmB, lead_full, kb * G::BK, n_blk * G::BN + static_cast<int>(rank) * 128);
          }
          __syncwarp();
          if (++s == G::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only; the warp stays converged, one elected lane drives both SMs' tensor cores =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(G::BM, G::BN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[as], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * G::BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * G::STAGE_BYTES);
          const uint64_t adesc = make_desc_sw128(a_addr);
          const uint64_t bdesc = make_desc_sw128(a_addr + G::A_BYTES);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < G::BK / G::UMMA_K; ++k)
              umma_bf16_2cta(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&empty_bar[s]);
          }
          __syncwarp();
          if (++s == G::STAGES) { s = 0; ph ^= 1u; }
        }
        if (elect_one_sync()) umma_commit_2cta(&tfull_bar[as]);
        __syncwarp();
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): this CTA's 128 accumulator rows =====
    const int ew = warp - 2;
    const int lane_grp = warp & 3;
    const int col_part = ew >> 2;                       // 0..2: the eight 32-column chunks are split 3 / 3 / 2
    const int c_begin = col_part * 3, c_end = (col_part == 2) ? 8 : c_begin + 3;
    float* stage = epi_stage + ew * (32 * 36);
    uint8_t* obuf = reinterpret_cast<uint8_t*>(stage);   // TMA-store path: two 2 KB tiles inside this warp's 4.5 KB staging region
    int ob = 0;
    const float sc = (epi.scale != nullptr) ? __ldg(epi.scale) : 1.0f;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      // The fp32 residual of the NEXT tile is pulled into L2 now -- one tile period (a whole main loop) ahead of its use, while
      // this thread would otherwise wait for the accumulator: the epilogue's residual loads then cost an L2 hit instead of an
      // HBM round trip (it was latency-bound on them: vit.out+res long_scoreboard 29 %, tensor pipe 48 -> 56 %).  Only for K <= 1024:
      // with a longer main loop the epilogue has slack anyway and the prefetched lines are evicted before they are used (ncu:
      // vit.fc2+res, K = 4096, read 2.0 GB instead of 1.5 GB per launch with the prefetch on).
      if (epi.resid != nullptr && epi.dbg != 4 && num_kb <= 16 && tile + num_clusters < num_tiles) {
        const int nt = tile + num_clusters;
        const int prow = (nt / n_blocks) * G::BM + static_cast<int>(rank) * 128 + lane_grp * 32 + lane;
        const int pcol = (nt % n_blocks) * G::BN;
        if (prow < M) {
          for (int c = c_begin; c < c_end; ++c)
            if (pcol + c * 32 < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(epi.resid + static_cast<size_t>(prow) * epi.ldr + pcol + c * 32));
        }
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row0 = m_blk * G::BM + static_cast<int>(rank) * 128 + lane_grp * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(as * G::BN);
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
        const int col0 = n_blk * G::BN + c * 32;
        if (epi.dbg != 2 && row0 < M && col0 < N) {
          if (epi.tma_store) {
            if (elect_one_sync()) tma_store_wait_read<1>(); // the store issued two chunks ago (same buffer) has read its tile
            __syncwarp();
            epi_tma_store(epi, &tmC, sc, row0, col0, v, obuf + ob * 2048, lane);
            ob ^= 1;
          } else {
            epi_store_coalesced(epi, sc, row0, col0, M, N, v, stage, lane);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  }
  if (epi.tma_store && warp >= 2) {
    if (elect_one_sync()) tma_store_wait_all();         // the leader lane's bulk stores have been written
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, G::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT reference GEMM (fp32 accumulate; T = float for the validation path, bf16 for cross-checks)
// 64x64 tile, 256 threads, 4x4 outputs per thread, K step 16.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, int ldw, int M_host,
                 const int* __restrict__ m_dev, int m_mul, int N, int K, GemmEpi epi) {
  __shared__ float As[16][64 + 1];
  __shared__ float Ws[16][64 + 1];
  const int M = dyn_rows(M_host, m_dev, m_mul);
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  if (m0 >= M) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
      const int r = idx >> 4, kk = idx & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? to_f32<T>(A[static_cast<size_t>(gm) * lda + gk]) : 0.0f;
      Ws[kk][r] = (gn < N && gk < K) ? to_f32<T>(W[static_cast<size_t>(gn) * ldw + gk]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float sc = (epi.scale != nullptr) ? __ldg(epi.scale) : 1.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    const int col0 = n0 + tx * 4;
    if (row < M && col0 < N) {
      float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
      epi_store<4>(epi, sc, row, col0, N, v);
    }
  }
}
