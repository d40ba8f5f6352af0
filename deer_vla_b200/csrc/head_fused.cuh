// One persistent kernel for a whole exit-head probe of a SMALL active set (<= 32 episodes): DeterministicDecoder's
// single-step forward (action_head.py:499-611) = max-pool over tokens -> LayerNormLSTM x layers (action_head.py:15-64) ->
// both MLP heads [Linear, LN, ReLU] x 2 -> Linear -> tanh / sigmoid (action_head.py:82-161,183-269).
//
// Why: with <= 32 rows every one of the head's seven weight matrices is a skinny (GEMV-like) contraction whose cost is the
// 82 MB of bf16 weights it streams, not its FLOPs.  As separate launches (15 per probe: pool, gather, 4 x [GEMM, cell], GEMM,
// LN, 2 x GEMM, LN, output) each matrix is streamed by the <= 64 CTAs its 128 x 64 output tiles give, behind a launch gap;
// here ALL SMs stream every matrix (every CTA owns 8 LSTM units / 16 or 8 MLP columns), the whole probe's weights are
// prefetched into L2 at kernel entry, the layers are separated by grid barriers instead of launches, and LayerNorm / the LSTM
// cell / ReLU are applied on the fly while the next operand is staged into shared memory.
//
// Math: mma.sync.m16n8k16 (bf16 x bf16 -> fp32).  The weights never touch shared memory: a thread's B fragment is one
// 16-byte global load of 8 consecutive k of its weight row (the k index inside a 32-wide chunk is permuted identically for A
// and B, which leaves the contraction unchanged), so every 32-byte sector fetched is fully used.  The activations (<= 32 rows,
// bf16) sit in shared memory with a 64-byte row skew (conflict-free 16-byte loads).  K is split over the 8 warps of a CTA in
// interleaved 32-wide chunks and reduced in a fixed order through shared memory: results are deterministic and independent
// of how many rows are active.
#pragma once
#include "common.cuh"

#define HF_MAX_LAYERS 8
#define HF_THREADS 256
#define HF_WARPS 8

struct HeadFusedParams {
  // inputs
  const float* x;            // [n, S, D] fp32 hidden state of the probed layer
  const int* valid_len;      // [n]
  const int* slot;           // [n] LSTM state slot of active row r
  const float* h_state;      // [slots, LL, H]
  const float* c_state;
  int S, D, H, LL, nstride, d0, d1;
  int mr;                    // rows the shared-memory operand holds (16 or 32)
  int lda;                   // shared-memory row pitch in elements (max K + 32)
  // weights
  const bf16* w_cat[HF_MAX_LAYERS];     // [4H, in_l + H]  ([W_ih | W_hh])
  const float* b_sum[HF_MAX_LAYERS];    // [4H]  b_ih + b_hh
  const float* ln_g[HF_MAX_LAYERS];     // [H]
  const float* ln_b[HF_MAX_LAYERS];
  const bf16* w_mlp1; const float *b_mlp1, *g_mlp1, *be_mlp1;           // [2*d0, H] (pose rows | gripper rows)
  const bf16 *w_mlp2_pose, *w_mlp2_grip; const float *b_mlp2, *g_mlp2, *be_mlp2;   // [d1, d0] each
  const float *w3_pose, *b3_pose, *w3_grip, *b3_grip;                  // [6, d1], [6], [1, d1], [1]
  // scratch / outputs
  float* pooled;             // [n, D] fp32 max-pooled feature (`.logits[2]`)
  float* mlp1;               // [n, 2*d0] pre-LayerNorm
  float* mlp2;               // [n, 2*d1] pre-LayerNorm
  float *cand_h, *cand_c;    // [LL, nstride, H] candidate state (stored h is PRE-LayerNorm, action_head.py:52-57)
  float *cur_pose, *cur_grip;
  unsigned long long* bar;   // grid barrier counter (monotonic; every launch adds gridDim.x * (LL + 3))
  int n_host; const int* n_dev;
  long long* dbg;            // optional [16]: %globaltimer (ns) of CTA 0 at every phase boundary (tuning only)
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// All CTAs of the (cooperatively launched, co-resident) grid meet here.  Bounded: a protocol bug traps instead of hanging.
__device__ __forceinline__ void hf_grid_barrier(unsigned long long* bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1ULL);
    const long long t0 = clock64();
    while (ld_acquire_u64(bar) < target) {
      __nanosleep(32);
      if (clock64() - t0 > 40000000000LL) __trap();     // ~20 s: far beyond any real wait, also under compute-sanitizer's slow-down
    }
    __threadfence();
  }
  __syncthreads();
}

// acc[mt][nt][4] += A_s[rows, k] * W[wrow(nt), k] over this warp's interleaved 32-wide k chunks, in batches of four chunks.
// wp[nt]: this thread's weight row pointer (row = column (lane >> 2) of n-tile nt), already offset by (lane & 3) * 8.
// as: this thread's operand pointer A_s + (lane >> 2) * lda + (lane & 3) * 8 (+ a column offset of the operand).
// The FIRST batch of a phase is loaded into registers before the grid barrier and the operand staging (the weights do not
// depend on them), so its HBM latency is hidden behind both.
template <int NT>
__device__ __forceinline__ void hf_load_batch(uint4 (&b)[4][NT], const bf16* (&wp)[NT], int c0, int kc) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = c0 + u * HF_WARPS;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b[u][nt] = (c < kc) ? ldg_nc_v4(wp[nt] + c * 32) : make_uint4(0u, 0u, 0u, 0u);
  }
}
template <int NT>
__device__ __forceinline__ void hf_mma_batch(const bf16* as, int lda, bool two_m, const uint4 (&b)[4][NT], int c0, int kc, float (&acc)[2][NT][4]) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = c0 + u * HF_WARPS;
    if (c < kc) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt == 0 || two_m) {
          const uint4 lo = *reinterpret_cast<const uint4*>(as + (mt * 16) * lda + c * 32);
          const uint4 hi = *reinterpret_cast<const uint4*>(as + (mt * 16 + 8) * lda + c * 32);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            mma_bf16_16816(acc[mt][nt], lo.x, hi.x, lo.y, hi.y, b[u][nt].x, b[u][nt].y);
            mma_bf16_16816(acc[mt][nt], lo.z, hi.z, lo.w, hi.w, b[u][nt].z, b[u][nt].w);
          }
        }
      }
    }
  }
}
// `b` holds the batch at c0 = warp (already loaded)
template <int NT>
__device__ __forceinline__ void hf_skinny_gemm(const bf16* as, int lda, bool two_m, const bf16* (&wp)[NT], int K, int warp,
                                               uint4 (&b)[4][NT], float (&acc)[2][NT][4]) {
  const int kc = K >> 5;                                  // 32-wide chunks
  for (int c0 = warp; c0 < kc; c0 += 4 * HF_WARPS) {
    if (c0 != warp) hf_load_batch<NT>(b, wp, c0, kc);
    hf_mma_batch<NT>(as, lda, two_m, b, c0, kc, acc);
  }
}

// per-warp partial tiles -> red[warp][row][NT*8]
template <int NT>
__device__ __forceinline__ void hf_store_partials(float* red, int mr, bool two_m, int warp, int lane, const float (&acc)[2][NT][4]) {
  const int g = lane >> 2, t = lane & 3;
  float* base = red + static_cast<size_t>(warp) * mr * (NT * 8);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    if (mt == 0 || two_m) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float* p0 = base + (mt * 16 + g) * (NT * 8) + nt * 8 + 2 * t;
        float* p1 = base + (mt * 16 + g + 8) * (NT * 8) + nt * 8 + 2 * t;
        p0[0] = acc[mt][nt][0]; p0[1] = acc[mt][nt][1];
        p1[0] = acc[mt][nt][2]; p1[1] = acc[mt][nt][3];
      }
    }
  }
}
__device__ __forceinline__ float hf_reduce(const float* red, int mr, int ncols, int r, int col) {
  float s = 0.0f;
#pragma unroll
  for (int w = 0; w < HF_WARPS; ++w) s += red[(static_cast<size_t>(w) * mr + r) * ncols + col];
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// Operand staging.  The bf16 operand of a phase ([rows, K], <= 32 rows) is needed in full by EVERY CTA (each owns a slice of the
// output columns).  If every CTA built it from the fp32 sources itself, the grid would pull rows x K x 4 B x 148 CTAs through
// L2 per phase (215 MB per probe at 28 rows: measured 150 us per probe, L2-bound) and normalise every row 148 times.  Instead
// the kernel runs in thread-block CLUSTERS: CTA k of a cluster stages the rows r = k (mod cluster size) -- one L2 read,
// one LayerNorm -- and writes the bf16 result into the shared memory of all CTAs of its cluster (st.shared::cluster); a
// cluster barrier publishes it.  L2 traffic and LayerNorm work drop by the cluster size (8 where the GPU can co-schedule
// enough clusters, else 4, else 1).
//
// Every global load of a staging job is issued BEFORE its first store: the compiler cannot hoist a load above a store it may
// alias, so a load-convert-store loop would pay one L2 round trip per iteration.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hf_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t hf_mapa(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void hf_st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t hf_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t hf_cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void hf_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void hf_st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// One warp, one job: `len` (<= 1024, % 256 == 0) fp32 values at src (global; may have been written by other CTAs: L2-coherent
// loads) -> [LayerNorm over exactly these len values (two-pass fp32, eps 1e-5) (+ReLU) when gamma != nullptr] -> bf16 -> byte
// offset `off` of the operand buffer of every CTA of the cluster (peer[d] = cluster address of that CTA's buffer).  A lane owns
// 8 consecutive values per 256-wide group, so a remote store carries 16 bytes.
__device__ __forceinline__ void hf_stage_job(const float* src, int len, const float* __restrict__ gamma, const float* __restrict__ beta,
                                             bool relu, uint32_t off, const uint32_t (&peer)[8], int cs, int lane) {
  float4 v[8], g[8], bt[8];
  const int ng = len >> 8;                                 // 256-wide groups; v[2i], v[2i+1] = the lane's 8 values of group i
  const bool ln = gamma != nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = i * 256 + lane * 8 + h * 4;
      v[2 * i + h] = (i < ng) ? __ldcg(reinterpret_cast<const float4*>(src + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[2 * i + h] = (ln && i < ng) ? __ldg(reinterpret_cast<const float4*>(gamma + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
      bt[2 * i + h] = (ln && i < ng) ? __ldg(reinterpret_cast<const float4*>(beta + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float mean = 0.0f, rstd = 1.0f;
  if (ln) {
    const float inv_len = 1.0f / static_cast<float>(len);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    mean = warp_sum(s) * inv_len;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if ((i >> 1) < ng) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    }
    rstd = rsqrtf(warp_sum(sq) * inv_len + 1e-5f);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < ng) {
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 2 * i + h;
        float4 y = v[k];
        if (ln) {
          y.x = (v[k].x - mean) * rstd * g[k].x + bt[k].x; y.y = (v[k].y - mean) * rstd * g[k].y + bt[k].y;
          y.z = (v[k].z - mean) * rstd * g[k].z + bt[k].z; y.w = (v[k].w - mean) * rstd * g[k].w + bt[k].w;
          if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
        }
        w[2 * h] = pack_bf16x2(y.x, y.y); w[2 * h + 1] = pack_bf16x2(y.z, y.w);
      }
      const uint32_t o = off + static_cast<uint32_t>((i * 256 + lane * 8) * 2);
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (d < cs) hf_st_cluster_v4(peer[d] + o, w[0], w[1], w[2], w[3]);
    }
  }
}

#define HF_MARK(i)                                                                                            \
  do {                                                                                                        \
    if (p.dbg != nullptr && cta == 0 && threadIdx.x == 0) {                                                   \
      unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); p.dbg[(i)] = (long long)_t; \
    }                                                                                                         \
  } while (0)

__global__ void __launch_bounds__(HF_THREADS, 1)
head_fused_kernel(const HeadFusedParams p) {
  extern __shared__ __align__(16) uint8_t hf_smem[];
  bf16* A_s = reinterpret_cast<bf16*>(hf_smem);
  float* red = reinterpret_cast<float*>(hf_smem);          // aliases A_s once a GEMM has consumed it
  __shared__ int s_slot[32];
  const int n = dyn_rows(p.n_host, p.n_dev, 1);
  if (n <= 0) return;                                      // uniform over the grid: nobody touches a barrier
  const int G = gridDim.x, cta = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int H = p.H, D = p.D, LL = p.LL, d0 = p.d0, d1 = p.d1, lda = p.lda, mr = p.mr;
  const bool two_m = n > 16;
  const int rows_used = two_m ? 32 : 16;
  const int cs = static_cast<int>(hf_cluster_size());      // 1, 4 or 8 (divides 16)
  const int crank = static_cast<int>(hf_cluster_rank());
  const int my_rows = rows_used / cs;                      // rows this CTA stages for its cluster: r = crank + cs * i
  uint32_t peer[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) peer[d] = hf_mapa(hf_smem_u32(A_s), static_cast<uint32_t>(d < cs ? d : 0));
  const unsigned long long per_launch = static_cast<unsigned long long>(G) * (LL + 3);
  unsigned long long bar_target = 0;
  if (threadIdx.x == 0) bar_target = (ld_acquire_u64(p.bar) / per_launch) * per_launch;     // < G arrivals of this launch so far
  int n_bar = 0;
  const int nu = H / 8;                                    // LSTM work items: 8 hidden units (x 4 gates = 32 weight rows)
  const int n1 = 2 * d0 / 16;                              // MLP1 items: 16 columns
  const int n2 = 2 * d1 / 8;                               // MLP2 items: 8 columns
  if (threadIdx.x < 32) s_slot[threadIdx.x] = (threadIdx.x < n) ? p.slot[threadIdx.x] : 0;
  hf_cluster_sync();                                       // every CTA of the cluster is running before anyone addresses a peer's shared memory

  HF_MARK(0);
  // ---- L2 prefetch of every weight row this CTA will stream, in phase order (82 MB over the grid; L2 holds 126 MB)
  for (int l = 0; l < LL; ++l) {
    const int K = (l == 0 ? D : H) + H;
    const int lines = (K * 2 + 127) / 128;
    for (int it = cta; it < nu; it += G)
      for (int i = threadIdx.x; i < 32 * lines; i += HF_THREADS) {
        const int rr = i / lines, ln = i % lines;
        prefetch_l2(reinterpret_cast<const char*>(p.w_cat[l] + static_cast<size_t>((rr >> 3) * H + it * 8 + (rr & 7)) * K) + ln * 128);
      }
  }
  {
    const int lines = (H * 2 + 127) / 128;
    for (int it = cta; it < n1; it += G)
      for (int i = threadIdx.x; i < 16 * lines; i += HF_THREADS)
        prefetch_l2(reinterpret_cast<const char*>(p.w_mlp1 + static_cast<size_t>(it * 16 + i / lines) * H) + (i % lines) * 128);
    const int lines2 = (d0 * 2 + 127) / 128;
    for (int it = cta; it < n2; it += G) {
      const int head = it / (d1 / 8), c0 = (it % (d1 / 8)) * 8;
      const bf16* w = head ? p.w_mlp2_grip : p.w_mlp2_pose;
      for (int i = threadIdx.x; i < 8 * lines2; i += HF_THREADS)
        prefetch_l2(reinterpret_cast<const char*>(w + static_cast<size_t>(c0 + i / lines2) * d0) + (i % lines2) * 128);
    }
  }

  // ---- phase 0: AdaptiveMaxPool1d over the valid tokens (action_head.py:519-520)
  {
    const int d4 = D >> 2;
    for (int idx = cta * HF_THREADS + threadIdx.x; idx < n * d4; idx += G * HF_THREADS) {
      const int r = idx / d4, j = (idx % d4) * 4;
      const int vl = p.valid_len[r];
      float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      const float* xr = p.x + static_cast<size_t>(r) * p.S * D + j;
      for (int s0 = 0; s0 < vl; s0 += 8) {                 // eight token rows in flight per thread
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = (s0 + u < vl) ? *reinterpret_cast<const float4*>(xr + static_cast<size_t>(s0 + u) * D) : m;
#pragma unroll
        for (int u = 0; u < 8; ++u) { m.x = fmaxf(m.x, v[u].x); m.y = fmaxf(m.y, v[u].y); m.z = fmaxf(m.z, v[u].z); m.w = fmaxf(m.w, v[u].w); }
      }
      *reinterpret_cast<float4*>(p.pooled + static_cast<size_t>(r) * D + j) = m;
    }
  }
#define HF_BARRIER()                                                          \
  do {                                                                        \
    HF_MARK(2 * n_bar + 1);                                                   \
    ++n_bar;                                                                  \
    hf_grid_barrier(p.bar, bar_target + static_cast<unsigned long long>(n_bar) * G); \
    HF_MARK(2 * n_bar);                                                       \
  } while (0)
  // rows >= n of the operand are zero: every CTA clears its own copy (peers only ever write rows < n)
#define HF_ZERO_TAIL(KLEN)                                                                               \
  do {                                                                                                   \
    for (int r = n + warp; r < rows_used; r += HF_WARPS)                                                 \
      for (int j = lane * 8; j < (KLEN); j += 256) *reinterpret_cast<uint4*>(A_s + static_cast<size_t>(r) * lda + j) = make_uint4(0u, 0u, 0u, 0u); \
  } while (0)
  // weight pointers / first batches: loaded ahead of the barrier that precedes their phase
  uint4 pre4[4][4];
  const bf16* wp4[4];
  const bool lstm_on = cta < nu;
#pragma unroll
  for (int q = 0; q < 4; ++q) wp4[q] = p.w_cat[0] + static_cast<size_t>(q * H + (lstm_on ? cta : 0) * 8 + g) * (D + H) + t * 8;
  if (lstm_on) hf_load_batch<4>(pre4, wp4, warp, (D + H) >> 5);
  HF_BARRIER();

  uint4 pre2[4][2];
  const bf16* wp2[2];
  // ---- LSTM layers
  for (int l = 0; l < LL; ++l) {
    const int in = (l == 0) ? D : H, K = in + H;
    // this thread's cell of the layer's epilogue: bias and carried c are fetched now, off the critical path
    const int e_r = threadIdx.x >> 3, e_u = threadIdx.x & 7;
    const bool e_on = (cta < nu) && e_r < n;
    float e_b[4] = {0.f, 0.f, 0.f, 0.f}, e_c = 0.0f;
    if (e_on) {
      const int j = cta * 8 + e_u;
#pragma unroll
      for (int q = 0; q < 4; ++q) e_b[q] = __ldg(p.b_sum[l] + q * H + j);
      e_c = p.c_state[(static_cast<size_t>(s_slot[e_r]) * LL + l) * H + j];
    }
    // operand [rows, in | H]: layer input (pooled feature / LayerNorm of the previous layer's h) | carried h of this layer.
    // jobs of this CTA: (row i, part): parts = in / 1024 pooled pieces (l == 0) or one LayerNorm row, then the h row
    {
      const int xparts = (l == 0) ? (D + 1023) / 1024 : 1;
      const int njobs = my_rows * (xparts + 1);
      for (int jb = warp; jb < njobs; jb += HF_WARPS) {
        const int i = jb / (xparts + 1), part = jb % (xparts + 1);
        const int r = crank + cs * i;
        if (r >= n) continue;                              // warp-uniform
        const uint32_t row_off = static_cast<uint32_t>(r) * lda * 2;
        if (part == xparts) {
          hf_stage_job(p.h_state + (static_cast<size_t>(s_slot[r]) * LL + l) * H, H, nullptr, nullptr, false, row_off + in * 2, peer, cs, lane);
        } else if (l == 0) {
          const int c0 = part * 1024, len = min(1024, D - c0);
          hf_stage_job(p.pooled + static_cast<size_t>(r) * D + c0, len, nullptr, nullptr, false, row_off + c0 * 2, peer, cs, lane);
        } else {
          hf_stage_job(p.cand_h + (static_cast<size_t>(l - 1) * p.nstride + r) * H, H, p.ln_g[l - 1], p.ln_b[l - 1], false, row_off, peer, cs, lane);
        }
      }
      HF_ZERO_TAIL(K);
    }
    hf_cluster_sync();
    if (l == 1) HF_MARK(16);
    if (lstm_on) {
      const int j0 = cta * 8;
      float acc[2][4][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.0f;
      hf_skinny_gemm<4>(A_s + static_cast<size_t>(g) * lda + t * 8, lda, two_m, wp4, K, warp, pre4, acc);
      __syncthreads();                                     // every warp is done reading A_s (red aliases it)
      if (l == 1) HF_MARK(17);
      hf_store_partials<4>(red, mr, two_m, warp, lane, acc);
      __syncthreads();
      // LSTM cell (gate order i, f, g, o): thread = (row, unit)
      if (e_on) {
        const int j = j0 + e_u;
        const float gi = hf_reduce(red, mr, 32, e_r, e_u) + e_b[0];
        const float gf = hf_reduce(red, mr, 32, e_r, 8 + e_u) + e_b[1];
        const float gg = hf_reduce(red, mr, 32, e_r, 16 + e_u) + e_b[2];
        const float go = hf_reduce(red, mr, 32, e_r, 24 + e_u) + e_b[3];
        const float c2 = sigmoidf_(gf) * e_c + sigmoidf_(gi) * tanhf(gg);
        const float h2 = sigmoidf_(go) * tanhf(c2);
        p.cand_c[(static_cast<size_t>(l) * p.nstride + e_r) * H + j] = c2;
        p.cand_h[(static_cast<size_t>(l) * p.nstride + e_r) * H + j] = h2;
      }
      if (l == 1) HF_MARK(18);
    }
    // next phase's first weight batch (the host guarantees one work item per CTA and phase: H / 8, 2*d0 / 16, 2*d1 / 8 <= gridDim.x)
    if (l + 1 < LL) {
#pragma unroll
      for (int q = 0; q < 4; ++q) wp4[q] = p.w_cat[l + 1] + static_cast<size_t>(q * H + (lstm_on ? cta : 0) * 8 + g) * (2 * H) + t * 8;
      if (lstm_on) hf_load_batch<4>(pre4, wp4, warp, (2 * H) >> 5);
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) wp2[q] = p.w_mlp1 + static_cast<size_t>((cta < n1 ? cta : 0) * 16 + q * 8 + g) * H + t * 8;
      if (cta < n1) hf_load_batch<2>(pre2, wp2, warp, H >> 5);
    }
    HF_BARRIER();
  }

  uint4 pre1[4][1];
  const bf16* wp1[1];
  // ---- MLP layer 1 of both heads: [n, H] x [2*d0, H]^T, operand = LayerNorm of the last LSTM layer's h
  {
    for (int i = warp; i < my_rows; i += HF_WARPS) {
      const int r = crank + cs * i;
      if (r >= n) continue;
      hf_stage_job(p.cand_h + (static_cast<size_t>(LL - 1) * p.nstride + r) * H, H, p.ln_g[LL - 1], p.ln_b[LL - 1], false,
                   static_cast<uint32_t>(r) * lda * 2, peer, cs, lane);
    }
    HF_ZERO_TAIL(H);
    hf_cluster_sync();
    if (cta < n1) {
      const int it = cta;
      float acc[2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.0f;
      hf_skinny_gemm<2>(A_s + static_cast<size_t>(g) * lda + t * 8, lda, two_m, wp2, H, warp, pre2, acc);
      __syncthreads();
      hf_store_partials<2>(red, mr, two_m, warp, lane, acc);
      __syncthreads();
      for (int idx = threadIdx.x; idx < n * 16; idx += HF_THREADS) {
        const int r = idx >> 4, cc = idx & 15, col = it * 16 + cc;
        p.mlp1[static_cast<size_t>(r) * 2 * d0 + col] = hf_reduce(red, mr, 16, r, cc) + __ldg(p.b_mlp1 + col);
      }
    }
    {
      const int it = cta < n2 ? cta : 0;
      wp1[0] = ((it / (d1 / 8)) ? p.w_mlp2_grip : p.w_mlp2_pose) + static_cast<size_t>((it % (d1 / 8)) * 8 + g) * d0 + t * 8;
      if (cta < n2) hf_load_batch<1>(pre1, wp1, warp, d0 >> 5);
    }
    HF_BARRIER();
  }

  // ---- MLP layer 2: per head [n, d0] x [d1, d0]^T, operand = ReLU(LayerNorm(mlp1 half))
  {
    for (int jb = warp; jb < 2 * my_rows; jb += HF_WARPS) {
      const int r = crank + cs * (jb >> 1), head = jb & 1;
      if (r >= n) continue;
      hf_stage_job(p.mlp1 + static_cast<size_t>(r) * 2 * d0 + head * d0, d0, p.g_mlp1 + head * d0, p.be_mlp1 + head * d0, true,
                   static_cast<uint32_t>(r) * lda * 2 + head * d0 * 2, peer, cs, lane);
    }
    HF_ZERO_TAIL(2 * d0);
    hf_cluster_sync();
    if (cta < n2) {
      const int it = cta;
      const int head = it / (d1 / 8), c0 = (it % (d1 / 8)) * 8;
      float acc[2][1][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][0][c] = 0.0f;
      hf_skinny_gemm<1>(A_s + static_cast<size_t>(g) * lda + head * d0 + t * 8, lda, two_m, wp1, d0, warp, pre1, acc);
      __syncthreads();
      hf_store_partials<1>(red, mr, two_m, warp, lane, acc);
      __syncthreads();
      for (int idx = threadIdx.x; idx < n * 8; idx += HF_THREADS) {
        const int r = idx >> 3, cc = idx & 7, col = head * d1 + c0 + cc;
        p.mlp2[static_cast<size_t>(r) * 2 * d1 + col] = hf_reduce(red, mr, 8, r, cc) + __ldg(p.b_mlp2 + col);
      }
    }
    HF_BARRIER();
  }

  // ---- output layers: ReLU(LayerNorm(mlp2 half)) . w3 -> tanh (6 pose dims) / sigmoid (gripper); one CTA per row
  for (int r = cta; r < n; r += G) {
    if (warp < 7) {
      const int head = (warp == 6) ? 1 : 0;
      const float* src = p.mlp2 + static_cast<size_t>(r) * 2 * d1 + head * d1;
      const float* gam = p.g_mlp2 + head * d1;
      const float* bet = p.be_mlp2 + head * d1;
      const float* w = head ? p.w3_grip : p.w3_pose + static_cast<size_t>(warp) * d1;
      float v[16], gv[16], bv[16], wv[16];
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = lane + 32 * i;
        const bool on = j < d1;
        v[i] = on ? __ldcg(src + j) : 0.0f;
        gv[i] = on ? __ldg(gam + j) : 0.0f; bv[i] = on ? __ldg(bet + j) : 0.0f; wv[i] = on ? __ldg(w + j) : 0.0f;
        s += v[i];
      }
      const float mean = warp_sum(s) / static_cast<float>(d1);
      float q = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float d = (lane + 32 * i < d1) ? v[i] - mean : 0.0f;
        q += d * d;
      }
      const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(d1) + 1e-5f);
      float acc = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (lane + 32 * i < d1) acc += fmaxf((v[i] - mean) * rstd * gv[i] + bv[i], 0.0f) * wv[i];
      acc = warp_sum(acc);
      if (lane == 0) {
        if (head) p.cur_grip[r] = sigmoidf_(acc + p.b3_grip[0]);
        else p.cur_pose[r * 6 + warp] = tanhf(acc + p.b3_pose[warp]);
      }
    }
  }
  HF_MARK(2 * n_bar + 1);
  // every launch adds exactly gridDim.x * (LL + 3) barrier arrivals (n_bar == LL + 3 here)
  hf_cluster_sync();                                       // no CTA of a cluster exits while a peer could still address its shared memory
#undef HF_BARRIER
#undef HF_ZERO_TAIL
#undef HF_MARK
}
