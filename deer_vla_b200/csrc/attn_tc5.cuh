// attn_tc5_kernel: the ViT attention (softmax(scale Q K^T) V, head_dim 64, no mask, Sk <= 264) on tcgen05 -- schedule v5.
//
// Why a new schedule.  ncu of the v3/v4 kernel (profiles/r01_attn_v4_ncu_summary.txt) shows it is neither MUFU- nor TMEM-bound
// but LATENCY-bound: the eight softmax warps march through the same dependent phases in lock-step (score load -> max ->
// bar.sync exchange with the partner warp -> exp2 -> pack -> store), two warps per scheduler, issue slots 45 % busy, MUFU pipe
// 31 % busy, and a quarter of all executed instructions is barrier polling.
//
// v5: TWO INDEPENDENT GROUPS of four warps.  Group g owns the tiles with (tile index & 1) == g and the TMEM buffer g, so the
// groups are naturally half a tile out of phase: while one waits for its PV / the next QK or drains its output, the other keeps
// the MUFU and TMEM-read pipes busy.  Inside a group ONE THREAD OWNS ONE FULL ROW (TMEM lane): no partner warp, no shared-memory
// exchange, no named barrier.  The 256 scores of the row arrive as eight 32-column chunks; chunk k+1 is in flight while chunk k
// goes through max -> exp2 -> pack, and its probabilities are written straight back IN PLACE (bf16 pairs over score columns the
// thread itself has already consumed), so a thread holds two chunks, not the row.
//
// Exact softmax without a first pass over the row: probabilities are taken relative to a running reference `ref` (scaled log2
// domain) that starts as the max of chunk 0 (and the tail keys) and is left alone as long as no later chunk max exceeds it by
// more than 2^60 -- then every stored p is <= 2^60, far inside bf16 / fp32 range, and O / l is scale invariant.  If a chunk max
// does exceed ref + 60 (logits > 41 nats apart inside one row), the warp takes a slow path: everything stored so far (this
// thread's own P columns, read back from tensor memory) and the running sum are multiplied by the exact power of two
// 2^(ref - M) and ref moves to M.  The branch is warp-uniform (tcgen05.ld / st are .sync.aligned): lanes that do not need it
// rescale by 1.
//
// Keys 256..Sk-1 (at most 8; the 257th token of a ViT image) and the 1-8 query rows that do not fill a 128-row tile are handled
// on CUDA cores exactly as in v3/v4.  Output rows are staged through shared memory so that global stores are row-contiguous.
// warps: 0 TMA producer, 1 MMA issuer + TMEM owner, 2-5 softmax group 0, 6-9 softmax group 1, 10 leftover rows.
#pragma once
#include "attn_tc.cuh"

struct AttnTc5 {
  using A = AttnTc;
  static constexpr int OST_ROW = 144;                            // staged output row: 128 B of bf16 + 16 B pad (odd multiple of 16 B)
  static constexpr int OFF_K = 0, OFF_V = 2 * A::KV_BYTES, OFF_Q = 4 * A::KV_BYTES;
  static constexpr int OFF_P = OFF_Q + 2 * A::Q_BYTES;           // fp32 scratch of the spare warp
  static constexpr int OFF_XTAIL = OFF_P + A::SKP * 4 + 64;      // [2 buf][MAX_TAIL][128] raw tail scores, then their probabilities
  static constexpr int OFF_OST = OFF_XTAIL + 2 * A::MAX_TAIL * 128 * 4;   // [8 warps][32 rows][OST_ROW]
  static constexpr int OFF_BAR = OFF_OST + 8 * 32 * OST_ROW;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int THREADS = 11 * 32;
  static constexpr int TMEM_COLS = 512;
  static constexpr float REF_SLACK = 60.0f;                      // log2 units a chunk max may exceed the reference before it moves
};
static_assert(AttnTc5::SMEM_BYTES <= 227 * 1024, "attn_tc5 shared memory");

__global__ void __launch_bounds__(AttnTc5::THREADS, 1)
attn_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const bf16* __restrict__ Qg, int ldq, bf16* __restrict__ Og, int ldo, int B, int H, int Sq, int Sk, float scale_log2,
                long long* __restrict__ dbg, int rev) {
  using A = AttnTc;
  using A5 = AttnTc5;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A5::OFF_BAR);
  uint64_t* kv_full = bars;          // [2]  K/V of an item landed
  uint64_t* kv_empty = bars + 2;     // [2]  MMA commit + every softmax warp once per tile of the item + spare warp
  uint64_t* q_full = bars + 4;       // [2]
  uint64_t* q_empty = bars + 6;      // [2]  MMA commit (QK done) + the 4 warps of the tile's group (tail scoring reads Q)
  uint64_t* s_full = bars + 8;       // [2]  per TMEM buffer = per group
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
  // rev: walk the (image, head) items from the LAST image down -- the q/k/v rows the preceding GEMM wrote last are still in L2,
  // and the rows written last here are the ones the following GEMM reads first ("snake" order)
  auto item_of = [&](int is) { const int i = static_cast<int>(blockIdx.x) + is * static_cast<int>(gridDim.x); return rev ? n_items - 1 - i : i; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 4 * n_sub + 2);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 5);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, A5::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer: K/V once per item (double buffered across items), Q once per tile.  The warp stays converged and one
    // elected lane issues (gemm.cuh: elect_one_sync) =====
    {
      for (int is = 0; is < my_items; ++is) {
        const int it = item_of(is);
        const int b = it / H, h = it % H;
        const int kb = is & 1;
        mbar_wait(&kv_empty[kb], ((is >> 1) & 1) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&kv_full[kb], 2 * A::KV_BYTES);
          uint8_t* sk = smem + A5::OFF_K + kb * A::KV_BYTES;
          uint8_t* sv = smem + A5::OFF_V + kb * A::KV_BYTES;
          tma_load_2d(sk, &tmK, &kv_full[kb], h * 64, b * Sk);
          tma_load_2d(sk + A::HALF_BYTES, &tmK, &kv_full[kb], h * 64, b * Sk + A::HALF_ROWS);
          tma_load_2d(sv, &tmV, &kv_full[kb], h * 64, b * Sk);
          tma_load_2d(sv + A::HALF_BYTES, &tmV, &kv_full[kb], h * 64, b * Sk + A::HALF_ROWS);
        }
        __syncwarp();
        for (int t = 0; t < n_sub; ++t) {
          const int n = is * n_sub + t, qb = n & 1;
          mbar_wait(&q_empty[qb], ((n >> 1) & 1) ^ 1u);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&q_full[qb], A::Q_BYTES);
            tma_load_2d(smem + A5::OFF_Q + qb * A::Q_BYTES, &tmQ, &q_full[qb], h * 64, b * Sq + t * 128);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  The two softmax groups are independent, so the issuer must not make one wait for the other: QK(nq) and
    // PV(np) are issued in whichever order their preconditions come true (a fixed QK(n+1)-then-PV(n) order left the group that
    // had finished its probabilities waiting until the OTHER group had drained its previous tile).  The warp stays converged --
    // lane 0's barrier tests decide for all lanes -- and one elected lane issues: from a `lane == 0` branch every tcgen05.mma sits in
    // a serialisation loop (~70 cycles each, 16 PV steps per tile). =====
    {
      constexpr uint32_t idesc_s = make_idesc_bf16_major(128, A::NMMA, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(128, 64, 0, 1);     // B = V, MN-major
      int nq = 0, np = 0;                                  // next QK / next PV to issue
      uint32_t idle = 0;
      while (np < total) {
        bool did = false;
        if (nq < total && nq < np + 2) {                   // at most two tiles in flight (two TMEM buffers)
          const int is = nq / n_sub, t = nq - is * n_sub, kb = is & 1, qb = nq & 1, u = nq & 1;
          const bool ok0 = (t != 0 || mbar_test(&kv_full[kb], (is >> 1) & 1)) && mbar_test(&q_full[qb], (nq >> 1) & 1) &&
                           (nq < 2 || mbar_test(&o_empty[u], ((nq >> 1) - 1) & 1));      // the group has drained its previous tile
          const bool ok = __shfl_sync(0xffffffffu, ok0 ? 1 : 0, 0) != 0;
          if (ok) {
            tc_fence_after();
            const uint32_t q_addr = smem_u32(smem + A5::OFF_Q + qb * A::Q_BYTES);
            const uint32_t k_addr = smem_u32(smem + A5::OFF_K + kb * A::KV_BYTES);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + static_cast<uint32_t>(256 * u), make_desc_sw128(q_addr) + static_cast<uint64_t>(2 * k),
                          make_desc_sw128(k_addr) + static_cast<uint64_t>(2 * k), idesc_s, k != 0 ? 1u : 0u);
              umma_commit(&s_full[u]);
              umma_commit(&q_empty[qb]);
            }
            __syncwarp();
            ++nq;
            did = true;
          }
        }
        if (np < nq) {
          const int is = np / n_sub, t = np - is * n_sub, kb = is & 1, u = np & 1;
          const bool okp = __shfl_sync(0xffffffffu, mbar_test(&p_full[u], (np >> 1) & 1) ? 1 : 0, 0) != 0;
          if (okp) {
            tc_fence_after();
            const uint32_t v_addr = smem_u32(smem + A5::OFF_V + kb * A::KV_BYTES);
            if (elect_one_sync()) {
#pragma unroll 1
              for (int k = 0; k < pv_steps; ++k)
                umma_bf16_ts(tmem_base + static_cast<uint32_t>(256 * u + 128), tmem_base + static_cast<uint32_t>(256 * u + 8 * k),
                             make_desc(v_addr + static_cast<uint32_t>(k) * 2048u, 1024u, 1024u), idesc_pv, k != 0 ? 1u : 0u);
              umma_commit(&o_full[u]);
              if (t == n_sub - 1) umma_commit(&kv_empty[kb]);
            }
            __syncwarp();
            ++np;
            did = true;
          }
        }
        if (did) idle = 0;
        else {
          __nanosleep(32);
          if (++idle > (1u << 26)) __trap();               // a protocol bug must surface as a launch failure, never as a hung GPU
        }
      }
    }
  } else if (warp < 10) {
    // ===== softmax groups =====
    const int g = (warp - 2) >> 2;                         // group = TMEM buffer = tile parity
    const int qd = warp & 3;                               // TMEM lane quarter
    const int rloc = qd * 32 + lane;                       // this thread's row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tbuf = tmem_base + lane_off + static_cast<uint32_t>(256 * g);
    float* xtail = reinterpret_cast<float*>(smem + A5::OFF_XTAIL) + g * (A::MAX_TAIL * 128);
    uint8_t* ost = smem + A5::OFF_OST + (warp - 2) * (32 * A5::OST_ROW);
    const int qb = g, u = g;
#ifdef DEER_ATTN_PROF
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool prof = (dbg != nullptr) && warp == 2 && lane == 0;
#define TICK5(i) do { if (prof) { const long long _t = clock64(); tacc[i] += _t - t0; t0 = _t; } } while (0)
    long long t0 = prof ? clock64() : 0;
#else
#define TICK5(i) do { } while (0)
#endif
    for (int n = g; n < total; n += 2) {
      const int is = n / n_sub, t = n - is * n_sub, kb = is & 1;
      const uint32_t par = static_cast<uint32_t>((n >> 1) & 1);
      const uint8_t* sk = smem + A5::OFF_K + kb * A::KV_BYTES;
      const uint8_t* sv = smem + A5::OFF_V + kb * A::KV_BYTES;
      mbar_wait(&s_full[u], par);
      tc_fence_after();
      TICK5(0);
      float sa[32], sb[32];
      tmem_ld_32x32(tbuf, sa);
      // tail keys (>= 256) on CUDA cores while the first chunk is in flight: this row's q (smem) . staged K rows
      float m = -INFINITY;
      {
        const uint8_t* sq = smem + A5::OFF_Q + qb * A::Q_BYTES;
#pragma unroll 1
        for (int j = 0; j < n_tail; ++j) {                 // raw scores parked in smem (each thread reads back only its own slot)
          const float sj = dot64_swz(sq, rloc, sk, A::NMMA + j);
          xtail[j * 128 + rloc] = sj;
          m = fmaxf(m, sj);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_empty[qb]);
      }
      float ref = 0.0f, l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
      // slow path (warp-uniform): move the reference to `mnew`, rescaling the probabilities already stored (chunks < k) and l
      auto rescale = [&](int k, float mnew, bool need) {
        const float f = need ? ex2_approx(ref - mnew) : 1.0f;
        if (need) ref = mnew;
        l0 *= f; l1 *= f; l2 *= f; l3 *= f;
        tmem_st_wait();                                    // the chunks written so far must have landed before they are read back
#pragma unroll 1
        for (int j = 0; j < k; ++j) {
          uint32_t w[16];
          float* wf = reinterpret_cast<float*>(w);
          tmem_ld_32x32_x16(tbuf + static_cast<uint32_t>(16 * j), *reinterpret_cast<float(*)[16]>(wf));
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float lo = __uint_as_float(w[i] << 16) * f, hi = __uint_as_float(w[i] & 0xffff0000u) * f;
            w[i] = pack_bf16x2(lo, hi);
          }
          tmem_st_32x32b_x16(tbuf + static_cast<uint32_t>(16 * j), w);
        }
      };
#define STAGE5(SX, K)                                                                                          \
      {                                                                                                        \
        if (32 * ((K) + 1) > n_mma) {                                                                          \
          _Pragma("unroll") for (int j = 0; j < 32; ++j) if (32 * (K) + j >= n_mma) SX[j] = -INFINITY;         \
        }                                                                                                      \
        float cm = ((K) == 0) ? m : -INFINITY;                                                                 \
        _Pragma("unroll") for (int j = 0; j < 32; j += 2) cm = fmaxf(cm, fmaxf(SX[j], SX[j + 1]));             \
        cm *= scale_log2;                                                                                      \
        if ((K) == 0) ref = cm;                                                                                \
        else {                                                                                                 \
          const bool need = (cm - ref) > A5::REF_SLACK;                                                        \
          if (__any_sync(0xffffffffu, need)) rescale((K), cm, need);                                           \
        }                                                                                                      \
        uint32_t pk[16];                                                                                       \
        _Pragma("unroll") for (int j = 0; j < 32; j += 4) {                                                    \
          const float p0 = ex2_approx(fmaf(SX[j], scale_log2, -ref)), p1 = ex2_approx(fmaf(SX[j + 1], scale_log2, -ref)); \
          const float p2 = ex2_approx(fmaf(SX[j + 2], scale_log2, -ref)), p3 = ex2_approx(fmaf(SX[j + 3], scale_log2, -ref)); \
          l0 += p0; l1 += p1; l2 += p2; l3 += p3;                                                              \
          pk[j >> 1] = pack_bf16x2(p0, p1);                                                                    \
          pk[(j >> 1) + 1] = pack_bf16x2(p2, p3);                                                              \
        }                                                                                                      \
        tmem_st_32x32b_x16(tbuf + static_cast<uint32_t>(16 * (K)), pk);                                        \
      }
      // chunk k+1 is requested before chunk k is processed; P chunk k lands on score columns [16k, 16k+16), all inside chunks <= k.
      // Rolled (two stages per iteration, one per ping-pong buffer): the fully unrolled body did not fit the instruction cache
      // (ncu: 19 % of the warp stall samples were "no instruction").
#pragma unroll 1
      for (int kk = 0; kk < 4; ++kk) {
        tmem_ld_wait(); tmem_ld_32x32(tbuf + static_cast<uint32_t>(64 * kk + 32), sb);
        STAGE5(sa, 2 * kk)
        tmem_ld_wait();
        if (kk < 3) tmem_ld_32x32(tbuf + static_cast<uint32_t>(64 * kk + 64), sa);
        STAGE5(sb, 2 * kk + 1)
      }
#undef STAGE5
      float l = (l0 + l1) + (l2 + l3);
#pragma unroll 1
      for (int j = 0; j < n_tail; ++j) {
        const float pj = ex2_approx(fmaf(xtail[j * 128 + rloc], scale_log2, -ref));
        l += pj;
        xtail[j * 128 + rloc] = pj;
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[u]);
      TICK5(1);
      // ---- output of this tile: (O + tail keys) / l -> bf16 -> staging -> row-contiguous global stores.  The other group
      // keeps the SM busy while this one waits for the PV MMA.
      const int it = item_of(is);
      const int b = it / H, h = it - b * H;
      const float inv = 1.0f / l;
      mbar_wait(&o_full[u], par);
      tc_fence_after();
      TICK5(2);
      {
        float va[32], vb[32];                              // the whole 64-column O row, both loads in flight together
        tmem_ld_32x32(tbuf + 128u, va);
        tmem_ld_32x32(tbuf + 160u, vb);
        tmem_ld_wait();
        for (int j = 0; j < n_tail; ++j) {
          const int kr = A::NMMA + j;
          const float pj = xtail[j * 128 + rloc];
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {                 // eight 16-byte chunks = 64 d; every lane reads the same address (broadcast)
            const uint4 y = *reinterpret_cast<const uint4*>(sv + kr * 128 + ((cc ^ (kr & 7)) << 4));
            const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(y2[e]);
              if (cc < 4) {
                va[cc * 8 + 2 * e] = fmaf(pj, f.x, va[cc * 8 + 2 * e]);
                va[cc * 8 + 2 * e + 1] = fmaf(pj, f.y, va[cc * 8 + 2 * e + 1]);
              } else {
                vb[(cc - 4) * 8 + 2 * e] = fmaf(pj, f.x, vb[(cc - 4) * 8 + 2 * e]);
                vb[(cc - 4) * 8 + 2 * e + 1] = fmaf(pj, f.y, vb[(cc - 4) * 8 + 2 * e + 1]);
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 o;
          o.x = pack_bf16x2(va[j] * inv, va[j + 1] * inv); o.y = pack_bf16x2(va[j + 2] * inv, va[j + 3] * inv);
          o.z = pack_bf16x2(va[j + 4] * inv, va[j + 5] * inv); o.w = pack_bf16x2(va[j + 6] * inv, va[j + 7] * inv);
          *reinterpret_cast<uint4*>(ost + lane * A5::OST_ROW + j * 2) = o;
          o.x = pack_bf16x2(vb[j] * inv, vb[j + 1] * inv); o.y = pack_bf16x2(vb[j + 2] * inv, vb[j + 3] * inv);
          o.z = pack_bf16x2(vb[j + 4] * inv, vb[j + 5] * inv); o.w = pack_bf16x2(vb[j + 6] * inv, vb[j + 7] * inv);
          *reinterpret_cast<uint4*>(ost + lane * A5::OST_ROW + 64 + j * 2) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&o_empty[u]);                          // TMEM buffer free: QK(n+2) may start while the rows are being stored
        mbar_arrive(&kv_empty[kb]);                        // last read of this item's K / V by this warp for this tile
      }
      {
        const int ch = lane & 7, rsub = lane >> 3;         // 8 lanes cover a row's 128 bytes, 4 rows per instruction
        bf16* obase = Og + (static_cast<size_t>(b) * Sq + t * 128 + qd * 32) * ldo + h * 64 + ch * 8;
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int r = itr * 4 + rsub;
          const uint4 o = *reinterpret_cast<const uint4*>(ost + r * A5::OST_ROW + ch * 16);
          if (t * 128 + qd * 32 + r < Sq) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(r) * ldo) = o;
        }
      }
      __syncwarp();                                        // the staging rows are rewritten by the next tile's output
      TICK5(3);
    }
#ifdef DEER_ATTN_PROF
    if (prof) for (int i = 0; i < 8; ++i) dbg[blockIdx.x * 8 + i] = tacc[i];
#endif
#undef TICK5
  } else {
    // ===== spare warp: the few query rows that do not fill a 128-row tile, on CUDA cores, against the smem K/V =====
    float* pbuf = reinterpret_cast<float*>(smem + A5::OFF_P);
    for (int is = 0; is < my_items; ++is) {
      const int it = item_of(is);
      const int b = it / H, h = it % H, kb = is & 1;
      mbar_wait(&kv_full[kb], (is >> 1) & 1);
      const uint8_t* sK = smem + A5::OFF_K + kb * A::KV_BYTES;
      const uint8_t* sV = smem + A5::OFF_V + kb * A::KV_BYTES;
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
        float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;  // lane owns d = 2*lane, 2*lane+1; two keys per iteration
        int j = 0;
        for (; j + 1 < Sk; j += 2) {
          const __nv_bfloat162 va = *reinterpret_cast<const __nv_bfloat162*>(sV + j * 128 + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4);
          const __nv_bfloat162 vb = *reinterpret_cast<const __nv_bfloat162*>(sV + (j + 1) * 128 + (((lane >> 2) ^ ((j + 1) & 7)) << 4) + (lane & 3) * 4);
          const float2 fa = __bfloat1622float2(va), fb = __bfloat1622float2(vb);
          const float pa = pbuf[j], pb = pbuf[j + 1];
          o0 = fmaf(pa, fa.x, o0); o1 = fmaf(pa, fa.y, o1);
          o2 = fmaf(pb, fb.x, o2); o3 = fmaf(pb, fb.y, o3);
        }
        if (j < Sk) {
          const __nv_bfloat162 va = *reinterpret_cast<const __nv_bfloat162*>(sV + j * 128 + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4);
          const float2 fa = __bfloat1622float2(va);
          const float pa = pbuf[j];
          o0 = fmaf(pa, fa.x, o0); o1 = fmaf(pa, fa.y, o1);
        }
        const float inv = 1.0f / l;
        *reinterpret_cast<uint32_t*>(Og + (static_cast<size_t>(b) * Sq + row) * ldo + h * 64 + 2 * lane) = pack_bf16x2((o0 + o2) * inv, (o1 + o3) * inv);
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
    tmem_dealloc(tmem_base, A5::TMEM_COLS);
  }
}
