// deer_engine: C-ABI implementation + orchestration of one DeeR action-step on a B200.
// See include/deer_engine.h for the contract and DESIGN.md for the data layout.
#include "../../include/deer_engine.h"
#include "attention.cuh"
#include "attn_tc.cuh"
#include "attn_tc5.cuh"
#include "gemm.cuh"
#include "head_fused.cuh"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <vector>

// ---------------------------------------------------------------------------------------------
// small host utilities
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int set_err(std::string* dst, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (dst) *dst = buf;
  g_last_error = buf;
  return code;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] with row pitch ld (elements); box = 64 x box_rows, 128B swizzle, OOB -> 0
static bool make_tmap(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// bf16 row-major output [rows, cols] with row pitch ld: box = 32 columns x 32 rows, 64-byte swizzle (the TMA-store epilogue's tile)
static bool make_tmap_c(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Launch with (optionally) programmatic stream serialization: the kernel may begin while its predecessor drains; the kernels
// launched this way all start with pdl_wait() (common.cuh).  DEER_PDL=0 turns it off (A/B), and it is never used during graph capture.
// "Snake" traversal: the streaming kernels between two GEMMs (LayerNorm rows, ViT attention items) walk their rows from the END, i.e.
// they start on what the producing GEMM wrote last (still in the 126 MB L2) and finish on what the consuming GEMM reads first.
// Row-independent kernels: results are bit-identical either way (DEER_SNAKE=0 for the A/B).
static bool g_snake = getenv("DEER_SNAKE") ? atoi(getenv("DEER_SNAKE")) != 0 : true;
static bool g_launch_overlap = getenv("DEER_PDL") ? atoi(getenv("DEER_PDL")) != 0 : true;
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (pdl && g_launch_overlap) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

struct TmapKey {
  const void* p; uint64_t rows, cols, ld; uint32_t box;
  bool operator<(const TmapKey& o) const { return std::tie(p, rows, cols, ld, box) < std::tie(o.p, o.rows, o.cols, o.ld, o.box); }
};

struct LaunchStats {
  int64_t total = 0, step = 0, gemm_step = 0;
  double gemm_flops_step = 0;
};

// ---------------------------------------------------------------------------------------------
// GEMM dispatch (shared by the engine and deer_op_gemm)
// ---------------------------------------------------------------------------------------------
struct GemmRec { cudaEvent_t e0, e1; int tag, epoch, m_host, m_mul, n, k; bool dyn; int in_es, out_es; bool has_res; };
struct GemmCtx {
  std::map<TmapKey, CUtensorMap> tmaps;
  bool attr_set[5] = {false, false, false, false, false};
  LaunchStats* stats = nullptr;
  std::string* err = nullptr;
  bool pdl = false;            // programmatic dependent launch allowed (off while a CUDA graph is being captured)
  // split-K workspace of the few-row launches (gemm_tc_splitk_kernel): fp32 partial tiles + one arrival counter per tile
  float* sk_ws = nullptr;
  int* sk_cnt = nullptr;
  size_t sk_ws_bytes = 0;
  bool sk_attr = false;
  bool alloc_splitk() {
    if (sk_ws) return true;
    const size_t bytes = size_t(16) << 20;
    if (cudaMalloc(&sk_ws, bytes) != cudaSuccess) { sk_ws = nullptr; return false; }
    if (cudaMalloc(&sk_cnt, 256 * sizeof(int)) != cudaSuccess || cudaMemset(sk_cnt, 0, 256 * sizeof(int)) != cudaSuccess) {
      cudaFree(sk_ws); sk_ws = nullptr; if (sk_cnt) cudaFree(sk_cnt); sk_cnt = nullptr; return false;
    }
    sk_ws_bytes = bytes;
    return true;
  }
  void free_splitk() {
    if (sk_ws) cudaFree(sk_ws);
    if (sk_cnt) cudaFree(sk_cnt);
    sk_ws = nullptr; sk_cnt = nullptr; sk_ws_bytes = 0;
  }
  // profiling
  bool profile = false;
  int tag = 0, epoch = 0;
  std::vector<GemmRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};

template <int BN, int BM = 128>
static cudaError_t launch_tc(GemmCtx& g, const CUtensorMap& ta, const CUtensorMap& tb, int M, const int* m_dev, int m_mul, int N, int K,
                             const GemmEpi& epi, cudaStream_t st, int idx) {
  using G = GemmTc<BN, BM>;
  if (!g.attr_set[idx]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    g.attr_set[idx] = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = std::max(1, std::min(tiles, DEER_NUM_SMS));
  return launch_k(gemm_tc_kernel<BN, BM>, dim3(grid), dim3(G::THREADS), G::SMEM_BYTES, st, g.pdl, ta, tb, M, m_dev, m_mul, N, K, epi);
}

static cudaError_t launch_splitk(GemmCtx& g, const CUtensorMap& ta, const CUtensorMap& tb, int M, const int* m_dev, int m_mul, int N, int K,
                                 const GemmEpi& epi, cudaStream_t st, int tiles, int splits, int kb_per) {
  using G = GemmTc<64>;
  if (!g.sk_attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    g.sk_attr = true;
  }
  return launch_k(gemm_tc_splitk_kernel, dim3(tiles, splits), dim3(G::THREADS), G::SMEM_BYTES, st, g.pdl, ta, tb, M, m_dev, m_mul, N, K, epi,
                  g.sk_ws, g.sk_cnt, kb_per);
}

static cudaError_t launch_tc2(GemmCtx& g, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, const int* m_dev, int m_mul,
                              int N, int K, const GemmEpi& epi, cudaStream_t st) {
  using G = GemmTc2;
  if (!g.attr_set[3]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    g.attr_set[3] = true;
  }
  const int tiles = ((M + 255) / 256) * ((N + 255) / 256);
  const int clusters = std::max(1, std::min(tiles, DEER_NUM_SMS / 2));
  return launch_k(gemm_tc2_kernel, dim3(2 * clusters), dim3(G::THREADS), G::SMEM_BYTES, st, g.pdl, ta, tb, tc, M, m_dev, m_mul, N, K, epi);
}

// dtype: 0 bf16, 1 fp32.  impl: DEER_GEMM_* (2 = force the 1-CTA tcgen05 kernels, 3 = 128-row tiles with split-K allowed, 4 = 128-row tiles only; for A/B comparisons).  a_rows: allocated rows of A (>= M) for the TMA bounds.
static int run_gemm(GemmCtx& g, int impl, int dtype, const void* A, int lda, int a_rows, const void* W, int ldw, int M, int N, int K,
                    const GemmEpi& epi, const int* m_dev, int m_mul, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return DEER_OK;
  cudaError_t e;
  GemmRec rec{};
  if (g.profile) {
    rec = GemmRec{g.get_event(), g.get_event(), g.tag, g.epoch, M, m_mul, N, K, m_dev != nullptr, dtype == 1 ? 4 : 2,
                  (dtype == 1 || epi.out_f32) ? 4 : 2, epi.resid != nullptr};
    cudaEventRecord(rec.e0, st);
  }
  if (dtype == 1 || impl == DEER_GEMM_SIMT) {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    if (dtype == 1)
      gemm_simt_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(W), ldw, M, m_dev, m_mul, N, K, epi);
    else
      gemm_simt_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(A), lda, reinterpret_cast<const bf16*>(W), ldw, M, m_dev, m_mul, N, K, epi);
    e = cudaGetLastError();
  } else {
    if ((lda & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
      return set_err(g.err, DEER_E_INVALID, "tcgen05 GEMM needs 16-byte aligned operands (lda=%d ldw=%d)", lda, ldw);
    // Tile configuration by a wave-count cost model (rounds up partial waves; per-SM tile cost / measured relative efficiency
    // of each kernel, profiles/r01_gemm_microbench.txt): the 2-CTA 256x256 kernel wins whenever its 74 cluster slots fill, but
    // a projection with few tiles (M = 896 rows of a 28-episode shard x N = 2048: 32 tiles on 74 pairs) runs 1.4x faster as
    // 112 one-CTA 128x128 tiles.  m_dev-clamped launches are sized for the host bound M (the first probes of a step).
    const int mt = (M + 127) / 128;
    auto waves = [](long tiles, int slots) { return (double)((tiles + slots - 1) / slots); };
    const double c2 = (N >= 256) ? waves((long)((M + 255) / 256) * ((N + 255) / 256), DEER_NUM_SMS / 2) * 32768.0 / 1.00 : 1e30;
    const double c256 = (N >= 256) ? waves((long)mt * ((N + 255) / 256), DEER_NUM_SMS) * 32768.0 / 0.83 : 1e30;
    const double c128 = (N >= 128) ? waves((long)mt * ((N + 127) / 128), DEER_NUM_SMS) * 16384.0 / 0.75 : 1e30;
    const double c64 = waves((long)mt * ((N + 63) / 64), DEER_NUM_SMS) * 8192.0 / 0.55;
    // 64 x 64 tiles (tcgen05.mma M = 64): such launches are bound by the bytes one SM pulls per k-block (16 KB against 24 KB for
    // 128 x 64), so a tile costs 2/3 of a 128 x 64 tile for half the rows -- it wins when it turns a sparse wave into a full one
    // (M = 514, N = 1024: 80 -> 144 tiles) and, with 12 stages in flight, on the <= 64-row launches of a 1-2 episode step.
    const double c6464 = waves((long)((M + 63) / 64) * ((N + 63) / 64), DEER_NUM_SMS) * 8192.0 / 0.55 * (2.0 / 3.0);
    static const bool old_pick = getenv("DEER_GEMM_OLD_PICK") != nullptr;
    int BN, BM = 128;
    bool use2;
    const bool one_cta = impl == 2 || impl == 3 || impl == 4;   // deer_op_gemm A/B selectors (include/deer_engine.h)
    if (old_pick) {
      use2 = !one_cta && N >= 256 && ((M + 255) / 256) * ((N + 255) / 256) >= 32;
      if (use2) BN = 128;
      else if (mt * ((N + 255) / 256) >= 2 * DEER_NUM_SMS && N >= 256) BN = 256;
      else if (mt * ((N + 127) / 128) >= DEER_NUM_SMS && N >= 128) BN = 128;
      else BN = 64;
    } else {
      static const bool half_tiles = getenv("DEER_GEMM_M64") ? atoi(getenv("DEER_GEMM_M64")) != 0 : true;
      const double best1 = std::min(c256, std::min(c128, c64));
      use2 = !one_cta && c2 <= best1;
      if (use2) BN = 128;                     // B box = this CTA's 128-row half of the 256-wide tile
      else BN = (c256 <= c128 && c256 <= c64) ? 256 : (c128 <= c64 ? 128 : 64);
      if (!use2 && half_tiles && impl != 3 && impl != 4 && c6464 < (one_cta ? best1 : std::min(best1, c2))) { BN = 64; BM = 64; }
    }
    TmapKey ka{A, (uint64_t)a_rows, (uint64_t)K, (uint64_t)lda, (uint32_t)BM};
    TmapKey kb{W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)BN};
    auto ia = g.tmaps.find(ka);
    if (ia == g.tmaps.end()) {
      CUtensorMap tm;
      if (!make_tmap(&tm, A, a_rows, K, lda, BM)) return set_err(g.err, DEER_E_CUDA, "cuTensorMapEncodeTiled failed for A (%d x %d ld %d)", a_rows, K, lda);
      ia = g.tmaps.emplace(ka, tm).first;
    }
    auto ib = g.tmaps.find(kb);
    if (ib == g.tmaps.end()) {
      CUtensorMap tm;
      if (!make_tmap(&tm, W, N, K, ldw, BN)) return set_err(g.err, DEER_E_CUDA, "cuTensorMapEncodeTiled failed for W (%d x %d ld %d)", N, K, ldw);
      ib = g.tmaps.emplace(kb, tm).first;
    }
    if (use2) {
      // bf16 output without a residual leaves through shared memory + TMA stores (DEER_GEMM_TMA_STORE=0: the transposing epilogue, A/B)
      const char* ts_env = getenv("DEER_GEMM_TMA_STORE");                // read per call: tests flip it inside one process
      const bool tma_store_on = ts_env ? atoi(ts_env) != 0 : true;
      GemmEpi e2 = epi;
      const CUtensorMap* tc = &ia->second;
      if (tma_store_on && !epi.out_f32 && epi.resid == nullptr && (N % 32) == 0 && (epi.ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(epi.C) & 15) == 0 &&
          (epi.bias == nullptr || (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0)) {
        TmapKey kc{epi.C, (uint64_t)M, (uint64_t)N, (uint64_t)epi.ldc, 3232u};       // rows = M: the hardware clips the last tile's rows
        auto ic = g.tmaps.find(kc);
        if (ic == g.tmaps.end()) {
          CUtensorMap tm;
          if (!make_tmap_c(&tm, epi.C, M, N, epi.ldc)) return set_err(g.err, DEER_E_CUDA, "cuTensorMapEncodeTiled failed for C (%d x %d ld %d)", M, N, epi.ldc);
          ic = g.tmaps.emplace(kc, tm).first;
        }
        tc = &ic->second;
        e2.tma_store = 1;
      }
      e = launch_tc2(g, ia->second, ib->second, *tc, M, m_dev, m_mul, N, K, e2, st);
    }
    else if (BN == 256) e = launch_tc<256>(g, ia->second, ib->second, M, m_dev, m_mul, N, K, epi, st, 0);
    else if (BN == 128) e = launch_tc<128>(g, ia->second, ib->second, M, m_dev, m_mul, N, K, epi, st, 1);
    else if (BM == 64) e = launch_tc<64, 64>(g, ia->second, ib->second, M, m_dev, m_mul, N, K, epi, st, 4);
    else {
      // Few tiles on 148 SMs (a 1-4 episode step: M <= 128 rows): splitting K over the idle SMs opens more TMA pipelines on the
      // weight matrix (llm.down at 32 rows: 33 -> 16 us, profiles/r02_splitk.txt).  OPT-IN (DEER_GEMM_SPLITK=1, or impl 3 through
      // deer_op_gemm): the partial sums change the fp32 summation order over K, so an episode's bits would depend on how many
      // episodes share its step -- the default keeps results bit-identical across batch sizes and shardings, and the whole 1-episode
      // step gains only 2.4 % (it is launch-latency bound).  A split must stream >= 8 k-blocks (>= 16 above 64 rows: the
      // workspace round trip grows with the rows) or the reduction costs more than it saves.
      static const bool splitk_on = getenv("DEER_GEMM_SPLITK") ? atoi(getenv("DEER_GEMM_SPLITK")) != 0 : false;
      const int tiles = mt * ((N + 63) / 64), num_kb = (K + 63) / 64;
      int splits = std::min(DEER_NUM_SMS / std::max(1, tiles), num_kb / (M <= 64 ? 8 : 16));
      int kb_per = 0;
      if ((splitk_on || impl == 3) && g.sk_ws != nullptr && splits >= 2 && tiles <= 256 && (N % 8) == 0) {
        kb_per = (num_kb + splits - 1) / splits;
        splits = (num_kb + kb_per - 1) / kb_per;
      }
      if (kb_per > 0 && splits >= 2 && size_t(splits) * M * N * sizeof(float) <= g.sk_ws_bytes)
        e = launch_splitk(g, ia->second, ib->second, M, m_dev, m_mul, N, K, epi, st, tiles, splits, kb_per);
      else
        e = launch_tc<64>(g, ia->second, ib->second, M, m_dev, m_mul, N, K, epi, st, 2);
    }
  }
  if (g.profile) { cudaEventRecord(rec.e1, st); g.recs.push_back(rec); }
  if (e != cudaSuccess) return set_err(g.err, DEER_E_CUDA, "GEMM launch failed: %s", cudaGetErrorString(e));
  if (g.stats) {
    g.stats->total++; g.stats->step++; g.stats->gemm_step++;
    g.stats->gemm_flops_step += 2.0 * M * N * K;
  }
  return DEER_OK;
}

static std::map<TmapKey, CUtensorMap> g_attn_tmaps;
static long long* g_attn_dbg = nullptr;       // optional per-CTA phase-cycle counters (deer_attn_debug)
static long long* g_head_dbg = nullptr;       // optional phase timestamps of the fused head probe (deer_head_debug)
static const CUtensorMap* attn_tmap(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box) {
  TmapKey k{ptr, rows, cols, ld, box};
  auto it = g_attn_tmaps.find(k);
  if (it == g_attn_tmaps.end()) {
    CUtensorMap tm;
    if (!make_tmap(&tm, ptr, rows, cols, ld, box)) return nullptr;
    if (g_attn_tmaps.size() > 4096) g_attn_tmaps.clear();
    it = g_attn_tmaps.emplace(k, tm).first;
  }
  return &it->second;
}

// impl: 0 = tensor-core path (tcgen05 kernel when the shape allows, else the mma.sync flash kernel), 1 = SIMT reference,
//       2 = on-chip small-attention kernel (LLM layers), 3 = force the mma.sync flash kernel
static int run_attention(int impl, int dtype, const AttnParams& p, std::string* err, LaunchStats* stats, cudaStream_t st, bool pdl = false) {
  if (p.B_host <= 0) return DEER_OK;
  const bool plain = dtype == 0 && p.hd == 64 && !p.causal && !p.alibi_slopes && !p.kv_len && !p.qmode && !p.n_dev && p.ldk == p.ldv;
  static const bool no_tc = getenv("DEER_ATTN_NO_TC") != nullptr;
  bool done = false;
  if (impl == 0 && plain && !no_tc && p.Sk <= AttnTc::MAX_SK && p.Sq > 8 && (p.ldq % 8) == 0 && (p.ldk % 8) == 0 && (p.ldo % 8) == 0) {
    const uint64_t width = (uint64_t)p.H * 64;
    const CUtensorMap* tq = attn_tmap(p.Q, (uint64_t)p.B_host * p.Sq, width, p.ldq, 128);
    const CUtensorMap* tk = attn_tmap(p.K, (uint64_t)p.B_host * p.Sk, width, p.ldk, AttnTc::HALF_ROWS);
    const CUtensorMap* tv = attn_tmap(p.V, (uint64_t)p.B_host * p.Sk, width, p.ldv, AttnTc::HALF_ROWS);
    if (tq && tk && tv) {
      static bool attr_done = false;
      if (!attr_done) {
        cudaFuncSetAttribute(attn_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnTc5::SMEM_BYTES);
        attr_done = true;
      }
      const int grid = std::min(p.B_host * p.H, DEER_NUM_SMS);
      launch_k(attn_tc5_kernel, dim3(grid), dim3(AttnTc5::THREADS), AttnTc5::SMEM_BYTES, st, pdl, *tq, *tk, *tv, reinterpret_cast<const bf16*>(p.Q), p.ldq,
               reinterpret_cast<bf16*>(p.O), p.ldo, p.B_host, p.H, p.Sq, p.Sk, p.scale * 1.4426950408889634f, g_attn_dbg, g_snake ? 1 : 0);
      done = true;
    }
  }
  if (done) {
  } else if ((impl == 0 || impl == 3) && plain) {
    dim3 grid((p.Sq + 63) / 64, p.H, p.B_host);
    launch_k(flash_attn_hd64_kernel, grid, dim3(128), 0, st, pdl, reinterpret_cast<const bf16*>(p.Q), p.ldq, reinterpret_cast<const bf16*>(p.K),
             reinterpret_cast<const bf16*>(p.V), p.ldk, reinterpret_cast<bf16*>(p.O), p.ldo, p.Sq, p.Sk, p.scale * 1.4426950408889634f);
  } else if ((impl == 2 || impl == 4) && dtype == 0 && p.Sq <= 32 && p.Sk <= 128 && (p.hd == 64 || p.hd == 128) && (p.ldq % 8) == 0 &&
             (p.ldk % 8) == 0 && (p.ldv % 8) == 0 && (p.ldo % 2) == 0 && ((reinterpret_cast<uintptr_t>(p.Q) | reinterpret_cast<uintptr_t>(p.K) |
             reinterpret_cast<uintptr_t>(p.V)) & 15) == 0 && (impl == 4 || getenv("DEER_ATTN_LLM_SIMT") == nullptr)) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(attn_llm_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      cudaFuncSetAttribute(attn_llm_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      attr_done = true;
    }
    dim3 grid(p.H, p.B_host);
    const size_t sm = attn_llm_mma_smem(p.Sk, p.hd);
    if (p.hd == 64) launch_k(attn_llm_mma_kernel<64>, grid, dim3(64), sm, st, pdl, p);
    else launch_k(attn_llm_mma_kernel<128>, grid, dim3(64), sm, st, pdl, p);
  } else if ((impl == 2 || impl == 4) && attn_small_smem(p.Sq, p.Sk, p.hd) <= 200 * 1024) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(attn_small_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(attn_small_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_done = true;
    }
    dim3 grid(p.H, p.B_host);
    const size_t sm = attn_small_smem(p.Sq, p.Sk, p.hd);
    if (dtype == 1) attn_small_kernel<float><<<grid, 256, sm, st>>>(p);
    else attn_small_kernel<bf16><<<grid, 256, sm, st>>>(p);
  } else {
    dim3 grid(p.Sq, p.H, p.B_host);
    const size_t sm = (p.hd + p.Sk + 8) * sizeof(float);
    if (dtype == 1) attn_simt_kernel<float><<<grid, 128, sm, st>>>(p);
    else attn_simt_kernel<bf16><<<grid, 128, sm, st>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(err, DEER_E_CUDA, "attention launch failed: %s", cudaGetErrorString(e));
  if (stats) { stats->total++; stats->step++; }
  return DEER_OK;
}

// ---------------------------------------------------------------------------------------------
// engine context
// ---------------------------------------------------------------------------------------------
struct Staged { float* dev = nullptr; std::vector<int64_t> shape; size_t numel = 0; };

struct VitLayerW { float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *b_in, *b_out, *b_fc, *b_proj; void *w_in, *w_out, *w_fc, *w_proj; };
struct PercLayerW { float *nm_g, *nm_b, *nl_g, *nl_b, *ff_g, *ff_b; void *w_q, *w_kv, *w_out, *w_ff1, *w_ff2; };
struct LlmLayerW {
  bool has_x = false;
  float *xn_g = nullptr, *xn_b = nullptr, *xff_g = nullptr, *xff_b = nullptr, *attn_gate_tanh = nullptr, *ff_gate_tanh = nullptr;
  void *xw_q = nullptr, *xw_kv = nullptr, *xw_out = nullptr, *xw_ff1 = nullptr, *xw_ff2 = nullptr;
  float *ln1_g, *ln2_g, *qln_g = nullptr, *kln_g = nullptr;
  void *w_qkv, *w_out, *w_up, *w_down;
};
struct LstmLayerW { void* w_cat; float *b_sum, *ln_g, *ln_b; int in; };

struct deer_ctx {
  deer_config cfg;
  int device = 0;
  std::string err;
  bool finalized = false, thresholds_set = false;
  int es = 2;                 // element size of T
  int dtype = 0;              // 0 bf16, 1 fp32
  // derived
  int np = 0, L = 0, Kp = 0, nl = 0, inner = 0, xin = 0, hd = 0;
  std::vector<int> exits;     // exit layer indices (flamingo_mpt.py:268-270)
  int ctrl_max_layer = 0, n_thr = 0;
  // weights
  std::map<std::string, Staged> staged;
  std::vector<std::string> required;
  std::vector<void*> allocs;
  int64_t weight_bytes = 0, workspace_bytes = 0;
  void* w_conv = nullptr; float *v_cls = nullptr, *v_pos = nullptr, *v_lnpre_g = nullptr, *v_lnpre_b = nullptr;
  std::vector<VitLayerW> vit;
  float* p_latents = nullptr; std::vector<PercLayerW> perc; float *p_norm_g = nullptr, *p_norm_b = nullptr;
  void* wte = nullptr;
  std::vector<LlmLayerW> llm;
  std::vector<LstmLayerW> lstm;
  void *w_mlp1 = nullptr, *w_mlp2_pose = nullptr, *w_mlp2_grip = nullptr, *w_mlp3_pose = nullptr, *w_mlp3_grip = nullptr;
  float *b_mlp1 = nullptr, *g_mlp1 = nullptr, *be_mlp1 = nullptr, *b_mlp2 = nullptr, *g_mlp2 = nullptr, *be_mlp2 = nullptr;
  float *b_mlp3 = nullptr, *g_mlp3 = nullptr, *be_mlp3 = nullptr;
  float *w3_pose = nullptr, *b3_pose = nullptr, *w3_grip = nullptr, *b3_grip = nullptr;
  float* alibi = nullptr; float* thr_dev = nullptr;
  // workspace
  void *im_cols = nullptr, *patch_out = nullptr, *v_ln = nullptr, *v_qkv = nullptr, *v_att = nullptr, *v_h = nullptr;
  float* vx = nullptr;
  float* lat = nullptr; void *p_kvin = nullptr, *p_latn = nullptr, *p_q = nullptr, *p_kv = nullptr, *p_att = nullptr, *p_ffh = nullptr;
  float* x[2] = {nullptr, nullptr}; void* media[2] = {nullptr, nullptr};
  void *l_ln = nullptr, *l_q = nullptr, *l_kv = nullptr, *l_att = nullptr, *l_h = nullptr, *l_qkv = nullptr;
  void *a0 = nullptr, *a_rest = nullptr, *head_in = nullptr, *y1 = nullptr;
  float *gates = nullptr, *cand_h = nullptr, *cand_c = nullptr, *mlp1 = nullptr, *mlp2 = nullptr, *y2 = nullptr, *cur_pose = nullptr, *cur_grip = nullptr;
  void* y2t = nullptr; float *mlp3 = nullptr, *y3 = nullptr;      // third hidden layer (old checkpoints)
  float *pooled = nullptr, *o_feat = nullptr;                    // max-pooled exit feature: per active row / per episode at its exit
  float *h_state = nullptr, *c_state = nullptr; int* cur_exit_id = nullptr;
  EpisodeMeta meta[2];
  uint8_t* exit_flag = nullptr; int* src_of_new = nullptr; int* n_active = nullptr;   // n_active[2]
  // output scratch (when the caller passes NULL) and host staging for deer_step_host
  float *o_pose = nullptr, *o_grip = nullptr, *o_action = nullptr, *o_delta = nullptr; int* o_exit = nullptr;
  float *s_rgb = nullptr, *s_grip = nullptr; int64_t* s_ids = nullptr; uint8_t* s_mask = nullptr; int *s_slots = nullptr, *s_ts = nullptr;
  int* h_nact = nullptr;      // pinned host int for host_control
  // taps
  float *tap_vit = nullptr, *tap_media = nullptr, *tap_hidden = nullptr, *tap_pooled = nullptr;
  int tap_n = 0, tap_S = 0;
  int vit_chunk = 0;           // images per ViT pass (0 = all at once); DEER_VIT_CHUNK overrides
  cudaStream_t own_stream = nullptr;
  // cross-stream ordering of consecutive calls on one ctx (workspace and episode state are shared)
  cudaEvent_t last_done = nullptr; cudaStream_t last_stream = nullptr; bool has_last = false;
  int* nact_hist = nullptr;            // device [16]: active count after each probe of the last step
  int n_probes_last = 0, n_last = 0;
  cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr;
  // CUDA-graph mode
  struct GraphEntry { cudaGraphExec_t exec = nullptr; cudaGraph_t graph = nullptr; int seen = 0; int64_t launches = 0, gemms = 0; double flops = 0; int probes = 0; };
  std::map<std::tuple<int, int, int, unsigned, int>, GraphEntry> graphs;
  cudaStream_t cap_streams[20] = {nullptr};
  int cap_level = -1;                 // >= 0 while a step is being captured
  int64_t graph_launches = 0, graphs_built = 0;
  // on-device image transform: Pillow coefficient tables per (in_size -> out_size), uint8 staging
  struct AxisTab { int* bounds = nullptr; int* kk = nullptr; int ksize = 0; int need = 0; };
  std::map<std::pair<int, int>, AxisTab> axis_tabs;
  uint8_t* u8_stage[2] = {nullptr, nullptr}; size_t u8_cap[2] = {0, 0};
  // generate (calibration) mode
  const int32_t* gen_commit_layer = nullptr; float *pend_h = nullptr, *pend_c = nullptr; uint8_t* pend_flag = nullptr;
  // fused exit-head probe (head_fused.cuh): persistent cooperative kernel for active sets of <= hf_mr episodes
  bool hf_ok = false; int hf_mr = 0, hf_lda = 0, hf_grid = 0, hf_cluster = 1; size_t hf_smem = 0; unsigned long long* hf_bar = nullptr;
  GemmCtx gemm;
  LaunchStats stats;
  int64_t steps = 0;
};

#define CK(...)                                                                                           \
  do {                                                                                                    \
    cudaError_t _e = (__VA_ARGS__);                                                                       \
    if (_e != cudaSuccess) return set_err(&c->err, DEER_E_CUDA, "%s failed: %s", #__VA_ARGS__, cudaGetErrorString(_e)); \
  } while (0)
#define RET(...)                  \
  do {                            \
    int _r = (__VA_ARGS__);       \
    if (_r != DEER_OK) return _r; \
  } while (0)
#define KLAUNCH_CHECK(name)                                                                                 \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess) return set_err(&c->err, DEER_E_CUDA, "%s launch failed: %s", name, cudaGetErrorString(_e)); \
    c->stats.total++; c->stats.step++;                                                                      \
  } while (0)

static int dalloc(deer_ctx* c, void** p, size_t bytes, bool weight, bool zero = false) {
  bytes = (bytes + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) return set_err(&c->err, DEER_E_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  c->allocs.push_back(*p);
  if (zero) cudaMemset(*p, 0, bytes);
  if (weight) c->weight_bytes += bytes; else c->workspace_bytes += bytes;
  return DEER_OK;
}

// ---------------------------------------------------------------------------------------------
// config validation / derived quantities
// ---------------------------------------------------------------------------------------------
static int derive(deer_ctx* c) {
  const deer_config& f = c->cfg;
  if (f.abi_version != DEER_ABI_VERSION) return set_err(&c->err, DEER_E_INVALID, "abi_version %d != %d", f.abi_version, DEER_ABI_VERSION);
  if (f.image_size % f.patch) return set_err(&c->err, DEER_E_INVALID, "image_size %% patch != 0");
  if (f.vit_width % f.vit_heads || f.vit_width / f.vit_heads != 64) return set_err(&c->err, DEER_E_INVALID, "ViT head_dim must be 64");
  if (f.perc_dim_head != 64 || f.xattn_dim_head != 64) return set_err(&c->err, DEER_E_INVALID, "perceiver / x-attn dim_head must be 64");
  if (f.vit_width > 1024 || f.vit_width % 8) return set_err(&c->err, DEER_E_INVALID, "vit_width must be <= 1024 and a multiple of 8");
  if (f.d_model % f.n_heads || f.d_model > 4096 || f.d_model % 8) return set_err(&c->err, DEER_E_INVALID, "bad d_model / n_heads");
  if (f.d_model / f.n_heads > 128) return set_err(&c->err, DEER_E_INVALID, "LLM head_dim must be <= 128");
  if (f.lstm_hidden > 2048 || f.lstm_hidden % 8 || f.lstm_layers < 1) return set_err(&c->err, DEER_E_INVALID, "bad lstm config");
  if (f.mlp_dim0 % 8 || f.mlp_dim1 % 8 || f.mlp_dim0 > 4096 || f.mlp_dim1 > 4096) return set_err(&c->err, DEER_E_INVALID, "bad mlp dims");
  if (f.mlp_dim2 < 0 || f.mlp_dim2 % 8 || f.mlp_dim2 > 4096) return set_err(&c->err, DEER_E_INVALID, "bad mlp_dim2");
  if (f.max_batch < 1 || f.max_batch > 1024) return set_err(&c->err, DEER_E_INVALID, "max_batch must be in [1,1024]");
  if (f.max_slots < 1 || f.max_seq < 1 || f.max_seq > 256) return set_err(&c->err, DEER_E_INVALID, "bad max_slots / max_seq");
  if (f.n_layers < 1 || f.exit_interval < 1 || f.max_layer < 1) return set_err(&c->err, DEER_E_INVALID, "bad exit config");
  if (f.threshold_type < 0 || f.threshold_type > 3) return set_err(&c->err, DEER_E_INVALID, "bad threshold_type");
  if (f.precision != DEER_PREC_BF16 && f.precision != DEER_PREC_FP32) return set_err(&c->err, DEER_E_INVALID, "bad precision");
  c->dtype = f.precision == DEER_PREC_FP32 ? 1 : 0;
  c->es = c->dtype ? 4 : 2;
  const int g = f.image_size / f.patch;
  c->np = g * g;
  c->L = c->np + 1;
  c->Kp = ((3 * f.patch * f.patch + 63) / 64) * 64;
  c->nl = f.num_latents;
  c->inner = f.perc_heads * f.perc_dim_head;
  c->xin = f.xattn_heads * f.xattn_dim_head;
  c->hd = f.d_model / f.n_heads;
  // exit list: range(interval-1, early_exit_layer, interval) + [n_layers-1]   (flamingo_mpt.py:248-250,268-270)
  c->exits.clear();
  const int eel = f.n_layers - 1;
  for (int i = f.exit_interval - 1; i < eel; i += f.exit_interval) c->exits.push_back(i);
  c->exits.push_back(f.n_layers - 1);
  c->ctrl_max_layer = std::min(f.max_layer - 1, c->exits.back());          // value_net.py:173
  c->n_thr = 0;
  for (int e : c->exits) if (e <= c->ctrl_max_layer) c->n_thr++;            // value_net.py:178
  if (c->exits[0] < 1) return set_err(&c->err, DEER_E_INVALID, "first exit must be > 0 (value_net.py:121)");
  return DEER_OK;
}

static void build_required(deer_ctx* c) {
  const deer_config& f = c->cfg;
  std::vector<std::string>& r = c->required;
  r.clear();
  char b[256];
  const char* v = "vision_encoder.visual.";
  for (const char* s : {"conv1.weight", "class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias"}) r.push_back(std::string(v) + s);
  for (int i = 0; i < f.vit_layers; ++i)
    for (const char* s : {"ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
                          "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias"}) {
      snprintf(b, sizeof(b), "%stransformer.resblocks.%d.%s", v, i, s);
      r.push_back(b);
    }
  r.push_back("perceiver.latents");
  for (int i = 0; i < f.perc_depth; ++i)
    for (const char* s : {"0.norm_media.weight", "0.norm_media.bias", "0.norm_latents.weight", "0.norm_latents.bias", "0.to_q.weight",
                          "0.to_kv.weight", "0.to_out.weight", "1.0.weight", "1.0.bias", "1.1.weight", "1.3.weight"}) {
      snprintf(b, sizeof(b), "perceiver.layers.%d.%s", i, s);
      r.push_back(b);
    }
  r.push_back("perceiver.norm.weight");
  r.push_back("perceiver.norm.bias");
  r.push_back("lang_encoder.transformer.wte.weight");
  // every block the reference builds (flamingo_mpt.py:198): calibration runs to n_layers-1 and exit_id may name any of them
  for (int i = 0; i < f.n_layers; ++i) {
    if ((i + 1) % f.cross_attn_every == 0)
      for (const char* s : {"attn.norm.weight", "attn.norm.bias", "attn.to_q.weight", "attn.to_kv.weight", "attn.to_out.weight", "attn_gate",
                            "ff.0.weight", "ff.0.bias", "ff.1.weight", "ff.3.weight", "ff_gate"}) {
        snprintf(b, sizeof(b), "lang_encoder.transformer.blocks.%d.gated_cross_attn_layer.%s", i, s);
        r.push_back(b);
      }
    if (f.qk_ln) {
      for (const char* s : {"ln_1.weight", "attn.Wqkv.weight", "attn.q_ln.weight", "attn.k_ln.weight", "attn.out_proj.weight", "ln_2.weight",
                            "mlp.mlp_up.weight", "mlp.mlp_down.weight"}) {
        snprintf(b, sizeof(b), "lang_encoder.transformer.blocks.%d.decoder_layer.%s", i, s);
        r.push_back(b);
      }
    } else {
      for (const char* s : {"norm_1.weight", "attn.Wqkv.weight", "attn.out_proj.weight", "norm_2.weight", "ffn.up_proj.weight", "ffn.down_proj.weight"}) {
        snprintf(b, sizeof(b), "lang_encoder.transformer.blocks.%d.decoder_layer.%s", i, s);
        r.push_back(b);
      }
    }
  }
  for (int l = 0; l < f.lstm_layers; ++l) {
    if (f.lstm_layernorm) {                     // LayerNormLSTM: ModuleList [LSTM, LayerNorm, Dropout] per layer (action_head.py:15-40)
      for (const char* s : {"weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"}) {
        snprintf(b, sizeof(b), "extra_exit.rnn.layers.%d.%s", 3 * l, s);
        r.push_back(b);
      }
      for (const char* s : {"weight", "bias"}) {
        snprintf(b, sizeof(b), "extra_exit.rnn.layers.%d.%s", 3 * l + 1, s);
        r.push_back(b);
      }
    } else {                                    // plain stacked nn.LSTM (action_head.py:71-79)
      for (const char* s : {"weight_ih", "weight_hh", "bias_ih", "bias_hh"}) {
        snprintf(b, sizeof(b), "extra_exit.rnn.%s_l%d", s, l);
        r.push_back(b);
      }
    }
  }
  // MLP heads in the 'layerwise' Sequential layout (Linear at 4j+1, LayerNorm at 4j+2, output Linear at 4n+1); the host
  // side renames the other dropout_mode layouts to this one (checkpoint.py)
  const int nh = f.mlp_dim2 > 0 ? 3 : 2;
  for (const char* head : {"actions", "gripper"}) {
    for (int j = 0; j <= nh; ++j)
      for (int part = 0; part < 2; ++part) {
        if (part == 1 && (j == nh || !f.mlp_layernorm)) continue;
        for (const char* s : {"weight", "bias"}) {
          snprintf(b, sizeof(b), "extra_exit.%s.mlp.%d.%s", head, 4 * j + 1 + part, s);
          r.push_back(b);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// C ABI: library level
// ---------------------------------------------------------------------------------------------
extern "C" int deer_abi_version(void) { return DEER_ABI_VERSION; }
#ifndef DEER_SRC_HASH
#define DEER_SRC_HASH "unhashed"
#endif
// The marker string lets build() read the hash out of the file without loading it.
extern "C" const char* deer_source_hash(void) {
  static const char marker[] = "DEER_SRC_HASH=" DEER_SRC_HASH;
  return marker + 14;
}
extern "C" const char* deer_build_info(void) {
  return "deer_engine sm_100a (tcgen05/TMEM/TMA GEMM + attention), built " __DATE__ " " __TIME__ " src=" DEER_SRC_HASH;
}
extern "C" const char* deer_last_error(deer_ctx* c) { return c ? c->err.c_str() : g_last_error.c_str(); }

extern "C" int deer_create(const deer_config* cfg, int device, deer_ctx** out) {
  if (!cfg || !out) return set_err(nullptr, DEER_E_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_err(nullptr, DEER_E_CUDA, "no CUDA device: the DeeR engine has no CPU fallback");
  if (device < 0 || device >= ndev) return set_err(nullptr, DEER_E_INVALID, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) return set_err(nullptr, DEER_E_CUDA, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
  deer_ctx* c = new deer_ctx();
  c->cfg = *cfg;
  c->device = device;
  int r = derive(c);
  if (r != DEER_OK) { g_last_error = c->err; delete c; return r; }
  cudaSetDevice(device);
  build_required(c);
  c->vit_chunk = getenv("DEER_VIT_CHUNK") ? atoi(getenv("DEER_VIT_CHUNK")) : 0;
  c->gemm.stats = &c->stats;
  c->gemm.err = &c->err;
  if (c->dtype == 0 && !c->gemm.alloc_splitk()) {
    r = set_err(&c->err, DEER_E_CUDA, "split-K workspace allocation failed");
    g_last_error = c->err; deer_destroy(c); return r;
  }
  const deer_config& f = c->cfg;
  const size_t B = f.max_batch, NI = 2 * B, S = f.max_seq, D = f.d_model, Wd = f.vit_width, es = c->es;
  const size_t L = c->L, np = c->np, nl = c->nl, H = f.lstm_hidden, LL = f.lstm_layers;
#define WS(ptr, bytes) do { r = dalloc(c, reinterpret_cast<void**>(&(ptr)), (bytes), false, true); if (r != DEER_OK) { g_last_error = c->err; deer_destroy(c); return r; } } while (0)
  WS(c->im_cols, NI * np * c->Kp * es);
  WS(c->patch_out, NI * np * Wd * es);
  WS(c->vx, NI * L * Wd * 4);
  WS(c->v_ln, NI * L * Wd * es);
  WS(c->v_qkv, NI * L * 3 * Wd * es);
  WS(c->v_att, NI * L * Wd * es);
  WS(c->v_h, NI * L * f.vit_mlp * es);
  WS(c->lat, NI * nl * Wd * 4);
  WS(c->p_kvin, NI * (np + nl) * Wd * es);
  WS(c->p_latn, NI * nl * Wd * es);
  WS(c->p_q, NI * nl * c->inner * es);
  WS(c->p_kv, NI * (np + nl) * 2 * c->inner * es);
  WS(c->p_att, NI * nl * c->inner * es);
  WS(c->p_ffh, NI * nl * f.perc_ff_mult * Wd * es);
  for (int i = 0; i < 2; ++i) {
    WS(c->x[i], B * S * D * 4);
    WS(c->media[i], B * 2 * nl * Wd * es);
    WS(c->meta[i].orig, B * 4); WS(c->meta[i].slot, B * 4); WS(c->meta[i].valid_len, B * 4);
    WS(c->meta[i].xmode, B * S); WS(c->meta[i].prev_pose, B * 6 * 4); WS(c->meta[i].prev_grip, B * 4); WS(c->meta[i].use_prev, B);
  }
  WS(c->l_ln, B * S * D * es);
  WS(c->l_q, B * S * c->xin * es);
  WS(c->l_kv, B * 2 * nl * 2 * c->xin * es);
  WS(c->l_att, B * S * std::max<size_t>(c->xin, D) * es);
  WS(c->l_h, B * S * std::max<size_t>(f.mlp_ratio, f.xattn_ff_mult) * D * es);
  WS(c->l_qkv, B * S * 3 * D * es);
  WS(c->a0, B * (D + H) * es);
  WS(c->a_rest, std::max<size_t>(1, LL - 1) * B * 2 * H * es);
  WS(c->head_in, B * H * es);
  WS(c->y1, B * 2 * f.mlp_dim0 * es);
  WS(c->gates, B * 4 * H * 4);
  WS(c->cand_h, LL * B * H * 4);
  WS(c->cand_c, LL * B * H * 4);
  WS(c->mlp1, B * 2 * f.mlp_dim0 * 4);
  WS(c->mlp2, B * 2 * f.mlp_dim1 * 4);
  WS(c->y2, B * 2 * f.mlp_dim1 * 4);
  if (f.mlp_dim2 > 0) { WS(c->y2t, B * 2 * f.mlp_dim1 * es); WS(c->mlp3, B * 2 * f.mlp_dim2 * 4); WS(c->y3, B * 2 * f.mlp_dim2 * 4); }
  WS(c->pooled, B * D * 4);
  WS(c->o_feat, B * D * 4);
  WS(c->cur_pose, B * 6 * 4);
  WS(c->cur_grip, B * 4);
  WS(c->h_state, (size_t)f.max_slots * LL * H * 4);
  WS(c->c_state, (size_t)f.max_slots * LL * H * 4);
  WS(c->cur_exit_id, (size_t)f.max_slots * 4);
  WS(c->exit_flag, B);
  WS(c->src_of_new, B * 4);
  WS(c->n_active, 2 * 4);
  WS(c->o_pose, B * 6 * 4); WS(c->o_grip, B * 4); WS(c->o_action, B * 7 * 4); WS(c->o_exit, B * 4);
  WS(c->o_delta, B * std::max(1, c->n_thr) * 4);
  WS(c->s_rgb, B * 3 * f.image_size * f.image_size * 4);
  WS(c->s_grip, B * 3 * f.image_size * f.image_size * 4);
  WS(c->s_ids, B * S * 8); WS(c->s_mask, B * S); WS(c->s_slots, B * 4); WS(c->s_ts, B * 4);
  WS(c->thr_dev, std::max(1, c->n_thr) * 4);
  WS(c->nact_hist, 16 * 4);
  WS(c->alibi, f.n_heads * 4);
  if (f.debug_taps) {
    WS(c->tap_vit, NI * np * Wd * 4);
    WS(c->tap_media, B * 2 * nl * Wd * 4);
    WS(c->tap_hidden, (size_t)f.n_layers * B * S * D * 4);
    WS(c->tap_pooled, B * D * 4);
  }
#undef WS
  if (cudaMallocHost(&c->h_nact, 4) != cudaSuccess) { set_err(&c->err, DEER_E_CUDA, "cudaMallocHost failed"); g_last_error = c->err; deer_destroy(c); return DEER_E_CUDA; }
  cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->last_done, cudaEventDisableTiming);
  // ALiBi slopes (MPT attention.alibi_bias / gen_slopes): 2^-(i * bias_max / Hp), interleaved when H is not a power of two
  {
    const int Hn = f.n_heads;
    int hp = 1;
    while (hp < Hn) hp <<= 1;
    std::vector<float> base(hp), sl(Hn);
    for (int i = 0; i < hp; ++i) base[i] = 1.0f / std::pow(2.0f, (float)(i + 1) * ((float)f.alibi_bias_max / (float)hp));
    if (hp == Hn) sl = base;
    else {
      std::vector<float> il;
      for (int i = 1; i < hp; i += 2) il.push_back(base[i]);
      for (int i = 0; i < hp; i += 2) il.push_back(base[i]);
      for (int i = 0; i < Hn; ++i) sl[i] = il[i];
    }
    cudaMemcpy(c->alibi, sl.data(), Hn * 4, cudaMemcpyHostToDevice);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_err(&c->err, DEER_E_CUDA, "init failed: %s", cudaGetErrorString(e)); g_last_error = c->err; deer_destroy(c); return DEER_E_CUDA; }
  *out = c;
  return DEER_OK;
}

extern "C" void deer_destroy(deer_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->staged) if (kv.second.dev) cudaFree(kv.second.dev);
  for (void* p : c->allocs) cudaFree(p);
  if (c->h_nact) cudaFreeHost(c->h_nact);
  for (auto& r : c->gemm.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : c->gemm.pool) cudaEventDestroy(e);
  c->gemm.free_splitk();
  if (c->ev_step0) cudaEventDestroy(c->ev_step0);
  if (c->ev_step1) cudaEventDestroy(c->ev_step1);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->last_done) cudaEventDestroy(c->last_done);
  for (auto& kv : c->graphs) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); if (kv.second.graph) cudaGraphDestroy(kv.second.graph); }
  for (auto s2 : c->cap_streams) if (s2) cudaStreamDestroy(s2);
  for (int k = 0; k < 2; ++k) if (c->u8_stage[k]) cudaFree(c->u8_stage[k]);
  delete c;
}

extern "C" int deer_num_required_weights(deer_ctx* c) { return c ? (int)c->required.size() : 0; }
extern "C" const char* deer_required_weight_name(deer_ctx* c, int i) {
  return (c && i >= 0 && i < (int)c->required.size()) ? c->required[i].c_str() : nullptr;
}
extern "C" int deer_num_exits(deer_ctx* c) { return c ? (int)c->exits.size() : 0; }
extern "C" int deer_exit_layer(deer_ctx* c, int i) { return (c && i >= 0 && i < (int)c->exits.size()) ? c->exits[i] : -1; }
extern "C" int deer_num_thresholds(deer_ctx* c) { return c ? c->n_thr : 0; }
extern "C" int deer_ctrl_max_layer(deer_ctx* c) { return c ? c->ctrl_max_layer : -1; }

extern "C" int deer_load_weight(deer_ctx* c, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!c || !name || !data || (ndim > 0 && !shape)) return set_err(c ? &c->err : nullptr, DEER_E_INVALID, "null argument");
  if (c->finalized) return set_err(&c->err, DEER_E_STATE, "weights already finalized");
  std::string nm(name);
  if (nm.rfind("module.", 0) == 0) nm = nm.substr(7);
  if (std::find(c->required.begin(), c->required.end(), nm) == c->required.end()) return 1;   // not on the hot path: skipped (strict=False)
  cudaSetDevice(c->device);
  Staged s;
  s.numel = 1;
  for (int i = 0; i < ndim; ++i) { s.shape.push_back(shape[i]); s.numel *= (size_t)shape[i]; }
  auto it = c->staged.find(nm);
  if (it != c->staged.end()) { cudaFree(it->second.dev); c->staged.erase(it); }
  CK(cudaMalloc(reinterpret_cast<void**>(&s.dev), std::max<size_t>(16, s.numel * 4)));
  CK(cudaMemcpy(s.dev, data, s.numel * 4, cudaMemcpyHostToDevice));
  c->staged[nm] = s;
  return DEER_OK;
}

// ---- weight packing helpers -------------------------------------------------------------------
static int need(deer_ctx* c, const std::string& name, std::initializer_list<int64_t> shape, Staged** out) {
  auto it = c->staged.find(name);
  if (it == c->staged.end()) return set_err(&c->err, DEER_E_MISSING, "missing weight '%s'", name.c_str());
  std::vector<int64_t> want(shape);
  if (it->second.shape != want) {
    std::string got;
    for (auto v : it->second.shape) got += std::to_string(v) + ",";
    std::string w;
    for (auto v : want) w += std::to_string(v) + ",";
    return set_err(&c->err, DEER_E_INVALID, "weight '%s' has shape [%s] expected [%s]", name.c_str(), got.c_str(), w.c_str());
  }
  *out = &it->second;
  return DEER_OK;
}
// fp32 parameter kept as fp32 (LayerNorm, bias, embeddings)
static int pack_f32(deer_ctx* c, const std::string& name, std::initializer_list<int64_t> shape, float** dst) {
  Staged* s;
  RET(need(c, name, shape, &s));
  RET(dalloc(c, reinterpret_cast<void**>(dst), s->numel * 4, true));
  CK(cudaMemcpy(*dst, s->dev, s->numel * 4, cudaMemcpyDeviceToDevice));
  return DEER_OK;
}
// GEMM weight [rows, cols] -> T at dst[:, col_off : col_off+cols] of a [rows, dst_ld] matrix (dst allocated when *dst == nullptr)
static int pack_mat(deer_ctx* c, const std::string& name, int64_t rows, int64_t cols, void** dst, int64_t dst_ld = 0, int64_t col_off = 0,
                    int64_t row_off = 0, int64_t dst_rows = 0, std::initializer_list<int64_t> shape_override = {}) {
  Staged* s;
  if (shape_override.size()) RET(need(c, name, shape_override, &s));
  else RET(need(c, name, {rows, cols}, &s));
  if (dst_ld == 0) dst_ld = cols;
  if (dst_rows == 0) dst_rows = rows;
  if (*dst == nullptr) RET(dalloc(c, dst, (size_t)dst_rows * dst_ld * c->es, true, true));
  const int grid = 1184;
  char* base = reinterpret_cast<char*>(*dst) + (size_t)row_off * dst_ld * c->es;
  if (c->dtype == 1) convert_kernel<float><<<grid, 256>>>(s->dev, reinterpret_cast<float*>(base), rows, cols, cols, dst_ld, col_off);
  else convert_kernel<bf16><<<grid, 256>>>(s->dev, reinterpret_cast<bf16*>(base), rows, cols, cols, dst_ld, col_off);
  CK(cudaGetLastError());
  return DEER_OK;
}
static int host_vec(deer_ctx* c, const std::string& name, std::initializer_list<int64_t> shape, std::vector<float>* out) {
  Staged* s;
  RET(need(c, name, shape, &s));
  out->resize(s->numel);
  CK(cudaMemcpy(out->data(), s->dev, s->numel * 4, cudaMemcpyDeviceToHost));
  return DEER_OK;
}
static int upload_f32(deer_ctx* c, const std::vector<float>& v, float** dst) {
  RET(dalloc(c, reinterpret_cast<void**>(dst), v.size() * 4, true));
  CK(cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return DEER_OK;
}

// The fused probe kernel serves the trained head configuration on the bf16 product path; anything else keeps the per-op path.
static int setup_head_fused(deer_ctx* c) {
  const deer_config& f = c->cfg;
  c->hf_ok = false;
  if (getenv("DEER_HEAD_FUSED") && atoi(getenv("DEER_HEAD_FUSED")) == 0) return DEER_OK;
  if (c->dtype != 0 || f.gemm_impl != DEER_GEMM_TCGEN05) return DEER_OK;
  if (!f.lstm_layernorm || !f.mlp_layernorm || f.mlp_dim2 > 0) return DEER_OK;
  const int D = f.d_model, H = f.lstm_hidden, d0 = f.mlp_dim0, d1 = f.mlp_dim1, LL = f.lstm_layers;
  if (H % 256 || d0 % 256 || d1 % 128 || D % 256 || d1 > 512 || H > 1024 || d0 > 1024 || LL > HF_MAX_LAYERS) return DEER_OK;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  if (!prop.cooperativeLaunch) return DEER_OK;
  const int kmax = std::max(std::max(D + H, 2 * H), 2 * d0);
  const int lda = kmax + 32;                       // 64-byte skew: conflict-free 16-byte shared loads (head_fused.cuh)
  int mr = 32;
  if ((size_t)mr * lda * 2 > 200 * 1024) mr = 16;
  if ((size_t)mr * lda * 2 > 200 * 1024) return DEER_OK;
  const size_t smem = std::max((size_t)mr * lda * 2, (size_t)HF_WARPS * mr * 32 * 4);
  if (cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return DEER_OK; }
  RET(dalloc(c, reinterpret_cast<void**>(&c->hf_bar), 256, false, true));
  // Cluster size: the largest of 8 / 4 / 1 for which the GPU can co-schedule >= 128 CTAs (one per SM: every CTA of the cooperative
  // grid must be resident) -- checked with the occupancy query AND a dry-run launch (n = 0: the kernel returns at once).
  const int want_cs = getenv("DEER_HEAD_CLUSTER") ? atoi(getenv("DEER_HEAD_CLUSTER")) : 8;
  const int need = std::max(std::max(H / 8, 2 * d0 / 16), 2 * d1 / 8);
  int G = 0, CS = 0;
  for (int cs : {8, 4, 1}) {
    if (cs > want_cs) continue;
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(HF_THREADS); lc.dynamicSmemBytes = smem; lc.gridDim = dim3((DEER_NUM_SMS / cs) * cs);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension; at[1].val.clusterDim.x = cs; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = (cs > 1) ? 2 : 1;
    int grid;
    if (cs > 1) {
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, head_fused_kernel, &lc) != cudaSuccess) { cudaGetLastError(); continue; }
      grid = cs * std::min(ncl, DEER_NUM_SMS / cs);
    } else {
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, head_fused_kernel, HF_THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); continue; }
      grid = std::min(DEER_NUM_SMS, prop.multiProcessorCount);
    }
    if (grid < need) continue;
    lc.gridDim = dim3(grid); lc.stream = c->own_stream;
    HeadFusedParams hp{};
    hp.n_host = 0; hp.bar = c->hf_bar;
    if (cudaLaunchKernelEx(&lc, head_fused_kernel, hp) != cudaSuccess || cudaStreamSynchronize(c->own_stream) != cudaSuccess) { cudaGetLastError(); continue; }
    G = grid; CS = cs;
    break;
  }
  if (G == 0) return DEER_OK;
  c->hf_cluster = CS;
  c->hf_mr = mr; c->hf_lda = lda; c->hf_grid = G; c->hf_smem = smem; c->hf_ok = true;
  return DEER_OK;
}

extern "C" int deer_finalize_weights(deer_ctx* c) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  if (c->finalized) return DEER_OK;
  cudaSetDevice(c->device);
  const deer_config& f = c->cfg;
  const int64_t Wd = f.vit_width, D = f.d_model, H = f.lstm_hidden, P = f.patch;
  char b[256];
  // ---- ViT
  const std::string v = "vision_encoder.visual.";
  RET(pack_mat(c, v + "conv1.weight", Wd, 3 * P * P, &c->w_conv, c->Kp, 0, 0, 0, {Wd, 3, P, P}));
  RET(pack_f32(c, v + "class_embedding", {Wd}, &c->v_cls));
  RET(pack_f32(c, v + "positional_embedding", {c->L, Wd}, &c->v_pos));
  RET(pack_f32(c, v + "ln_pre.weight", {Wd}, &c->v_lnpre_g));
  RET(pack_f32(c, v + "ln_pre.bias", {Wd}, &c->v_lnpre_b));
  c->vit.resize(f.vit_layers);
  for (int i = 0; i < f.vit_layers; ++i) {
    snprintf(b, sizeof(b), "%stransformer.resblocks.%d.", v.c_str(), i);
    const std::string p(b);
    VitLayerW& w = c->vit[i];
    memset(&w, 0, sizeof(w));
    RET(pack_f32(c, p + "ln_1.weight", {Wd}, &w.ln1_g)); RET(pack_f32(c, p + "ln_1.bias", {Wd}, &w.ln1_b));
    RET(pack_f32(c, p + "ln_2.weight", {Wd}, &w.ln2_g)); RET(pack_f32(c, p + "ln_2.bias", {Wd}, &w.ln2_b));
    RET(pack_mat(c, p + "attn.in_proj_weight", 3 * Wd, Wd, &w.w_in)); RET(pack_f32(c, p + "attn.in_proj_bias", {3 * Wd}, &w.b_in));
    RET(pack_mat(c, p + "attn.out_proj.weight", Wd, Wd, &w.w_out)); RET(pack_f32(c, p + "attn.out_proj.bias", {Wd}, &w.b_out));
    RET(pack_mat(c, p + "mlp.c_fc.weight", f.vit_mlp, Wd, &w.w_fc)); RET(pack_f32(c, p + "mlp.c_fc.bias", {f.vit_mlp}, &w.b_fc));
    RET(pack_mat(c, p + "mlp.c_proj.weight", Wd, f.vit_mlp, &w.w_proj)); RET(pack_f32(c, p + "mlp.c_proj.bias", {Wd}, &w.b_proj));
  }
  // ---- Perceiver
  RET(pack_f32(c, "perceiver.latents", {c->nl, Wd}, &c->p_latents));
  c->perc.resize(f.perc_depth);
  for (int i = 0; i < f.perc_depth; ++i) {
    snprintf(b, sizeof(b), "perceiver.layers.%d.", i);
    const std::string p(b);
    PercLayerW& w = c->perc[i];
    memset(&w, 0, sizeof(w));
    RET(pack_f32(c, p + "0.norm_media.weight", {Wd}, &w.nm_g)); RET(pack_f32(c, p + "0.norm_media.bias", {Wd}, &w.nm_b));
    RET(pack_f32(c, p + "0.norm_latents.weight", {Wd}, &w.nl_g)); RET(pack_f32(c, p + "0.norm_latents.bias", {Wd}, &w.nl_b));
    RET(pack_mat(c, p + "0.to_q.weight", c->inner, Wd, &w.w_q));
    RET(pack_mat(c, p + "0.to_kv.weight", 2 * c->inner, Wd, &w.w_kv));
    RET(pack_mat(c, p + "0.to_out.weight", Wd, c->inner, &w.w_out));
    RET(pack_f32(c, p + "1.0.weight", {Wd}, &w.ff_g)); RET(pack_f32(c, p + "1.0.bias", {Wd}, &w.ff_b));
    RET(pack_mat(c, p + "1.1.weight", f.perc_ff_mult * Wd, Wd, &w.w_ff1));
    RET(pack_mat(c, p + "1.3.weight", Wd, f.perc_ff_mult * Wd, &w.w_ff2));
  }
  RET(pack_f32(c, "perceiver.norm.weight", {Wd}, &c->p_norm_g));
  RET(pack_f32(c, "perceiver.norm.bias", {Wd}, &c->p_norm_b));
  // ---- LLM
  RET(pack_mat(c, "lang_encoder.transformer.wte.weight", f.vocab, D, &c->wte));
  c->llm.resize(f.n_layers);
  for (int i = 0; i < f.n_layers; ++i) {
    LlmLayerW& w = c->llm[i];
    w.has_x = ((i + 1) % f.cross_attn_every == 0);
    if (w.has_x) {
      snprintf(b, sizeof(b), "lang_encoder.transformer.blocks.%d.gated_cross_attn_layer.", i);
      const std::string p(b);
      RET(pack_f32(c, p + "attn.norm.weight", {D}, &w.xn_g)); RET(pack_f32(c, p + "attn.norm.bias", {D}, &w.xn_b));
      RET(pack_mat(c, p + "attn.to_q.weight", c->xin, D, &w.xw_q));
      RET(pack_mat(c, p + "attn.to_kv.weight", 2 * c->xin, Wd, &w.xw_kv));
      RET(pack_mat(c, p + "attn.to_out.weight", D, c->xin, &w.xw_out));
      RET(pack_f32(c, p + "ff.0.weight", {D}, &w.xff_g)); RET(pack_f32(c, p + "ff.0.bias", {D}, &w.xff_b));
      RET(pack_mat(c, p + "ff.1.weight", f.xattn_ff_mult * D, D, &w.xw_ff1));
      RET(pack_mat(c, p + "ff.3.weight", D, f.xattn_ff_mult * D, &w.xw_ff2));
      std::vector<float> g1, g2;
      RET(host_vec(c, p + "attn_gate", {1}, &g1)); RET(host_vec(c, p + "ff_gate", {1}, &g2));
      g1[0] = std::tanh(g1[0]); g2[0] = std::tanh(g2[0]);      // helpers.py:272,275
      RET(upload_f32(c, g1, &w.attn_gate_tanh)); RET(upload_f32(c, g2, &w.ff_gate_tanh));
    }
    snprintf(b, sizeof(b), "lang_encoder.transformer.blocks.%d.decoder_layer.", i);
    const std::string p(b);
    if (f.qk_ln) {
      RET(pack_f32(c, p + "ln_1.weight", {D}, &w.ln1_g)); RET(pack_f32(c, p + "ln_2.weight", {D}, &w.ln2_g));
      {                                                    // [q_ln | k_ln] weights back to back: one LayerNorm launch covers both
        std::vector<float> gq, gk;
        RET(host_vec(c, p + "attn.q_ln.weight", {D}, &gq)); RET(host_vec(c, p + "attn.k_ln.weight", {D}, &gk));
        gq.insert(gq.end(), gk.begin(), gk.end());
        RET(upload_f32(c, gq, &w.qln_g));
        w.kln_g = w.qln_g + D;
      }
      RET(pack_mat(c, p + "mlp.mlp_up.weight", f.mlp_ratio * D, D, &w.w_up));
      RET(pack_mat(c, p + "mlp.mlp_down.weight", D, f.mlp_ratio * D, &w.w_down));
    } else {
      RET(pack_f32(c, p + "norm_1.weight", {D}, &w.ln1_g)); RET(pack_f32(c, p + "norm_2.weight", {D}, &w.ln2_g));
      RET(pack_mat(c, p + "ffn.up_proj.weight", f.mlp_ratio * D, D, &w.w_up));
      RET(pack_mat(c, p + "ffn.down_proj.weight", D, f.mlp_ratio * D, &w.w_down));
    }
    RET(pack_mat(c, p + "attn.Wqkv.weight", 3 * D, D, &w.w_qkv));
    RET(pack_mat(c, p + "attn.out_proj.weight", D, D, &w.w_out));
  }
  // ---- action head: LSTM weights fused as [W_ih | W_hh], biases summed
  c->lstm.resize(f.lstm_layers);
  for (int l = 0; l < f.lstm_layers; ++l) {
    LstmLayerW& w = c->lstm[l];
    memset(&w, 0, sizeof(w));
    w.in = (l == 0) ? (int)D : (int)H;
    std::string n_ih, n_hh, n_bi, n_bh;
    if (f.lstm_layernorm) {
      snprintf(b, sizeof(b), "extra_exit.rnn.layers.%d.", 3 * l);
      const std::string p(b);
      n_ih = p + "weight_ih_l0"; n_hh = p + "weight_hh_l0"; n_bi = p + "bias_ih_l0"; n_bh = p + "bias_hh_l0";
    } else {
      const std::string p = "extra_exit.rnn.", sfx = "_l" + std::to_string(l);
      n_ih = p + "weight_ih" + sfx; n_hh = p + "weight_hh" + sfx; n_bi = p + "bias_ih" + sfx; n_bh = p + "bias_hh" + sfx;
    }
    RET(pack_mat(c, n_ih, 4 * H, w.in, &w.w_cat, w.in + H, 0));
    RET(pack_mat(c, n_hh, 4 * H, H, &w.w_cat, w.in + H, w.in));
    std::vector<float> bi, bh;
    RET(host_vec(c, n_bi, {4 * H}, &bi)); RET(host_vec(c, n_bh, {4 * H}, &bh));
    for (size_t k = 0; k < bi.size(); ++k) bi[k] += bh[k];
    RET(upload_f32(c, bi, &w.b_sum));
    if (f.lstm_layernorm) {
      snprintf(b, sizeof(b), "extra_exit.rnn.layers.%d.", 3 * l + 1);
      RET(pack_f32(c, std::string(b) + "weight", {H}, &w.ln_g)); RET(pack_f32(c, std::string(b) + "bias", {H}, &w.ln_b));
    }
  }
  {
    const int64_t d0 = f.mlp_dim0, d1 = f.mlp_dim1, d2 = f.mlp_dim2, dl = d2 > 0 ? d2 : d1;
    const std::string pa = "extra_exit.actions.mlp.", pg = "extra_exit.gripper.mlp.";
    RET(pack_mat(c, pa + "1.weight", d0, H, &c->w_mlp1, H, 0, 0, 2 * d0));
    RET(pack_mat(c, pg + "1.weight", d0, H, &c->w_mlp1, H, 0, d0, 2 * d0));
    RET(pack_mat(c, pa + "5.weight", d1, d0, &c->w_mlp2_pose));
    RET(pack_mat(c, pg + "5.weight", d1, d0, &c->w_mlp2_grip));
    auto cat2 = [&](const std::string& a, const std::string& g, int64_t n, float** dst) -> int {
      std::vector<float> va, vg;
      RET(host_vec(c, a, {n}, &va)); RET(host_vec(c, g, {n}, &vg));
      va.insert(va.end(), vg.begin(), vg.end());
      return upload_f32(c, va, dst);
    };
    RET(cat2(pa + "1.bias", pg + "1.bias", d0, &c->b_mlp1));
    RET(cat2(pa + "5.bias", pg + "5.bias", d1, &c->b_mlp2));
    if (f.mlp_layernorm) {
      RET(cat2(pa + "2.weight", pg + "2.weight", d0, &c->g_mlp1)); RET(cat2(pa + "2.bias", pg + "2.bias", d0, &c->be_mlp1));
      RET(cat2(pa + "6.weight", pg + "6.weight", d1, &c->g_mlp2)); RET(cat2(pa + "6.bias", pg + "6.bias", d1, &c->be_mlp2));
    }
    if (d2 > 0) {
      RET(pack_mat(c, pa + "9.weight", d2, d1, &c->w_mlp3_pose));
      RET(pack_mat(c, pg + "9.weight", d2, d1, &c->w_mlp3_grip));
      RET(cat2(pa + "9.bias", pg + "9.bias", d2, &c->b_mlp3));
      if (f.mlp_layernorm) { RET(cat2(pa + "10.weight", pg + "10.weight", d2, &c->g_mlp3)); RET(cat2(pa + "10.bias", pg + "10.bias", d2, &c->be_mlp3)); }
    }
    const std::string io = std::to_string(d2 > 0 ? 13 : 9);
    RET(pack_f32(c, pa + io + ".weight", {6, dl}, &c->w3_pose)); RET(pack_f32(c, pa + io + ".bias", {6}, &c->b3_pose));
    RET(pack_f32(c, pg + io + ".weight", {1, dl}, &c->w3_grip)); RET(pack_f32(c, pg + io + ".bias", {1}, &c->b3_grip));
  }
  RET(setup_head_fused(c));
  CK(cudaDeviceSynchronize());
  for (auto& kv : c->staged) if (kv.second.dev) cudaFree(kv.second.dev);
  c->staged.clear();
  c->finalized = true;
  return DEER_OK;
}

extern "C" int deer_set_thresholds(deer_ctx* c, const float* thr, int n) {
  if (!c || !thr) return set_err(c ? &c->err : nullptr, DEER_E_INVALID, "null argument");
  if (n != c->n_thr) return set_err(&c->err, DEER_E_INVALID, "got %d thresholds, expected %d (value_net.py:178-179)", n, c->n_thr);
  cudaSetDevice(c->device);
  if (c->has_last) CK(cudaStreamSynchronize(c->last_stream));     // a step still in flight reads the old thresholds
  CK(cudaMemcpy(c->thr_dev, thr, n * 4, cudaMemcpyHostToDevice));
  c->thresholds_set = true;
  return DEER_OK;
}

// Calls on one ctx may arrive on different streams (caller's stream for deer_step, the engine's own stream for the
// host-buffer entry points, per-subtask resets in between).  The workspace and the per-slot state are shared, so every
// call is ordered after the previous one with an event instead of relying on the caller to synchronise.
static void order_after_last(deer_ctx* c, cudaStream_t st) {
  if (c->has_last && c->last_stream != st) cudaStreamWaitEvent(st, c->last_done, 0);
}
static void mark_done(deer_ctx* c, cudaStream_t st) {
  cudaEventRecord(c->last_done, st);
  c->last_stream = st; c->has_last = true;
}

extern "C" int deer_reset(deer_ctx* c, const int32_t* slots, int n) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  cudaSetDevice(c->device);
  const size_t per = (size_t)c->cfg.lstm_layers * c->cfg.lstm_hidden * 4;
  // stream-ordered after the last step (wherever it ran); before any step, on the engine's own stream
  cudaStream_t st = c->has_last ? c->last_stream : c->own_stream;
  if (n < 0) {
    CK(cudaMemsetAsync(c->h_state, 0, per * c->cfg.max_slots, st));
    CK(cudaMemsetAsync(c->c_state, 0, per * c->cfg.max_slots, st));
    CK(cudaMemsetAsync(c->cur_exit_id, 0, 4 * (size_t)c->cfg.max_slots, st));
    mark_done(c, st);
    return DEER_OK;
  }
  if (!slots) return set_err(&c->err, DEER_E_INVALID, "null slots");
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= c->cfg.max_slots) return set_err(&c->err, DEER_E_INVALID, "slot %d out of range", slots[i]);
  for (int i = 0; i < n; ++i) {
    CK(cudaMemsetAsync(reinterpret_cast<char*>(c->h_state) + per * slots[i], 0, per, st));
    CK(cudaMemsetAsync(reinterpret_cast<char*>(c->c_state) + per * slots[i], 0, per, st));
    CK(cudaMemsetAsync(c->cur_exit_id + slots[i], 0, 4, st));
  }
  mark_done(c, st);
  return DEER_OK;
}

// ---------------------------------------------------------------------------------------------
// step pieces (templated on the operand type T)
// ---------------------------------------------------------------------------------------------
template <typename T> struct StepImpl {
  deer_ctx* c;
  cudaStream_t st;
  int n;              // episodes in this call
  int S;
  int cur = 0;        // ping-pong index of x / media / meta / n_active

  T* tp(void* p) { return reinterpret_cast<T*>(p); }
  const int* nact() const { return c->n_active + cur; }

  int gemm(const void* A, int lda, int a_rows, const void* W, int ldw, int M, int N, int K, void* C, int ldc, int out_f32, const float* bias,
           int act, const float* scale, const float* resid, int ldr, const int* m_dev, int m_mul) {
    GemmEpi e{C, ldc, (c->dtype == 1) ? 1 : out_f32, bias, act, scale, resid, ldr, 0};
    return run_gemm(c->gemm, c->cfg.gemm_impl, c->dtype, A, lda, a_rows, W, ldw, M, N, K, e, m_dev, m_mul, st);
  }
  template <typename TIn, typename TOut>
  int ln(const TIn* in, int in_ld, const float* g, const float* b, int groups, int D, TOut* out, int out_ld, LnMap map, TOut* out2, int out2_ld,
         int relu, int rows, const int* n_dev, int n_mul, LnMap in_map = LnMap{0, 0, 0}) {
    if (rows <= 0) return DEER_OK;
    const bool vec_ok = (D % 128 == 0) && (in_ld % 4 == 0) && (out_ld % 4 == 0) && (out2 == nullptr || out2_ld % 4 == 0);
    const int nvec = D / 128;
    const int wgrid = (rows + 3) / 4;
#define LNW(NV) launch_k(layernorm_warp_kernel<TIn, TOut, NV>, dim3(wgrid), dim3(128), 0, st, c->gemm.pdl, in, in_ld, in_map, g, b, groups, out, out_ld, map, out2, out2_ld, relu, rows, n_dev, n_mul, g_snake ? 1 : 0)
    if (vec_ok && nvec == 1) LNW(1);
    else if (vec_ok && nvec == 2) LNW(2);
    else if (vec_ok && nvec == 4) LNW(4);
    else if (vec_ok && nvec == 8) LNW(8);
    else if (vec_ok && nvec == 16) LNW(16);
    else if (vec_ok && nvec == 32) LNW(32);
    else {
      const int grid = std::min(rows, DEER_NUM_SMS * 16);
      layernorm_kernel<TIn, TOut><<<grid, 128, 0, st>>>(in, in_ld, in_map, g, b, groups, D, out, out_ld, map, out2, out2_ld, relu, rows, n_dev, n_mul);
    }
#undef LNW
    KLAUNCH_CHECK("layernorm");
    return DEER_OK;
  }

  // ---- ViT-L/14 x2 + Perceiver x2 -> media (flamingo_mpt.py:609-668)
  int vision(const float* rgb, const float* grip) {
    const deer_config& f = c->cfg;
    const int NI = 2 * n, np = c->np, L = c->L, Wd = f.vit_width, NImax = 2 * f.max_batch;
    if (f.patch == 14) im2col_kernel<T, 14><<<NI * np, 128, 0, st>>>(rgb, grip, tp(c->im_cols), n, f.image_size, f.patch, c->Kp);
    else im2col_kernel<T, 0><<<NI * np, 128, 0, st>>>(rgb, grip, tp(c->im_cols), n, f.image_size, f.patch, c->Kp);
    KLAUNCH_CHECK("im2col");
    RET(gemm(c->im_cols, c->Kp, NImax * np, c->w_conv, c->Kp, NI * np, Wd, c->Kp, c->patch_out, Wd, 0, nullptr, ACT_NONE, nullptr, nullptr, 0, nullptr, 0));
    if (Wd == 1024)
      vit_assemble_warp_kernel<T, 8><<<(NI * L + 3) / 4, 128, 0, st>>>(tp(c->patch_out), c->v_cls, c->v_pos, c->v_lnpre_g, c->v_lnpre_b, c->vx, NI, L);
    else if (Wd == 128)
      vit_assemble_warp_kernel<T, 1><<<(NI * L + 3) / 4, 128, 0, st>>>(tp(c->patch_out), c->v_cls, c->v_pos, c->v_lnpre_g, c->v_lnpre_b, c->vx, NI, L);
    else
      vit_assemble_kernel<T><<<NI * L, 128, 0, st>>>(tp(c->patch_out), c->v_cls, c->v_pos, c->v_lnpre_g, c->v_lnpre_b, c->vx, NI, L, Wd);
    KLAUNCH_CHECK("vit_assemble");
    const int Mmax = NImax * L;
    const LnMap id{0, 0, 0};
    // The 24 ViT blocks run over CHUNKS of images: the chunk's residual stream and the (re-used) ln / qkv / attention /
    // MLP scratch stay resident in the 126 MB L2 from one GEMM to the next instead of round-tripping through HBM.
    const int chunk = (c->vit_chunk > 0) ? std::min(c->vit_chunk, NI) : NI;
    for (int i0 = 0; i0 < NI; i0 += chunk) {
      const int nc = std::min(chunk, NI - i0);
      const int M = nc * L;
      float* vxc = c->vx + (size_t)i0 * L * Wd;
      for (int i = 0; i < f.vit_layers; ++i) {
        const VitLayerW& w = c->vit[i];
        RET(ln<float, T>(vxc, Wd, w.ln1_g, w.ln1_b, 1, Wd, tp(c->v_ln), Wd, id, nullptr, 0, 0, M, nullptr, 0));
        RET(gemm(c->v_ln, Wd, Mmax, w.w_in, Wd, M, 3 * Wd, Wd, c->v_qkv, 3 * Wd, 0, w.b_in, ACT_NONE, nullptr, nullptr, 0, nullptr, 0));
        AttnParams a{};
        a.Q = c->v_qkv; a.K = tp(c->v_qkv) + Wd; a.V = tp(c->v_qkv) + 2 * Wd; a.O = c->v_att;
        a.ldq = a.ldk = a.ldv = 3 * Wd; a.ldo = Wd; a.Sq = a.Sk = L; a.H = f.vit_heads; a.hd = 64; a.scale = 0.125f; a.B_host = nc;
        RET(run_attention(0, c->dtype, a, &c->err, &c->stats, st, c->gemm.pdl));
        RET(gemm(c->v_att, Wd, Mmax, w.w_out, Wd, M, Wd, Wd, vxc, Wd, 1, w.b_out, ACT_NONE, nullptr, vxc, Wd, nullptr, 0));
        RET(ln<float, T>(vxc, Wd, w.ln2_g, w.ln2_b, 1, Wd, tp(c->v_ln), Wd, id, nullptr, 0, 0, M, nullptr, 0));
        RET(gemm(c->v_ln, Wd, Mmax, w.w_fc, Wd, M, f.vit_mlp, Wd, c->v_h, f.vit_mlp, 0, w.b_fc, ACT_QUICKGELU, nullptr, nullptr, 0, nullptr, 0));
        RET(gemm(c->v_h, f.vit_mlp, Mmax, w.w_proj, f.vit_mlp, M, Wd, f.vit_mlp, vxc, Wd, 1, w.b_proj, ACT_NONE, nullptr, vxc, Wd, nullptr, 0));
      }
    }
    if (f.debug_taps) {
      vit_tokens_kernel<<<1184, 256, 0, st>>>(c->vx, c->tap_vit, NI, L, Wd);
      KLAUNCH_CHECK("vit_tokens");
    }
    // ---- Perceiver resampler (helpers.py:107-132), shared weights for both cameras
    const int nl = c->nl, inner = c->inner, kvr = np + nl;
    broadcast_rows_kernel<<<1184, 256, 0, st>>>(c->p_latents, c->lat, (size_t)nl * Wd, (size_t)NI * nl * Wd);
    KLAUNCH_CHECK("latents_init");
    c->gemm.tag = 1;
    for (int i = 0; i < f.perc_depth; ++i) {
      const PercLayerW& w = c->perc[i];
      // norm_media(x): read the patch tokens straight out of the ViT residual stream (skip CLS), write rows [0,np) of each kv block
      RET(ln<float, T>(c->vx, Wd, w.nm_g, w.nm_b, 1, Wd, tp(c->p_kvin), Wd, LnMap{np, kvr, 0}, nullptr, 0, 0, NI * np, nullptr, 0, LnMap{np, L, 1}));
      RET(ln<float, T>(c->lat, Wd, w.nl_g, w.nl_b, 1, Wd, tp(c->p_kvin), Wd, LnMap{nl, kvr, np}, tp(c->p_latn), Wd, 0, NI * nl, nullptr, 0));
      RET(gemm(c->p_latn, Wd, NImax * nl, w.w_q, Wd, NI * nl, inner, Wd, c->p_q, inner, 0, nullptr, ACT_NONE, nullptr, nullptr, 0, nullptr, 0));
      RET(gemm(c->p_kvin, Wd, NImax * kvr, w.w_kv, Wd, NI * kvr, 2 * inner, Wd, c->p_kv, 2 * inner, 0, nullptr, ACT_NONE, nullptr, nullptr, 0, nullptr, 0));
      AttnParams a{};
      a.Q = c->p_q; a.K = c->p_kv; a.V = tp(c->p_kv) + inner; a.O = c->p_att;
      a.ldq = inner; a.ldk = a.ldv = 2 * inner; a.ldo = inner; a.Sq = nl; a.Sk = kvr; a.H = f.perc_heads; a.hd = 64; a.scale = 0.125f; a.B_host = NI;
      RET(run_attention(0, c->dtype, a, &c->err, &c->stats, st, c->gemm.pdl));
      RET(gemm(c->p_att, inner, NImax * nl, w.w_out, inner, NI * nl, Wd, inner, c->lat, Wd, 1, nullptr, ACT_NONE, nullptr, c->lat, Wd, nullptr, 0));
      RET(ln<float, T>(c->lat, Wd, w.ff_g, w.ff_b, 1, Wd, tp(c->p_latn), Wd, id, nullptr, 0, 0, NI * nl, nullptr, 0));
      const int ffd = f.perc_ff_mult * Wd;
      RET(gemm(c->p_latn, Wd, NImax * nl, w.w_ff1, Wd, NI * nl, ffd, Wd, c->p_ffh, ffd, 0, nullptr, ACT_GELU, nullptr, nullptr, 0, nullptr, 0));
      RET(gemm(c->p_ffh, ffd, NImax * nl, w.w_ff2, ffd, NI * nl, Wd, ffd, c->lat, Wd, 1, nullptr, ACT_NONE, nullptr, c->lat, Wd, nullptr, 0));
    }
    // final norm, written straight into media[e, cam*nl + j, :]  (cat along the latent dim, flamingo_mpt.py:661)
    for (int cam = 0; cam < 2; ++cam)
      RET(ln<float, T>(c->lat + (size_t)cam * n * nl * Wd, Wd, c->p_norm_g, c->p_norm_b, 1, Wd, tp(c->media[cur]), Wd, LnMap{nl, 2 * nl, cam * nl},
                       nullptr, 0, 0, n * nl, nullptr, 0));
    if (f.debug_taps) {
      tap_scatter_kernel<T><<<dim3(n, 8), 256, 0, st>>>(tp(c->media[cur]), c->tap_media, nullptr, (size_t)2 * nl * Wd, n, nullptr);
      KLAUNCH_CHECK("tap_media");
    }
    return DEER_OK;
  }

  // ---- one FlamingoLayer: gated x-attn (if any) then the MPT block (flamingo_lm.py:46-83)
  int llm_layer(int li) {
    const deer_config& f = c->cfg;
    const LlmLayerW& w = c->llm[li];
    const int D = f.d_model, M = n * S, Bm = f.max_batch, Mmax = Bm * f.max_seq, nl2 = 2 * c->nl, Wd = f.vit_width, xin = c->xin;
    float* x = c->x[cur];
    const int* nd = nact();
    const LnMap id{0, 0, 0};
    if (w.has_x) {
      RET(ln<float, T>(x, D, w.xn_g, w.xn_b, 1, D, tp(c->l_ln), D, id, nullptr, 0, 0, M, nd, S));
      RET(gemm(c->l_ln, D, Mmax, w.xw_q, D, M, xin, D, c->l_q, xin, 0, nullptr, ACT_NONE, nullptr, nullptr, 0, nd, S));
      RET(gemm(c->media[cur], Wd, Bm * nl2, w.xw_kv, Wd, n * nl2, 2 * xin, Wd, c->l_kv, 2 * xin, 0, nullptr, ACT_NONE, nullptr, nullptr, 0, nd, nl2));
      AttnParams a{};
      a.Q = c->l_q; a.K = c->l_kv; a.V = tp(c->l_kv) + xin; a.O = c->l_att;
      a.ldq = xin; a.ldk = a.ldv = 2 * xin; a.ldo = xin; a.Sq = S; a.Sk = nl2; a.H = f.xattn_heads; a.hd = 64; a.scale = 0.125f;
      a.qmode = c->meta[cur].xmode; a.B_host = n; a.n_dev = nd;
      RET(run_attention(2, c->dtype, a, &c->err, &c->stats, st, c->gemm.pdl));
      RET(gemm(c->l_att, xin, Mmax, w.xw_out, xin, M, D, xin, x, D, 1, nullptr, ACT_NONE, w.attn_gate_tanh, x, D, nd, S));
      RET(ln<float, T>(x, D, w.xff_g, w.xff_b, 1, D, tp(c->l_ln), D, id, nullptr, 0, 0, M, nd, S));
      const int ffd = f.xattn_ff_mult * D;
      RET(gemm(c->l_ln, D, Mmax, w.xw_ff1, D, M, ffd, D, c->l_h, ffd, 0, nullptr, ACT_GELU, nullptr, nullptr, 0, nd, S));
      RET(gemm(c->l_h, ffd, Mmax, w.xw_ff2, ffd, M, D, ffd, x, D, 1, nullptr, ACT_NONE, w.ff_gate_tanh, x, D, nd, S));
    }
    // MPT block: pre-LN, bias-free
    RET(ln<float, T>(x, D, w.ln1_g, nullptr, 1, D, tp(c->l_ln), D, id, nullptr, 0, 0, M, nd, S));
    RET(gemm(c->l_ln, D, Mmax, w.w_qkv, D, M, 3 * D, D, c->l_qkv, 3 * D, 0, nullptr, ACT_NONE, nullptr, nullptr, 0, nd, S));
    if (f.qk_ln) {   // LayerNorm over the full d_model of q and of k before the head split (attn_qk_ln)
      // ONE launch for both: logical row r = 2 * token + part (part 0 = q, 1 = k) lives at row 3 * token + part of the [.., D]-pitched
      // view of the fused qkv buffer; gamma group r % 2
      const LnMap qk{2, 3, 0};
      RET(ln<T, T>(tp(c->l_qkv), D, w.qln_g, nullptr, 2, D, tp(c->l_qkv), D, qk, nullptr, 0, 0, 2 * M, nd, 2 * S, qk));
    }
    AttnParams a{};
    a.Q = c->l_qkv; a.K = tp(c->l_qkv) + D; a.V = tp(c->l_qkv) + 2 * D; a.O = c->l_att;
    a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D; a.Sq = a.Sk = S; a.H = f.n_heads; a.hd = c->hd; a.scale = 1.0f / std::sqrt((float)c->hd);
    a.causal = 1; a.alibi_slopes = c->alibi; a.kv_len = c->meta[cur].valid_len; a.B_host = n; a.n_dev = nd;
    RET(run_attention(2, c->dtype, a, &c->err, &c->stats, st, c->gemm.pdl));
    RET(gemm(c->l_att, D, Mmax, w.w_out, D, M, D, D, x, D, 1, nullptr, ACT_NONE, nullptr, x, D, nd, S));
    RET(ln<float, T>(x, D, w.ln2_g, nullptr, 1, D, tp(c->l_ln), D, id, nullptr, 0, 0, M, nd, S));
    const int ffd = f.mlp_ratio * D;
    RET(gemm(c->l_ln, D, Mmax, w.w_up, D, M, ffd, D, c->l_h, ffd, 0, nullptr, ACT_GELU, nullptr, nullptr, 0, nd, S));
    RET(gemm(c->l_h, ffd, Mmax, w.w_down, ffd, M, D, ffd, x, D, 1, nullptr, ACT_NONE, nullptr, x, D, nd, S));
    if (f.debug_taps) {
      tap_scatter_kernel<float><<<dim3(n, 8), 256, 0, st>>>(x, c->tap_hidden + (size_t)li * c->tap_n * S * D, c->meta[cur].orig, (size_t)S * D, n, nd);
      KLAUNCH_CHECK("tap_hidden");
    }
    return DEER_OK;
  }

  // ---- DeterministicDecoder single-step forward on x (action_head.py:499-611) -> cur_pose / cur_grip / candidate (h,c)
  int head_eval() {
    struct TagGuard { GemmCtx& g; int old; TagGuard(GemmCtx& g_) : g(g_), old(g_.tag) { g.tag = 3; } ~TagGuard() { g.tag = old; } } tg(c->gemm);
    const deer_config& f = c->cfg;
    const int D = f.d_model, H = f.lstm_hidden, LL = f.lstm_layers, Bm = f.max_batch, d0 = f.mlp_dim0, d1 = f.mlp_dim1;
    const int* nd = nact();
    const EpisodeMeta& m = c->meta[cur];
    if (c->hf_ok && n <= c->hf_mr) {
      HeadFusedParams hp{};
      hp.x = c->x[cur]; hp.valid_len = m.valid_len; hp.slot = m.slot; hp.h_state = c->h_state; hp.c_state = c->c_state;
      hp.S = S; hp.D = D; hp.H = H; hp.LL = LL; hp.nstride = Bm; hp.d0 = d0; hp.d1 = d1; hp.mr = c->hf_mr; hp.lda = c->hf_lda;
      for (int l = 0; l < LL; ++l) {
        hp.w_cat[l] = reinterpret_cast<const bf16*>(c->lstm[l].w_cat); hp.b_sum[l] = c->lstm[l].b_sum;
        hp.ln_g[l] = c->lstm[l].ln_g; hp.ln_b[l] = c->lstm[l].ln_b;
      }
      hp.w_mlp1 = reinterpret_cast<const bf16*>(c->w_mlp1); hp.b_mlp1 = c->b_mlp1; hp.g_mlp1 = c->g_mlp1; hp.be_mlp1 = c->be_mlp1;
      hp.w_mlp2_pose = reinterpret_cast<const bf16*>(c->w_mlp2_pose); hp.w_mlp2_grip = reinterpret_cast<const bf16*>(c->w_mlp2_grip);
      hp.b_mlp2 = c->b_mlp2; hp.g_mlp2 = c->g_mlp2; hp.be_mlp2 = c->be_mlp2;
      hp.w3_pose = c->w3_pose; hp.b3_pose = c->b3_pose; hp.w3_grip = c->w3_grip; hp.b3_grip = c->b3_grip;
      hp.pooled = c->pooled; hp.mlp1 = c->mlp1; hp.mlp2 = c->mlp2; hp.cand_h = c->cand_h; hp.cand_c = c->cand_c;
      hp.cur_pose = c->cur_pose; hp.cur_grip = c->cur_grip; hp.bar = c->hf_bar; hp.n_host = n; hp.n_dev = nd; hp.dbg = g_head_dbg;
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(c->hf_grid); lc.blockDim = dim3(HF_THREADS); lc.dynamicSmemBytes = c->hf_smem; lc.stream = st;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
      at[1].id = cudaLaunchAttributeClusterDimension; at[1].val.clusterDim.x = c->hf_cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = (c->hf_cluster > 1) ? 2 : 1;
      const cudaError_t le = cudaLaunchKernelEx(&lc, head_fused_kernel, hp);
      if (le != cudaSuccess) return set_err(&c->err, DEER_E_CUDA, "head_fused launch failed: %s", cudaGetErrorString(le));
      c->stats.total++; c->stats.step++;
      if (f.debug_taps) {
        tap_scatter_kernel<float><<<dim3(n, 1), 256, 0, st>>>(c->pooled, c->tap_pooled, nullptr, (size_t)D, n, nd);
        KLAUNCH_CHECK("tap_pooled");
      }
      return DEER_OK;
    }
    launch_k(maxpool_kernel<T>, dim3((D + 255) / 256, n), dim3(256), 0, st, c->gemm.pdl, c->x[cur], m.valid_len, tp(c->a0), D + H, c->pooled, S, D, n, nd);
    KLAUNCH_CHECK("maxpool");
    if (f.debug_taps) {
      tap_scatter_kernel<float><<<dim3(n, 1), 256, 0, st>>>(c->pooled, c->tap_pooled, nullptr, (size_t)D, n, nd);
      KLAUNCH_CHECK("tap_pooled");
    }
    launch_k(lstm_gather_h_kernel<T>, dim3(n, LL), dim3(256), 0, st, c->gemm.pdl, c->h_state, m.slot, tp(c->a0), D + H, D, tp(c->a_rest), 2 * H, LL, H, Bm, n, nd);
    KLAUNCH_CHECK("lstm_gather_h");
    for (int l = 0; l < LL; ++l) {
      const LstmLayerW& w = c->lstm[l];
      const int K = w.in + H;
      const void* A = (l == 0) ? c->a0 : static_cast<const void*>(reinterpret_cast<char*>(c->a_rest) + (size_t)(l - 1) * Bm * 2 * H * c->es);
      RET(gemm(A, K, Bm, w.w_cat, K, n, 4 * H, K, c->gates, 4 * H, 1, w.b_sum, ACT_NONE, nullptr, nullptr, 0, nd, 1));
      T* y = (l + 1 < LL) ? tp(c->a_rest) + (size_t)l * Bm * 2 * H : tp(c->head_in);
      const int y_ld = (l + 1 < LL) ? 2 * H : H;
      launch_k(lstm_cell_kernel<T>, dim3(n), dim3(256), 0, st, c->gemm.pdl, c->gates, c->c_state, m.slot, l, LL, c->cand_h, c->cand_c, w.ln_g, w.ln_b, y, y_ld, H, Bm, n, nd);
      KLAUNCH_CHECK("lstm_cell");
    }
    // both MLP heads: [Linear, (LN), ReLU] x (2 or 3), the first layer fused across the heads, then the tiny output layers.
    // With LayerNorm the GEMM leaves fp32 pre-activations and the LN kernel applies LN+ReLU; without it ReLU rides the epilogue.
    const int d2 = f.mlp_dim2;
    const bool lnm = f.mlp_layernorm != 0;
    const LnMap idm{0, 0, 0};
    if (lnm) {
      RET(gemm(c->head_in, H, Bm, c->w_mlp1, H, n, 2 * d0, H, c->mlp1, 2 * d0, 1, c->b_mlp1, ACT_NONE, nullptr, nullptr, 0, nd, 1));
      RET(ln<float, T>(c->mlp1, d0, c->g_mlp1, c->be_mlp1, 2, d0, tp(c->y1), d0, idm, nullptr, 0, 1, 2 * n, nd, 2));
    } else {
      RET(gemm(c->head_in, H, Bm, c->w_mlp1, H, n, 2 * d0, H, c->y1, 2 * d0, 0, c->b_mlp1, ACT_RELU, nullptr, nullptr, 0, nd, 1));
    }
    const float* y_last = c->y2;
    int d_last = d1;
    {
      const bool last = (d2 == 0);
      void* dst = lnm ? static_cast<void*>(c->mlp2) : (last ? static_cast<void*>(c->y2) : c->y2t);
      const int of32 = (lnm || last) ? 1 : 0;
      const int esz = (of32 || c->dtype == 1) ? 4 : 2;
      RET(gemm(c->y1, 2 * d0, Bm, c->w_mlp2_pose, d0, n, d1, d0, dst, 2 * d1, of32, c->b_mlp2, lnm ? ACT_NONE : ACT_RELU, nullptr, nullptr, 0, nd, 1));
      RET(gemm(tp(c->y1) + d0, 2 * d0, Bm, c->w_mlp2_grip, d0, n, d1, d0, reinterpret_cast<char*>(dst) + (size_t)d1 * esz, 2 * d1, of32, c->b_mlp2 + d1,
               lnm ? ACT_NONE : ACT_RELU, nullptr, nullptr, 0, nd, 1));
      if (lnm) {
        if (last) RET(ln<float, float>(c->mlp2, d1, c->g_mlp2, c->be_mlp2, 2, d1, c->y2, d1, idm, nullptr, 0, 1, 2 * n, nd, 2));
        else RET(ln<float, T>(c->mlp2, d1, c->g_mlp2, c->be_mlp2, 2, d1, tp(c->y2t), d1, idm, nullptr, 0, 1, 2 * n, nd, 2));
      }
    }
    if (d2 > 0) {
      void* dst = lnm ? static_cast<void*>(c->mlp3) : static_cast<void*>(c->y3);
      RET(gemm(c->y2t, 2 * d1, Bm, c->w_mlp3_pose, d1, n, d2, d1, dst, 2 * d2, 1, c->b_mlp3, lnm ? ACT_NONE : ACT_RELU, nullptr, nullptr, 0, nd, 1));
      RET(gemm(tp(c->y2t) + d1, 2 * d1, Bm, c->w_mlp3_grip, d1, n, d2, d1, reinterpret_cast<float*>(dst) + d2, 2 * d2, 1, c->b_mlp3 + d2,
               lnm ? ACT_NONE : ACT_RELU, nullptr, nullptr, 0, nd, 1));
      if (lnm) RET(ln<float, float>(c->mlp3, d2, c->g_mlp3, c->be_mlp3, 2, d2, c->y3, d2, idm, nullptr, 0, 1, 2 * n, nd, 2));
      y_last = c->y3; d_last = d2;
    }
    launch_k(head_out_kernel, dim3(n), dim3(256), 0, st, c->gemm.pdl, y_last, d_last, c->w3_pose, c->b3_pose, c->w3_grip, c->b3_grip, c->cur_pose, c->cur_grip, n, nd);
    KLAUNCH_CHECK("head_out");
    return DEER_OK;
  }
};

template <typename T>
static int step_impl(deer_ctx* c, int n, const int32_t* slots, const float* rgb, const float* grip, const int64_t* ids, const uint8_t* mask,
                     int S, const int32_t* timesteps, int static_exit, uint32_t flags, float* pose_out, float* grip_out, float* action_out,
                     int32_t* exit_out, float* delta_out, cudaStream_t st) {
  const deer_config& f = c->cfg;
  StepImpl<T> s{c, st, n, S, 0};
  const bool gen = (flags & DEER_FLAG_GENERATE) != 0;
  const bool full_depth = gen || (flags & DEER_FLAG_FULL_DEPTH) != 0;
  const int nthr_step = gen ? (int)c->exits.size() : c->n_thr;           // generate mode reports a delta for EVERY exit
  if (!pose_out) pose_out = c->o_pose;
  if (!grip_out) grip_out = c->o_grip;
  if (!action_out) action_out = c->o_action;
  if (!exit_out) exit_out = c->o_exit;
  if (!delta_out) delta_out = c->o_delta;
  c->tap_n = n; c->tap_S = S;
  init_episodes_kernel<<<(n + 127) / 128, 128, 0, s.st>>>(c->meta[0], c->n_active, slots, ids, mask, timesteps, n, S, f.media_token_id,
                                                          f.steps_per_stage, nthr_step, exit_out, delta_out, f.max_slots);
  KLAUNCH_CHECK("init_episodes");
  if (gen) CK(cudaMemsetAsync(c->pend_flag, 0, n, s.st));
  c->n_probes_last = 0; c->n_last = n;
  c->gemm.epoch = 0;
  if (c->gemm.profile) {
    for (auto& r : c->gemm.recs) { c->gemm.pool.push_back(r.e0); c->gemm.pool.push_back(r.e1); }
    c->gemm.recs.clear();
    if (!c->ev_step0) { cudaEventCreate(&c->ev_step0); cudaEventCreate(&c->ev_step1); }
    cudaEventRecord(c->ev_step0, s.st);
  }
  c->gemm.tag = 0;
  static const bool pdl_in_graph = getenv("DEER_PDL_GRAPH") ? atoi(getenv("DEER_PDL_GRAPH")) != 0 : true;   // programmatic edges inside the captured graph (B=1: 3.03 -> 2.89 ms)
  c->gemm.pdl = (c->cap_level < 0 || pdl_in_graph) && !c->gemm.profile;
  RET(s.vision(rgb, grip));
  c->gemm.tag = 2;
  if (c->dtype == 1) embed_kernel<float><<<n * S, 256, 0, s.st>>>(ids, reinterpret_cast<const float*>(c->wte), c->x[0], n * S, f.d_model, f.vocab);
  else embed_kernel<bf16><<<n * S, 256, 0, s.st>>>(ids, reinterpret_cast<const bf16*>(c->wte), c->x[0], n * S, f.d_model, f.vocab);
  KLAUNCH_CHECK("embed");
  const int last_layer = (static_exit >= 0) ? static_exit : (gen ? c->exits.back() : c->ctrl_max_layer);
  const int first_exit = c->exits[0];
  const int first_prev = gen ? 0 : first_exit - 1;        // generate mode compares the first exit with layer 0 (value_net.py:136)
  for (int li = 0; li <= last_layer; ++li) {
    RET(s.llm_layer(li));
    bool is_exit = false;
    int exit_idx = -1;
    if (static_exit >= 0) {
      is_exit = (li == static_exit);
    } else {
      for (size_t k = 0; k < c->exits.size(); ++k)
        if (c->exits[k] == li) { is_exit = true; exit_idx = (int)k; }
      if (li == first_prev) {
        // first exit compares with the head applied to the PREVIOUS layer's features (value_net.py:122-125)
        RET(s.head_eval());
        copy_pose_kernel<<<(n * 6 + 127) / 128, 128, 0, s.st>>>(c->cur_pose, c->meta[s.cur].prev_pose, n, s.nact());
        KLAUNCH_CHECK("copy_pose");
      }
    }
    if (!is_exit) continue;
    RET(s.head_eval());
    GateParams g{};
    g.layer = li; g.exit_idx = (exit_idx >= 0 && exit_idx < nthr_step) ? exit_idx : -1;
    g.forced = (static_exit >= 0) ? 1 : (li >= last_layer ? 1 : 0);
    if (gen) { g.commit_layer = c->gen_commit_layer; g.pend_flag = c->pend_flag; }
    g.full_depth = full_depth ? 1 : 0; g.has_prev = (static_exit >= 0) ? 0 : 1;
    g.threshold_type = f.threshold_type; g.n_thr = nthr_step; g.thresholds = c->thr_dev;
    g.cur_pose = c->cur_pose; g.cur_grip = c->cur_grip; g.cur_exit_id = c->cur_exit_id;
    g.pose_out = pose_out; g.grip_out = grip_out; g.action_out = action_out; g.exit_out = exit_out; g.delta_out = delta_out;
    g.exit_flag = c->exit_flag; g.src_of_new = c->src_of_new;
    g.ensemble = (f.action_ensemble && static_exit < 0 && !gen) ? 1 : 0; g.probe_idx = c->n_probes_last;
    const int nxt = s.cur ^ 1;
    // graph capture: the layers after this exit go into a conditional IF node; the gate kernel sets its condition
    const bool wrap_cond = (c->cap_level >= 0) && !g.forced;
    if (wrap_cond) {
      cudaStreamCaptureStatus cst; unsigned long long cid; cudaGraph_t cg; const cudaGraphNode_t* deps; size_t nd;
      CK(cudaStreamGetCaptureInfo_v2(s.st, &cst, &cid, &cg, &deps, &nd));
      CK(cudaGraphConditionalHandleCreate(&g.cond, cg, 0, cudaGraphCondAssignDefault));
      g.has_cond = 1;
    }
    exit_gate_kernel<<<1, 1024, 0, s.st>>>(g, c->meta[s.cur], c->n_active + s.cur, c->n_active + nxt,
                                           c->nact_hist + std::min(c->n_probes_last, 15));
    KLAUNCH_CHECK("exit_gate");
    c->n_probes_last++;
    c->gemm.epoch = c->n_probes_last;
    if (gen) {
      park_state_kernel<<<dim3(n, f.lstm_layers), 256, 0, s.st>>>(c->exit_flag, c->meta[s.cur].orig, c->cand_h, c->cand_c, c->pend_h, c->pend_c,
                                                               f.lstm_layers, f.lstm_hidden, f.max_batch, n);
      KLAUNCH_CHECK("park_state");
      if (g.forced) {
        commit_pending_kernel<<<dim3(n, f.lstm_layers), 256, 0, s.st>>>(c->pend_flag, slots, c->pend_h, c->pend_c, c->h_state, c->c_state,
                                                                     f.lstm_layers, f.lstm_hidden, n);
        KLAUNCH_CHECK("commit_pending");
        break;
      }
      // nobody leaves before the last layer: no compaction, only action_list[-1] moves on (value_net.py:127)
      copy_pose_kernel<<<(n * 6 + 127) / 128, 128, 0, s.st>>>(c->cur_pose, c->meta[s.cur].prev_pose, n, s.nact());
      KLAUNCH_CHECK("copy_pose");
      continue;
    }
    commit_state_kernel<<<dim3(n, f.lstm_layers), 256, 0, s.st>>>(c->exit_flag, c->meta[s.cur].slot, c->cand_h, c->cand_c, c->h_state, c->c_state,
                                                               f.lstm_layers, f.lstm_hidden, f.max_batch, n, c->n_active + s.cur,
                                                               c->meta[s.cur].orig, c->pooled, c->o_feat, f.d_model);
    KLAUNCH_CHECK("commit_state");
    if (g.forced) break;
    compact_kernel<T><<<dim3(n, 8), 256, 0, s.st>>>(c->src_of_new, c->x[s.cur], c->x[nxt], (size_t)S * f.d_model, s.tp(c->media[s.cur]), s.tp(c->media[nxt]),
                                                  (size_t)2 * c->nl * f.vit_width, c->meta[s.cur], c->meta[nxt], c->cur_pose, c->cur_grip, S, n, c->n_active + nxt);
    KLAUNCH_CHECK("compact");
    s.cur = nxt;
    if (wrap_cond) {
      cudaStreamCaptureStatus cst; unsigned long long cid; cudaGraph_t cg; const cudaGraphNode_t* deps; size_t nd;
      CK(cudaStreamGetCaptureInfo_v2(s.st, &cst, &cid, &cg, &deps, &nd));
      cudaGraphNodeParams np = {};
      np.type = cudaGraphNodeTypeConditional;
      np.conditional.handle = g.cond;
      np.conditional.type = cudaGraphCondTypeIf;
      np.conditional.size = 1;
      cudaGraphNode_t cnode;
      CK(cudaGraphAddNode(&cnode, cg, deps, nd, &np));
      CK(cudaStreamUpdateCaptureDependencies(s.st, &cnode, 1, cudaStreamSetCaptureDependencies));
      cudaGraph_t body = np.conditional.phGraph_out[0];
      c->cap_level++;
      if (c->cap_level >= 19) return set_err(&c->err, DEER_E_INVALID, "too many nested exits for graph capture");
      if (!c->cap_streams[c->cap_level]) CK(cudaStreamCreateWithFlags(&c->cap_streams[c->cap_level], cudaStreamNonBlocking));
      CK(cudaStreamBeginCaptureToGraph(c->cap_streams[c->cap_level], body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
      s.st = c->cap_streams[c->cap_level];
    }
    if (f.host_control) {
      // the reference's per-probe device->host sync (value_net.py:293): stop LAUNCHING once every episode has exited
      CK(cudaMemcpyAsync(c->h_nact, c->n_active + s.cur, 4, cudaMemcpyDeviceToHost, s.st));
      CK(cudaStreamSynchronize(s.st));
      if (*c->h_nact == 0) break;
      s.n = *c->h_nact;                       // exact-sized launches from here on
    }
  }
  if (c->gemm.profile) cudaEventRecord(c->ev_step1, s.st);
  return DEER_OK;
}

static int validate_step(deer_ctx* c, int n, int S, int static_exit, uint32_t flags) {
  if (!c->finalized) return set_err(&c->err, DEER_E_STATE, "weights not finalized");
  if (static_exit < 0 && !(flags & DEER_FLAG_FULL_DEPTH) && !c->thresholds_set)
    return set_err(&c->err, DEER_E_STATE, "Please set thresholds before calling forward (value_net.py:279)");
  if (n < 1 || n > c->cfg.max_batch) return set_err(&c->err, DEER_E_INVALID, "n=%d outside [1,%d]", n, c->cfg.max_batch);
  if (S < 1 || S > c->cfg.max_seq) return set_err(&c->err, DEER_E_INVALID, "S=%d outside [1,%d]", S, c->cfg.max_seq);
  if (static_exit >= c->cfg.n_layers)
    return set_err(&c->err, DEER_E_INVALID, "exit_id %d out of range [0, %d) (flamingo_mpt.py:449)", static_exit, c->cfg.n_layers);
  return DEER_OK;
}

static int step_dispatch(deer_ctx* c, int n, const int32_t* slots, const float* rgb, const float* gripper, const int64_t* ids,
                         const uint8_t* mask, int S, const int32_t* timesteps, int static_exit, uint32_t flags, float* pose_out,
                         float* gripper_out, float* action_out, int32_t* exit_out, float* delta_out, cudaStream_t st) {
  if (c->dtype == 1)
    return step_impl<float>(c, n, slots, rgb, gripper, ids, mask, S, timesteps, static_exit, flags, pose_out, gripper_out, action_out, exit_out, delta_out, st);
  return step_impl<bf16>(c, n, slots, rgb, gripper, ids, mask, S, timesteps, static_exit, flags, pose_out, gripper_out, action_out, exit_out, delta_out, st);
}

// CUDA-graph mode: inputs are staged into engine-owned buffers (fixed addresses), the whole step -- including the nested
// conditional IF nodes around every post-exit segment -- is replayed with one cudaGraphLaunch, outputs are copied out.
static int step_graph(deer_ctx* c, int n, const int32_t* slots, const float* rgb, const float* gripper, const int64_t* ids,
                      const uint8_t* mask, int S, const int32_t* timesteps, int static_exit, uint32_t flags, float* pose_out,
                      float* gripper_out, float* action_out, int32_t* exit_out, float* delta_out, cudaStream_t st) {
  const size_t img = (size_t)3 * c->cfg.image_size * c->cfg.image_size * 4;
  auto stage = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    return (dst == src) ? cudaSuccess : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
  };
  CK(stage(c->s_rgb, rgb, n * img));
  CK(stage(c->s_grip, gripper, n * img));
  CK(stage(c->s_ids, ids, (size_t)n * S * 8));
  CK(stage(c->s_mask, mask, (size_t)n * S));
  CK(stage(c->s_slots, slots, (size_t)n * 4));
  if (timesteps) CK(stage(c->s_ts, timesteps, (size_t)n * 4));
  auto key = std::make_tuple(n, S, static_exit, flags, timesteps ? 1 : 0);
  deer_ctx::GraphEntry& ge = c->graphs[key];
  int r = DEER_OK;
  if (ge.exec == nullptr && ge.seen == 0) {
    // first use of this shape: run eagerly once (sets function attributes, builds the TMA descriptors)
    ge.seen = 1;
    r = step_dispatch(c, n, c->s_slots, c->s_rgb, c->s_grip, c->s_ids, c->s_mask, S, timesteps ? c->s_ts : nullptr, static_exit, flags,
                      c->o_pose, c->o_grip, c->o_action, c->o_exit, c->o_delta, st);
    if (r != DEER_OK) return r;
  } else {
    if (ge.exec == nullptr) {
      CK(cudaStreamSynchronize(st));
      CK(cudaGraphCreate(&ge.graph, 0));
      if (!c->cap_streams[0]) CK(cudaStreamCreateWithFlags(&c->cap_streams[0], cudaStreamNonBlocking));
      CK(cudaStreamBeginCaptureToGraph(c->cap_streams[0], ge.graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
      c->cap_level = 0;
      r = step_dispatch(c, n, c->s_slots, c->s_rgb, c->s_grip, c->s_ids, c->s_mask, S, timesteps ? c->s_ts : nullptr, static_exit, flags,
                        c->o_pose, c->o_grip, c->o_action, c->o_exit, c->o_delta, c->cap_streams[0]);
      // close the nested captures from the innermost conditional body outwards
      cudaError_t ce = cudaSuccess;
      for (int lvl = c->cap_level; lvl >= 0; --lvl) {
        cudaGraph_t out = nullptr;
        cudaError_t e2 = cudaStreamEndCapture(c->cap_streams[lvl], &out);
        if (e2 != cudaSuccess && ce == cudaSuccess) ce = e2;
      }
      c->cap_level = -1;
      if (r != DEER_OK) return r;
      if (ce != cudaSuccess) return set_err(&c->err, DEER_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
      CK(cudaGraphInstantiate(&ge.exec, ge.graph, 0));
      ge.launches = c->stats.step; ge.gemms = c->stats.gemm_step; ge.flops = c->stats.gemm_flops_step; ge.probes = c->n_probes_last;
      c->graphs_built++;
    }
    CK(cudaGraphLaunch(ge.exec, st));
    c->graph_launches++;
    c->stats.step = ge.launches; c->stats.gemm_step = ge.gemms; c->stats.gemm_flops_step = ge.flops; c->stats.total += ge.launches;
    c->n_probes_last = ge.probes; c->n_last = n; c->tap_n = n; c->tap_S = S;
  }
  if (pose_out && pose_out != c->o_pose) CK(cudaMemcpyAsync(pose_out, c->o_pose, (size_t)n * 6 * 4, cudaMemcpyDeviceToDevice, st));
  if (gripper_out && gripper_out != c->o_grip) CK(cudaMemcpyAsync(gripper_out, c->o_grip, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  if (action_out && action_out != c->o_action) CK(cudaMemcpyAsync(action_out, c->o_action, (size_t)n * 7 * 4, cudaMemcpyDeviceToDevice, st));
  if (exit_out && exit_out != c->o_exit) CK(cudaMemcpyAsync(exit_out, c->o_exit, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  if (delta_out && delta_out != c->o_delta && c->n_thr > 0) CK(cudaMemcpyAsync(delta_out, c->o_delta, (size_t)n * c->n_thr * 4, cudaMemcpyDeviceToDevice, st));
  return DEER_OK;
}

extern "C" int deer_step(deer_ctx* c, int n, const int32_t* slots, const float* rgb, const float* gripper, const int64_t* ids,
                         const uint8_t* mask, int S, const int32_t* timesteps, int static_exit, uint32_t flags, float* pose_out,
                         float* gripper_out, float* action_out, int32_t* exit_out, float* delta_out, void* stream) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  if (!slots || !rgb || !gripper || !ids || !mask) return set_err(&c->err, DEER_E_INVALID, "null input pointer");
  RET(validate_step(c, n, S, static_exit, flags));
  cudaSetDevice(c->device);
  c->stats.step = 0; c->stats.gemm_step = 0; c->stats.gemm_flops_step = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  order_after_last(c, st);
  int r;
  if (c->cfg.use_graph && !c->cfg.host_control && !c->gemm.profile)
    r = step_graph(c, n, slots, rgb, gripper, ids, mask, S, timesteps, static_exit, flags, pose_out, gripper_out, action_out, exit_out, delta_out, st);
  else
    r = step_dispatch(c, n, slots, rgb, gripper, ids, mask, S, timesteps, static_exit, flags, pose_out, gripper_out, action_out, exit_out, delta_out, st);
  mark_done(c, st);
  if (r == DEER_OK) c->steps++;
  return r;
}

extern "C" int deer_step_generate(deer_ctx* c, int n, const int32_t* slots, const float* rgb, const float* gripper, const int64_t* ids,
                                  const uint8_t* mask, int S, const int32_t* commit_layer, float* delta_out, void* stream) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  if (!slots || !rgb || !gripper || !ids || !mask || !commit_layer || !delta_out) return set_err(&c->err, DEER_E_INVALID, "null pointer");
  RET(validate_step(c, n, S, -1, DEER_FLAG_FULL_DEPTH));
  if (c->exits[0] == 0) return set_err(&c->err, DEER_E_INVALID, "generate mode needs 0 not in exit_list (value_net.py:135)");
  cudaSetDevice(c->device);
  if (!c->pend_h) {
    const size_t per = (size_t)c->cfg.max_batch * c->cfg.lstm_layers * c->cfg.lstm_hidden * 4;
    RET(dalloc(c, reinterpret_cast<void**>(&c->pend_h), per, false));
    RET(dalloc(c, reinterpret_cast<void**>(&c->pend_c), per, false));
    RET(dalloc(c, reinterpret_cast<void**>(&c->pend_flag), c->cfg.max_batch, false));
  }
  c->stats.step = 0; c->stats.gemm_step = 0; c->stats.gemm_flops_step = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  order_after_last(c, st);
  c->gen_commit_layer = commit_layer;
  int r = step_dispatch(c, n, slots, rgb, gripper, ids, mask, S, nullptr, -1, DEER_FLAG_FULL_DEPTH | DEER_FLAG_GENERATE, nullptr, nullptr,
                        nullptr, nullptr, delta_out, st);
  c->gen_commit_layer = nullptr;
  mark_done(c, st);
  if (r == DEER_OK) c->steps++;
  return r;
}

extern "C" int deer_step_host(deer_ctx* c, int n, const int32_t* slots, const float* rgb, const float* gripper, const int64_t* ids,
                              const uint8_t* mask, int S, const int32_t* timesteps, int static_exit, uint32_t flags, float* pose_out,
                              float* gripper_out, float* action_out, int32_t* exit_out, float* delta_out) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  if (!slots || !rgb || !gripper || !ids || !mask) return set_err(&c->err, DEER_E_INVALID, "null input pointer");
  RET(validate_step(c, n, S, static_exit, flags));
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= c->cfg.max_slots) return set_err(&c->err, DEER_E_INVALID, "slot %d out of range", slots[i]);
  cudaSetDevice(c->device);
  cudaStream_t st = c->own_stream;
  order_after_last(c, st);        // a previous step (any stream) may still read the staging buffers
  const size_t img = (size_t)3 * c->cfg.image_size * c->cfg.image_size * 4;
  CK(cudaMemcpyAsync(c->s_rgb, rgb, n * img, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->s_grip, gripper, n * img, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->s_ids, ids, (size_t)n * S * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->s_mask, mask, (size_t)n * S, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->s_slots, slots, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  if (timesteps) CK(cudaMemcpyAsync(c->s_ts, timesteps, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  RET(deer_step(c, n, c->s_slots, c->s_rgb, c->s_grip, c->s_ids, c->s_mask, S, timesteps ? c->s_ts : nullptr, static_exit, flags,
                c->o_pose, c->o_grip, c->o_action, c->o_exit, c->o_delta, st));
  if (pose_out) CK(cudaMemcpyAsync(pose_out, c->o_pose, (size_t)n * 6 * 4, cudaMemcpyDeviceToHost, st));
  if (gripper_out) CK(cudaMemcpyAsync(gripper_out, c->o_grip, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (action_out) CK(cudaMemcpyAsync(action_out, c->o_action, (size_t)n * 7 * 4, cudaMemcpyDeviceToHost, st));
  if (exit_out) CK(cudaMemcpyAsync(exit_out, c->o_exit, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (delta_out && c->n_thr > 0) CK(cudaMemcpyAsync(delta_out, c->o_delta, (size_t)n * c->n_thr * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return DEER_OK;
}

extern "C" int deer_get_state(deer_ctx* c, int slot, float* h_out, float* c_out) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  if (slot < 0 || slot >= c->cfg.max_slots) return set_err(&c->err, DEER_E_INVALID, "slot out of range");
  cudaSetDevice(c->device);
  const size_t per = (size_t)c->cfg.lstm_layers * c->cfg.lstm_hidden * 4;
  CK(cudaDeviceSynchronize());
  if (h_out) CK(cudaMemcpy(h_out, reinterpret_cast<char*>(c->h_state) + per * slot, per, cudaMemcpyDeviceToHost));
  if (c_out) CK(cudaMemcpy(c_out, reinterpret_cast<char*>(c->c_state) + per * slot, per, cudaMemcpyDeviceToHost));
  return DEER_OK;
}

extern "C" int deer_set_state(deer_ctx* c, int slot, const float* h, const float* cc) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  if (slot < 0 || slot >= c->cfg.max_slots) return set_err(&c->err, DEER_E_INVALID, "slot out of range");
  if (!h || !cc) return set_err(&c->err, DEER_E_INVALID, "null state pointer");
  cudaSetDevice(c->device);
  const size_t per = (size_t)c->cfg.lstm_layers * c->cfg.lstm_hidden * 4;
  cudaStream_t st = c->has_last ? c->last_stream : c->own_stream;
  CK(cudaMemcpyAsync(reinterpret_cast<char*>(c->h_state) + per * slot, h, per, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(reinterpret_cast<char*>(c->c_state) + per * slot, cc, per, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));                 // the host buffers may be pageable / freed by the caller right away
  mark_done(c, st);
  return DEER_OK;
}

extern "C" int deer_get_features(deer_ctx* c, int n, float* out, int out_is_host, void* stream) {
  if (!c || !out) return set_err(c ? &c->err : nullptr, DEER_E_INVALID, "null argument");
  if (n < 1 || n > c->n_last) return set_err(&c->err, DEER_E_INVALID, "n=%d but the last step had %d episodes", n, c->n_last);
  cudaSetDevice(c->device);
  const size_t bytes = (size_t)n * c->cfg.d_model * 4;
  if (out_is_host) {
    cudaStream_t st = c->has_last ? c->last_stream : c->own_stream;
    CK(cudaMemcpyAsync(out, c->o_feat, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  } else {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    order_after_last(c, st);
    CK(cudaMemcpyAsync(out, c->o_feat, bytes, cudaMemcpyDeviceToDevice, st));
    mark_done(c, st);
  }
  return DEER_OK;
}

extern "C" int deer_get_tap(deer_ctx* c, const char* name, int layer, float* out, int64_t capacity, int64_t* shape_out) {
  if (!c || !name || !out || !shape_out) return set_err(c ? &c->err : nullptr, DEER_E_INVALID, "null argument");
  if (!c->cfg.debug_taps) return set_err(&c->err, DEER_E_STATE, "engine was created without debug_taps");
  cudaSetDevice(c->device);
  const deer_config& f = c->cfg;
  const int64_t n = c->tap_n, S = c->tap_S;
  const float* src = nullptr;
  int64_t sh[4] = {0, 0, 0, 0};
  std::string nm(name);
  if (nm == "vit") { src = c->tap_vit; sh[0] = 2 * n; sh[1] = c->np; sh[2] = f.vit_width; }
  else if (nm == "media") { src = c->tap_media; sh[0] = n; sh[1] = 2 * c->nl; sh[2] = f.vit_width; }
  else if (nm == "hidden") {
    if (layer < 0 || layer >= f.n_layers) return set_err(&c->err, DEER_E_INVALID, "layer out of range");
    src = c->tap_hidden + (size_t)layer * n * S * f.d_model; sh[0] = n; sh[1] = S; sh[2] = f.d_model;
  } else if (nm == "pooled") { src = c->tap_pooled; sh[0] = n; sh[1] = f.d_model; sh[2] = 1; }
  else return set_err(&c->err, DEER_E_INVALID, "unknown tap '%s'", name);
  const int64_t numel = sh[0] * sh[1] * sh[2];
  if (numel > capacity) return set_err(&c->err, DEER_E_INVALID, "tap needs %lld floats, capacity %lld", (long long)numel, (long long)capacity);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, src, numel * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) shape_out[i] = sh[i];
  return DEER_OK;
}

extern "C" int deer_set_launch_overlap(int enable) {
  const int prev = g_launch_overlap ? 1 : 0;
  g_launch_overlap = enable != 0;
  return prev;
}

extern "C" int deer_set_row_order(int snake) {
  const int prev = g_snake ? 1 : 0;
  g_snake = snake != 0;
  return prev;
}

extern "C" int deer_set_profile(deer_ctx* c, int enable) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  c->gemm.profile = enable != 0;
  return DEER_OK;
}

extern "C" int deer_get_profile(deer_ctx* c, deer_profile* out) {
  if (!c || !out) return set_err(c ? &c->err : nullptr, DEER_E_INVALID, "null argument");
  cudaSetDevice(c->device);
  CK(cudaDeviceSynchronize());
  memset(out, 0, sizeof(*out));
  int hist[16] = {0};
  CK(cudaMemcpy(hist, c->nact_hist, sizeof(hist), cudaMemcpyDeviceToHost));
  out->n_probes = c->n_probes_last;
  for (int i = 0; i < 16; ++i) out->active_after[i] = (i < c->n_probes_last) ? hist[i] : 0;
  for (const GemmRec& r : c->gemm.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    int M = r.m_host;
    if (r.dyn) {
      const int act = (r.epoch == 0) ? c->n_last : hist[std::min(r.epoch - 1, 15)];
      M = std::min(M, act * r.m_mul);
    }
    out->gemm_ms[r.tag] += ms;
    out->gemm_flops[r.tag] += 2.0 * M * r.n * r.k;
    // algorithmic HBM bytes: A and W read once, C written once, fp32 residual read once
    out->gemm_bytes[r.tag] += ((double)M * r.k + (double)r.n * r.k) * r.in_es + (double)M * r.n * (r.out_es + (r.has_res ? 4 : 0));
    out->gemm_launches[r.tag]++;
  }
  if (c->ev_step0 && c->gemm.profile) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_step0, c->ev_step1) == cudaSuccess) out->step_ms = ms;
  }
  return DEER_OK;
}

extern "C" int deer_get_stats(deer_ctx* c, deer_stats* out) {
  if (!c || !out) return set_err(c ? &c->err : nullptr, DEER_E_INVALID, "null argument");
  out->kernel_launches_total = c->stats.total;
  out->kernel_launches_last_step = c->stats.step;
  out->gemm_launches_last_step = c->stats.gemm_step;
  out->gemm_flops_last_step = c->stats.gemm_flops_step;
  out->steps = c->steps;
  out->weight_bytes = c->weight_bytes;
  out->workspace_bytes = c->workspace_bytes;
  return DEER_OK;
}

// ---------------------------------------------------------------------------------------------
// single-operator entry points
// ---------------------------------------------------------------------------------------------
static GemmCtx g_op_gemm;
static std::string g_op_err;

extern "C" int deer_op_gemm(int impl, int dtype, const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                            const float* bias, int act, const float* scale, const float* resid, int out_f32, const int32_t* m_dev, int m_mul,
                            void* stream) {
  if (!A || !W || !C) return set_err(nullptr, DEER_E_INVALID, "null pointer");
  g_op_gemm.err = &g_op_err;
  if (dtype == 0) g_op_gemm.alloc_splitk();             // best effort: without the workspace the few-row launches stay one CTA per tile
  static int dbg = getenv("DEER_EPI_DBG") ? atoi(getenv("DEER_EPI_DBG")) : 0;
  GemmEpi e{C, ldc, dtype == 1 ? 1 : out_f32, bias, act, scale, resid, ldc, dbg};
  int r = run_gemm(g_op_gemm, impl, dtype, A, lda, M, W, ldw, M, N, K, e, m_dev, m_mul, reinterpret_cast<cudaStream_t>(stream));
  if (g_op_gemm.tmaps.size() > 4096) g_op_gemm.tmaps.clear();
  return r;
}

extern "C" int deer_op_attention(int impl, int dtype, const void* Q, int ldq, const void* K, const void* V, int ldk, void* O, int ldo, int B, int H,
                                 int Sq, int Sk, int head_dim, float scale, void* stream) {
  if (!Q || !K || !V || !O) return set_err(nullptr, DEER_E_INVALID, "null pointer");
  if ((impl == 0 || impl == 3) && (dtype != 0 || head_dim != 64)) return set_err(nullptr, DEER_E_INVALID, "tensor-core attention needs bf16 and head_dim 64");
  AttnParams a{};
  a.Q = Q; a.K = K; a.V = V; a.O = O; a.ldq = ldq; a.ldk = a.ldv = ldk; a.ldo = ldo; a.Sq = Sq; a.Sk = Sk; a.H = H; a.hd = head_dim;
  a.scale = scale; a.B_host = B;
  return run_attention(impl, dtype, a, &g_op_err, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

// the same with the LLM-side masks: ALiBi key bias, causal, per-episode valid key count, x-attn media classes, dynamic batch
extern "C" int deer_op_attention_masked(int impl, int dtype, const void* Q, int ldq, const void* K, const void* V, int ldk, void* O, int ldo,
                                        int B, int H, int Sq, int Sk, int head_dim, float scale, int causal, const float* alibi_slopes,
                                        const int32_t* kv_len, const uint8_t* qmode, const int32_t* n_dev, void* stream) {
  if (!Q || !K || !V || !O) return set_err(nullptr, DEER_E_INVALID, "null pointer");
  if (impl != 1 && impl != 2 && impl != 4) return set_err(nullptr, DEER_E_INVALID, "masked attention: impl 1 (simt), 2 (auto small) or 4 (mma)");
  if (impl == 4 && (dtype != 0 || Sq > 32 || Sk > 128 || (head_dim != 64 && head_dim != 128)))
    return set_err(nullptr, DEER_E_INVALID, "mma LLM attention needs bf16, Sq<=32, Sk<=128, head_dim 64/128");
  AttnParams a{};
  a.Q = Q; a.K = K; a.V = V; a.O = O; a.ldq = ldq; a.ldk = a.ldv = ldk; a.ldo = ldo; a.Sq = Sq; a.Sk = Sk; a.H = H; a.hd = head_dim;
  a.scale = scale; a.B_host = B; a.causal = causal; a.alibi_slopes = alibi_slopes; a.kv_len = kv_len; a.qmode = qmode; a.n_dev = n_dev;
  return run_attention(impl, dtype, a, &g_op_err, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

// bring-up probe for the tensor-core attention: O = P @ V with P in TMEM (A operand) and V as an MN-major B operand
extern "C" int deer_op_pv_probe(const void* P, const void* V, float* O, int K, int lbo_bytes, int sbo_bytes, int kstep_bytes, void* stream) {
  if (!P || !V || !O || K % 16 || K > 256) return set_err(nullptr, DEER_E_INVALID, "bad pv probe arguments");
  CUtensorMap tm;
  if (!make_tmap(&tm, V, K, 64, 64, (uint32_t)K)) return set_err(nullptr, DEER_E_CUDA, "tensor map for V failed");
  const int smem = 256 * 128 + 1024 + 64;
  cudaFuncSetAttribute(pv_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pv_probe_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm, reinterpret_cast<const bf16*>(P), O, K, (uint32_t)lbo_bytes,
                                                                           (uint32_t)sbo_bytes, (uint32_t)kstep_bytes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(nullptr, DEER_E_CUDA, "pv probe launch failed: %s", cudaGetErrorString(e));
  return DEER_OK;
}

// tuning aid: enable / read per-CTA cycle counters of the attention kernel's softmax warp (5 phases x up to 512 CTAs)
extern "C" int deer_attn_debug(int enable, long long* host_out, int n) {
  if (enable && !g_attn_dbg) { cudaMalloc(reinterpret_cast<void**>(&g_attn_dbg), 512 * 8 * sizeof(long long)); cudaMemset(g_attn_dbg, 0, 512 * 8 * sizeof(long long)); }
  if (host_out && g_attn_dbg) { cudaDeviceSynchronize(); cudaMemcpy(host_out, g_attn_dbg, std::min(n, 512 * 8) * sizeof(long long), cudaMemcpyDeviceToHost); }
  if (!enable && g_attn_dbg) { cudaFree(g_attn_dbg); g_attn_dbg = nullptr; }
  return DEER_OK;
}

// tuning aid: %globaltimer (ns) of CTA 0 at the phase boundaries of the LAST fused head probe: [0] entry, then per grid barrier
// k = 1..layers+3 the pair ([2k-1] arrival, [2k] release), [2*(layers+3)+1] exit
extern "C" int deer_head_debug(int enable, long long* host_out, int n) {
  if (enable && !g_head_dbg) { cudaMalloc(reinterpret_cast<void**>(&g_head_dbg), 64 * sizeof(long long)); cudaMemset(g_head_dbg, 0, 64 * sizeof(long long)); }
  if (host_out && g_head_dbg) { cudaDeviceSynchronize(); cudaMemcpy(host_out, g_head_dbg, std::min(n, 64) * sizeof(long long), cudaMemcpyDeviceToHost); }
  if (!enable && g_head_dbg) { cudaFree(g_head_dbg); g_head_dbg = nullptr; }
  return DEER_OK;
}

#ifdef DEER_GEMM_PROF
// instrumented build only: per-k-block clock64 of CTA 0 (first tile) of the last gemm_tc_kernel launch
extern "C" int deer_gemm_debug(long long* host_out, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_gemm_prof, std::min(n, 4 * 256) * sizeof(long long)) == cudaSuccess ? DEER_OK : DEER_E_CUDA;
}
#endif

// ---------------------------------------------------------------------------------------------
// on-device image transform
// ---------------------------------------------------------------------------------------------
static double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
// Pillow src/libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc (box = the whole axis)
static int get_axis(deer_ctx* c, int in_size, int out_size, ResizeAxis* out) {
  auto key = std::make_pair(in_size, out_size);
  auto it = c->axis_tabs.find(key);
  if (it == c->axis_tabs.end()) {
    deer_ctx::AxisTab t;
    t.need = (in_size != out_size);
    if (t.need) {
      const double scale = (double)in_size / out_size;
      const double filterscale = scale < 1.0 ? 1.0 : scale;
      const double support = 2.0 * filterscale;
      t.ksize = (int)std::ceil(support) * 2 + 1;
      std::vector<int> bounds(2 * out_size), kk((size_t)out_size * t.ksize, 0);
      std::vector<double> w(t.ksize);
      const double ss = 1.0 / filterscale;
      for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) { w[x] = pil_bicubic((x + xmin - center + 0.5) * ss); ww += w[x]; }
        for (int x = 0; x < xmax; ++x) {
          const double v = (ww != 0.0) ? w[x] / ww : w[x];
          kk[(size_t)xx * t.ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
        }
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
      }
      RET(dalloc(c, reinterpret_cast<void**>(&t.bounds), bounds.size() * 4, false));
      RET(dalloc(c, reinterpret_cast<void**>(&t.kk), kk.size() * 4, false));
      CK(cudaMemcpy(t.bounds, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(t.kk, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice));
    }
    it = c->axis_tabs.emplace(key, t).first;
  }
  out->bounds = it->second.bounds; out->kk = it->second.kk; out->ksize = it->second.ksize; out->need = it->second.need;
  return DEER_OK;
}

extern "C" int deer_preprocess_u8(deer_ctx* c, const uint8_t* frames, int n, int H, int W, float* out, void* stream) {
  if (!c || !frames || !out) return set_err(c ? &c->err : nullptr, DEER_E_INVALID, "null argument");
  if (n < 1 || H < 1 || W < 1 || H > 8192 || W > 8192) return set_err(&c->err, DEER_E_INVALID, "bad image dims");
  cudaSetDevice(c->device);
  const int size = c->cfg.image_size;
  // torchvision Resize(size:int): smaller edge -> size, other edge -> int(size * long / short); then CenterCrop(size)
  int nh, nw;
  if (H <= W) { nh = size; nw = (int)((double)size * W / H); } else { nw = size; nh = (int)((double)size * H / W); }
  const int top = (int)std::nearbyint((nh - size) / 2.0), left = (int)std::nearbyint((nw - size) / 2.0);
  ResizeAxis ax, ay;
  RET(get_axis(c, W, nw, &ax));
  RET(get_axis(c, H, nh, &ay));
  dim3 grid((size * size + 255) / 256, n);
  preprocess_u8_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(frames, n, H, W, ax, ay, top, left, size, out);
  KLAUNCH_CHECK("preprocess_u8");
  return DEER_OK;
}

extern "C" int deer_step_host_u8(deer_ctx* c, int n, const int32_t* slots, const uint8_t* rgb, int Hs, int Ws, const uint8_t* gripper,
                                 int Hg, int Wg, const int64_t* ids, const uint8_t* mask, int S, const int32_t* timesteps,
                                 int static_exit, uint32_t flags, float* pose_out, float* gripper_out, float* action_out,
                                 int32_t* exit_out, float* delta_out) {
  if (!c) return set_err(nullptr, DEER_E_INVALID, "null ctx");
  if (!slots || !rgb || !gripper || !ids || !mask) return set_err(&c->err, DEER_E_INVALID, "null input pointer");
  RET(validate_step(c, n, S, static_exit, flags));
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= c->cfg.max_slots) return set_err(&c->err, DEER_E_INVALID, "slot %d out of range", slots[i]);
  cudaSetDevice(c->device);
  cudaStream_t st = c->own_stream;
  const size_t bytes[2] = {(size_t)n * Hs * Ws * 3, (size_t)n * Hg * Wg * 3};
  for (int k = 0; k < 2; ++k)
    if (c->u8_cap[k] < bytes[k]) {
      if (c->u8_stage[k]) cudaFree(c->u8_stage[k]);
      const size_t cap = std::max(bytes[k], (size_t)c->cfg.max_batch * (k == 0 ? Hs * Ws : Hg * Wg) * 3);
      CK(cudaMalloc(reinterpret_cast<void**>(&c->u8_stage[k]), cap));
      c->u8_cap[k] = cap;
    }
  order_after_last(c, st);        // a previous step (any stream) may still read the staging buffers
  CK(cudaMemcpyAsync(c->u8_stage[0], rgb, bytes[0], cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->u8_stage[1], gripper, bytes[1], cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->s_ids, ids, (size_t)n * S * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->s_mask, mask, (size_t)n * S, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->s_slots, slots, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  if (timesteps) CK(cudaMemcpyAsync(c->s_ts, timesteps, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  RET(deer_preprocess_u8(c, c->u8_stage[0], n, Hs, Ws, c->s_rgb, st));
  RET(deer_preprocess_u8(c, c->u8_stage[1], n, Hg, Wg, c->s_grip, st));
  RET(deer_step(c, n, c->s_slots, c->s_rgb, c->s_grip, c->s_ids, c->s_mask, S, timesteps ? c->s_ts : nullptr, static_exit, flags,
                c->o_pose, c->o_grip, c->o_action, c->o_exit, c->o_delta, st));
  if (pose_out) CK(cudaMemcpyAsync(pose_out, c->o_pose, (size_t)n * 6 * 4, cudaMemcpyDeviceToHost, st));
  if (gripper_out) CK(cudaMemcpyAsync(gripper_out, c->o_grip, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (action_out) CK(cudaMemcpyAsync(action_out, c->o_action, (size_t)n * 7 * 4, cudaMemcpyDeviceToHost, st));
  if (exit_out) CK(cudaMemcpyAsync(exit_out, c->o_exit, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (delta_out && c->n_thr > 0) CK(cudaMemcpyAsync(delta_out, c->o_delta, (size_t)n * c->n_thr * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return DEER_OK;
}
