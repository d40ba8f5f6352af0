// Bandwidth-bound kernels of the DeeR step: patch gather, token assembly, LayerNorm, embedding,
// max-pool, LSTM cell, head output + exit gate, active-set compaction.
// Residual streams are fp32; T (bf16 on the product path, float on the validation path) is the
// GEMM operand type.  Row counts that depend on the number of still-active episodes are read
// from device memory (n_dev) so that no host synchronisation is needed inside a step.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// block reductions (blockDim.x <= 1024, multiple of 32)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}

// ---------------------------------------------------------------------------------------------
// im2col for the 14x14/stride-14 patch-embed conv (open_clip VisionTransformer.conv1, bias-free):
// out[(img*np + p), k] with k = c*P*P + ky*P + kx  (== conv weight [out,c,ky,kx] flattened), zero
// padded to Kp.  Image order: all rgb frames, then all gripper frames.
// ---------------------------------------------------------------------------------------------
// PC > 0: patch size known at compile time (14 for ViT-L/14: every division below becomes a multiply-shift); PC = 0: runtime P.
template <typename T, int PC>
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ rgb, const float* __restrict__ grip, T* __restrict__ out,
              int n, int H, int P_rt, int Kp) {
  const int P = PC > 0 ? PC : P_rt;
  const int g = H / P, np = g * g;
  const int K = 3 * P * P;
  // 256 threads walk k (fastest) of two operand rows per CTA pass: consecutive threads read consecutive pixels of a patch line
  // and write consecutive operand elements
  const int rows = 2 * n * np;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int img = row / np, p = row % np;
    const float* src = (img < n ? rgb + static_cast<size_t>(img) * 3 * H * H : grip + static_cast<size_t>(img - n) * 3 * H * H);
    const int py = p / g, px = p % g;
    const float* base = src + static_cast<size_t>(py * P) * H + px * P;
    T* orow = out + static_cast<size_t>(row) * Kp;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
      float v = 0.0f;
      if (k < K) {
        const int c = k / (P * P), rem = k - c * (P * P), ky = rem / P, kx = rem - ky * P;
        v = __ldg(base + (static_cast<size_t>(c) * H + ky) * H + kx);
      }
      orow[k] = from_f32<T>(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm family.  One CTA (128 threads) per row, row kept in registers (D <= 4096), two-pass
// statistics in fp32, eps 1e-5.
//   in row r at  in + r*in_ld ;  gamma/beta indexed by (r % groups)*D + j
//   out row index = (r / gin) * gout + goff + (r % gin)   (gin=0: identity)
// ---------------------------------------------------------------------------------------------
struct LnMap { int gin, gout, goff; };
__device__ __forceinline__ size_t ln_out_row(int r, const LnMap& m) {
  return m.gin ? static_cast<size_t>(r / m.gin) * m.gout + m.goff + (r % m.gin) : static_cast<size_t>(r);
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(128)
layernorm_kernel(const TIn* in, int in_ld, LnMap in_map, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int groups, int D, TOut* out, int out_ld, LnMap map, TOut* out2, int out2_ld,
                 int relu, int rows_host, const int* __restrict__ n_dev, int n_mul) {
  __shared__ float red[32];
  const int rows = dyn_rows(rows_host, n_dev, n_mul);
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const TIn* x = in + ln_out_row(r, in_map) * in_ld;   // (in == out is allowed: the row is held in registers)
    float v[32];
    float s = 0.0f;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
      const int j = threadIdx.x + it * 128;
      v[it] = (j < D) ? to_f32<TIn>(x[j]) : 0.0f;
      s += v[it];
    }
    const float mean = block_sum(s, red) / D;
    float q = 0.0f;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
      const int j = threadIdx.x + it * 128;
      const float d = (j < D) ? v[it] - mean : 0.0f;
      q += d * d;
    }
    const float rstd = rsqrtf(block_sum(q, red) / D + 1e-5f);
    const float* g = gamma + static_cast<size_t>(groups > 1 ? (r % groups) : 0) * D;
    const float* bt = beta ? beta + static_cast<size_t>(groups > 1 ? (r % groups) : 0) * D : nullptr;
    TOut* o = out + ln_out_row(r, map) * out_ld;
    TOut* o2 = out2 ? out2 + static_cast<size_t>(r) * out2_ld : nullptr;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
      const int j = threadIdx.x + it * 128;
      if (j < D) {
        float y = (v[it] - mean) * rstd * g[j] + (bt ? bt[j] : 0.0f);
        if (relu) y = fmaxf(y, 0.0f);
        o[j] = from_f32<TOut>(y);
        if (o2) o2[j] = from_f32<TOut>(y);
      }
    }
    __syncthreads();
  }
}

// Fast path: one WARP per row, D = NVEC*128, row held in registers as float4s, no block barriers.
// Same two-pass fp32 statistics and the same row maps as layernorm_kernel (in == out allowed).
template <typename T> struct Vec4IO;
template <> struct Vec4IO<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4IO<bf16> {
  static __device__ __forceinline__ float4 ld(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  static __device__ __forceinline__ void st(bf16* p, float4 v) {
    uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

template <typename TIn, typename TOut, int NVEC>
__global__ void __launch_bounds__(128)
layernorm_warp_kernel(const TIn* in, int in_ld, LnMap in_map, const float* __restrict__ gamma, const float* __restrict__ beta,
                      int groups, TOut* out, int out_ld, LnMap map, TOut* out2, int out2_ld,
                      int relu, int rows_host, const int* __restrict__ n_dev, int n_mul, int rev) {
  constexpr int D = NVEC * 128;
  pdl_wait();
  pdl_trigger();
  const int rows = dyn_rows(rows_host, n_dev, n_mul);
  const int lane = threadIdx.x & 31;
  int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  if (rev) r = rows - 1 - r;          // last rows first: they are what the producing GEMM left in L2 ("snake" order)
  const TIn* x = in + ln_out_row(r, in_map) * in_ld;
  float4 v[NVEC];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    v[i] = Vec4IO<TIn>::ld(x + (lane + 32 * i) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
  const size_t goff = static_cast<size_t>(groups > 1 ? (r % groups) : 0) * D;
  TOut* o = out + ln_out_row(r, map) * out_ld;
  TOut* o2 = out2 ? out2 + static_cast<size_t>(r) * out2_ld : nullptr;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int j = (lane + 32 * i) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + goff + j);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta) b = *reinterpret_cast<const float4*>(beta + goff + j);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x; y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z; y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
    Vec4IO<TOut>::st(o + j, y);
    if (o2) Vec4IO<TOut>::st(o2 + j, y);
  }
}

// ViT token assembly + ln_pre: x[img, 0] = cls + pos[0]; x[img, 1+p] = patch[img*np+p] + pos[1+p]; x = ln_pre(x)
// One WARP per token row (4 rows per CTA), the row held in registers as float4s (D = NVEC * 128), two-pass fp32 statistics.
template <typename T, int NVEC>
__global__ void __launch_bounds__(128)
vit_assemble_warp_kernel(const T* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ x, int n_img, int L) {
  constexpr int D = NVEC * 128;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n_img * L) return;
  const int img = r / L, tok = r % L;
  float4 v[NVEC];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int j = (lane + 32 * i) * 4;
    float4 t = (tok == 0) ? *reinterpret_cast<const float4*>(cls + j)
                          : Vec4IO<T>::ld(patch + (static_cast<size_t>(img) * (L - 1) + (tok - 1)) * D + j);
    const float4 pe = *reinterpret_cast<const float4*>(pos + static_cast<size_t>(tok) * D + j);
    t.x += pe.x; t.y += pe.y; t.z += pe.z; t.w += pe.w;
    v[i] = t;
    s += (t.x + t.y) + (t.z + t.w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int j = (lane + 32 * i) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + j), b = *reinterpret_cast<const float4*>(beta + j);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x; y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z; y.w = (v[i].w - mean) * rstd * g.w + b.w;
    *reinterpret_cast<float4*>(x + static_cast<size_t>(r) * D + j) = y;
  }
}
// general widths: one CTA per token row
template <typename T>
__global__ void __launch_bounds__(128)
vit_assemble_kernel(const T* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ x,
                    int n_img, int L, int D) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int img = r / L, tok = r % L;
  float v[8];
  float s = 0.0f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int j = threadIdx.x + it * 128;
    float t = 0.0f;
    if (j < D) {
      t = (tok == 0) ? cls[j] : to_f32<T>(patch[(static_cast<size_t>(img) * (L - 1) + (tok - 1)) * D + j]);
      t += pos[static_cast<size_t>(tok) * D + j];
    }
    v[it] = t;
    s += t;
  }
  const float mean = block_sum(s, red) / D;
  float q = 0.0f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int j = threadIdx.x + it * 128;
    const float d = (j < D) ? v[it] - mean : 0.0f;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum(q, red) / D + 1e-5f);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int j = threadIdx.x + it * 128;
    if (j < D) x[static_cast<size_t>(r) * D + j] = (v[it] - mean) * rstd * gamma[j] + beta[j];
  }
}

// drop CLS: tokens[img, p, :] (T) = x[img, 1+p, :]   (perceiver media input; open_clip tokens = x[:,1:])
__global__ void vit_tokens_kernel(const float* __restrict__ x, float* __restrict__ tok_f32, int n_img, int L, int D) {
  const size_t total = static_cast<size_t>(n_img) * (L - 1) * D;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int j = i % D;
    const size_t rp = i / D;
    const int p = rp % (L - 1);
    const size_t img = rp / (L - 1);
    tok_f32[i] = x[(img * L + 1 + p) * D + j];
  }
}

// latents[img, :, :] = perceiver.latents (helpers.py:128)
__global__ void broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t per, size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i % per];
}

// ---------------------------------------------------------------------------------------------
// per-step episode bookkeeping
// ---------------------------------------------------------------------------------------------
struct EpisodeMeta {          // all arrays sized max_batch; "r" = rank inside the current active set
  int* orig;                  // r -> index in this call's batch
  int* slot;                  // r -> LSTM state slot
  int* valid_len;             // r -> number of real tokens
  uint8_t* xmode;             // [r, S] x-attn class per token (0 no media yet, 1 attend, 2 more than one <image>)
  float* prev_pose;           // [r, 6] previous exit's pose
  float* prev_grip;           // [r] previous exit's gripper output (action ensemble, value_net.py:92-95)
  uint8_t* use_prev;          // r -> 1 when timestep % steps_per_stage != 0 (reuse cur_exit_id, value_net.py:285)
};

__global__ void init_episodes_kernel(EpisodeMeta m, int* n_active, const int* __restrict__ slots, const int64_t* __restrict__ ids,
                                     const uint8_t* __restrict__ mask, const int* __restrict__ timesteps, int n, int S,
                                     int media_id, int sps, int n_thr, int* __restrict__ exit_dev, float* __restrict__ delta_dev,
                                     int max_slots) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) *n_active = n;
  if (r >= n) return;
  m.orig[r] = r;
  const int sl = slots[r];                     // device-resident slot ids cannot be validated on the host: never index out of bounds
  m.slot[r] = sl < 0 ? 0 : (sl >= max_slots ? max_slots - 1 : sl);
  int vl = 0, tt = 0;
  for (int s = 0; s < S; ++s) {
    vl += mask[r * S + s] ? 1 : 0;
    tt += (ids[r * S + s] == media_id) ? 1 : 0;                  // text_time = cumsum(media_locations), helpers.py:208
    m.xmode[r * S + s] = static_cast<uint8_t>(tt > 2 ? 2 : tt);
  }
  m.valid_len[r] = vl;
  m.use_prev[r] = (timesteps != nullptr && sps > 1 && (timesteps[r] % sps) != 0) ? 1 : 0;
  exit_dev[r] = -1;
  for (int k = 0; k < n_thr; ++k) delta_dev[r * n_thr + k] = __int_as_float(0x7fc00000);
}

// x[r*S+s, :] = wte[ids[r,s], :]   (mosaic_gpt_3b.py:341-343; ALiBi => no position embedding)
template <typename T>
__global__ void embed_kernel(const int64_t* __restrict__ ids, const T* __restrict__ wte, float* __restrict__ x, int rows, int D, int vocab) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  long long id = ids[r];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const T* w = wte + static_cast<size_t>(id) * D;
  for (int j = threadIdx.x; j < D; j += blockDim.x) x[static_cast<size_t>(r) * D + j] = to_f32<T>(w[j]);
}

// AdaptiveMaxPool1d(1) over the (valid) token positions (action_head.py:519-520) -> T row of the LSTM input buffer
template <typename T>
__global__ void maxpool_kernel(const float* __restrict__ x, const int* __restrict__ valid_len, T* __restrict__ out, int out_ld,
                               float* __restrict__ out_f32, int S, int D, int n_host, const int* __restrict__ n_dev) {
  pdl_wait();
  pdl_trigger();
  const int n = dyn_rows(n_host, n_dev, 1);
  const int r = blockIdx.y;
  if (r >= n) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  const int vl = valid_len[r];
  float m = -INFINITY;
  for (int s = 0; s < vl; ++s) m = fmaxf(m, x[(static_cast<size_t>(r) * S + s) * D + j]);
  out[static_cast<size_t>(r) * out_ld + j] = from_f32<T>(m);
  if (out_f32) out_f32[static_cast<size_t>(r) * D + j] = m;
}

// ---------------------------------------------------------------------------------------------
// LSTM (LayerNormLSTM, action_head.py:15-64): per layer  gates = [x | h_prev] @ [W_ih | W_hh]^T + (b_ih+b_hh)
// ---------------------------------------------------------------------------------------------
// copy the carried h of every layer into the GEMM A operands: A_l[r, in_l + j] = h_state[slot[r], l, j]
template <typename T>
__global__ void lstm_gather_h_kernel(const float* __restrict__ h_state, const int* __restrict__ slot, T* __restrict__ a0, int ld0, int off0,
                                     T* __restrict__ a_rest, int ld_rest, int layers, int Hh, int nstride, int n_host, const int* __restrict__ n_dev) {
  pdl_wait();
  pdl_trigger();
  const int n = dyn_rows(n_host, n_dev, 1);
  const int r = blockIdx.x, l = blockIdx.y;
  if (r >= n) return;
  const float* h = h_state + (static_cast<size_t>(slot[r]) * layers + l) * Hh;
  T* dst = (l == 0) ? a0 + static_cast<size_t>(r) * ld0 + off0
                    : a_rest + (static_cast<size_t>(l - 1) * nstride + r) * ld_rest + Hh;
  for (int j = threadIdx.x; j < Hh; j += blockDim.x) dst[j] = from_f32<T>(h[j]);
}

// LSTM cell pointwise (gate order i,f,g,o) + LayerNorm(h') -> next layer's x (T).  One CTA (256 thr) per episode.
template <typename T>
__global__ void __launch_bounds__(256)
lstm_cell_kernel(const float* __restrict__ gates, const float* __restrict__ c_state, const int* __restrict__ slot, int layer, int layers,
                 float* __restrict__ cand_h, float* __restrict__ cand_c, const float* __restrict__ gamma, const float* __restrict__ beta,
                 T* __restrict__ y, int y_ld, int Hh, int nstride, int n_host, const int* __restrict__ n_dev) {
  __shared__ float red[32];
  pdl_wait();
  pdl_trigger();
  const int n = dyn_rows(n_host, n_dev, 1);
  const int r = blockIdx.x;
  if (r >= n) return;
  const float* g = gates + static_cast<size_t>(r) * 4 * Hh;
  const float* cp = c_state + (static_cast<size_t>(slot[r]) * layers + layer) * Hh;
  float hv[8];
  float s = 0.0f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int j = threadIdx.x + it * 256;
    float h2 = 0.0f;
    if (j < Hh) {
      const float ig = sigmoidf_(g[j]), fg = sigmoidf_(g[Hh + j]), gg = tanhf(g[2 * Hh + j]), og = sigmoidf_(g[3 * Hh + j]);
      const float c2 = fg * cp[j] + ig * gg;
      h2 = og * tanhf(c2);
      cand_c[(static_cast<size_t>(layer) * nstride + r) * Hh + j] = c2;
      cand_h[(static_cast<size_t>(layer) * nstride + r) * Hh + j] = h2;
    }
    hv[it] = h2;
    s += h2;
  }
  if (gamma == nullptr) {                      // plain stacked nn.LSTM (lstm_layernorm=False): the next layer sees h itself
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int j = threadIdx.x + it * 256;
      if (j < Hh) y[static_cast<size_t>(r) * y_ld + j] = from_f32<T>(hv[it]);
    }
    return;
  }
  const float mean = block_sum(s, red) / Hh;
  float q = 0.0f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int j = threadIdx.x + it * 256;
    const float d = (j < Hh) ? hv[it] - mean : 0.0f;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum(q, red) / Hh + 1e-5f);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int j = threadIdx.x + it * 256;
    if (j < Hh) y[static_cast<size_t>(r) * y_ld + j] = from_f32<T>((hv[it] - mean) * rstd * gamma[j] + beta[j]);
  }
}

// final Linear(+tanh / sigmoid) of both MLP heads (action_head.py:114-116, 215-217).
// y2 [n, 2*d1] fp32 (pose half | gripper half).  One CTA of 256 threads per episode: warp k<6 -> pose_k, warp 6 -> gripper.
__global__ void __launch_bounds__(256)
head_out_kernel(const float* __restrict__ y2, int d1, const float* __restrict__ w_pose, const float* __restrict__ b_pose,
                const float* __restrict__ w_grip, const float* __restrict__ b_grip, float* __restrict__ pose, float* __restrict__ grip,
                int n_host, const int* __restrict__ n_dev) {
  pdl_wait();
  pdl_trigger();
  const int n = dyn_rows(n_host, n_dev, 1);
  const int r = blockIdx.x;
  if (r >= n) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp > 6) return;
  const float* x = y2 + static_cast<size_t>(r) * 2 * d1 + (warp == 6 ? d1 : 0);
  const float* w = (warp == 6) ? w_grip : w_pose + static_cast<size_t>(warp) * d1;
  float acc = 0.0f;
  for (int j = lane; j < d1; j += 32) acc += x[j] * w[j];
  acc = warp_sum(acc);
  if (lane == 0) {
    if (warp == 6) grip[r] = sigmoidf_(acc + b_grip[0]);
    else pose[r * 6 + warp] = tanhf(acc + b_pose[warp]);
  }
}

// ---------------------------------------------------------------------------------------------
// the exit gate (ExitController.forward value_net.py:277-297 + ActionValueNet get_delta 105-117) for every
// active episode at once, followed by the survivor scan.  Single CTA of 1024 threads (max_batch <= 1024).
// ---------------------------------------------------------------------------------------------
struct GateParams {
  int layer;                 // LLM layer index of this exit
  int exit_idx;              // index into thresholds / delta columns (-1: no threshold for this exit)
  int forced;                // layer >= ctrl_max_layer (value_net.py:293) or static exit
  int full_depth;            // calibration: never exit before `forced`
  int has_prev;              // 0: static exit mode, no delta
  int threshold_type;
  int n_thr;
  const float* thresholds;   // device [n_thr]
  const float* cur_pose;     // [r,6]
  const float* cur_grip;     // [r]
  int* cur_exit_id;          // [slots] value_net.py:294
  const int* commit_layer;   // generate mode: [orig] exit layer whose candidate LSTM state becomes the carried one (or nullptr)
  uint8_t* pend_flag;        // generate mode: [orig] set when that candidate has been captured
  // outputs indexed by orig
  float* pose_out; float* grip_out; float* action_out; int* exit_out; float* delta_out;
  // compaction
  uint8_t* exit_flag;        // [r] this episode leaves now
  int* src_of_new;           // [r'] source rank of survivor r'
  // CUDA-graph control flow: when has_cond, the conditional IF node that wraps every later layer runs iff survivors remain
  cudaGraphConditionalHandle cond;
  int has_cond;
  // action ensemble (eval_utils.py:457-462): action_out = mean of the last two entries of this step's action_list
  int ensemble;
  int probe_idx;             // probes evaluated before this one in the step (= len(action_list) before the append)
};

__global__ void __launch_bounds__(1024)
exit_gate_kernel(GateParams g, EpisodeMeta m, const int* __restrict__ n_cur, int* __restrict__ n_next, int* __restrict__ hist) {
  __shared__ int wsum[32];
  __shared__ int total;
  const int n = *n_cur;
  const int r = threadIdx.x;
  int survive = 0;
  if (r < n) {
    const float* cp = g.cur_pose + r * 6;
    float delta = 0.0f;
    if (g.has_prev) {
      const float* pp = m.prev_pose + r * 6;
      if (g.threshold_type == 3) {                                  // cosine (value_net.py:13-21,114)
        float n1 = 0.f, n2 = 0.f, dp = 0.f;
        for (int k = 0; k < 6; ++k) { n1 += cp[k] * cp[k]; n2 += pp[k] * pp[k]; dp += cp[k] * pp[k]; }
        delta = 1.0f - dp / (fmaxf(sqrtf(n1), 1e-5f) * fmaxf(sqrtf(n2), 1e-5f));
      } else {
        float acc = 0.0f;
        for (int k = 0; k < 6; ++k) {
          const float d = fabsf(cp[k] - pp[k]);
          if (g.threshold_type == 0) acc += d * d;
          else if (g.threshold_type == 1) acc += d;
          else acc = fmaxf(acc, d);
        }
        delta = (g.threshold_type == 0) ? sqrtf(acc / 6.0f) : (g.threshold_type == 1 ? acc / 6.0f : acc);
      }
    }
    const int o = m.orig[r];
    const int sl = m.slot[r];
    bool ex;
    if (g.full_depth) {
      ex = g.forced != 0;
      if (g.exit_idx >= 0 && g.delta_out) g.delta_out[o * g.n_thr + g.exit_idx] = delta;
    } else if (m.use_prev[r]) {
      ex = g.forced || g.layer >= g.cur_exit_id[sl];                // value_net.py:285-286
    } else {
      if (g.exit_idx >= 0 && g.has_prev && g.delta_out) g.delta_out[o * g.n_thr + g.exit_idx] = delta;
      const bool leq = (g.exit_idx >= 0 && g.has_prev) ? (delta <= g.thresholds[g.exit_idx]) : false;
      ex = leq || g.forced;
      if (ex) g.cur_exit_id[sl] = g.layer;                          // value_net.py:294
    }
    if (ex) {
      const float gr = g.cur_grip[r];
      if (g.pose_out) for (int k = 0; k < 6; ++k) g.pose_out[o * 6 + k] = cp[k];
      if (g.grip_out) g.grip_out[o] = gr;
      if (g.action_out) {                                           // eval_utils.py:458,464
        // ensemble: torch.stack(action_list[-2:]).mean(0) -- one entry only when this is the step's first probe
        const bool ens = g.ensemble && g.probe_idx > 0 && g.has_prev && !m.use_prev[r];
        const float ga = ens ? (gr + m.prev_grip[r]) / 2.0f : gr;
        for (int k = 0; k < 6; ++k) g.action_out[o * 7 + k] = ens ? (cp[k] + m.prev_pose[r * 6 + k]) / 2.0f : cp[k];
        g.action_out[o * 7 + 6] = (ga > 0.5f) ? 1.0f : -1.0f;
      }
      if (g.exit_out) g.exit_out[o] = g.layer;
    }
    if (g.commit_layer) {                                           // flamingo_mpt.py:484-492: a drawn exit's features feed the history
      const int hit = (g.commit_layer[o] == g.layer) ? 1 : 0;
      g.exit_flag[r] = hit;
      if (hit) g.pend_flag[o] = 1;
    } else {
      g.exit_flag[r] = ex ? 1 : 0;
    }
    survive = ex ? 0 : 1;
  }
  // exclusive scan of `survive` over the CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = survive;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    wsum[lane] = wi - w;
    if (lane == 31) total = wi;
  }
  __syncthreads();
  const int pos = wsum[warp] + incl - survive;
  if (survive) g.src_of_new[pos] = r;
  if (threadIdx.x == 0) {
    *n_next = total;
    *hist = total;
    if (g.has_cond) cudaGraphSetConditional(g.cond, total > 0 ? 1u : 0u);
  }
}

// commit the candidate LSTM state of the episodes that exit now (update_hidden_state=True, action_head.py:554-556)
__global__ void commit_state_kernel(const uint8_t* __restrict__ exit_flag, const int* __restrict__ slot, const float* __restrict__ cand_h,
                                    const float* __restrict__ cand_c, float* __restrict__ h_state, float* __restrict__ c_state,
                                    int layers, int Hh, int nstride, int n_host, const int* __restrict__ n_dev,
                                    const int* __restrict__ orig, const float* __restrict__ pooled, float* __restrict__ feat_out, int D) {
  const int n = dyn_rows(n_host, n_dev, 1);
  const int r = blockIdx.x, l = blockIdx.y;
  if (r >= n || !exit_flag[r]) return;
  const size_t src = (static_cast<size_t>(l) * nstride + r) * Hh;
  const size_t dst = (static_cast<size_t>(slot[r]) * layers + l) * Hh;
  for (int j = threadIdx.x; j < Hh; j += blockDim.x) {
    h_state[dst + j] = cand_h[src + j];
    c_state[dst + j] = cand_c[src + j];
  }
  if (l == 0 && feat_out != nullptr) {         // `.logits[2]`: the pooled feature of the exit layer (action_head.py:608-609)
    const float* ps = pooled + static_cast<size_t>(r) * D;
    float* fo = feat_out + static_cast<size_t>(orig[r]) * D;
    for (int j = threadIdx.x; j < D; j += blockDim.x) fo[j] = ps[j];
  }
}

// generate mode: candidates captured at the drawn exit are parked per row (state layout) and committed after the last probe,
// so that every probe of the frame starts from the same carried state
__global__ void park_state_kernel(const uint8_t* __restrict__ exit_flag, const int* __restrict__ orig, const float* __restrict__ cand_h,
                                  const float* __restrict__ cand_c, float* __restrict__ pend_h, float* __restrict__ pend_c,
                                  int layers, int Hh, int nstride, int n) {
  const int r = blockIdx.x, l = blockIdx.y;
  if (r >= n || !exit_flag[r]) return;
  const size_t src = (static_cast<size_t>(l) * nstride + r) * Hh;
  const size_t dst = (static_cast<size_t>(orig[r]) * layers + l) * Hh;
  for (int j = threadIdx.x; j < Hh; j += blockDim.x) { pend_h[dst + j] = cand_h[src + j]; pend_c[dst + j] = cand_c[src + j]; }
}
__global__ void commit_pending_kernel(const uint8_t* __restrict__ pend_flag, const int* __restrict__ slots, const float* __restrict__ pend_h,
                                      const float* __restrict__ pend_c, float* __restrict__ h_state, float* __restrict__ c_state,
                                      int layers, int Hh, int n) {
  const int o = blockIdx.x, l = blockIdx.y;
  if (o >= n || !pend_flag[o]) return;
  const size_t src = (static_cast<size_t>(o) * layers + l) * Hh;
  const size_t dst = (static_cast<size_t>(slots[o]) * layers + l) * Hh;
  for (int j = threadIdx.x; j < Hh; j += blockDim.x) { h_state[dst + j] = pend_h[src + j]; c_state[dst + j] = pend_c[src + j]; }
}

// gather survivors into the other half of the ping-pong buffers
template <typename T>
__global__ void compact_kernel(const int* __restrict__ src_of_new, const float* __restrict__ x_in, float* __restrict__ x_out, size_t x_per,
                               const T* __restrict__ media_in, T* __restrict__ media_out, size_t media_per,
                               EpisodeMeta mi, EpisodeMeta mo, const float* __restrict__ cur_pose, const float* __restrict__ cur_grip, int S,
                               int n_host, const int* __restrict__ n_next) {
  const int n = dyn_rows(n_host, n_next, 1);
  const int r2 = blockIdx.x;
  if (r2 >= n) return;
  const int r = src_of_new[r2];
  const float4* xi = reinterpret_cast<const float4*>(x_in + static_cast<size_t>(r) * x_per);
  float4* xo = reinterpret_cast<float4*>(x_out + static_cast<size_t>(r2) * x_per);
  for (size_t i = blockIdx.y * static_cast<size_t>(blockDim.x) + threadIdx.x; i < x_per / 4; i += static_cast<size_t>(gridDim.y) * blockDim.x) xo[i] = xi[i];
  const uint4* mi4 = reinterpret_cast<const uint4*>(media_in + static_cast<size_t>(r) * media_per);
  uint4* mo4 = reinterpret_cast<uint4*>(media_out + static_cast<size_t>(r2) * media_per);
  const size_t mvec = media_per * sizeof(T) / 16;
  for (size_t i = blockIdx.y * static_cast<size_t>(blockDim.x) + threadIdx.x; i < mvec; i += static_cast<size_t>(gridDim.y) * blockDim.x) mo4[i] = mi4[i];
  if (blockIdx.y == 0) {
    if (threadIdx.x == 0) {
      mo.orig[r2] = mi.orig[r];
      mo.slot[r2] = mi.slot[r];
      mo.valid_len[r2] = mi.valid_len[r];
      mo.use_prev[r2] = mi.use_prev[r];
      mo.prev_grip[r2] = cur_grip[r];
    }
    if (threadIdx.x < 6) mo.prev_pose[r2 * 6 + threadIdx.x] = cur_pose[r * 6 + threadIdx.x];   // action_list[-1], value_net.py:127
    for (int s = threadIdx.x; s < S; s += blockDim.x) mo.xmode[r2 * S + s] = mi.xmode[r * S + s];
  }
}

// prev_pose[r] = pose[r]  (first exit: pseudo previous action from the previous layer's features, value_net.py:122-125)
__global__ void copy_pose_kernel(const float* __restrict__ pose, float* __restrict__ prev_pose, int n_host, const int* __restrict__ n_dev) {
  const int n = dyn_rows(n_host, n_dev, 1);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * 6) prev_pose[i] = pose[i];
}

// debug taps: dst[orig[r], :] = src[r, :]
template <typename TSrc>
__global__ void tap_scatter_kernel(const TSrc* __restrict__ src, float* __restrict__ dst, const int* __restrict__ orig, size_t per,
                                   int n_host, const int* __restrict__ n_dev) {
  const int n = dyn_rows(n_host, n_dev, 1);
  const int r = blockIdx.x;
  if (r >= n) return;
  const size_t o = orig ? orig[r] : r;
  for (size_t i = blockIdx.y * static_cast<size_t>(blockDim.x) + threadIdx.x; i < per; i += static_cast<size_t>(gridDim.y) * blockDim.x)
    dst[o * per + i] = to_f32<TSrc>(src[static_cast<size_t>(r) * per + i]);
}

// fp32 -> T conversion (weight packing)
template <typename T>
__global__ void convert_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t rows, size_t cols, size_t src_ld, size_t dst_ld, size_t dst_col_off) {
  const size_t total = rows * cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / cols, c = i % cols;
    dst[r * dst_ld + dst_col_off + c] = from_f32<T>(src[r * src_ld + c]);
  }
}

// ---------------------------------------------------------------------------------------------
// On-device image transform (replaces the host PIL path in front of the hot path: data.py:898-902 -> open_clip
// image_transform): Resize(size, BICUBIC) -> CenterCrop(size) -> /255 -> Normalize(CLIP mean, std), uint8 HWC in,
// fp32 CHW out.  The resize reproduces Pillow's 8-bit resampler BIT-EXACTLY: horizontal pass then vertical pass, 22-bit
// fixed-point coefficients (tables built on the host with Pillow's own double arithmetic), int32 accumulation starting at
// 1<<21, arithmetic shift, clamp to [0,255], uint8 round-trip between the passes.  One thread per output pixel.
// ---------------------------------------------------------------------------------------------
struct ResizeAxis { const int* bounds; const int* kk; int ksize; int need; };   // bounds[out][2] = (min, count); kk[out][ksize]

__device__ __forceinline__ int clip8_fixed(int v) {
  v >>= 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ img, int n, int H, int W, ResizeAxis ax, ResizeAxis ay, int top, int left, int size,
                     float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= size * size || b >= n) return;
  const int oy = idx / size, ox = idx % size;
  const int ry = oy + top, rx = ox + left;                      // coordinates in the resized (pre-crop) image
  const uint8_t* im = img + static_cast<size_t>(b) * H * W * 3;
  int ymin = ry, yn = 1, xmin = rx, xn = 1;
  if (ay.need) { ymin = ay.bounds[2 * ry]; yn = ay.bounds[2 * ry + 1]; }
  if (ax.need) { xmin = ax.bounds[2 * rx]; xn = ax.bounds[2 * rx + 1]; }
  int v0 = 1 << 21, v1 = 1 << 21, v2 = 1 << 21;
  int h0 = 0, h1 = 0, h2 = 0;
  for (int y = 0; y < yn; ++y) {
    const uint8_t* row = im + static_cast<size_t>(ymin + y) * W * 3;
    if (ax.need) {
      int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
      for (int x = 0; x < xn; ++x) {
        const int k = ax.kk[rx * ax.ksize + x];
        const uint8_t* px = row + (xmin + x) * 3;
        a0 += px[0] * k; a1 += px[1] * k; a2 += px[2] * k;
      }
      h0 = clip8_fixed(a0); h1 = clip8_fixed(a1); h2 = clip8_fixed(a2);
    } else {
      h0 = row[rx * 3]; h1 = row[rx * 3 + 1]; h2 = row[rx * 3 + 2];
    }
    if (ay.need) {
      const int k = ay.kk[ry * ay.ksize + y];
      v0 += h0 * k; v1 += h1 * k; v2 += h2 * k;
    }
  }
  if (ay.need) { h0 = clip8_fixed(v0); h1 = clip8_fixed(v1); h2 = clip8_fixed(v2); }
  // ToTensor (/255) and Normalize, IEEE fp32 like torch
  const float m0 = 0.48145466f, m1 = 0.4578275f, m2 = 0.40821073f, s0 = 0.26862954f, s1 = 0.26130258f, s2 = 0.27577711f;
  float* o = out + static_cast<size_t>(b) * 3 * size * size + static_cast<size_t>(oy) * size + ox;
  o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(h0), 255.0f), m0), s0);
  o[static_cast<size_t>(size) * size] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(h1), 255.0f), m1), s1);
  o[2 * static_cast<size_t>(size) * size] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(h2), 255.0f), m2), s2);
}
