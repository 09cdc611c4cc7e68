// SIMT (fp32 FFMA) form of the RBM-family kernels: used for small / ragged layers (the reference's
// own test shapes are 6x5, tests/test_rbm.py:10-16), for the real-valued layer kinds (BGRBM / GBRBM /
// GGRBM / Concrete units), and as the general form of the CD gradient. Large binary RBMs go through
// the tcgen05 path (gemm_tc.cu / cd_chain.cu). Launches with few output tiles (the reference's 328 x 1024 x 500) split K
// over a thread-block cluster and reduce through distributed shared memory (simt_gemm.cuh).
#include "context.h"
#include "simt_gemm.cuh"

namespace db {
namespace {

// ------------------------------------------------------------------------------------------------
// K1 (SIMT): out = sample(act(col_scale * ((A / a_div) @ op(W)) + bias))
// ------------------------------------------------------------------------------------------------
struct LayerFwdParams {
  const float* A; int lda; const float* a_div; int a_div_mode;
  const float* W; int ldw;
  const float* col_scale; int col_scale_mode; const float* bias;
  int M, N, K; int act; int sample_kind; float inv_temp;
  const float* noise_scale; int noise_scale_mode; const float* injected; RngArgs rng;
  float* out_act; float* out_sample;
};

struct ALoadLayer {
  static constexpr bool k_contig = true;
  const float* A; int lda; const float* div; int div_mode; int M, K;
  __device__ __forceinline__ float operator()(int m, int k) const {
    if (m >= M || k >= K) return 0.0f;
    float v = __ldg(A + (size_t)m * lda + k);
    if (div_mode == DB_BCAST_ROW) v = v / __ldg(div + k);
    else if (div_mode == DB_BCAST_FULL) v = v / __ldg(div + (size_t)m * lda + k);
    return v;
  }
  __host__ __device__ bool vec_ok() const {
    return lda % 4 == 0 && aligned16(A) && (div_mode == DB_BCAST_NONE || aligned16(div));
  }
  __device__ __forceinline__ float4 load4(int m, int k) const {     // k .. k+3 of row m
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (k + 3 >= K) return make_float4((*this)(m, k), (*this)(m, k + 1), (*this)(m, k + 2), (*this)(m, k + 3));
    float4 v = __ldg(reinterpret_cast<const float4*>(A + (size_t)m * lda + k));
    if (div_mode != DB_BCAST_NONE) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(div + (div_mode == DB_BCAST_ROW ? (size_t)k : (size_t)m * lda + k)));
      v.x = v.x / d.x; v.y = v.y / d.y; v.z = v.z / d.z; v.w = v.w / d.w;
    }
    return v;
  }
};
template <bool TRANS>
struct BLoadW {
  static constexpr bool k_contig = TRANS;
  const float* W; int ldw; int K, N;
  __device__ __forceinline__ float operator()(int k, int n) const {
    if (k >= K || n >= N) return 0.0f;
    return TRANS ? __ldg(W + (size_t)n * ldw + k) : __ldg(W + (size_t)k * ldw + n);
  }
  __host__ __device__ bool vec_ok() const { return ldw % 4 == 0 && aligned16(W); }
  __device__ __forceinline__ float4 load4(int k, int n) const {     // TRANS: k .. k+3 of column n; else n .. n+3 of row k
    if (k >= K || n >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (TRANS) {
      if (k + 3 >= K) return make_float4((*this)(k, n), (*this)(k + 1, n), (*this)(k + 2, n), (*this)(k + 3, n));
      return __ldg(reinterpret_cast<const float4*>(W + (size_t)n * ldw + k));
    }
    if (n + 3 >= N) return make_float4((*this)(k, n), (*this)(k, n + 1), (*this)(k, n + 2), (*this)(k, n + 3));
    return __ldg(reinterpret_cast<const float4*>(W + (size_t)k * ldw + n));
  }
};

__device__ __forceinline__ float draw_sample(int kind, float val, float inv_temp, float noise, const float* injected,
                                             const RngArgs& rng, int m, int n, size_t idx) {
  if (kind == DB_SAMPLE_BERNOULLI) {
    float u = injected ? __ldg(injected + idx) : philox_uniform_at(rng, rng.row_offset + (uint32_t)m, (uint32_t)n);
    return val > u ? 1.0f : 0.0f;
  } else if (kind == DB_SAMPLE_CONCRETE) {
    float u = injected ? __ldg(injected + idx) : philox_uniform_at(rng, rng.row_offset + (uint32_t)m, (uint32_t)n);
    return concrete_sample(val, u, inv_temp);
  } else if (kind == DB_SAMPLE_GAUSSIAN) {
    float e = injected ? __ldg(injected + idx) : philox_normal_at(rng, rng.row_offset + (uint32_t)m, (uint32_t)n);
    return fmaf(noise, e, val);
  }
  return val;
}

// epilogue of one output element: scale, bias, activation, sample
__device__ __forceinline__ void layer_epilogue(const LayerFwdParams& p, const RngArgs& rng, int m, int n, float x) {
  if (m >= p.M || n >= p.N) return;
  const size_t idx = (size_t)m * p.N + n;
  if (p.col_scale_mode == DB_BCAST_ROW) x *= __ldg(p.col_scale + n);
  else if (p.col_scale_mode == DB_BCAST_FULL) x *= __ldg(p.col_scale + idx);
  if (p.bias) x += __ldg(p.bias + n);
  const float val = p.act == DB_ACT_SIGMOID ? sigmoidf_(x) : x;
  if (p.out_act) p.out_act[idx] = val;
  if (p.out_sample) {
    float ns = 0.0f;
    if (p.noise_scale_mode == DB_BCAST_ROW) ns = __ldg(p.noise_scale + n);
    else if (p.noise_scale_mode == DB_BCAST_FULL) ns = __ldg(p.noise_scale + idx);
    p.out_sample[idx] = draw_sample(p.sample_kind, val, p.inv_temp, ns, p.injected, rng, m, n, idx);
  }
}

// grid (N tiles, M tiles, splits); splits > 1: launched as clusters (1,1,splits) that split K (simt_gemm.cuh).
// VEC: float4 operand loads (alignment checked by the launcher)
template <bool TRANS, class T, bool VEC>
__global__ void __launch_bounds__(256) layer_forward_kernel(const LayerFwdParams p, int splits) {
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  ALoadLayer a{p.A, p.lda, p.a_div, p.a_div_mode, p.M, p.K};
  BLoadW<TRANS> b{p.W, p.ldw, p.K, p.N};
  float acc[T::TM][T::TN];
  int k_begin, k_end;
  simt_split_range<T>(p.K, splits, k_begin, k_end);
  if (VEC) simt_gemm_tile_vec<T>(sm, a, b, m0, n0, k_begin, k_end, acc);
  else simt_gemm_tile<T>(sm, a, b, m0, n0, k_begin, k_end, acc);
  const RngArgs rng = rng_resolved(p.rng);
  simt_cluster_reduce<T>(red, acc, splits, [&](int r, int c, float x) { layer_epilogue(p, rng, m0 + r, n0 + c, x); });
}

__global__ void sample_kernel(const float* __restrict__ P, int M, int N, int kind, float inv_temp,
                              const float* __restrict__ noise_scale, int ns_mode, const float* __restrict__ injected,
                              RngArgs rng_in, float* __restrict__ out) {
  const RngArgs rng = rng_resolved(rng_in);
  const size_t total = (size_t)M * N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx / N), n = (int)(idx % N);
    float ns = 0.0f;
    if (ns_mode == DB_BCAST_ROW) ns = __ldg(noise_scale + n);
    else if (ns_mode == DB_BCAST_FULL) ns = __ldg(noise_scale + idx);
    out[idx] = draw_sample(kind, __ldg(P + idx), inv_temp, ns, injected, rng, m, n, idx);
  }
}

// ------------------------------------------------------------------------------------------------
// free energy row reduction: one warp per datapoint
// ------------------------------------------------------------------------------------------------
__global__ void free_energy_reduce_kernel(int gaussian_hidden, const float* __restrict__ v, const float* __restrict__ x,
                                          int N, int V, int H, const float* __restrict__ b, const float* __restrict__ c,
                                          const float* __restrict__ sigma, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float s = 0.0f;
  for (int i = lane; i < V; i += 32) s = fmaf(__ldg(v + (size_t)row * V + i), __ldg(b + i), s);
  for (int j = lane; j < H; j += 32) {
    const float xj = __ldg(x + (size_t)row * H + j);
    if (gaussian_hidden) {
      // bgrbm.py:143-151: 0.5*x^2 + (c/sigma)*x
      s += 0.5f * xj * xj + (__ldg(c + j) / __ldg(sigma + j)) * xj;
    } else {
      s += softplusf_(xj + __ldg(c + j));   // rbm.py:388-389
    }
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = -s;
}

// ------------------------------------------------------------------------------------------------
// CD gradient (general form)
// ------------------------------------------------------------------------------------------------
// q = sigmoid(x + c) (Bernoulli hidden) or x + c/sigma (Gaussian hidden); in place on x [N,H]
__global__ void cd_hidden_term_kernel(int gaussian_hidden, float* __restrict__ x, int N, int H, const float* __restrict__ c,
                                      const float* __restrict__ sigma) {
  const size_t total = (size_t)N * H;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % H);
    const float xv = x[idx];
    x[idx] = gaussian_hidden ? xv + __ldg(c + j) / __ldg(sigma + j) : sigmoidf_(xv + __ldg(c + j));
  }
}

// dW[v,h] = -inv_n * sum_n (v0[n,v] q0[n,h] - vk[n,v] qk[n,h])   as one GEMM with K = 2N
struct ALoadVT {
  static constexpr bool k_contig = false;   // for fixed k (=n) memory is contiguous along m (=v)
  const float* v0; const float* vk; int N, V;
  __device__ __forceinline__ float operator()(int m, int k) const {
    if (m >= V || k >= 2 * N) return 0.0f;
    // k = 2n + phase: data and model terms alternate so the fp32 partial sums of the difference stay small
    return (k & 1) ? __ldg(vk + (size_t)(k >> 1) * V + m) : __ldg(v0 + (size_t)(k >> 1) * V + m);
  }
  __host__ __device__ bool vec_ok() const { return V % 4 == 0 && aligned16(v0) && aligned16(vk); }
  __device__ __forceinline__ float4 load4(int m, int k) const {     // m .. m+3 of row k
    if (m >= V || k >= 2 * N) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (m + 3 >= V) return make_float4((*this)(m, k), (*this)(m + 1, k), (*this)(m + 2, k), (*this)(m + 3, k));
    return __ldg(reinterpret_cast<const float4*>(((k & 1) ? vk : v0) + (size_t)(k >> 1) * V + m));
  }
};
struct BLoadQ {
  static constexpr bool k_contig = false;
  const float* q0; const float* qk; int N, H;
  __device__ __forceinline__ float operator()(int k, int n) const {
    if (n >= H || k >= 2 * N) return 0.0f;
    return (k & 1) ? -__ldg(qk + (size_t)(k >> 1) * H + n) : __ldg(q0 + (size_t)(k >> 1) * H + n);
  }
  __host__ __device__ bool vec_ok() const { return H % 4 == 0 && aligned16(q0) && aligned16(qk); }
  __device__ __forceinline__ float4 load4(int k, int n) const {     // n .. n+3 of row k
    if (n >= H || k >= 2 * N) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (n + 3 >= H) return make_float4((*this)(k, n), (*this)(k, n + 1), (*this)(k, n + 2), (*this)(k, n + 3));
    const float4 v = __ldg(reinterpret_cast<const float4*>(((k & 1) ? qk : q0) + (size_t)(k >> 1) * H + n));
    return (k & 1) ? make_float4(-v.x, -v.y, -v.z, -v.w) : v;
  }
};
template <class T, bool VEC>
__global__ void __launch_bounds__(256) cd_dw_kernel(const float* v0, const float* vk, const float* q0, const float* qk,
                                                    int N, int V, int H, float scale, float* __restrict__ dW, int splits) {
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  ALoadVT a{v0, vk, N, V};
  BLoadQ b{q0, qk, N, H};
  float acc[T::TM][T::TN];
  int k_begin, k_end;
  simt_split_range<T>(2 * N, splits, k_begin, k_end);
  if (VEC) simt_gemm_tile_vec<T>(sm, a, b, m0, n0, k_begin, k_end, acc);
  else simt_gemm_tile<T>(sm, a, b, m0, n0, k_begin, k_end, acc);
  simt_cluster_reduce<T>(red, acc, splits, [&](int r, int c, float x) {
    if (m0 + r < V && n0 + c < H) dW[(size_t)(m0 + r) * H + n0 + c] = scale * x;
  });
}

// Column sums of two differences in one launch: out1[c] += scale * sum_n (a1 - b1)[n,c] over C1 columns and the same for
// (a2, b2, C2, out2); outputs must be zeroed first. The 328-row batches of the reference leave this latency-bound (bytes
// in flight, not bandwidth), so the CTA is narrow and tall: 256 threads = 8 columns x 32 row lanes (rows lane, lane + 32,
// ...; four independent partial sums per thread), lanes summed in order through shared memory; grid (ceil(C1/8) +
// ceil(C2/8), row chunks) — with one row chunk the result is deterministic (one atomicAdd onto 0 per column).
__global__ void __launch_bounds__(256) colsum_diff2_kernel(const float* __restrict__ a1, const float* __restrict__ b1, int C1,
                                                           float* __restrict__ out1, const float* __restrict__ a2,
                                                           const float* __restrict__ b2, int C2, float* __restrict__ out2,
                                                           int N, float scale) {
  __shared__ float part[32][9];
  const int blocks1 = (C1 + 7) / 8;
  const bool second = (int)blockIdx.x >= blocks1;
  const float* __restrict__ a = second ? a2 : a1;
  const float* __restrict__ b = second ? b2 : b1;
  float* __restrict__ out = second ? out2 : out1;
  const int C = second ? C2 : C1;
  const int cl = threadIdx.x & 7, lane = threadIdx.x >> 3;
  const int c = ((int)blockIdx.x - (second ? blocks1 : 0)) * 8 + cl;
  const int rows_per = (N + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(N, r0 + rows_per);
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  if (c < C) {
    int n = r0 + lane;
    for (; n + 96 < r1; n += 128) {
      s0 += __ldg(a + (size_t)n * C + c) - __ldg(b + (size_t)n * C + c);
      s1 += __ldg(a + (size_t)(n + 32) * C + c) - __ldg(b + (size_t)(n + 32) * C + c);
      s2 += __ldg(a + (size_t)(n + 64) * C + c) - __ldg(b + (size_t)(n + 64) * C + c);
      s3 += __ldg(a + (size_t)(n + 96) * C + c) - __ldg(b + (size_t)(n + 96) * C + c);
    }
    for (; n < r1; n += 32) s0 += __ldg(a + (size_t)n * C + c) - __ldg(b + (size_t)n * C + c);
  }
  part[lane][cl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (lane == 0 && c < C) {
    float t = part[0][cl];
#pragma unroll
    for (int q = 1; q < 32; ++q) t += part[q][cl];
    atomicAdd(out + c, scale * t);
  }
}

// Gaussian-hidden bias / sigma gradients from d[j] = inv_n * sum_n (x0 - xk)[n,j]  (held in gc on entry)
//   dc = -d/sigma ; dsigma = +(c/sigma^2) d ; dopt_sigma = dsigma * sigmoid(opt_sigma)   
__global__ void cd_gaussian_hidden_grads_kernel(int H, const float* __restrict__ c, const float* __restrict__ opt_sigma,
                                                float* __restrict__ gc, float* __restrict__ gsig) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  const float d = -gc[j];   // colsum kernel stored -inv_n * sum(x0 - xk)
  const float os = __ldg(opt_sigma + j);
  const float sig = softplusf_(os);
  gc[j] = -d / sig;
  gsig[j] = (__ldg(c + j) / (sig * sig)) * d * sigmoidf_(os);
}

__global__ void softplus_vec_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = softplusf_(x[i]);
}

// ------------------------------------------------------------------------------------------------
// optimizer
// ------------------------------------------------------------------------------------------------
__global__ void optimizer_kernel(db_opt_t o, float lr_t, float* __restrict__ p, const float* __restrict__ g,
                                 float* __restrict__ st, float* __restrict__ st2, size_t n) {
  if (o.clock && o.kind == DB_OPT_ADAM) {      // step number from the device clock (graph replay): bias correction here
    __shared__ float lr_sh;
    if (threadIdx.x == 0) {
      const double t = (double)o.step + (double)__ldg((const unsigned long long*)o.clock + 1);
      lr_sh = (float)((double)o.lr * sqrt(1.0 - pow((double)o.beta2, t)) / (1.0 - pow((double)o.beta1, t)));
    }
    __syncthreads();
    lr_t = lr_sh;
  }
  static_assert(DB_OPT_SGD == 0 && DB_OPT_MOMENTUM == 1 && DB_OPT_ADAM == 2, "opt_apply (common.cuh) uses these values");
  const OptRule rule{o.kind, o.lr, lr_t, o.momentum, o.beta1, o.beta2, o.eps, o.weight_decay};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float s1 = o.kind != DB_OPT_SGD ? st[i] : 0.0f, s2 = o.kind == DB_OPT_ADAM ? st2[i] : 0.0f;
    p[i] = opt_apply(rule, p[i], __ldg(g + i), s1, s2);
    if (o.kind != DB_OPT_SGD) st[i] = s1;
    if (o.kind == DB_OPT_ADAM) st2[i] = s2;
  }
}

// ------------------------------------------------------------------------------------------------
// dist.py reductions
// ------------------------------------------------------------------------------------------------
__global__ void mse_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t total, float scale,
                           float* __restrict__ out) {
  float s = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float d = __ldg(a + i) - __ldg(b + i);
    s = fmaf(d, d, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s * scale);
}
__global__ void cross_entropy_kernel(const float* __restrict__ probs, const float* __restrict__ t, size_t total, float scale,
                                     float* __restrict__ out) {
  float s = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float e0 = 1.0e-6f;                              // dist.py:25 (10e-7)
    float p = fminf(fmaxf(__ldg(probs + i), e0), 1.0f - e0);
    float x = logf(p / (1.0f - p));                        // dist.py:27
    float z = __ldg(t + i);
    s += fmaxf(x, 0.0f) - x * z + log1pf(__expf(-fabsf(x)));   // sigmoid_cross_entropy_with_logits [TF]
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s * scale);
}

inline int ew_grid(size_t n, int num_sms) {
  size_t blocks = (n + 255) / 256;
  size_t cap = (size_t)num_sms * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

int launch_layer_forward(db_handle_t h, cudaStream_t st, const LayerFwdParams& p, int transW) {
  // (128 x 64 tiles with 8 x 4 micro-tiles measured SLOWER at the reference's sizes, 328 x 1024 x 500: 31.8 vs 24.1 us per
  // launch — fewer, fatter CTAs, and the kernel is latency- not FMA-bound)
  using T = TileSmall;
  dim3 grid((p.N + T::BN - 1) / T::BN, (p.M + T::BM - 1) / T::BM);
  const int splits = simt_splits((long)grid.x * grid.y, p.K, h->num_sms);
  grid.z = splits;
  const ALoadLayer a{p.A, p.lda, p.a_div, p.a_div_mode, p.M, p.K};
  const bool vec = a.vec_ok() && BLoadW<false>{p.W, p.ldw, p.K, p.N}.vec_ok();
  auto kernel = transW ? (vec ? layer_forward_kernel<true, T, true> : layer_forward_kernel<true, T, false>)
                       : (vec ? layer_forward_kernel<false, T, true> : layer_forward_kernel<false, T, false>);
  simt_launch_split(kernel, grid, 256, st, p, splits);
  return check_launch(h, "layer_forward");
}

}  // namespace db

using namespace db;

extern "C" int db_layer_forward(db_handle_t h, db_stream_t stream, const float* A, int lda, const float* a_div,
                                int a_div_mode, const float* W, int ldw, int transW, const float* col_scale,
                                int col_scale_mode, const float* bias, int M, int N, int K, int act, int sample_kind,
                                float temperature, const float* noise_scale, int noise_scale_mode,
                                const float* injected, const db_rng_t* rng, float* out_act, float* out_sample) {
  DB_REQUIRE(h, A && W && M > 0 && N > 0 && K > 0, "null operand or empty shape");
  DB_REQUIRE(h, out_act || out_sample, "no output requested");
  DB_REQUIRE(h, (a_div_mode == DB_BCAST_NONE) == (a_div == nullptr), "a_div / a_div_mode mismatch");
  DB_REQUIRE(h, (col_scale_mode == DB_BCAST_NONE) == (col_scale == nullptr), "col_scale / mode mismatch");
  DB_REQUIRE(h, sample_kind >= DB_SAMPLE_NONE && sample_kind <= DB_SAMPLE_GAUSSIAN, "bad sample_kind");
  DB_REQUIRE(h, sample_kind != DB_SAMPLE_CONCRETE || temperature > 0.0f, "Concrete units need temperature > 0");
  DB_REQUIRE(h, sample_kind != DB_SAMPLE_GAUSSIAN || noise_scale != nullptr, "Gaussian sampling needs noise_scale");
  DB_REQUIRE(h, sample_kind == DB_SAMPLE_NONE || injected || rng, "sampling needs injected draws or an rng");
  LayerFwdParams p;
  p.A = A; p.lda = lda; p.a_div = a_div; p.a_div_mode = a_div_mode; p.W = W; p.ldw = ldw;
  p.col_scale = col_scale; p.col_scale_mode = col_scale_mode; p.bias = bias; p.M = M; p.N = N; p.K = K;
  p.act = act; p.sample_kind = sample_kind; p.inv_temp = temperature > 0.0f ? 1.0f / temperature : 0.0f;
  p.noise_scale = noise_scale; p.noise_scale_mode = noise_scale ? noise_scale_mode : DB_BCAST_NONE;
  p.injected = injected; to_rng(rng, p.rng); p.out_act = out_act; p.out_sample = out_sample;
  return launch_layer_forward(h, (cudaStream_t)stream, p, transW);
}

extern "C" int db_sample(db_handle_t h, db_stream_t stream, const float* P, int M, int N, int sample_kind,
                         float temperature, const float* noise_scale, int noise_scale_mode, const float* injected,
                         const db_rng_t* rng, float* out) {
  DB_REQUIRE(h, P && out && M > 0 && N > 0, "null operand or empty shape");
  DB_REQUIRE(h, sample_kind >= DB_SAMPLE_NONE && sample_kind <= DB_SAMPLE_GAUSSIAN, "bad sample_kind");
  DB_REQUIRE(h, sample_kind != DB_SAMPLE_CONCRETE || temperature > 0.0f, "Concrete units need temperature > 0");
  DB_REQUIRE(h, sample_kind != DB_SAMPLE_GAUSSIAN || noise_scale != nullptr, "Gaussian sampling needs noise_scale");
  DB_REQUIRE(h, sample_kind == DB_SAMPLE_NONE || injected || rng, "sampling needs injected draws or an rng");
  RngArgs r; to_rng(rng, r);
  const size_t total = (size_t)M * N;
  sample_kernel<<<ew_grid(total, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
      P, M, N, sample_kind, temperature > 0.0f ? 1.0f / temperature : 0.0f, noise_scale,
      noise_scale ? noise_scale_mode : DB_BCAST_NONE, injected, r, out);
  return check_launch(h, "sample");
}

static bool gaussian_hidden(int kind) { return kind == DB_LAYER_BGRBM || kind == DB_LAYER_GGRBM; }

extern "C" int db_free_energy(db_handle_t h, db_stream_t stream, int layer_kind, const float* v, int N, int V, int H,
                              const float* W, const float* b, const float* c, const float* sigma_h, float* workspace,
                              float* out) {
  DB_REQUIRE(h, v && W && b && c && workspace && out && N > 0 && V > 0 && H > 0, "null operand or empty shape");
  DB_REQUIRE(h, !gaussian_hidden(layer_kind) || sigma_h, "Gaussian hidden units need sigma_h");
  cudaStream_t st = (cudaStream_t)stream;
  LayerFwdParams p{};
  p.A = v; p.lda = V; p.W = W; p.ldw = H; p.M = N; p.N = H; p.K = V; p.act = DB_ACT_LINEAR;
  p.out_act = workspace;
  int rc = launch_layer_forward(h, st, p, 0);
  if (rc) return rc;
  free_energy_reduce_kernel<<<(N + 7) / 8, 256, 0, st>>>(gaussian_hidden(layer_kind) ? 1 : 0, v, workspace, N, V, H, b, c,
                                                       sigma_h, out);
  return check_launch(h, "free_energy_reduce");
}

extern "C" size_t db_cd_param_count(int layer_kind, int V, int H) {
  size_t n = (size_t)V * H + V + H;
  if (gaussian_hidden(layer_kind)) n += H;
  return n;
}

// hp0 / hpk: hidden probabilities sigmoid(v W + c) at the two ends of the chain when the caller already has them
// (Bernoulli-hidden layers: they are what the chain sampled h_0 and h_k from), else recomputed here
static int cd_gradients_impl(db_handle_t h, cudaStream_t st, int layer_kind, const float* v0, const float* vk,
                             const float* hp0, const float* hpk, int N, int V, int H, const float* params,
                             float inv_global_n, float* workspace, float* grad) {
  const float* W = params;
  const float* c = params + (size_t)V * H + V;
  const float* opt_sigma = c + H;
  float* gW = grad; float* gb = grad + (size_t)V * H; float* gc = gb + V; float* gsig = gc + H;
  const bool gh = gaussian_hidden(layer_kind);
  const float* q0 = hp0; const float* qk = hpk;
  float* sigma = nullptr;
  if (gh) {
    // effective sigma_h = softplus(opt_sigma_h) (bgrbm.py:51), staged in the gsig slot until the end
    sigma = gsig;
    softplus_vec_kernel<<<(H + 255) / 256, 256, 0, st>>>(opt_sigma, H, sigma);
  }
  const int row_chunks = N >= 8192 ? 8 : 1;
  cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)(V + H), st);
  auto colsums = [&](bool visible, bool hidden) {      // db = -inv_n sum (v0 - vk), dc = -inv_n sum (q0 - qk)
    const int bv = visible ? (V + 7) / 8 : 0, bh = hidden ? (H + 7) / 8 : 0;
    colsum_diff2_kernel<<<dim3(bv + bh, row_chunks), 256, 0, st>>>(v0, vk, visible ? V : 0, gb, q0, qk, hidden ? H : 0, gc, N,
                                                                 -inv_global_n);
  };
  if (!hp0) {
    float* w0 = workspace; float* wk = workspace + (size_t)N * H;
    for (int side = 0; side < 2; ++side) {
      LayerFwdParams p{};
      p.A = side ? vk : v0; p.lda = V; p.W = W; p.ldw = H; p.M = N; p.N = H; p.K = V;
      if (gh) { p.act = DB_ACT_LINEAR; }                       // x = v W; c / sigma is added below
      else { p.act = DB_ACT_SIGMOID; p.bias = c; }             // q = sigmoid(v W + c) in the GEMM epilogue
      p.out_act = side ? wk : w0;
      int rc = launch_layer_forward(h, st, p, 0);
      if (rc) return rc;
    }
    q0 = w0; qk = wk;
    if (gh) {
      // bias sums need x0 - xk, i.e. before c / sigma is added (the shift cancels in the difference anyway)
      colsums(true, true);
      const size_t nh = (size_t)N * H;
      cd_hidden_term_kernel<<<ew_grid(nh, h->num_sms), 256, 0, st>>>(1, w0, N, H, c, sigma);
      cd_hidden_term_kernel<<<ew_grid(nh, h->num_sms), 256, 0, st>>>(1, wk, N, H, c, sigma);
    }
  }
  if (!gh) colsums(true, true);
  {
    using T = TileSmall;
    dim3 g((H + T::BN - 1) / T::BN, (V + T::BM - 1) / T::BM);
    const int splits = simt_splits((long)g.x * g.y, 2 * N, h->num_sms);
    g.z = splits;
    const bool vec = ALoadVT{v0, vk, N, V}.vec_ok() && BLoadQ{q0, qk, N, H}.vec_ok();
    simt_launch_split(vec ? cd_dw_kernel<T, true> : cd_dw_kernel<T, false>, g, 256, st, v0, vk, q0, qk, N, V, H,
                      -inv_global_n, gW, splits);
  }
  if (gh) cd_gaussian_hidden_grads_kernel<<<(H + 255) / 256, 256, 0, st>>>(H, c, opt_sigma, gc, gsig);
  return check_launch(h, "cd_gradients");
}

extern "C" int db_cd_gradients(db_handle_t h, db_stream_t stream, int layer_kind, const float* v0, const float* vk, int N,
                               int V, int H, const float* params, float inv_global_n, float* workspace, float* grad) {
  DB_REQUIRE(h, v0 && vk && params && workspace && grad && N > 0 && V > 0 && H > 0, "null operand or empty shape");
  return cd_gradients_impl(h, (cudaStream_t)stream, layer_kind, v0, vk, nullptr, nullptr, N, V, H, params, inv_global_n,
                           workspace, grad);
}

extern "C" int db_cd_gradients_hidden(db_handle_t h, db_stream_t stream, int layer_kind, const float* v0, const float* vk,
                                      const float* hp0, const float* hpk, int N, int V, int H, const float* params,
                                      float inv_global_n, float* grad) {
  DB_REQUIRE(h, v0 && vk && hp0 && hpk && params && grad && N > 0 && V > 0 && H > 0, "null operand or empty shape");
  DB_REQUIRE(h, !gaussian_hidden(layer_kind), "Gaussian-hidden layers recompute their hidden term (db_cd_gradients)");
  return cd_gradients_impl(h, (cudaStream_t)stream, layer_kind, v0, vk, hp0, hpk, N, V, H, params, inv_global_n, nullptr,
                           grad);
}

__global__ void clock_advance_kernel(unsigned long long* clock, unsigned long long draws, unsigned long long steps) {
  clock[0] += draws;
  clock[1] += steps;
}

extern "C" int db_clock_advance(db_handle_t h, db_stream_t stream, uint64_t* clock, uint64_t draws, uint64_t steps) {
  DB_REQUIRE(h, clock, "null clock");
  clock_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)clock, draws, steps);
  return check_launch(h, "clock_advance");
}

extern "C" int db_optimizer_step(db_handle_t h, db_stream_t stream, const db_opt_t* opt, float* params, const float* grad,
                                 float* state, size_t n_total) {
  const size_t begin = 0, count = n_total;
  DB_REQUIRE(h, opt && params && grad && n_total > 0, "null operand or empty parameter vector");
  DB_REQUIRE(h, opt->kind == DB_OPT_SGD || state, "optimizer state required");
  float lr_t = opt->lr;
  if (opt->kind == DB_OPT_ADAM) {
    DB_REQUIRE(h, opt->step >= 1, "Adam needs step >= 1");
    // lr * sqrt(1 - b2^t) / (1 - b1^t), epsilon outside the root [TF AdamOptimizer, not vendored]
    const double b1t = pow((double)opt->beta1, opt->step), b2t = pow((double)opt->beta2, opt->step);
    lr_t = (float)((double)opt->lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  }
  float* m = state ? state + begin : nullptr;
  float* v = (state && opt->kind == DB_OPT_ADAM) ? state + n_total + begin : nullptr;
  optimizer_kernel<<<ew_grid(count, h->num_sms), 256, 0, (cudaStream_t)stream>>>(*opt, lr_t, params + begin, grad + begin, m, v, count);
  return check_launch(h, "optimizer_step");
}

extern "C" int db_mse(db_handle_t h, db_stream_t stream, const float* a, const float* b, int N, int D, float* out) {
  DB_REQUIRE(h, a && b && out && N > 0 && D > 0, "null operand or empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(float), st);
  const size_t total = (size_t)N * D;
  mse_kernel<<<ew_grid(total, h->num_sms), 256, 0, st>>>(a, b, total, 1.0f / (float)N, out);
  return check_launch(h, "mse");
}

extern "C" int db_cross_entropy(db_handle_t h, db_stream_t stream, const float* probs, const float* targets, int N, int D,
                                int reduce_mean, float* out) {
  DB_REQUIRE(h, probs && targets && out && N > 0 && D > 0, "null operand or empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(float), st);
  const size_t total = (size_t)N * D;
  cross_entropy_kernel<<<ew_grid(total, h->num_sms), 256, 0, st>>>(probs, targets, total,
                                                                   reduce_mean ? 1.0f / (float)N : 1.0f, out);
  return check_launch(h, "cross_entropy");
}

extern "C" int db_softplus(db_handle_t h, db_stream_t stream, const float* x, int n, float* out) {
  DB_REQUIRE(h, x && out && n > 0, "null operand or empty shape");
  softplus_vec_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  return check_launch(h, "softplus");
}
