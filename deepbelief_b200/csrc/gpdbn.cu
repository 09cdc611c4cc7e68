// Backward pass of the GPDBN stack (gprbm.py:247-256 loss, minimised by TF autodiff in gprbm.py:386-396):
// the general fp32 GEMM entry, the Concrete-unit and cross-entropy derivatives, and the fused elementwise
// stages of the GPRBM top layer. The GP block itself (K, L, K^-1, pred_var, Hm, dX) lives in gp.cu.
//
// Top-layer algebra (alpha = alpha_h[d], r = sqrt(pred_var[n]), sigma_h = alpha r — gprbm.py:114-115):
//   H1 = sigma_h (x + eps1) + c           x = a W          (bgrbm.py:79-82, 107-110)
//   H2 = (H1 - mean_n H1) / alpha                           (gprbm.py:146-147)
//   H  = Hm alpha + mean_n H1 + sigma_h eps2                (gprbm.py:225-231)
//   y  = H / sigma_h                                         (bgrbm.py:126)
#include "context.h"
#include "gemm_simt.cuh"

namespace db {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// dz = g * d concrete/dp * dp/dz, with a = concrete(p, u, T) saved by the forward pass (rbm.py:243-275):
//   da/dp = a (1-a) / T * (1/(p+e0) + 1/(1-p+e0)),  dp/dz = p (1-p)
__global__ void concrete_backward_kernel(const float* __restrict__ g, const float* __restrict__ p, const float* __restrict__ a,
                                         size_t n, float inv_temp, float* __restrict__ dz) {
  const float e0 = 1.0e-6f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pv = __ldg(p + i), av = __ldg(a + i);
    dz[i] = __ldg(g + i) * av * (1.0f - av) * inv_temp * (1.0f / (pv + e0) + 1.0f / (1.0f - pv + e0)) * pv * (1.0f - pv);
  }
}

// dz = g * p (1 - p): backward of tf.sigmoid when the probabilities themselves are propagated (sample=False)
__global__ void sigmoid_backward_kernel(const float* __restrict__ g, const float* __restrict__ p, size_t n, float* __restrict__ dz) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pv = __ldg(p + i);
    dz[i] = __ldg(g + i) * pv * (1.0f - pv);
  }
}

// d/dprobs of dist.cross_entropy (dist.py:24-35): clip to [1e-6, 1-1e-6] (gradient passes inside the interval only),
// logits = log(pc/(1-pc)), sigmoid CE  ->  -t/pc + (1-t)/(1-pc)
__global__ void cross_entropy_grad_kernel(const float* __restrict__ probs, const float* __restrict__ t, size_t n, float scale,
                                          float* __restrict__ out) {
  const float e0 = 1.0e-6f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float p = __ldg(probs + i), tv = __ldg(t + i);
    const float pc = fminf(fmaxf(p, e0), 1.0f - e0);
    const bool inside = p >= e0 && p <= 1.0f - e0;
    out[i] = inside ? scale * (-tv / pc + (1.0f - tv) / (1.0f - pc)) : 0.0f;
  }
}

// out[c] = beta * out[c] + scale * sum_n A[n,c] : one CTA per 32 columns (x) and row slice (y), 8 row lanes.
// gridDim.y > 1 (tall matrices with few columns, e.g. the bias gradients over a batch of 5000): the slices add their
// partial sums with atomics into out, which colsum_scale_kernel has already replaced by beta * out.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int M, int N, float scale, float beta,
                                                     float* __restrict__ out, int rows_per_slice) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const int r_begin = blockIdx.y * rows_per_slice, r_end = min(M, r_begin + rows_per_slice);
  float s = 0.0f;
  if (c < N) for (int r = r_begin + ry; r < r_end; r += 8) s += __ldg(A + (size_t)r * N + c);
  part[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < N) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
    if (gridDim.y > 1) atomicAdd(out + c, scale * t);
    else out[c] = (beta != 0.0f ? beta * out[c] : 0.0f) + scale * t;
  }
}
__global__ void colsum_scale_kernel(float* __restrict__ out, int N, float beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < N) out[c] = beta != 0.0f ? beta * out[c] : 0.0f;
}

// ---- top layer, forward --------------------------------------------------------------------------------
// one CTA per hidden unit d: H1[:,d], its mean, H2[:,d]
__global__ void __launch_bounds__(256) top_up_kernel(const float* __restrict__ x, const float* __restrict__ pv,
                                                     const float* __restrict__ alpha_h, const float* __restrict__ c,
                                                     const float* __restrict__ eps1, int N, int D, float* __restrict__ H1,
                                                     float* __restrict__ H1_mean, float* __restrict__ H2) {
  __shared__ float red[8];
  const int d = blockIdx.x;
  const float al = __ldg(alpha_h + d), cd = __ldg(c + d);
  float s = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = (size_t)n * D + d;
    const float h = al * sqrtf(__ldg(pv + n)) * (__ldg(x + i) + __ldg(eps1 + i)) + cd;
    H1[i] = h;
    s += h;
  }
  const float mean = block_sum(s, red) / (float)N;
  if (threadIdx.x == 0) H1_mean[d] = mean;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = (size_t)n * D + d;
    H2[i] = (H1[i] - mean) / al;
  }
}
// H = Hm alpha + mean + alpha r eps2 ; y = H / (alpha r)
__global__ void top_mid_kernel(const float* __restrict__ Hm, const float* __restrict__ pv, const float* __restrict__ alpha_h,
                               const float* __restrict__ H1_mean, const float* __restrict__ eps2, int N, int D,
                               float* __restrict__ H, float* __restrict__ y) {
  const size_t total = (size_t)N * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / D), d = (int)(i % D);
    const float al = __ldg(alpha_h + d), sg = al * sqrtf(__ldg(pv + n));
    const float h = fmaf(__ldg(Hm + i), al, __ldg(H1_mean + d)) + sg * __ldg(eps2 + i);
    if (H) H[i] = h;
    y[i] = h / sg;
  }
}

// ---- top layer, backward -------------------------------------------------------------------------------
// from g_y: g_Hm = g_H alpha, gsig = -g_y y / sigma + g_H eps2, ga[d] = sum_n g_H Hm, gmean[d] = sum_n g_H   (g_H = g_y / sigma)
__global__ void __launch_bounds__(256) top_back1_kernel(const float* __restrict__ g_y, const float* __restrict__ y,
                                                        const float* __restrict__ Hm, const float* __restrict__ pv,
                                                        const float* __restrict__ alpha_h, const float* __restrict__ eps2, int N,
                                                        int D, float* __restrict__ g_Hm, float* __restrict__ gsig,
                                                        float* __restrict__ ga, float* __restrict__ gmean) {
  __shared__ float red[8];
  const int d = blockIdx.x;
  const float al = __ldg(alpha_h + d);
  float sa = 0.0f, sm = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = (size_t)n * D + d;
    const float sg = al * sqrtf(__ldg(pv + n));
    const float gy = __ldg(g_y + i), gH = gy / sg;
    g_Hm[i] = gH * al;
    gsig[i] = -gy * __ldg(y + i) / sg + gH * __ldg(eps2 + i);
    sa = fmaf(gH, __ldg(Hm + i), sa);
    sm += gH;
  }
  sa = block_sum(sa, red);
  sm = block_sum(sm, red);
  if (threadIdx.x == 0) { ga[d] = sa; gmean[d] = sm; }
}
// from g_H2: g_H1 (centred), g_x = g_H1 sigma, gsig += g_H1 (x + eps1), g_c[d], g_alpha[d] total
__global__ void __launch_bounds__(256) top_back2_kernel(const float* __restrict__ g_H2, const float* __restrict__ H2,
                                                        const float* __restrict__ x, const float* __restrict__ eps1,
                                                        const float* __restrict__ pv, const float* __restrict__ alpha_h,
                                                        const float* __restrict__ ga_in, const float* __restrict__ gmean_in, int N,
                                                        int D, float* __restrict__ g_x, float* __restrict__ gsig,
                                                        float* __restrict__ g_c, float* __restrict__ g_alpha) {
  __shared__ float red[8];
  const int d = blockIdx.x;
  const float al = __ldg(alpha_h + d);
  float s1 = 0.0f, s2 = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = (size_t)n * D + d;
    const float g = __ldg(g_H2 + i);
    s1 += g / al;
    s2 = fmaf(g, __ldg(H2 + i), s2);
  }
  s1 = block_sum(s1, red);                       // sum_n g_H2 / alpha
  s2 = block_sum(s2, red);                       // sum_n g_H2 o H2
  const float gmean_total = __ldg(gmean_in + d) - s1;
  float sc = 0.0f, sal = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = (size_t)n * D + d;
    const float r = sqrtf(__ldg(pv + n));
    const float gH1 = __ldg(g_H2 + i) / al + gmean_total / (float)N;
    g_x[i] = gH1 * al * r;
    const float gs = gsig[i] + gH1 * (__ldg(x + i) + __ldg(eps1 + i));
    gsig[i] = gs;
    sc += gH1;
    sal = fmaf(gs, r, sal);
  }
  sc = block_sum(sc, red);
  sal = block_sum(sal, red);
  if (threadIdx.x == 0) {
    g_c[d] = sc;
    g_alpha[d] = __ldg(ga_in + d) - s2 / al + sal;
  }
}
// g_pv[n] = (sum_d gsig[n,d] alpha_d) / (2 r_n) ; g_opt_alpha_h[d] = g_alpha[d] sigmoid(opt_alpha_h[d])  (one warp per row)
__global__ void top_back3_kernel(const float* __restrict__ gsig, const float* __restrict__ pv, const float* __restrict__ alpha_h,
                                 const float* __restrict__ g_alpha, const float* __restrict__ opt_alpha_h, int N, int D,
                                 float* __restrict__ g_pv, float* __restrict__ g_opt_alpha_h) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0)
    for (int d = threadIdx.x; d < D; d += blockDim.x) g_opt_alpha_h[d] = __ldg(g_alpha + d) * sigmoidf_(__ldg(opt_alpha_h + d));
  if (row >= N) return;
  float s = 0.0f;
  for (int d = lane; d < D; d += 32) s = fmaf(__ldg(gsig + (size_t)row * D + d), __ldg(alpha_h + d), s);
  s = warp_sum(s);
  if (lane == 0) g_pv[row] = s / (2.0f * sqrtf(__ldg(pv + row)));
}

inline int ew_blocks(size_t n, int num_sms) {
  size_t blocks = (n + 255) / 256, cap = (size_t)num_sms * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace
}  // namespace db

using namespace db;

extern "C" int db_gemm(db_handle_t h, db_stream_t stream, int transA, int transB, int M, int N, int K, float alpha,
                       const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc) {
  DB_REQUIRE(h, A && B && C && M > 0 && N > 0 && K > 0, "null operand or empty shape");
  DB_REQUIRE(h, lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "leading dimension too small");
  return run_gemm_f32(h, (cudaStream_t)stream, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

extern "C" int db_concrete_backward(db_handle_t h, db_stream_t stream, const float* g, const float* p, const float* a, int M,
                                    int N, float temperature, float* dz) {
  DB_REQUIRE(h, g && p && a && dz && M > 0 && N > 0 && temperature > 0.0f, "null operand, empty shape or temperature <= 0");
  const size_t n = (size_t)M * N;
  concrete_backward_kernel<<<ew_blocks(n, h->num_sms), 256, 0, (cudaStream_t)stream>>>(g, p, a, n, 1.0f / temperature, dz);
  return check_launch(h, "concrete_backward");
}

extern "C" int db_sigmoid_backward(db_handle_t h, db_stream_t stream, const float* g, const float* p, int M, int N, float* dz) {
  DB_REQUIRE(h, g && p && dz && M > 0 && N > 0, "null operand or empty shape");
  const size_t n = (size_t)M * N;
  sigmoid_backward_kernel<<<ew_blocks(n, h->num_sms), 256, 0, (cudaStream_t)stream>>>(g, p, n, dz);
  return check_launch(h, "sigmoid_backward");
}

extern "C" int db_cross_entropy_grad(db_handle_t h, db_stream_t stream, const float* probs, const float* targets, int M, int N,
                                     float scale, float* out) {
  DB_REQUIRE(h, probs && targets && out && M > 0 && N > 0, "null operand or empty shape");
  const size_t n = (size_t)M * N;
  cross_entropy_grad_kernel<<<ew_blocks(n, h->num_sms), 256, 0, (cudaStream_t)stream>>>(probs, targets, n, scale, out);
  return check_launch(h, "cross_entropy_grad");
}

// ---- SBM_Lower (sbm_lower.py:59-117): four overlapping patches of the side x side image share one weight matrix ----
// Patch p reads the P x P window at (row offset, column offset) = (0,0), (0,pad), (pad,pad), (pad,0) with
// P = side/2 + overlap/2, pad = P - overlap (sbm_lower.py:63-67) and feeds hidden units [p H4, (p+1) H4). The layer
// is an RBM whose dense weight W_eff[V, 4 H4] holds the shared matrix scattered four times and zeros elsewhere, so the
// CD kernels run unchanged; the gradient of the shared matrix is the sum of the four scattered blocks.
namespace db { namespace {
__device__ __forceinline__ void sbm_offsets(int p, int pad, int& roff, int& coff) {
  roff = (p >= 2) ? pad : 0;
  coff = (p == 1 || p == 2) ? pad : 0;
}
__global__ void sbm_expand_kernel(const float* __restrict__ Ws, int side, int P, int pad, int H4, float* __restrict__ Weff) {
  const int H = 4 * H4;
  const size_t total = (size_t)side * side * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i / H), col = (int)(i % H), p = col / H4, jj = col % H4;
    int roff, coff;
    sbm_offsets(p, pad, roff, coff);
    const int r = v / side - roff, c = v % side - coff;
    Weff[i] = (r >= 0 && r < P && c >= 0 && c < P) ? __ldg(Ws + (size_t)(r * P + c) * H4 + jj) : 0.0f;
  }
}
__global__ void sbm_gather_kernel(const float* __restrict__ gWeff, int side, int P, int pad, int H4, float* __restrict__ gWs) {
  const int H = 4 * H4;
  const size_t total = (size_t)P * P * H4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i / H4), jj = (int)(i % H4), r = pix / P, c = pix % P;
    float s = 0.0f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int roff, coff;
      sbm_offsets(p, pad, roff, coff);
      s += __ldg(gWeff + (size_t)((r + roff) * side + c + coff) * H + p * H4 + jj);
    }
    gWs[i] = s;
  }
}
inline bool sbm_shape_ok(int side, int overlap, int num_h) {
  return side > 0 && num_h > 0 && overlap >= 0 && side % 2 == 0 && overlap % 2 == 0 && num_h % 4 == 0 && overlap <= side;
}
} }  // namespace db::(anonymous)

extern "C" int db_sbm_expand(db_handle_t h, db_stream_t stream, const float* W_shared, int side, int side_overlap, int num_h,
                             float* W_eff) {
  DB_REQUIRE(h, W_shared && W_eff, "null operand");
  DB_REQUIRE(h, sbm_shape_ok(side, side_overlap, num_h), "side and side_overlap must be even, num_h a multiple of 4");   // sbm_lower.py:24-26
  const int P = side / 2 + side_overlap / 2;
  const size_t total = (size_t)side * side * num_h;
  const size_t blocks = (total + 255) / 256, cap = (size_t)h->num_sms * 8;
  sbm_expand_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(W_shared, side, P, P - side_overlap,
                                                                                        num_h / 4, W_eff);
  return check_launch(h, "sbm_expand");
}

extern "C" int db_sbm_gather(db_handle_t h, db_stream_t stream, const float* gW_eff, int side, int side_overlap, int num_h,
                             float* gW_shared) {
  DB_REQUIRE(h, gW_eff && gW_shared, "null operand");
  DB_REQUIRE(h, sbm_shape_ok(side, side_overlap, num_h), "side and side_overlap must be even, num_h a multiple of 4");
  const int P = side / 2 + side_overlap / 2;
  const size_t total = (size_t)P * P * (num_h / 4);
  const size_t blocks = (total + 255) / 256, cap = (size_t)h->num_sms * 8;
  sbm_gather_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(gW_eff, side, P, P - side_overlap,
                                                                                        num_h / 4, gW_shared);
  return check_launch(h, "sbm_gather");
}

extern "C" int db_colsum(db_handle_t h, db_stream_t stream, const float* A, int M, int N, float scale, float beta, float* out) {
  DB_REQUIRE(h, A && out && M > 0 && N > 0, "null operand or empty shape");
  const int col_tiles = (N + 31) / 32;
  int slices = 1;
  if (M >= 2048 && col_tiles < h->num_sms) {
    slices = (2 * h->num_sms + col_tiles - 1) / col_tiles;
    if (slices > M / 256) slices = M / 256;
    if (slices < 1) slices = 1;
  }
  const int rows_per_slice = (M + slices - 1) / slices;
  if (slices > 1) colsum_scale_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(out, N, beta);
  colsum_kernel<<<dim3(col_tiles, slices), 256, 0, (cudaStream_t)stream>>>(A, M, N, scale, beta, out, rows_per_slice);
  return check_launch(h, "colsum");
}

extern "C" int db_gpdbn_top_forward_up(db_handle_t h, db_stream_t stream, const float* x, const float* pred_var,
                                       const float* alpha_h, const float* c, const float* eps1, int N, int D, float* H1,
                                       float* H1_mean, float* H2) {
  DB_REQUIRE(h, x && pred_var && alpha_h && c && eps1 && H1 && H1_mean && H2 && N > 0 && D > 0, "null operand or empty shape");
  top_up_kernel<<<D, 256, 0, (cudaStream_t)stream>>>(x, pred_var, alpha_h, c, eps1, N, D, H1, H1_mean, H2);
  return check_launch(h, "gpdbn_top_forward_up");
}

extern "C" int db_gpdbn_top_forward_mid(db_handle_t h, db_stream_t stream, const float* Hm, const float* pred_var,
                                        const float* alpha_h, const float* H1_mean, const float* eps2, int N, int D, float* H,
                                        float* y) {
  DB_REQUIRE(h, Hm && pred_var && alpha_h && H1_mean && eps2 && y && N > 0 && D > 0, "null operand or empty shape");
  top_mid_kernel<<<ew_blocks((size_t)N * D, h->num_sms), 256, 0, (cudaStream_t)stream>>>(Hm, pred_var, alpha_h, H1_mean, eps2, N,
                                                                                          D, H, y);
  return check_launch(h, "gpdbn_top_forward_mid");
}

extern "C" int db_gpdbn_top_backward1(db_handle_t h, db_stream_t stream, const float* g_y, const float* y, const float* Hm,
                                      const float* pred_var, const float* alpha_h, const float* eps2, int N, int D, float* g_Hm,
                                      float* gsig, float* ga, float* gmean) {
  DB_REQUIRE(h, g_y && y && Hm && pred_var && alpha_h && eps2 && g_Hm && gsig && ga && gmean && N > 0 && D > 0, "null operand");
  top_back1_kernel<<<D, 256, 0, (cudaStream_t)stream>>>(g_y, y, Hm, pred_var, alpha_h, eps2, N, D, g_Hm, gsig, ga, gmean);
  return check_launch(h, "gpdbn_top_backward1");
}

extern "C" int db_gpdbn_top_backward2(db_handle_t h, db_stream_t stream, const float* g_H2, const float* H2, const float* x,
                                      const float* eps1, const float* pred_var, const float* alpha_h, const float* opt_alpha_h,
                                      const float* ga, const float* gmean, int N, int D, float* g_x, float* gsig, float* g_c,
                                      float* g_alpha_scratch, float* g_opt_alpha_h, float* g_pv) {
  DB_REQUIRE(h, g_H2 && H2 && x && eps1 && pred_var && alpha_h && opt_alpha_h && ga && gmean && g_x && gsig && g_c &&
                    g_alpha_scratch && g_opt_alpha_h && g_pv && N > 0 && D > 0, "null operand");
  cudaStream_t st = (cudaStream_t)stream;
  top_back2_kernel<<<D, 256, 0, st>>>(g_H2, H2, x, eps1, pred_var, alpha_h, ga, gmean, N, D, g_x, gsig, g_c, g_alpha_scratch);
  top_back3_kernel<<<(N + 7) / 8, 256, 0, st>>>(gsig, pred_var, alpha_h, g_alpha_scratch, opt_alpha_h, N, D, g_pv, g_opt_alpha_h);
  return check_launch(h, "gpdbn_top_backward2");
}

// ---- test-time latent search (gplvm.py:470-541) ---------------------------------------------------------
// One CTA per restart r: pm[r,:] -> min/max normalisation (gplvm.py:293-296), cross entropy against the test
// point (dist.py:24-35, mean over the single row) scaled by 1/Dim, plus log_var_scaling * log(pred_var[r]);
// writes loss[r], the normalised and the [0,1]-clipped prediction, and d loss / d pm[r,:] and d loss / d pred_var[r].
namespace db { namespace {
__global__ void __launch_bounds__(256) testgen_rowloss_kernel(const float* __restrict__ pm, const float* __restrict__ target,
                                                             const float* __restrict__ pv, int R, int D, float log_var_scaling,
                                                             float* __restrict__ loss, float* __restrict__ pm_norm,
                                                             float* __restrict__ pm_clip, float* __restrict__ g_pm,
                                                             float* __restrict__ g_pv) {
  __shared__ float s_val[8]; __shared__ int s_idx[8];
  __shared__ float s_mn, s_mx; __shared__ int s_imn, s_imx;
  __shared__ float red[8];
  const int r = blockIdx.x;
  const float* row = pm + (size_t)r * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // min / max with the first index attaining them (tf.reduce_min/max route the gradient there)
  float mn = INFINITY, mx = -INFINITY; int imn = 0x7fffffff, imx = 0x7fffffff;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float v = __ldg(row + d);
    if (v < mn) { mn = v; imn = d; }
    if (v > mx) { mx = v; imx = d; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float omn = __shfl_xor_sync(0xffffffffu, mn, o); const int oi = __shfl_xor_sync(0xffffffffu, imn, o);
    if (omn < mn || (omn == mn && oi < imn)) { mn = omn; imn = oi; }
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o); const int oj = __shfl_xor_sync(0xffffffffu, imx, o);
    if (omx > mx || (omx == mx && oj < imx)) { mx = omx; imx = oj; }
  }
  if (lane == 0) { s_val[warp] = mn; s_idx[warp] = imn; }
  __syncthreads();
  if (threadIdx.x == 0) { float b = s_val[0]; int bi = s_idx[0]; for (int w = 1; w < 8; ++w) if (s_val[w] < b || (s_val[w] == b && s_idx[w] < bi)) { b = s_val[w]; bi = s_idx[w]; } s_mn = b; s_imn = bi; }
  __syncthreads();
  if (lane == 0) { s_val[warp] = mx; s_idx[warp] = imx; }
  __syncthreads();
  if (threadIdx.x == 0) { float b = s_val[0]; int bi = s_idx[0]; for (int w = 1; w < 8; ++w) if (s_val[w] > b || (s_val[w] == b && s_idx[w] < bi)) { b = s_val[w]; bi = s_idx[w]; } s_mx = b; s_imx = bi; }
  __syncthreads();
  mn = s_mn; mx = s_mx;
  const float rng = mx - mn, inv_rng = 1.0f / rng, inv_dim = 1.0f / (float)D;
  const float e0 = 1.0e-6f;
  float ce = 0.0f, sum_g = 0.0f, sum_gp = 0.0f;     // sum_d g_n, sum_d g_n * pn   (g_n = d loss / d pm_norm)
  for (int d = threadIdx.x; d < D; d += 256) {
    const float v = __ldg(row + d), t = __ldg(target + d);
    const float pn = (v - mn) * inv_rng;
    const float pc = fminf(fmaxf(pn, e0), 1.0f - e0);
    const float x = logf(pc / (1.0f - pc));
    ce += fmaxf(x, 0.0f) - x * t + log1pf(__expf(-fabsf(x)));
    const bool inside = pn >= e0 && pn <= 1.0f - e0;
    const float gn = inside ? inv_dim * (-t / pc + (1.0f - t) / (1.0f - pc)) : 0.0f;
    if (pm_norm) pm_norm[(size_t)r * D + d] = pn;
    if (pm_clip) pm_clip[(size_t)r * D + d] = fminf(fmaxf(v, 0.0f), 1.0f);
    g_pm[(size_t)r * D + d] = gn * inv_rng;
    sum_g += gn; sum_gp = fmaf(gn, pn, sum_gp);
  }
  auto bsum = [&](float v) { v = warp_sum(v); __syncthreads(); if (lane == 0) red[warp] = v; __syncthreads(); float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; return t; };
  ce = bsum(ce); sum_g = bsum(sum_g); sum_gp = bsum(sum_gp);
  __syncthreads();
  if (threadIdx.x == 0) {
    // pn = (v - mn) / (mx - mn): d/dmn = (pn - 1)/rng, d/dmx = -pn/rng, routed to the arg-min / arg-max entries
    g_pm[(size_t)r * D + s_imn] += (sum_gp - sum_g) * inv_rng;
    g_pm[(size_t)r * D + s_imx] += -sum_gp * inv_rng;
    const float v = __ldg(pv + r);
    loss[r] = ce * inv_dim + log_var_scaling * logf(v);
    g_pv[r] = log_var_scaling / v;
  }
}
// pred_var[r] = alpha - sum_i Kxs[i,r] Cv[i,r] ; flags negative entries
__global__ void testgen_pred_var_kernel(const float* __restrict__ Kxs, const float* __restrict__ Cv, int N, int R,
                                        const float* __restrict__ hyp, float* __restrict__ pv, int* __restrict__ info) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float s = 0.0f;
  for (int i = 0; i < N; ++i) s = fmaf(__ldg(Kxs + (size_t)i * R + r), __ldg(Cv + (size_t)i * R + r), s);
  const float v = __ldg(hyp) - s;
  pv[r] = v;
  if (v < 0.0f && info) atomicCAS(info, 0, r + 1);
}
// g_x[r,q] = -(1/gamma^2) sum_i (gK[i,r] - 2 g_pv[r] Cv[i,r]) Kxs[i,r] (x[r,q] - X[i,q])      (one warp per restart)
__global__ void testgen_grad_x_kernel(const float* __restrict__ gK, const float* __restrict__ Cv, const float* __restrict__ Kxs,
                                      const float* __restrict__ g_pv, const float* __restrict__ X, const float* __restrict__ x,
                                      int N, int R, int Q, const float* __restrict__ hyp, float* __restrict__ g_x) {
  const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  if (r >= R) return;
  const float inv_g2 = 1.0f / (__ldg(hyp + 1) * __ldg(hyp + 1)), gp = __ldg(g_pv + r);
  for (int q = 0; q < Q; ++q) {
    const float xq = __ldg(x + (size_t)r * Q + q);
    float s = 0.0f;
    for (int i = lane; i < N; i += 32) {
      const size_t idx = (size_t)i * R + r;
      const float w = (__ldg(gK + idx) - 2.0f * gp * __ldg(Cv + idx)) * __ldg(Kxs + idx);
      s = fmaf(w, xq - __ldg(X + (size_t)i * Q + q), s);
    }
    s = warp_sum(s);
    if (lane == 0) g_x[(size_t)r * Q + q] = -inv_g2 * s;
  }
}
}}  // namespace db::<anon>

extern "C" int db_testgen_rowloss(db_handle_t h, db_stream_t stream, const float* pm, const float* target, const float* pred_var,
                                  int R, int D, float log_var_scaling, float* loss, float* pm_norm, float* pm_clip, float* g_pm,
                                  float* g_pv) {
  DB_REQUIRE(h, pm && target && pred_var && loss && g_pm && g_pv && R > 0 && D > 1, "null operand or bad shape");
  db::testgen_rowloss_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(pm, target, pred_var, R, D, log_var_scaling, loss, pm_norm, pm_clip,
                                                                   g_pm, g_pv);
  return check_launch(h, "testgen_rowloss");
}

extern "C" int db_testgen_pred_var(db_handle_t h, db_stream_t stream, const float* Kxs, const float* Cv, int N, int R,
                                   const float* hyp, float* pred_var, int* info) {
  DB_REQUIRE(h, Kxs && Cv && hyp && pred_var && N > 0 && R > 0, "null operand or bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (info) cudaMemsetAsync(info, 0, sizeof(int), st);
  db::testgen_pred_var_kernel<<<(R + 127) / 128, 128, 0, st>>>(Kxs, Cv, N, R, hyp, pred_var, info);
  return check_launch(h, "testgen_pred_var");
}

extern "C" int db_testgen_grad_x(db_handle_t h, db_stream_t stream, const float* gK, const float* Cv, const float* Kxs,
                                 const float* g_pv, const float* X, const float* x, int N, int R, int Q, const float* hyp,
                                 float* g_x) {
  DB_REQUIRE(h, gK && Cv && Kxs && g_pv && X && x && hyp && g_x && N > 0 && R > 0 && Q > 0, "null operand or bad shape");
  db::testgen_grad_x_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(gK, Cv, Kxs, g_pv, X, x, N, R, Q, hyp, g_x);
  return check_launch(h, "testgen_grad_x");
}

// ---- GPRBM.test_generalisation (gprbm.py:468-562): batched over R restarts x S samples ----------------------
namespace db { namespace {
// per restart r: m = mean over its S down-propagated samples, loss[r] = CE(m, target)/Dim + lvs * log(pv[r])
// (gprbm.py:497-499), g_Vs[(r,s),:] = d loss[r] / d sample, model_point[r,:] = m
__global__ void __launch_bounds__(256) testgen_meanloss_kernel(const float* __restrict__ Vs, const float* __restrict__ target,
                                                              const float* __restrict__ pv, int R, int S, int Dim,
                                                              float log_var_scaling, float* __restrict__ loss,
                                                              float* __restrict__ model_point, float* __restrict__ g_Vs) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const float e0 = 1.0e-6f, inv_dim = 1.0f / (float)Dim, inv_s = 1.0f / (float)S;
  float ce = 0.0f;
  for (int d = threadIdx.x; d < Dim; d += 256) {
    float m = 0.0f;
    for (int s = 0; s < S; ++s) m += __ldg(Vs + ((size_t)r * S + s) * Dim + d);
    m *= inv_s;
    const float t = __ldg(target + d);
    const float pc = fminf(fmaxf(m, e0), 1.0f - e0);
    const float x = logf(pc / (1.0f - pc));
    ce += fmaxf(x, 0.0f) - x * t + log1pf(__expf(-fabsf(x)));
    const bool inside = m >= e0 && m <= 1.0f - e0;
    const float g = inside ? inv_dim * inv_s * (-t / pc + (1.0f - t) / (1.0f - pc)) : 0.0f;
    if (model_point) model_point[(size_t)r * Dim + d] = m;
    for (int s = 0; s < S; ++s) g_Vs[((size_t)r * S + s) * Dim + d] = g;
  }
  ce = warp_sum(ce);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ce;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss[r] = t * inv_dim + log_var_scaling * logf(__ldg(pv + r));
  }
}
// per restart r: g_Hm[r,d] = sum_s gHm_rows[(r,s),d] ; g_pv[r] = sum_{s,d} gsig_rows[(r,s),d] alpha_h[d] / (2 sqrt(pv[r])) + lvs / pv[r]
__global__ void __launch_bounds__(256) testgen_reduce_samples_kernel(const float* __restrict__ gHm_rows, const float* __restrict__ gsig_rows,
                                                                    const float* __restrict__ alpha_h, const float* __restrict__ pv,
                                                                    int R, int S, int D, float log_var_scaling,
                                                                    float* __restrict__ g_Hm, float* __restrict__ g_pv) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  float acc = 0.0f;
  for (int d = threadIdx.x; d < D; d += 256) {
    float sh = 0.0f, ss = 0.0f;
    for (int s = 0; s < S; ++s) {
      const size_t i = ((size_t)r * S + s) * D + d;
      sh += __ldg(gHm_rows + i); ss += __ldg(gsig_rows + i);
    }
    g_Hm[(size_t)r * D + d] = sh;
    acc = fmaf(ss, __ldg(alpha_h + d), acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w];
    const float v = __ldg(pv + r);
    g_pv[r] = t / (2.0f * sqrtf(v)) + log_var_scaling / v;
  }
}
}}  // namespace db::<anon>

// ---- GPRBM.test_held_out_data (gprbm.py:440-466) and GPRBM.interp_test (gprbm.py:614-689) ---------------------
namespace db { namespace {
// loss[r] = mean_m sum_d (T[m,d] - Vs[r,d])^2 — dist.mse (dist.py:5-21) with the [M,Dim] test set as `predictions` and the
// single down-propagated row of restart r broadcast as `targets`; g_Vs[r,:] = d loss[r] / d Vs[r,:]
__global__ void __launch_bounds__(256) heldout_mse_kernel(const float* __restrict__ Vs, const float* __restrict__ T, int R, int M,
                                                         int Dim, float* __restrict__ loss, float* __restrict__ g_Vs) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const float inv_m = 1.0f / (float)M;
  float acc = 0.0f;
  for (int d = threadIdx.x; d < Dim; d += 256) {
    const float v = __ldg(Vs + (size_t)r * Dim + d);
    float sq = 0.0f, gs = 0.0f;
    for (int m = 0; m < M; ++m) {
      const float diff = __ldg(T + (size_t)m * Dim + d) - v;
      sq = fmaf(diff, diff, sq);
      gs += diff;
    }
    acc += sq;
    g_Vs[(size_t)r * Dim + d] = -2.0f * gs * inv_m;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss[r] = t * inv_m;
  }
}
// objective = mean over the P+1 segments start -> Xi[0] -> ... -> Xi[P-1] -> end of |segment|^2 + lambda * mean_p log pv[p]
// (gprbm.py:651-660); g_x [P,Q] = gradient of the segment term, g_pv [P] = lambda / (P pv[p]). One block (P is tens of points).
__global__ void __launch_bounds__(256) interp_objective_kernel(const float* __restrict__ Xi, const float* __restrict__ xs,
                                                              const float* __restrict__ xe, const float* __restrict__ pv, int P, int Q,
                                                              float lambda, float* __restrict__ objective, float* __restrict__ g_x,
                                                              float* __restrict__ g_pv) {
  __shared__ float red[8];
  const float inv_seg = 1.0f / (float)(P + 1), inv_p = 1.0f / (float)P;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < (P + 1) * Q; i += 256) {           // segment s = i / Q runs from point s-1 to point s
    const int s = i / Q, q = i - s * Q;
    const float a = s == 0 ? __ldg(xs + q) : __ldg(Xi + (size_t)(s - 1) * Q + q);
    const float b = s == P ? __ldg(xe + q) : __ldg(Xi + (size_t)s * Q + q);
    acc = fmaf(b - a, b - a, acc);
  }
  acc *= inv_seg;
  for (int p = threadIdx.x; p < P; p += 256) {
    const float v = __ldg(pv + p);
    acc = fmaf(lambda * inv_p, logf(v), acc);
    g_pv[p] = lambda * inv_p / v;
  }
  for (int i = threadIdx.x; i < P * Q; i += 256) {
    const int p = i / Q, q = i - p * Q;
    const float x = __ldg(Xi + i);
    const float prev = p == 0 ? __ldg(xs + q) : __ldg(Xi + i - Q);
    const float next = p == P - 1 ? __ldg(xe + q) : __ldg(Xi + i + Q);
    g_x[i] = 2.0f * inv_seg * (2.0f * x - prev - next);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w];
    *objective = t;
  }
}
}}  // namespace db::<anon>

extern "C" int db_heldout_mse(db_handle_t h, db_stream_t stream, const float* Vs, const float* T, int R, int M, int Dim,
                              float* loss, float* g_Vs) {
  DB_REQUIRE(h, Vs && T && loss && g_Vs && R > 0 && M > 0 && Dim > 0, "null operand or bad shape");
  db::heldout_mse_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(Vs, T, R, M, Dim, loss, g_Vs);
  return check_launch(h, "heldout_mse");
}

extern "C" int db_interp_objective(db_handle_t h, db_stream_t stream, const float* Xi, const float* x_start, const float* x_end,
                                   const float* pred_var, int P, int Q, float lambda, float* objective, float* g_x, float* g_pv) {
  DB_REQUIRE(h, Xi && x_start && x_end && pred_var && objective && g_x && g_pv && P > 0 && Q > 0, "null operand or bad shape");
  db::interp_objective_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(Xi, x_start, x_end, pred_var, P, Q, lambda, objective, g_x, g_pv);
  return check_launch(h, "interp_objective");
}

extern "C" int db_testgen_meanloss(db_handle_t h, db_stream_t stream, const float* Vs, const float* target, const float* pred_var,
                                   int R, int S, int Dim, float log_var_scaling, float* loss, float* model_point, float* g_Vs) {
  DB_REQUIRE(h, Vs && target && pred_var && loss && g_Vs && R > 0 && S > 0 && Dim > 0, "null operand or bad shape");
  db::testgen_meanloss_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(Vs, target, pred_var, R, S, Dim, log_var_scaling, loss, model_point, g_Vs);
  return check_launch(h, "testgen_meanloss");
}

extern "C" int db_testgen_reduce_samples(db_handle_t h, db_stream_t stream, const float* gHm_rows, const float* gsig_rows,
                                         const float* alpha_h, const float* pred_var, int R, int S, int D, float log_var_scaling,
                                         float* g_Hm, float* g_pv) {
  DB_REQUIRE(h, gHm_rows && gsig_rows && alpha_h && pred_var && g_Hm && g_pv && R > 0 && S > 0 && D > 0, "null operand or bad shape");
  db::testgen_reduce_samples_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(gHm_rows, gsig_rows, alpha_h, pred_var, R, S, D,
                                                                         log_var_scaling, g_Hm, g_pv);
  return check_launch(h, "testgen_reduce_samples");
}
