// General fp32 SIMT GEMM  C[M,N] = alpha * op(A) * op(B) + beta * C  (row-major, leading dimensions in
// elements). Used for the backward products of the GPDBN stack (dW = A^T dZ, dA = dZ W^T: what TF autodiff
// derives from tf.matmul, rbm.py:129,144) and the K^-1 products of the GP block. Built on the functor-loader
// tile core of simt_gemm.cuh.
#pragma once
#include "context.h"
#include "simt_gemm.cuh"

namespace db {

template <bool TRANS>
struct GemmLoadA {   // a(m,k) of op(A): A is [M,K] (TRANS=0) or [K,M] (TRANS=1)
  static constexpr bool k_contig = !TRANS;
  const float* A; int lda; int M, K;
  __device__ __forceinline__ float operator()(int m, int k) const {
    if (m >= M || k >= K) return 0.0f;
    return TRANS ? __ldg(A + (size_t)k * lda + m) : __ldg(A + (size_t)m * lda + k);
  }
  __host__ __device__ bool vec_ok() const { return lda % 4 == 0 && aligned16(A); }
  __device__ __forceinline__ float4 load4(int m, int k) const {     // TRANS: m .. m+3 of row k; else k .. k+3 of row m
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (TRANS) {
      if (m + 3 >= M) return make_float4((*this)(m, k), (*this)(m + 1, k), (*this)(m + 2, k), (*this)(m + 3, k));
      return __ldg(reinterpret_cast<const float4*>(A + (size_t)k * lda + m));
    }
    if (k + 3 >= K) return make_float4((*this)(m, k), (*this)(m, k + 1), (*this)(m, k + 2), (*this)(m, k + 3));
    return __ldg(reinterpret_cast<const float4*>(A + (size_t)m * lda + k));
  }
};
template <bool TRANS>
struct GemmLoadB {   // b(k,n) of op(B): B is [K,N] (TRANS=0) or [N,K] (TRANS=1)
  static constexpr bool k_contig = TRANS;
  const float* B; int ldb; int K, N;
  __device__ __forceinline__ float operator()(int k, int n) const {
    if (k >= K || n >= N) return 0.0f;
    return TRANS ? __ldg(B + (size_t)n * ldb + k) : __ldg(B + (size_t)k * ldb + n);
  }
  __host__ __device__ bool vec_ok() const { return ldb % 4 == 0 && aligned16(B); }
  __device__ __forceinline__ float4 load4(int k, int n) const {     // TRANS: k .. k+3 of column n; else n .. n+3 of row k
    if (k >= K || n >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (TRANS) {
      if (k + 3 >= K) return make_float4((*this)(k, n), (*this)(k + 1, n), (*this)(k + 2, n), (*this)(k + 3, n));
      return __ldg(reinterpret_cast<const float4*>(B + (size_t)n * ldb + k));
    }
    if (n + 3 >= N) return make_float4((*this)(k, n), (*this)(k, n + 1), (*this)(k, n + 2), (*this)(k, n + 3));
    return __ldg(reinterpret_cast<const float4*>(B + (size_t)k * ldb + n));
  }
};

// gridDim.z > 1: split-K. Slice z covers k in [z * k_per, (z + 1) * k_per) and adds alpha * partial into C with atomics;
// the caller has already replaced C by beta * C (gemm_scale_c_kernel). The order of the fp32 additions is then not fixed.
template <class T, bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                       float* __restrict__ C, int ldc, int M, int N, int K, float alpha,
                                                       float beta, int k_per) {
  __shared__ typename T::Smem sm;
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  GemmLoadA<TA> a{A, lda, M, K};
  GemmLoadB<TB> b{B, ldb, K, N};
  float acc[T::TM][T::TN];
  const bool split = gridDim.z > 1;
  const int k_begin = split ? blockIdx.z * k_per : 0;
  const int k_end = split ? min(K, k_begin + k_per) : K;
  simt_gemm_tile<T>(sm, a, b, m0, n0, k_begin, k_end, acc);
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < T::TN; ++j) {
      const int n = simt_col<T>(n0, j);
      if (n >= N) continue;
      float* c = C + (size_t)m * ldc + n;
      if (split) atomicAdd(c, alpha * acc[i][j]);
      else *c = beta != 0.0f ? fmaf(beta, *c, alpha * acc[i][j]) : alpha * acc[i][j];
    }
  }
}
// 64 x 64 tiles, K split over a thread-block cluster (1,1,splits) and reduced through distributed shared memory in rank
// order (simt_gemm.cuh): deterministic, beta applied in the epilogue. The backward products of the horses-sized stacks have
// one to eight output tiles and K = 100-328 — every k-block is a dependent round trip to memory, 17-34 us per launch.
template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(256) gemm_f32_cluster_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                               int ldb, float* __restrict__ C, int ldc, int M, int N, int K,
                                                               float alpha, float beta, int splits) {
  using T = TileSmall;
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  GemmLoadA<TA> a{A, lda, M, K};
  GemmLoadB<TB> b{B, ldb, K, N};
  float acc[T::TM][T::TN];
  int kb, ke;
  simt_split_range<T>(K, splits, kb, ke);
  if (VEC) simt_gemm_tile_vec<T>(sm, a, b, m0, n0, kb, ke, acc);
  else simt_gemm_tile<T>(sm, a, b, m0, n0, kb, ke, acc);
  simt_cluster_reduce<T>(red, acc, splits, [&](int r, int c, float x) {
    const int m = m0 + r, n = n0 + c;
    if (m < M && n < N) {
      float* cp = C + (size_t)m * ldc + n;
      *cp = beta != 0.0f ? fmaf(beta, *cp, alpha * x) : alpha * x;
    }
  });
}

static __global__ void gemm_scale_c_kernel(float* __restrict__ C, int ldc, int M, int N, float beta) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float* c = C + (i / N) * (size_t)ldc + (i % N);
    *c = beta != 0.0f ? beta * *c : 0.0f;
  }
}

template <class T, bool TA, bool TB>
inline void launch_gemm_f32(cudaStream_t st, int num_sms, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M,
                            int N, int K, float alpha, float beta) {
  dim3 grid((N + T::BN - 1) / T::BN, (M + T::BM - 1) / T::BM);
  // few output tiles and a long K (dW = X^T dZ over the batch, K^-1 H2 at N = 5000): split K so that every SM has work
  const int tiles = (int)(grid.x * grid.y);
  int splits = 1;
  if (2 * tiles <= num_sms && K >= 1024) {
    splits = (2 * num_sms + tiles - 1) / tiles;
    const int max_splits = K / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int k_per = K;
  if (splits > 1) {
    k_per = ((K + splits - 1) / splits + T::BK - 1) / T::BK * T::BK;
    splits = (K + k_per - 1) / k_per;
  }
  if (splits > 1) {
    const size_t total = (size_t)M * N;
    const int blocks = (int)((total + 255) / 256 < (size_t)num_sms * 8 ? (total + 255) / 256 : (size_t)num_sms * 8);
    gemm_scale_c_kernel<<<blocks, 256, 0, st>>>(C, ldc, M, N, beta);
    grid.z = splits;
  } else if (T::BM == TileSmall::BM && T::BN == TileSmall::BN && T::BK == TileSmall::BK) {
    const int cs = simt_splits(tiles, K, num_sms, 32);
    const bool vec = GemmLoadA<TA>{A, lda, M, K}.vec_ok() && GemmLoadB<TB>{B, ldb, K, N}.vec_ok();
    grid.z = cs;
    simt_launch_split(vec ? gemm_f32_cluster_kernel<TA, TB, true> : gemm_f32_cluster_kernel<TA, TB, false>, grid, 256, st, A, lda, B,
                      ldb, C, ldc, M, N, K, alpha, beta, cs);
    return;
  }
  gemm_f32_kernel<T, TA, TB><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, alpha, beta, k_per);
}

// large tiles only when both output dimensions can fill them and there are enough tiles for 148 SMs
inline int run_gemm_f32(db_handle_t h, cudaStream_t st, int transA, int transB, int M, int N, int K, float alpha,
                        const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc) {
  const bool large = M >= 512 && N >= 512 && ((size_t)M * N) / (128 * 128) >= (size_t)h->num_sms;
#define DB_GEMM_CASE(TA, TB)                                                                              \
  do {                                                                                                     \
    if (large) launch_gemm_f32<TileLarge, TA, TB>(st, h->num_sms, A, lda, B, ldb, C, ldc, M, N, K, alpha, beta);       \
    else launch_gemm_f32<TileSmall, TA, TB>(st, h->num_sms, A, lda, B, ldb, C, ldc, M, N, K, alpha, beta);             \
  } while (0)
  if (!transA && !transB) DB_GEMM_CASE(false, false);
  else if (transA && !transB) DB_GEMM_CASE(true, false);
  else if (!transA && transB) DB_GEMM_CASE(false, true);
  else DB_GEMM_CASE(true, true);
#undef DB_GEMM_CASE
  return check_launch(h, "gemm_f32");
}

}  // namespace db
