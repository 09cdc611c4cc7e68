// fp32 SIMT GEMM tile core shared by the small-layer RBM kernels (layer_simt.cu) and the GP chain
// (gp.cu). 256 threads compute a BM x BN tile, each thread a TM x TN micro-tile, operands staged
// through padded shared memory with register double-buffering of the global loads.
//
// Operands are abstracted as loaders  a(m, k) / b(k, n)  that return 0 outside the matrix, so the
// callers can fuse prologues (bgrbm.py:126 h/sigma_h), transposed or triangular views, and on-the-fly
// operands (e.g. the RBF Gram entries) without materialising them.
#pragma once
#include <cooperative_groups.h>
#include <utility>
#include "common.cuh"

namespace db {

template <int BM_, int BN_, int BK_, int TM_, int TN_>
struct SimtTile {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int THREADS = (BM / TM) * (BN / TN);
  static constexpr int PAD = 4;
  static constexpr int A_PER_THREAD = BM * BK / THREADS;
  static constexpr int B_PER_THREAD = BN * BK / THREADS;
  static_assert(THREADS == 256, "tile shapes are chosen for 256 threads");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 smem reads");
  struct Smem {
    float a[BK][BM + PAD];
    float b[BK][BN + PAD];
  };
};

// acc is zero-initialised unless ZERO_INIT=false (then the products are added to its contents).
// ALoad: float operator()(int m, int k) const;  static constexpr bool k_contig (global memory is
// contiguous along k for fixed m). BLoad: float operator()(int k, int n) const; static constexpr bool
// k_contig likewise.
template <class T, class ALoad, class BLoad, bool ZERO_INIT = true>
__device__ __forceinline__ void simt_gemm_tile(typename T::Smem& sm, const ALoad& a, const BLoad& b, int m0, int n0,
                                               int k_begin, int k_end, float (&acc)[T::TM][T::TN]) {
  const int tid = threadIdx.x;
  const int tx = tid % (T::BN / T::TN);
  const int ty = tid / (T::BN / T::TN);
  if (ZERO_INIT) {
#pragma unroll
    for (int i = 0; i < T::TM; ++i)
#pragma unroll
      for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.0f;
  }

  float ra[T::A_PER_THREAD], rb[T::B_PER_THREAD];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < T::A_PER_THREAD; ++i) {
      const int idx = tid + i * T::THREADS;
      int mm, kk;
      if (ALoad::k_contig) { kk = idx % T::BK; mm = idx / T::BK; } else { mm = idx % T::BM; kk = idx / T::BM; }
      const int k = k0 + kk;
      ra[i] = (k < k_end) ? a(m0 + mm, k) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < T::B_PER_THREAD; ++i) {
      const int idx = tid + i * T::THREADS;
      int nn, kk;
      if (BLoad::k_contig) { kk = idx % T::BK; nn = idx / T::BK; } else { nn = idx % T::BN; kk = idx / T::BN; }
      const int k = k0 + kk;
      rb[i] = (k < k_end) ? b(k, n0 + nn) : 0.0f;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < T::A_PER_THREAD; ++i) {
      const int idx = tid + i * T::THREADS;
      int mm, kk;
      if (ALoad::k_contig) { kk = idx % T::BK; mm = idx / T::BK; } else { mm = idx % T::BM; kk = idx / T::BM; }
      sm.a[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < T::B_PER_THREAD; ++i) {
      const int idx = tid + i * T::THREADS;
      int nn, kk;
      if (BLoad::k_contig) { kk = idx % T::BK; nn = idx / T::BK; } else { nn = idx % T::BN; kk = idx / T::BN; }
      sm.b[kk][nn] = rb[i];
    }
  };

  if (k_begin >= k_end) return;
  gload(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += T::BK) {
    __syncthreads();           // previous tile fully consumed
    sstore();
    __syncthreads();
    if (k0 + T::BK < k_end) gload(k0 + T::BK);   // overlaps with the FMAs below
#pragma unroll
    for (int kk = 0; kk < T::BK; ++kk) {
      float av[T::TM], bv[T::TN];
      // micro-tile rows: TM/4 groups of 4, strided by BM/(TM/4) to keep smem reads conflict-free
#pragma unroll
      for (int g = 0; g < T::TM / 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.a[kk][g * (T::BM / (T::TM / 4)) + ty * 4]);
        av[4 * g + 0] = v.x; av[4 * g + 1] = v.y; av[4 * g + 2] = v.z; av[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < T::TN / 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.b[kk][g * (T::BN / (T::TN / 4)) + tx * 4]);
        bv[4 * g + 0] = v.x; bv[4 * g + 1] = v.y; bv[4 * g + 2] = v.z; bv[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < T::TM; ++i)
#pragma unroll
        for (int j = 0; j < T::TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
}

// Vector-load form of simt_gemm_tile: each thread moves whole float4s of each operand per k-block (instead of
// bounds-checked scalars), which removes most of the non-FMA instructions of the launch-bound small-layer kernels
// (ncu: 50 % of the issued instructions were FMAs before). Same accumulation order as simt_gemm_tile: identical bits.
// Loaders provide  float4 load4(m, k)  /  float4 load4(k, n): four consecutive elements along the loader's contiguous
// direction (k if k_contig, else m / n), zero outside the matrix; the caller checks `vec_ok()` (alignment) first.
template <class T, class ALoad, class BLoad>
__device__ __forceinline__ void simt_gemm_tile_vec(typename T::Smem& sm, const ALoad& a, const BLoad& b, int m0, int n0,
                                                   int k_begin, int k_end, float (&acc)[T::TM][T::TN]) {
  static_assert(T::BK % 4 == 0 && (T::BM * T::BK / 4) % T::THREADS == 0 && (T::BN * T::BK / 4) % T::THREADS == 0,
                "whole float4s per thread");
  constexpr int AV = T::BM * T::BK / 4 / T::THREADS, BV = T::BN * T::BK / 4 / T::THREADS;
  const int tid = threadIdx.x;
  const int tx = tid % (T::BN / T::TN);
  const int ty = tid / (T::BN / T::TN);
#pragma unroll
  for (int i = 0; i < T::TM; ++i)
#pragma unroll
    for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.0f;
  if (k_begin >= k_end) return;
  // quad q of a k-contiguous operand: (row q / (BK/4), k 4 (q % (BK/4))); otherwise (k q / (rows/4), row 4 (q % (rows/4)))
  auto quad = [](bool k_contig, int rows, int q, int& row, int& k) {
    if (k_contig) { row = q / (T::BK / 4); k = (q % (T::BK / 4)) * 4; }
    else { k = q / (rows / 4); row = (q % (rows / 4)) * 4; }
  };
  auto clip = [&](float4 v, int k) {      // k-contiguous quads may straddle the end of this CTA's K range
    if (k + 3 >= k_end) {
      if (k + 0 >= k_end) v.x = 0.0f;
      if (k + 1 >= k_end) v.y = 0.0f;
      if (k + 2 >= k_end) v.z = 0.0f;
      v.w = 0.0f;
    }
    return v;
  };
  float4 ra[AV], rb[BV];
  auto gload = [&](int k0) {
#pragma unroll
    for (int v = 0; v < AV; ++v) {
      int row, kq; quad(ALoad::k_contig, T::BM, tid + v * T::THREADS, row, kq);
      const int k = k0 + kq;
      if (ALoad::k_contig) ra[v] = clip(a.load4(m0 + row, k), k);
      else ra[v] = k < k_end ? a.load4(m0 + row, k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int v = 0; v < BV; ++v) {
      int row, kq; quad(BLoad::k_contig, T::BN, tid + v * T::THREADS, row, kq);
      const int k = k0 + kq;
      if (BLoad::k_contig) rb[v] = clip(b.load4(k, n0 + row), k);
      else rb[v] = k < k_end ? b.load4(k, n0 + row) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int v = 0; v < AV; ++v) {
      int row, kq; quad(ALoad::k_contig, T::BM, tid + v * T::THREADS, row, kq);
      if (ALoad::k_contig) {
        sm.a[kq + 0][row] = ra[v].x; sm.a[kq + 1][row] = ra[v].y; sm.a[kq + 2][row] = ra[v].z; sm.a[kq + 3][row] = ra[v].w;
      } else {
        *reinterpret_cast<float4*>(&sm.a[kq][row]) = ra[v];
      }
    }
#pragma unroll
    for (int v = 0; v < BV; ++v) {
      int row, kq; quad(BLoad::k_contig, T::BN, tid + v * T::THREADS, row, kq);
      if (BLoad::k_contig) {
        sm.b[kq + 0][row] = rb[v].x; sm.b[kq + 1][row] = rb[v].y; sm.b[kq + 2][row] = rb[v].z; sm.b[kq + 3][row] = rb[v].w;
      } else {
        *reinterpret_cast<float4*>(&sm.b[kq][row]) = rb[v];
      }
    }
  };
  gload(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += T::BK) {
    __syncthreads();
    sstore();
    __syncthreads();
    if (k0 + T::BK < k_end) gload(k0 + T::BK);
#pragma unroll
    for (int kk = 0; kk < T::BK; ++kk) {
      float av[T::TM], bv[T::TN];
#pragma unroll
      for (int g = 0; g < T::TM / 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.a[kk][g * (T::BM / (T::TM / 4)) + ty * 4]);
        av[4 * g + 0] = v.x; av[4 * g + 1] = v.y; av[4 * g + 2] = v.z; av[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < T::TN / 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.b[kk][g * (T::BN / (T::TN / 4)) + tx * 4]);
        bv[4 * g + 0] = v.x; bv[4 * g + 1] = v.y; bv[4 * g + 2] = v.z; bv[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < T::TM; ++i)
#pragma unroll
        for (int j = 0; j < T::TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
}

inline __host__ __device__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// global (m, n) of micro-tile element (i, j) for the calling thread
template <class T>
__device__ __forceinline__ int simt_row(int m0, int i) {
  const int ty = threadIdx.x / (T::BN / T::TN);
  return m0 + (i / 4) * (T::BM / (T::TM / 4)) + ty * 4 + (i % 4);
}
template <class T>
__device__ __forceinline__ int simt_col(int n0, int j) {
  const int tx = threadIdx.x % (T::BN / T::TN);
  return n0 + (j / 4) * (T::BN / (T::TN / 4)) + tx * 4 + (j % 4);
}

// ---- cluster split-K ---------------------------------------------------------------------------------------
// The reference-sized layers (328 x 1024 x 500) give a 64 x 64 tile grid of ~50 CTAs for 148 SMs, each CTA walking
// the whole K serially: 4 % of the thread slots, 4-6 TFLOP/s. Such launches run as thread-block clusters (1,1,S):
// the S CTAs of a cluster share one output tile and split K; every CTA parks its micro-tile in its own shared memory,
// and after a cluster barrier CTA r sums rows [r BM/S, (r+1) BM/S) of the S partial tiles through distributed shared
// memory — in rank order, so the result does not depend on scheduling — and runs the epilogue on them.

// S in {1,2,4,8}: enough CTAs for two per SM, at least min_k (64) columns of K each
inline int simt_splits(long tiles, int K, int num_sms, int min_k = 64) {
  int s = 1;
  while (s < 8 && tiles * s < 2L * num_sms && K / (2 * s) >= min_k) s *= 2;
  return s;
}
// this CTA's K range [k_begin, k_end) of a K split S ways (multiples of BK)
template <class T>
__device__ __forceinline__ void simt_split_range(int K, int splits, int& k_begin, int& k_end) {
  const int rank = splits > 1 ? (int)cooperative_groups::this_cluster().block_rank() : 0;
  const int chunk = ((K + splits - 1) / splits + T::BK - 1) / T::BK * T::BK;
  k_begin = min(K, rank * chunk);
  k_end = min(K, k_begin + chunk);
}
// red: BM*BN floats of shared memory (16-byte aligned); epi(m_local, n_local, value) is called once per tile element by
// exactly one thread of the cluster, consecutive threads on consecutive columns (coalesced stores). splits == 1 (a plain
// launch) takes the same route through shared memory: one copy of the epilogue code, row-contiguous output.
template <class T, class Epi>
__device__ __forceinline__ void simt_cluster_reduce(float* red, const float (&acc)[T::TM][T::TN], int splits, Epi&& epi) {
  namespace cg = cooperative_groups;
  const int tid = threadIdx.x;
  const int tx = tid % (T::BN / T::TN), ty = tid / (T::BN / T::TN);
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int row = (i / 4) * (T::BM / (T::TM / 4)) + ty * 4 + (i % 4);
#pragma unroll
    for (int g = 0; g < T::TN / 4; ++g) {
      const int col = g * (T::BN / (T::TN / 4)) + tx * 4;
      *reinterpret_cast<float4*>(red + row * T::BN + col) =
          make_float4(acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]);
    }
  }
  if (splits == 1) {
    __syncthreads();
    for (int idx = tid; idx < T::BM * T::BN; idx += T::THREADS) epi(idx / T::BN, idx % T::BN, red[idx]);
    return;
  }
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  const int rank = (int)cluster.block_rank();
  const int rows_per = T::BM / splits;
  const float* src[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) src[q] = q < splits ? cluster.map_shared_rank(red, q) : red;
  for (int idx = tid; idx < rows_per * T::BN; idx += T::THREADS) {
    const int row = rank * rows_per + idx / T::BN, col = idx % T::BN;
    float s = src[0][row * T::BN + col];
#pragma unroll
    for (int q = 1; q < 8; ++q)
      if (q < splits) s += src[q][row * T::BN + col];
    epi(row, col, s);
  }
  cluster.sync();     // nobody leaves while its tile is still being read
}

// launch `kernel` on `grid` (z = splits) as clusters of (1,1,splits); splits == 1 is a plain launch
template <class... KArgs, class... Args>
inline cudaError_t simt_launch_split(void (*kernel)(KArgs...), dim3 grid, int threads, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = grid.z;
  cfg.attrs = at; cfg.numAttrs = grid.z > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// small layers / ragged shapes: more CTAs. (Measured and dropped for the RBM kernels at 328 x 1024 x 500: 32-column
// k-blocks, 23.2 vs 21.9 us per forward launch and 47 vs 32 us for dW; 128 x 64 tiles with 8 x 4 micro-tiles, 31.8 us.)
using TileSmall = SimtTile<64, 64, 16, 4, 4>;
using TileMid = SimtTile<128, 64, 16, 8, 4>;      // RBM layers with >= 65 rows: 32 FMAs per 3 shared-memory loads     // small layers / ragged shapes: more CTAs
using TileLarge = SimtTile<128, 128, 8, 8, 8>;    // GP chain at N ~ 5000

}  // namespace db
