// GP path: RBF Gram (K3), blocked Cholesky (K4), triangular solves / inverse (K5) and the fused
// log-marginal-likelihood + analytic gradient (K6) behind gplvm.py:61-86, 210-288 and gprbm.py:193-245.
// Nothing here calls cuBLAS / cuSOLVER. Two forms of the same evaluation live in this file:
//
// fp32 SIMT form (N < 512, odd shapes, the reference's own N = 328 runs) — N points, D outputs, Q latent dims:
//   hyp      = softplus(opt) (+ jitter)                                    gplvm.py:39,49,152,219-221
//   K        = alpha * exp(-|xi-xj|^2 / (2 gamma^2)) + noise * I           gplvm.py:61-86, 218-224
//   L        = chol(K)           right-looking, 64-wide panels             gplvm.py:212
//   Linv     = L^-1              recursive block inversion (log-depth, batched GEMMs)
//   S        = Linv Y ; |S|^2 ; logdet = 2 sum log Lii                     gplvm.py:226-236
//   A        = Linv^T S = K^-1 Y
//   G        = 0.5 (D Linv^T Linv - A A^T), W = G o K_f  -> dX, d alpha, d gamma, d noise over the lower tile pairs
//
// tensor-core form (N >= 512, N and D multiples of 4; 3 x fp16 products when multiples of 8, else 3xTF32):
//   L        = one potrf_step_kernel launch per 64-wide column block + one tcgen05 trailing update per 512-wide outer block
//   U        = L^-T block row by block row (run_trinv_tc_f16) — on two auxiliary streams BESIDE the factorisation when the
//              step launches leave SMs free (TrinvOverlap, run_trinv_f16_side_block)
//   K^-1     = U U^T (exact fp32 diagonal, tr K^-1 from it), mirrored and converted in one pass
//   A        = K^-1 Y, sum Y o A and |A|^2 formed by the transposition that produces A;  -0.5 A A^T
//   dX, d alpha, d gamma, d noise: one pass over the tiles of G = (D/2) K^-1 - 0.5 A A^T (gp_contract_fused_kernel), the
//              hyper-parameter sums from fp64 closed forms (gp_fix_trace_kernel)
// Algorithmic work: N^3 + 3 N^2 D FLOP (SURVEY.md §8 d).
#include "context.h"
#include "simt_gemm.cuh"
#include "gemm_simt.cuh"
#include "gemm_tf32.cuh"
#include "gemm_f16x3.cuh"
#include "ptx.cuh"
#include <cstdlib>
#include <mutex>

namespace db {
namespace {

constexpr int NB = 64;          // Cholesky panel width / diagonal block size
constexpr int MAXQ = 8;         // latent dimensionality supported by the fused gradient kernel

// scalars accumulated on the device during one evaluation
struct GpAccum {
  float ssq;        // |L^-1 Y|_F^2
  float logdet;     // 2 sum log Lii
  float trG;        // tr G
  float sumW;       // sum_ij G_ij Kf_ij
  float sumWr2;     // sum_ij G_ij Kf_ij |xi-xj|^2
  float xsq;        // |X|_F^2
  float pad[2];
};

__global__ void effective_hyp_kernel(const float* __restrict__ opt, int flags, float fixed_noise, float jitter,
                                     float* __restrict__ hyp) {
  // alpha=False / gamma=False are the constant 1.0 (gplvm.py:31-32,41-42)
  hyp[0] = (flags & 1) ? softplusf_(opt[0]) : 1.0f;
  hyp[1] = (flags & 2) ? softplusf_(opt[1]) : 1.0f;
  // jitter only when the noise variance is trainable (gplvm.py:141-152,219-221)
  hyp[2] = (flags & 4) ? softplusf_(opt[2]) + jitter : fixed_noise;
}

// ---- K3: Gram -------------------------------------------------------------------------------------
// rows x 256 tile per CTA (rows = 128 for big matrices: one resident wave of CTAs, the latent staging amortised over 32
// rows per thread; 32 for small ones: more CTAs); thread = 4 consecutive columns x rows/4 rows -> 16-byte coalesced stores.
// Squared distances use the difference form (exact 0 on the diagonal; the reference's expanded form
// gplvm.py:67-82 leaves +-2e-6 there in fp32, SURVEY.md §7 hard part 3).
// The column latents are staged q-major ([Q][256]) so that a thread's four columns are one conflict-free 16-byte read,
// and for Q <= 4 they are read ONCE into registers (they do not depend on the row). The first version kept them
// column-major and re-read them per row: an 8-way bank conflict per read, 64 shared-memory wavefronts per warp and row —
// that, not HBM, bounded the kernel (51 us for the 100 MB of N = 5000, 1.96 TB/s).
// 2^x, flush-to-zero: no range fix-up around MUFU.EX2 (__expf costs 8 instructions per entry without -ftz)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// QT: the latent dimension as a compile-time constant (1..4; the reference's Q is 2), 0 = any Q at run time. ncu on the
// run-time-Q version: 75 % of the issue slots busy at 39 % occupancy and only 42 of the 100 MB reaching DRAM inside the
// kernel's duration — it was bound by instructions (predicated q loops, index arithmetic, per-element diagonal tests), not
// by memory.
template <int QT>
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int M,
                                                   int Q_rt, const float* __restrict__ hyp, int add_noise,
                                                   float* __restrict__ out, int ldo, int rows) {
  extern __shared__ __align__(16) float gsm[];
  const int Q = QT > 0 ? QT : Q_rt;
  float* ys = gsm;                 // [Q][256]
  float* xs = gsm + 256 * Q;       // [rows][Q]
  const int i0 = blockIdx.y * rows, j0 = blockIdx.x * 256;
  const float alpha = __ldg(hyp), inv_gamma = 1.0f / __ldg(hyp + 1), noise = __ldg(hyp + 2);
  const float* Ysrc = Y ? Y : X;
  for (int t = threadIdx.x; t < rows * Q; t += 256) {
    const int r = i0 + t / Q;
    xs[t] = r < N ? __ldg(X + (size_t)r * Q + t % Q) * inv_gamma : 0.0f;
  }
  for (int t = threadIdx.x; t < 256 * Q; t += 256) {
    const int c = t / Q, q = t % Q, r = j0 + c;
    ys[q * 256 + c] = r < M ? __ldg(Ysrc + (size_t)r * Q + q) * inv_gamma : 0.0f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = j0 + 4 * tx;
  if (j >= M) return;
  constexpr int QR = QT > 0 ? QT : 4;            // latent dimensions held in registers
  float4 yr[QR];
#pragma unroll
  for (int q = 0; q < QR; ++q)
    yr[q] = (QT > 0 || q < Q) ? *reinterpret_cast<const float4*>(ys + q * 256 + 4 * tx) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec_store = j + 3 < M && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const bool square = Y == nullptr;
  const float diag_value = add_noise ? alpha + noise : alpha;
  const float lg_alpha = log2f(alpha);           // alpha exp(-d/2) = 2^(log2 alpha - d log2(e)/2): one FFMA + one EX2 per entry
  const int nr = min(rows / 4, (N - i0 - ty + 3) / 4);          // rows i0 + ty + 4 rr < N
  float* o = out + (size_t)(i0 + ty) * ldo + j;
  const float* xrow = xs + ty * Q;
#pragma unroll 2
  for (int rr = 0; rr < nr; ++rr, o += (size_t)4 * ldo, xrow += 4 * Q) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < QR; ++q) {
      if (QT > 0 || q < Q) {
        const float xv = xrow[q];
        float t;
        t = xv - yr[q].x; d[0] = fmaf(t, t, d[0]);
        t = xv - yr[q].y; d[1] = fmaf(t, t, d[1]);
        t = xv - yr[q].z; d[2] = fmaf(t, t, d[2]);
        t = xv - yr[q].w; d[3] = fmaf(t, t, d[3]);
      }
    }
    if (QT == 0) {
      for (int q = QR; q < Q; ++q) {
        const float xv = xrow[q];
        const float4 yv = *reinterpret_cast<const float4*>(ys + q * 256 + 4 * tx);
        float t;
        t = xv - yv.x; d[0] = fmaf(t, t, d[0]);
        t = xv - yv.y; d[1] = fmaf(t, t, d[1]);
        t = xv - yv.z; d[2] = fmaf(t, t, d[2]);
        t = xv - yv.w; d[3] = fmaf(t, t, d[3]);
      }
    }
    float kv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) kv[e] = ex2_ftz(fmaf(d[e], -0.5f * 1.4426950408889634f, lg_alpha));   // alpha exp(-d/2)
    const int dj = i0 + ty + 4 * rr - j;                          // the diagonal element of this row, if it is one of the four
    if (square && (unsigned)dj < 4u) {                            // r = 0 there: exactly alpha (+ noise)
#pragma unroll
      for (int e = 0; e < 4; ++e) if (dj == e) kv[e] = diag_value;
    }
    if (vec_store) {
      *reinterpret_cast<float4*>(o) = make_float4(kv[0], kv[1], kv[2], kv[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (j + e < M) o[e] = kv[e];
    }
  }
}

// ---- K4: Cholesky ---------------------------------------------------------------------------------
// Two-level right-looking factorisation. Column blocks are NB = 64 wide; OB = 512 columns form an outer
// block. Per column block ONE launch (potrf_colblock_kernel) factorises the 64x64 diagonal block and solves the
// panel below it; a narrow SYRK then updates only the remaining columns of the outer block (K = 64), and once
// per outer block a wide SYRK (K = 512) updates the rest of the matrix — so the strictly sequential part is
// ~80 short launches and the bulk of the N^3/3 flops runs as deep-K GEMM tiles.
constexpr int OB = 512;
constexpr int LT_PITCH = 72;   // floats; rows 16-byte aligned for float4 reads, and (8 t + g) mod 32 is a bijection for the
                               // mma.sync fragment reads (t = lane & 3, g = lane >> 2): conflict-free

// Shared-memory layout of a 64x64 block: S[c][r] holds element (row r, column c), so "my row's element c" is
// lane-contiguous (thread r reads S[c][r]) and "row c2's element c" (S[c][c2]) is a broadcast; 8 consecutive
// columns of one row of L^T, S[k][c0..c0+7], are two aligned float4.
// Left-looking in chunks of 8 columns: the rank-c0 update of a chunk is a k-loop of 8 FMAs per 3 shared loads,
// only the 8x8 triangular part of each chunk is unrolled (small code: the whole factorisation stays in the
// instruction cache, unlike a fully unrolled 64-column version).

// acc[i] = Own[c0+i][r] - sum_{k<c0} Own[k][r] * S[k][c0+i]     (Own == S for the diagonal block)
__device__ __forceinline__ void chunk_update(float (&acc)[8], const float (*Own)[LT_PITCH], const float (*S)[LT_PITCH], int c0,
                                             int r) {
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = Own[c0 + i][r];
#pragma unroll 4
  for (int k = 0; k < c0; ++k) {
    const float x = Own[k][r];
    const float4 l0 = *reinterpret_cast<const float4*>(&S[k][c0]);
    const float4 l1 = *reinterpret_cast<const float4*>(&S[k][c0 + 4]);
    acc[0] = fmaf(-x, l0.x, acc[0]); acc[1] = fmaf(-x, l0.y, acc[1]); acc[2] = fmaf(-x, l0.z, acc[2]); acc[3] = fmaf(-x, l0.w, acc[3]);
    acc[4] = fmaf(-x, l1.x, acc[4]); acc[5] = fmaf(-x, l1.y, acc[5]); acc[6] = fmaf(-x, l1.z, acc[6]); acc[7] = fmaf(-x, l1.w, acc[7]);
  }
}

// grid.x = number of 64-row blocks from row j down; block 0 is the diagonal block. 64 threads, one matrix row each.
// Every CTA factorises the diagonal block itself (identical arithmetic) instead of waiting for another launch.
__global__ void __launch_bounds__(NB) potrf_colblock_kernel(float* __restrict__ A, int lda, int N, int j, int* __restrict__ info) {
  __shared__ __align__(16) float S[NB][LT_PITCH];   // diagonal block: A on entry, L on exit (lower part)
  __shared__ __align__(16) float P[NB][LT_PITCH];   // this CTA's panel rows (blockIdx.x > 0)
  __shared__ float dinv[NB];
  __shared__ int s_fail;
  const int r = threadIdx.x;
  const int nb = min(NB, N - j);
  // one matrix row per thread; loads are issued 16 at a time so their latencies overlap
  const int prow = j + blockIdx.x * NB + r;
  {
    const float* src = A + (size_t)(j + r) * lda + j;
    const bool vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && nb == NB;   // j is a multiple of 64
#pragma unroll 1
    for (int kb = 0; kb < NB; kb += 16) {
      float t[16];
      if (vec) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 q = (kb + 4 * u <= r) ? __ldg(reinterpret_cast<const float4*>(src + kb) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
          t[4 * u] = q.x; t[4 * u + 1] = q.y; t[4 * u + 2] = q.z; t[4 * u + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = (r < nb && kb + u <= r) ? __ldg(src + kb + u) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int k = kb + u;
        S[k][r] = (r < nb && k <= r) ? t[u] : (k == r ? 1.0f : 0.0f);   // identity padding
      }
    }
  }
  if (blockIdx.x > 0) {
    const float* src = A + (size_t)prow * lda + j;
    const bool vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && nb == NB && prow < N;
#pragma unroll 1
    for (int kb = 0; kb < NB; kb += 16) {
      float t[16];
      if (vec) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(src + kb) + u);
          t[4 * u] = q.x; t[4 * u + 1] = q.y; t[4 * u + 2] = q.z; t[4 * u + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = (prow < N && kb + u < nb) ? __ldg(src + kb + u) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) P[kb + u][r] = t[u];
    }
  }
  if (r == 0) s_fail = 0;
  __syncthreads();
  for (int c0 = 0; c0 < NB; c0 += 8) {
    float acc[8];
    chunk_update(acc, S, S, c0, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float s = acc[i];
#pragma unroll
      for (int m = 0; m < i; ++m) s = fmaf(-acc[m], S[c0 + m][c], s);    // acc[m] already holds L[r][c0+m]
      if (r == c) {
        if (!(s > 0.0f) && s_fail == 0) s_fail = c + 1;
        const float d = sqrtf(s);
        S[c][c] = d;
        dinv[c] = 1.0f / d;
      }
      __syncthreads();
      if (r > c) { acc[i] = s * dinv[c]; S[c][r] = acc[i]; }
      __syncthreads();
    }
  }
  if (blockIdx.x == 0) {
    if (r < nb) {
      float* dst = A + (size_t)(j + r) * lda + j;
      for (int k = 0; k <= r; ++k) dst[k] = S[k][r];
    }
    if (r == 0 && s_fail && s_fail <= nb) atomicCAS(info, 0, j + s_fail);
    return;
  }
  // panel rows: x L^T = a by forward substitution, same chunking, no synchronisation needed
  for (int c0 = 0; c0 < NB; c0 += 8) {
    float acc[8];
    chunk_update(acc, P, S, c0, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = acc[i];
#pragma unroll
      for (int m = 0; m < i; ++m) s = fmaf(-acc[m], S[c0 + m][c0 + i], s);
      acc[i] = s * dinv[c0 + i];
      P[c0 + i][r] = acc[i];
    }
  }
  if (prow < N) {
    float* dst = A + (size_t)prow * lda + j;
    if (((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && nb == NB) {
#pragma unroll 4
      for (int k = 0; k < NB; k += 4)
        *reinterpret_cast<float4*>(dst + k) = make_float4(P[k][r], P[k + 1][r], P[k + 2][r], P[k + 3][r]);
    } else {
      for (int k = 0; k < nb; ++k) dst[k] = P[k][r];
    }
  }
}

// ---- one launch per column block: fused look-ahead update + factorisation + panel solve + narrow update ----------
// Launch for column block j of the outer block [b0, b1), jprev = j - 64 (or -1 when j == b0). Every CTA owns its SM
// (STEP_SMEM below); 256 threads.
//   strip CTAs  (blockIdx.x <  strips): 64-row strip s of the rows from j down (s = 0 is the diagonal block). Each one
//     applies panel jprev (rank-64) to its own block of column block j and (redundantly) to the diagonal block - one
//     128 x 64 x 64 product on the tensor cores (tcgen05.mma, 3xTF32, accumulator in TMEM) -, factorises the diagonal block (redundantly:
//     identical arithmetic in every CTA, no inter-CTA dependency; warps 0-1, 8x8 sub-blocks in registers) and solves its
//     strip (warps 2-3, chunk c as soon as chunk c of L is final).
//   update CTAs (blockIdx.x >= strips, at most one per remaining SM): loop over the 128 x 64 tiles of the columns
//     [j + 64, b1): tile -= panel jprev product, i.e. the right-looking narrow update of the PREVIOUS panel rides along
//     in the same launch, off the critical path (next tile's operands in flight during the product).
// Column block cb has therefore received panel p < cb - 64 from the launch of step p + 64 and panel cb - 64 in its
// own launch. S and P (factorisation operands) are stored transposed, T[k][r] = element (row r, column k); the blocks
// of the previous panel (tensor-core operands only) row-major.
// Measured per step at N = 5000 (clock64 in strip CTA 1, build with -DDB_POTRF_TIMING and run with DB_POTRF_DEBUG=1):
// load 7.5k, update 2.5k (4.9k with mma.sync, 14.5k with fp32 FMAs), factorisation 15.5k cycles; 16 us per step.
constexpr int STEP_THREADS = 256;
// Requested dynamic shared memory: more than half an SM, so every CTA of a step launch owns its SM. The strip CTAs are
// the critical path; measured with two update CTAs sharing their SM, their load and product phases ran 3x slower.
constexpr int STEP_SMEM = 200 * 1024;

// 64x64 block at (row0, col0): global -> registers (PER float4 per thread of a GROUP-thread group), then registers ->
// T[k][r]. Split in two so a CTA can issue the loads of all its blocks before the first shared store.
// Rows >= rows_valid / columns >= cols_valid read as zero.
template <int GROUP>
struct BlockRegs { float4 v[1024 / GROUP]; };
template <int GROUP>
__device__ __forceinline__ void block_load(BlockRegs<GROUP>& b, int tid, const float* __restrict__ A, int lda, int row0, int col0,
                                           int rows_valid, int cols_valid, bool vec) {
  constexpr int PER = 1024 / GROUP, QS = GROUP / 64;
  const int r = tid & 63, q = tid >> 6;
  const float* src = A + (size_t)(row0 + r) * lda + col0;
  if (vec && cols_valid == NB) {
#pragma unroll
    for (int i = 0; i < PER; ++i)
      b.v[i] = r < rows_valid ? __ldcg(reinterpret_cast<const float4*>(src) + q + QS * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = 4 * (q + QS * i);
      b.v[i].x = (r < rows_valid && c + 0 < cols_valid) ? __ldcg(src + c + 0) : 0.0f;
      b.v[i].y = (r < rows_valid && c + 1 < cols_valid) ? __ldcg(src + c + 1) : 0.0f;
      b.v[i].z = (r < rows_valid && c + 2 < cols_valid) ? __ldcg(src + c + 2) : 0.0f;
      b.v[i].w = (r < rows_valid && c + 3 < cols_valid) ? __ldcg(src + c + 3) : 0.0f;
    }
  }
}
template <int GROUP>
__device__ __forceinline__ void block_store_T(float (*T)[LT_PITCH], int tid, const BlockRegs<GROUP>& b) {
  constexpr int PER = 1024 / GROUP, QS = GROUP / 64;
  const int r = tid & 63, q = tid >> 6;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = 4 * (q + QS * i);
    T[c][r] = b.v[i].x; T[c + 1][r] = b.v[i].y; T[c + 2][r] = b.v[i].z; T[c + 3][r] = b.v[i].w;
  }
}

// Operands that only feed the tensor-core products are loaded coalesced (16 threads read one 256-byte row; a warp-wide
// float4 load that touches 32 different rows, as the transposed loader above does, costs ~66 cycles of L1 replays) and
// stored as 64-byte-swizzled K-major tiles, the layout TMA SWIZZLE_64B would produce and gemm_tf32.cu feeds to
// tcgen05.mma: k-block kb (16 floats) of an R-row operand is a [R][64 B] tile at kb * R * 64 bytes, and 16-byte chunk c of
// row r sits at chunk position c ^ ((r >> 1) & 3). Every value is stored twice: x itself (the tensor core reads its top 19
// bits = tf32(x)) and the remainder x - tf32(x).
struct BlockRegsRM { float4 v[4]; };
__device__ __forceinline__ void block_load_rm(BlockRegsRM& b, int tid, const float* __restrict__ A, int lda, int row0, int col0,
                                              int rows_valid, bool vec) {
  const int c4 = tid & 15, rr = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = rr + 16 * i;
    const float* src = A + (size_t)(row0 + r) * lda + col0 + 4 * c4;
    if (r < rows_valid) {
      if (vec) b.v[i] = __ldcg(reinterpret_cast<const float4*>(src));
      else b.v[i] = make_float4(__ldcg(src), __ldcg(src + 1), __ldcg(src + 2), __ldcg(src + 3));
    } else {
      b.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
__device__ __forceinline__ float tf32_rem(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// rows [row_off, row_off + 64) of an R-row operand
__device__ __forceinline__ void block_store_sw(uint8_t* hi, uint8_t* lo, int R, int row_off, int tid, const BlockRegsRM& b) {
  const int c4 = tid & 15, rr = tid >> 4;
  const int kb = c4 >> 2, chunk = c4 & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row_off + rr + 16 * i;
    const int off = kb * R * 64 + r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4);
    const float4 v = b.v[i];
    *reinterpret_cast<float4*>(hi + off) = v;
    *reinterpret_cast<float4*>(lo + off) = make_float4(tf32_rem(v.x), tf32_rem(v.y), tf32_rem(v.z), tf32_rem(v.w));
  }
}
// K-major operand tile with SWIZZLE_64B: rows 64 B apart, 8-row groups 512 B apart (same descriptor as gemm_tf32.cu)
__device__ __forceinline__ uint64_t step_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)4u << 61;
  return d;
}
// D[128, 64] (TMEM: lane = row, column = column) = A[128, 64] B[64, 64]^T, 3xTF32 (lo*hi + hi*lo + hi*hi), issued by ONE
// thread: 4 k-blocks x 2 k-steps x 3 tcgen05.mma (M = 128, N = 64, K = 8); the commit arrives on `bar`.
__device__ __forceinline__ void step_issue_product(uint32_t tmem_d, const uint8_t* a_hi, const uint8_t* a_lo, const uint8_t* b_hi,
                                                   const uint8_t* b_lo, int b_rows, uint64_t* bar) {
  constexpr uint32_t idesc = ptx::umma_idesc(/*TF32*/ 2, 128, 64);
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    const uint64_t dah = step_desc_sw64(ptx::smem_u32(a_hi + kb * 128 * 64)), dal = step_desc_sw64(ptx::smem_u32(a_lo + kb * 128 * 64));
    const uint64_t dbh = step_desc_sw64(ptx::smem_u32(b_hi + kb * b_rows * 64)), dbl = step_desc_sw64(ptx::smem_u32(b_lo + kb * b_rows * 64));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);
      ptx::mma_tf32_ss(tmem_d, dal + koff, dbh + koff, idesc, (kb | k) ? 1u : 0u);
      ptx::mma_tf32_ss(tmem_d, dah + koff, dbl + koff, idesc, 1u);
      ptx::mma_tf32_ss(tmem_d, dah + koff, dbh + koff, idesc, 1u);
    }
  }
  ptx::mma_commit(bar);
}

// Cholesky of the 8x8 diagonal sub-block at c0 entirely in registers (every calling thread redundantly):
// l[i][m] (m <= i) and inv[i] = 1 / l[i][i]. rsqrt + one Newton step: no shared-memory round trip per column.
__device__ __forceinline__ void factor8(const float (*blk)[8], float (&l)[8][8], float (&inv)[8], int c0, int& fail) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 lo = *reinterpret_cast<const float4*>(&blk[i][0]);
    const float4 hi = *reinterpret_cast<const float4*>(&blk[i][4]);
    l[i][0] = lo.x; l[i][1] = lo.y; l[i][2] = lo.z; l[i][3] = lo.w;
    l[i][4] = hi.x; l[i][5] = hi.y; l[i][6] = hi.z; l[i][7] = hi.w;
  }
#pragma unroll
  for (int jc = 0; jc < 8; ++jc) {
    float s = l[jc][jc];
#pragma unroll
    for (int m = 0; m < jc; ++m) s = fmaf(-l[jc][m], l[jc][m], s);
    if (!(s > 0.0f) && fail == 0) fail = c0 + jc + 1;
    float y;                                               // one MUFU (rsqrtf adds a denormal-range fix-up to the dependent chain;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));   //  s is a pivot of magnitude ~noise variance or the step has failed)
    y = y * fmaf(-0.5f * s * y, y, 1.5f);
    inv[jc] = y;
    l[jc][jc] = s * y;
#pragma unroll
    for (int i = jc + 1; i < 8; ++i) {
      float t = l[i][jc];
#pragma unroll
      for (int m = 0; m < jc; ++m) t = fmaf(-l[i][m], l[jc][m], t);
      l[i][jc] = t * y;
    }
  }
}

#ifdef DB_POTRF_TIMING
#define STEP_MARK(slot, cond) do { if (dbg != nullptr && blockIdx.x == 1 && (cond)) dbg[slot] = clock64(); } while (0)
#else
#define STEP_MARK(slot, cond) do { } while (0)
#endif
__global__ void __launch_bounds__(STEP_THREADS) potrf_step_kernel(float* __restrict__ A, int lda, int N, int j, int jprev, int b1,
                                                                  int strips, int* __restrict__ info, long long* __restrict__ dbg) {
#if defined(DB_SM100)
  extern __shared__ __align__(16) uint8_t step_smem_raw[];
  // 1 KB aligned: the swizzle pattern of the operand tiles is a function of the shared-memory address bits
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(step_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = smem;                       // [4 k-blocks][128 rows][64 B]: rows 0-63 diagonal / first strip, 64-127 own / second strip
  uint8_t* a_lo = smem + 32 * 1024;
  uint8_t* b_hi = smem + 64 * 1024;           // [4][64][64 B] (update CTAs; the strip CTAs reuse rows 0-63 of a_*)
  uint8_t* b_lo = smem + 80 * 1024;
  float* step_smem = reinterpret_cast<float*>(smem + 96 * 1024);
  float (*S)[LT_PITCH] = reinterpret_cast<float (*)[LT_PITCH]>(step_smem);                        // diagonal block
  float (*P)[LT_PITCH] = reinterpret_cast<float (*)[LT_PITCH]>(step_smem + NB * LT_PITCH);        // own block
  float* stage = step_smem;                   // update CTAs: [128][STAGE_PITCH] product tile for the coalesced read-modify-write
  __shared__ float dinv[NB];
  __shared__ __align__(16) float dblk[2][8][8];   // updated 8x8 diagonal sub-block, double-buffered by chunk parity
  __shared__ int s_fail;
  __shared__ volatile int s_progress;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const bool vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool use_tc = jprev >= 0;             // launch-uniform: products with the previous panel exist
  if (use_tc) {
    if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 64); ptx::tmem_relinquish(); }
    if (t == 0) { ptx::mbar_init(&mma_bar, 1); ptx::fence_barrier_init(); }
  }
  // Programmatic dependent launch: the next step's CTAs may be scheduled as soon as every CTA of this launch has
  // started; they block here until the previous launch has completed and its writes are visible.
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  STEP_MARK(0, t == 0);
  uint32_t mma_phase = 0;

  if ((int)blockIdx.x >= strips) {
    // ---------------- update CTA: loops over the (128 rows x 64 columns) tiles of the columns [j + 64, b1);
    // tile -= (128 rows of panel jprev) (rows cb.. of panel jprev)^T, next tile's operands in flight meanwhile ----------------
    const int ncb = (b1 - (j + NB) + NB - 1) / NB;
    const int stride = (int)gridDim.x - strips;
    // tile index -> (column block cb, first row); column block cb has ceil((N - cb) / 128) tiles starting at row cb
    auto locate = [&](int tile, int& cb, int& row0) -> bool {
      int rem = tile;
      cb = j + NB;
      for (int cbi = 0; cbi < ncb; ++cbi, cb += NB) {
        const int nt = (N - cb + 2 * NB - 1) / (2 * NB);
        if (rem < nt) { row0 = cb + rem * 2 * NB; return true; }
        rem -= nt;
      }
      return false;
    };
    int tile = (int)blockIdx.x - strips;
    int cb, row0;
    bool have = locate(tile, cb, row0);
    BlockRegsRM r0, r1, rd;
    if (have) {
      block_load_rm(r0, t, A, lda, row0, jprev, N - row0, vec);
      block_load_rm(r1, t, A, lda, row0 + NB, jprev, N - row0 - NB, vec);
      block_load_rm(rd, t, A, lda, cb, jprev, N - cb, vec);
    }
    ptx::tc_fence_before();
    __syncthreads();                                        // TMEM address and barrier are visible
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    constexpr int STAGE_PITCH = 68;
    while (have) {
      block_store_sw(a_hi, a_lo, 128, 0, t, r0);
      block_store_sw(a_hi, a_lo, 128, NB, t, r1);
      block_store_sw(b_hi, b_lo, NB, 0, t, rd);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (t == 0) {
        ptx::tc_fence_after();
        step_issue_product(tmem_d, a_hi, a_lo, b_hi, b_lo, NB, &mma_bar);
      }
      const int cur_cb = cb, cur_row0 = row0;
      tile += stride;
      have = locate(tile, cb, row0);                        // CTA-uniform
      if (have) {                                           // next tile's operands in flight during the product
        block_load_rm(r0, t, A, lda, row0, jprev, N - row0, vec);
        block_load_rm(r1, t, A, lda, row0 + NB, jprev, N - row0 - NB, vec);
        block_load_rm(rd, t, A, lda, cb, jprev, N - cb, vec);
      }
      // the tile's current values, coalesced: 16 threads per row
      const int c4 = t & 15, rr = t >> 4;
      const int ncols = min(NB, b1 - cur_cb);
      float4 cur[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = cur_row0 + rr + 16 * i;
        const float* src = A + (size_t)row * lda + cur_cb + 4 * c4;
        if (row < N && vec && ncols == NB) cur[i] = __ldcg(reinterpret_cast<const float4*>(src));
        else {
          cur[i].x = (row < N && 4 * c4 + 0 < ncols) ? __ldcg(src + 0) : 0.0f;
          cur[i].y = (row < N && 4 * c4 + 1 < ncols) ? __ldcg(src + 1) : 0.0f;
          cur[i].z = (row < N && 4 * c4 + 2 < ncols) ? __ldcg(src + 2) : 0.0f;
          cur[i].w = (row < N && 4 * c4 + 3 < ncols) ? __ldcg(src + 3) : 0.0f;
        }
      }
      ptx::mbar_wait(&mma_bar, mma_phase);
      mma_phase ^= 1;
      ptx::tc_fence_after();
      {   // TMEM -> registers -> row-major staging tile: warp w reads lanes 32 (w & 3).., columns 32 (w >> 2)..
        uint32_t r[32];
        const int lrow = 32 * (warp & 3) + lane, col0 = 32 * (warp >> 2);
        ptx::tmem_ld_32x32(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(stage + lrow * STAGE_PITCH + col0 + 4 * q) =
              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
      }
      ptx::tc_fence_before();
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int lrow = rr + 16 * i, row = cur_row0 + lrow;
        if (row >= N) continue;
        const float4 d = *reinterpret_cast<const float4*>(stage + lrow * STAGE_PITCH + 4 * c4);
        float* dst = A + (size_t)row * lda + cur_cb + 4 * c4;
        const int col = cur_cb + 4 * c4;
        if (vec && ncols == NB && col + 3 <= row) {
          *reinterpret_cast<float4*>(dst) = make_float4(cur[i].x - d.x, cur[i].y - d.y, cur[i].z - d.z, cur[i].w - d.w);
        } else {
          if (4 * c4 + 0 < ncols && col + 0 <= row) dst[0] = cur[i].x - d.x;
          if (4 * c4 + 1 < ncols && col + 1 <= row) dst[1] = cur[i].y - d.y;
          if (4 * c4 + 2 < ncols && col + 2 <= row) dst[2] = cur[i].z - d.z;
          if (4 * c4 + 3 < ncols && col + 3 <= row) dst[3] = cur[i].w - d.w;
        }
      }
      __syncthreads();                                      // staging tile and operand tiles are free again
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_d, 64); }
    return;
  }

  // ---------------- strip CTA ----------------
  const int strip = blockIdx.x;
  const int nb = min(NB, N - j);
  const int row0 = j + strip * NB;
  if (vec && nb == NB) {
    // Full 64-column blocks (every step but a ragged last one): the two factorisation operands are read COALESCED (16 threads
    // per 256-byte row, like the tensor-core operands) into a row-major staging tile of pitch 65 and transposed from there.
    // The direct transposed load below touches 32 different rows per warp instruction (~66 cycles of L1 replays each, 8 such
    // instructions per thread): it made the load phase 7.5k of the step's ~30k cycles.
    float (*T2s)[65] = reinterpret_cast<float (*)[65]>(smem + 136 * 1024);
    float (*T2p)[65] = T2s + NB;
    BlockRegsRM rs2, rp2, rdp, rop;
    block_load_rm(rs2, t, A, lda, j, j, nb, vec);
    if (strip > 0) block_load_rm(rp2, t, A, lda, row0, j, N - row0, vec);
    if (use_tc) {
      block_load_rm(rdp, t, A, lda, j, jprev, nb, vec);
      block_load_rm(rop, t, A, lda, row0, jprev, strip > 0 ? N - row0 : 0, vec);
    }
    {
      const int c4 = t & 15, rr = t >> 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rr + 16 * i;
        T2s[r][4 * c4] = rs2.v[i].x; T2s[r][4 * c4 + 1] = rs2.v[i].y; T2s[r][4 * c4 + 2] = rs2.v[i].z; T2s[r][4 * c4 + 3] = rs2.v[i].w;
        if (strip > 0) {
          T2p[r][4 * c4] = rp2.v[i].x; T2p[r][4 * c4 + 1] = rp2.v[i].y; T2p[r][4 * c4 + 2] = rp2.v[i].z; T2p[r][4 * c4 + 3] = rp2.v[i].w;
        }
      }
    }
    if (use_tc) {
      block_store_sw(a_hi, a_lo, 128, 0, t, rdp);
      block_store_sw(a_hi, a_lo, 128, NB, t, rop);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    {
      const int r = t & 63, q = t >> 6;        // lanes along r: conflict-free reads of pitch 65 and writes of T[c][r]
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int c = 16 * q + e;
        S[c][r] = T2s[r][c];
        if (strip > 0) P[c][r] = T2p[r][c];
      }
    }
  } else {
    BlockRegs<256> rs, rp;
    BlockRegsRM rdp, rop;
    block_load<256>(rs, t, A, lda, j, j, nb, nb, vec);
    if (strip > 0) block_load<256>(rp, t, A, lda, row0, j, N - row0, nb, vec);
    if (use_tc) {
      block_load_rm(rdp, t, A, lda, j, jprev, nb, vec);
      block_load_rm(rop, t, A, lda, row0, jprev, strip > 0 ? N - row0 : 0, vec);
    }
    block_store_T<256>(S, t, rs);
    if (strip > 0) block_store_T<256>(P, t, rp);
    if (use_tc) {
      block_store_sw(a_hi, a_lo, 128, 0, t, rdp);           // rows 0-63: previous panel, diagonal rows (also the B operand)
      block_store_sw(a_hi, a_lo, 128, NB, t, rop);          // rows 64-127: previous panel, own rows (zeros for strip 0)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  if (t == 0) { s_fail = 0; s_progress = 0; }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  STEP_MARK(1, t == 0);
  if (use_tc) {
    // [S; P] -= [Dp; Op] Dp^T : one 128 x 64 x 64 product on the tensor cores (tcgen05, 3xTF32), accumulator in TMEM
    const uint32_t tmem_d = tmem_slot;
    if (t == 0) step_issue_product(tmem_d, a_hi, a_lo, a_hi, a_lo, 128, &mma_bar);
    ptx::mbar_wait(&mma_bar, mma_phase);
    mma_phase ^= 1;
    ptx::tc_fence_after();
    if ((warp & 3) < 2 || strip > 0) {                      // lanes 0-63 -> S, lanes 64-127 -> P
      uint32_t r[32];
      const int lrow = 32 * (warp & 3) + lane, col0 = 32 * (warp >> 2);
      ptx::tmem_ld_32x32(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0, r);
      ptx::tmem_ld_wait();
      float (*T)[LT_PITCH] = lrow < NB ? S : P;
      const int rloc = lrow & (NB - 1);
#pragma unroll
      for (int q = 0; q < 32; ++q) T[col0 + q][rloc] -= __uint_as_float(r[q]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_d, 64); }
  }
  if (nb < NB) {   // identity padding of the ragged last block
    for (int e = t; e < NB * NB; e += STEP_THREADS) {
      const int k = e >> 6, r = e & 63;
      if (r >= nb || k >= nb) S[k][r] = (k == r) ? 1.0f : 0.0f;
    }
    __syncthreads();
  }

  STEP_MARK(2, t == 0);
  if (t < NB) {
    // ---- factorisation of the diagonal block: thread r owns row r; left-looking in chunks of 8 columns ----
    const int r = t;
    int fail = 0;
    for (int c0 = 0; c0 < NB; c0 += 8) {
      float acc[8];
      chunk_update(acc, S, S, c0, r);
      const int par = (c0 >> 3) & 1;
      if (r >= c0 && r < c0 + 8) {     // rows of the 8x8 diagonal sub-block publish their updated values
        *reinterpret_cast<float4*>(&dblk[par][r - c0][0]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(&dblk[par][r - c0][4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
      float l[8][8], inv[8];
      factor8(dblk[par], l, inv, c0, fail);
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = acc[i];
#pragma unroll
        for (int m = 0; m < i; ++m) s = fmaf(-x[m], l[i][m], s);
        x[i] = s * inv[i];
        S[c0 + i][r] = (c0 + i <= r) ? x[i] : 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) if (r == c0 + i) dinv[r] = inv[i];
      asm volatile("bar.sync 1, 64;" ::: "memory");     // chunk c0 of L is final
      if (r == 0) { __threadfence_block(); s_progress = c0 / 8 + 1; }
    }
    if (r == 0 && fail) s_fail = fail;
    STEP_MARK(3, t == 0);
  } else if (t < 2 * NB && strip > 0) {
    // ---- panel rows: x L^T = a by forward substitution, chunk c as soon as chunk c of L is final ----
    const int r = t - NB;
    for (int c0 = 0; c0 < NB; c0 += 8) {
      while (s_progress < c0 / 8) __nanosleep(64);
      __threadfence_block();
      float acc[8];
      chunk_update(acc, P, S, c0, r);
      while (s_progress < c0 / 8 + 1) __nanosleep(64);    // polling without a pause steals shared-memory issue slots from the factorising warps
      __threadfence_block();
      float l[8][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const float4 lo = *reinterpret_cast<const float4*>(&S[c0 + m][c0]);
        const float4 hi = *reinterpret_cast<const float4*>(&S[c0 + m][c0 + 4]);
        l[0][m] = lo.x; l[1][m] = lo.y; l[2][m] = lo.z; l[3][m] = lo.w;
        l[4][m] = hi.x; l[5][m] = hi.y; l[6][m] = hi.z; l[7][m] = hi.w;
      }
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = acc[i];
#pragma unroll
        for (int m = 0; m < i; ++m) s = fmaf(-x[m], l[i][m], s);
        x[i] = s * dinv[c0 + i];
        P[c0 + i][r] = x[i];
      }
    }
  }
  STEP_MARK(4, t == NB);
  __syncthreads();
  // ---- write back: the diagonal CTA stores L_jj (lower part), the others their solved strip ----
  {
    const int r = t & 63, q = t >> 6;
    float (*Src)[LT_PITCH] = strip == 0 ? S : P;
    const int row = row0 + r;
    if (row < N) {
      float* dst = A + (size_t)row * lda + j;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = 4 * (q + 4 * i);
        if (strip > 0 && vec && nb == NB) {
          *reinterpret_cast<float4*>(dst + c) = make_float4(Src[c][r], Src[c + 1][r], Src[c + 2][r], Src[c + 3][r]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c + e < nb && (strip > 0 || c + e <= r)) dst[c + e] = Src[c + e][r];
        }
      }
    }
    if (strip == 0 && t == 0 && s_fail && s_fail <= nb) atomicCAS(info, 0, j + s_fail);
  }
  STEP_MARK(5, t == 0);
#ifdef DB_POTRF_TIMING
  if (dbg != nullptr && blockIdx.x == 1 && t == 0) { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); dbg[6] = (long long)g; }
#endif
#endif  // DB_SM100
}

// C[i, jc] -= sum_{k in [k0,k1)} A[i,k] A[jc,k]  for i in [i0,N), jc in [j0,j1), jc <= i   (lower part only)
struct LoadSyrkA {
  static constexpr bool k_contig = true;
  const float* P; int lda; int rows, kk;
  __device__ __forceinline__ float operator()(int m, int k) const { return (m < rows && k < kk) ? __ldg(P + (size_t)m * lda + k) : 0.0f; }
};
struct LoadSyrkB {
  static constexpr bool k_contig = true;
  const float* P; int lda; int rows, kk;
  __device__ __forceinline__ float operator()(int k, int n) const { return (n < rows && k < kk) ? __ldg(P + (size_t)n * lda + k) : 0.0f; }
};
__global__ void __launch_bounds__(256) syrk_block_kernel(float* __restrict__ A, int lda, int N, int i0, int j0, int j1, int k0,
                                                         int k1) {
  using T = TileLarge;
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  if (j0 + n0 > i0 + m0 + T::BM - 1) return;            // tile entirely above the diagonal
  __shared__ typename T::Smem sm;
  LoadSyrkA a{A + (size_t)i0 * lda + k0, lda, N - i0, k1 - k0};
  LoadSyrkB b{A + (size_t)j0 * lda + k0, lda, j1 - j0, k1 - k0};
  float acc[T::TM][T::TN];
  simt_gemm_tile<T>(sm, a, b, m0, n0, 0, k1 - k0, acc);
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int row = i0 + simt_row<T>(m0, i);
    if (row >= N) continue;
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int col = j0 + simt_col<T>(n0, jj);
      if (col < j1 && col <= row) A[(size_t)row * lda + col] -= acc[i][jj];
    }
  }
}

// ---- K5: triangular solves ------------------------------------------------------------------------
// diagonal-block solve of the blocked substitution: B_i <- L_ii^-1 B_i (trans=0) or L_ii^-T B_i (trans=1)
__global__ void __launch_bounds__(256) trsm_diag_kernel(const float* __restrict__ L, int ldl, int N, int i0, float* __restrict__ B,
                                                        int ldb, int nrhs, int trans) {
  __shared__ float D[NB][NB + 1];
  const int nb = min(NB, N - i0);
  for (int t = threadIdx.x; t < NB * NB; t += 256) {
    const int r = t / NB, c = t % NB;
    D[r][c] = (r < nb && c <= r) ? __ldg(L + (size_t)(i0 + r) * ldl + i0 + c) : 0.0f;
  }
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= nrhs) return;
  float x[NB];
  if (!trans) {
#pragma unroll 1
    for (int r = 0; r < nb; ++r) {
      float s = B[(size_t)(i0 + r) * ldb + col];
      for (int k = 0; k < r; ++k) s = fmaf(-D[r][k], x[k], s);
      x[r] = s / D[r][r];
    }
  } else {
#pragma unroll 1
    for (int r = nb - 1; r >= 0; --r) {
      float s = B[(size_t)(i0 + r) * ldb + col];
      for (int k = r + 1; k < nb; ++k) s = fmaf(-D[k][r], x[k], s);
      x[r] = s / D[r][r];
    }
  }
  for (int r = 0; r < nb; ++r) B[(size_t)(i0 + r) * ldb + col] = x[r];
}
// B[rows] -= Lblk * B_i : trans=0 rows below block i use L[rows, i]; trans=1 rows above use L[i, rows]^T
struct LoadTrsmA {
  const float* L; int ldl; int trans; int r0, i0, nrows, nb;
  static constexpr bool k_contig = true;   // exact only for trans=0; correctness does not depend on it
  __device__ __forceinline__ float operator()(int m, int k) const {
    if (m >= nrows || k >= nb) return 0.0f;
    return trans ? __ldg(L + (size_t)(i0 + k) * ldl + r0 + m) : __ldg(L + (size_t)(r0 + m) * ldl + i0 + k);
  }
};
struct LoadTrsmB {
  const float* B; int ldb; int i0, nb, nrhs;
  static constexpr bool k_contig = false;
  __device__ __forceinline__ float operator()(int k, int n) const {
    return (k < nb && n < nrhs) ? B[(size_t)(i0 + k) * ldb + n] : 0.0f;
  }
};
__global__ void __launch_bounds__(256) trsm_update_kernel(const float* __restrict__ L, int ldl, int i0, int nb, int r0, int nrows,
                                                          float* __restrict__ B, int ldb, int nrhs, int trans) {
  using T = TileSmall;
  __shared__ typename T::Smem sm;
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  LoadTrsmA a{L, ldl, trans, r0, i0, nrows, nb};
  LoadTrsmB b{B, ldb, i0, nb, nrhs};
  float acc[T::TM][T::TN];
  simt_gemm_tile<T>(sm, a, b, m0, n0, 0, nb, acc);
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= nrows) continue;
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (n < nrhs) B[(size_t)(r0 + m) * ldb + n] -= acc[i][jj];
    }
  }
}

// ---- triangular inverse ---------------------------------------------------------------------------
// level 0: invert every 64x64 diagonal block (one CTA each; thread c solves column c of L_bb X = I by forward
// substitution). The solution lives in shared memory (a register array indexed by the row went to local memory: 44 us per
// launch) and every thread walks the same (r, k) range — entries above the diagonal are exact zeros, so the extra terms
// change nothing — with the dot product unrolled by eight over four accumulators.
__global__ void __launch_bounds__(64) trinv_diag_kernel(const float* __restrict__ L, int ldl, int N, float* __restrict__ Li, int ldi) {
  __shared__ float D[NB][NB + 1];
  __shared__ float Xs[NB][NB + 1];      // Xs[r][c] = X[r, c]
  const int i0 = blockIdx.x * NB;
  const int nb = min(NB, N - i0);
  for (int t = threadIdx.x; t < NB * NB; t += 64) {
    const int r = t / NB, c = t % NB;
    D[r][c] = (r < nb && c <= r) ? __ldg(L + (size_t)(i0 + r) * ldl + i0 + c) : 0.0f;
  }
  __shared__ float rdiag[NB];
  __syncthreads();
  const int c = threadIdx.x;
  rdiag[c] = c < nb ? 1.0f / D[c][c] : 0.0f;             // the division leaves the serial chain
  for (int r = 0; r < NB; ++r) Xs[r][c] = 0.0f;          // rows not solved yet read as zero: the k loop needs no tail
  __syncthreads();
  for (int r = 0; r < nb; ++r) {
    float s[4] = {(r == c) ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f};
    for (int k = 0; k < r; k += 8) {                      // k .. k+7 <= 63; D[r][k > r] = 0 and Xs[k >= r][c] = 0
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u & 3] = fmaf(-D[r][k + u], Xs[k + u][c], s[u & 3]);
    }
    const float x = ((s[0] + s[1]) + (s[2] + s[3])) * rdiag[r];
    Xs[r][c] = x;                        // only this thread reads column c again
    if (c <= r) Li[(size_t)(i0 + r) * ldi + i0 + c] = x;
  }
}
// level with half-size s: for pair p, rows R = [(2p+1)s, min(N,(2p+2)s)), cols C = [2ps, (2p+1)s):
//   T = L[R, C] * Linv[C, C]          (Linv[C,C] lower triangular: k >= column)
//   Linv[R, C] = -Linv[R, R] * T      (Linv[R,R] lower triangular: k <= row)
struct LoadMat {      // element (m,k) of a row-major sub-matrix
  static constexpr bool k_contig = true;
  const float* P; int ld; int rows, cols;
  __device__ __forceinline__ float operator()(int m, int k) const { return (m < rows && k < cols) ? __ldg(P + (size_t)m * ld + k) : 0.0f; }
};
struct LoadMatB {     // element (k,n) of a row-major sub-matrix
  static constexpr bool k_contig = false;
  const float* P; int ld; int rows, cols;
  __device__ __forceinline__ float operator()(int k, int n) const { return (k < rows && n < cols) ? __ldg(P + (size_t)k * ld + n) : 0.0f; }
};
template <int STEP, class T>
__global__ void __launch_bounds__(256) trinv_level_kernel(const float* __restrict__ L, int ldl, int N, int s, float* __restrict__ Li,
                                                          int ldi, float* __restrict__ Tmp, int ldt) {
  __shared__ typename T::Smem sm;
  const int p = blockIdx.z;
  const int c0 = 2 * p * s, r0 = c0 + s;
  if (r0 >= N) return;
  const int nr = min(s, N - r0);
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  if (m0 >= nr || n0 >= s) return;
  float acc[T::TM][T::TN];
  if (STEP == 0) {
    LoadMat a{L + (size_t)r0 * ldl + c0, ldl, nr, s};
    LoadMatB b{Li + (size_t)c0 * ldi + c0, ldi, s, s};
    simt_gemm_tile<T>(sm, a, b, m0, n0, n0 /*k >= first column of the tile*/, s, acc);
  } else {
    LoadMat a{Li + (size_t)r0 * ldi + r0, ldi, nr, nr};
    LoadMatB b{Tmp + (size_t)r0 * ldt + c0, ldt, nr, s};
    simt_gemm_tile<T>(sm, a, b, m0, n0, 0, min(nr, m0 + T::BM) /*k <= last row of the tile*/, acc);
  }
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= nr) continue;
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (n >= s) continue;
      if (STEP == 0) Tmp[(size_t)(r0 + m) * ldt + c0 + n] = acc[i][jj];
      else Li[(size_t)(r0 + m) * ldi + c0 + n] = -acc[i][jj];
    }
  }
}

// The same level with the K range of every tile split over a thread-block cluster (simt_gemm.cuh); the pair index is folded
// into blockIdx.y (grid: (column tiles, row tiles * pairs, splits)). At N = 328 the three levels are six launches of 1-16
// tiles with K = 64-256: 7-18 us each, a quarter of an LML+gradient evaluation.
template <int STEP>
__global__ void __launch_bounds__(256) trinv_level_split_kernel(const float* __restrict__ L, int ldl, int N, int s,
                                                                float* __restrict__ Li, int ldi, float* __restrict__ Tmp, int ldt,
                                                                int tiles_y, int splits) {
  using T = TileSmall;
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  const int p = blockIdx.y / tiles_y, by = blockIdx.y % tiles_y;
  const int c0 = 2 * p * s, r0 = c0 + s;
  if (r0 >= N) return;                       // uniform over the cluster (same x, y)
  const int nr = min(s, N - r0);
  const int m0 = by * T::BM, n0 = blockIdx.x * T::BN;
  if (m0 >= nr || n0 >= s) return;
  float acc[T::TM][T::TN];
  int kb, ke;
  if (STEP == 0) {
    LoadMat a{L + (size_t)r0 * ldl + c0, ldl, nr, s};
    LoadMatB b{Li + (size_t)c0 * ldi + c0, ldi, s, s};
    simt_split_range<T>(s - n0, splits, kb, ke);                       // k >= first column of the tile
    simt_gemm_tile<T>(sm, a, b, m0, n0, n0 + kb, n0 + ke, acc);
  } else {
    LoadMat a{Li + (size_t)r0 * ldi + r0, ldi, nr, nr};
    LoadMatB b{Tmp + (size_t)r0 * ldt + c0, ldt, nr, s};
    simt_split_range<T>(min(nr, m0 + T::BM), splits, kb, ke);          // k <= last row of the tile
    simt_gemm_tile<T>(sm, a, b, m0, n0, kb, ke, acc);
  }
  simt_cluster_reduce<T>(red, acc, splits, [&](int r, int c, float x) {
    const int m = m0 + r, n = n0 + c;
    if (m < nr && n < s) {
      if (STEP == 0) Tmp[(size_t)(r0 + m) * ldt + c0 + n] = x;
      else Li[(size_t)(r0 + m) * ldi + c0 + n] = -x;
    }
  });
}

// ---- triangular products with Linv ------------------------------------------------------------------
// out = Linv * B (trans=0, k <= row) or Linv^T * B (trans=1, k >= row); optional sum of squares of out
template <int TRANS>
struct LoadLinvA {
  static constexpr bool k_contig = (TRANS == 0);
  const float* Li; int ldi; int N;
  __device__ __forceinline__ float operator()(int m, int k) const {
    if (m >= N || k >= N) return 0.0f;
    return TRANS ? __ldg(Li + (size_t)k * ldi + m) : __ldg(Li + (size_t)m * ldi + k);
  }
};
template <int TRANS, class T>
__global__ void __launch_bounds__(256) trmm_kernel(const float* __restrict__ Li, int ldi, int N, const float* __restrict__ B, int ldb,
                                                   int nrhs, float* __restrict__ out, int ldo, float* __restrict__ ssq) {
  __shared__ typename T::Smem sm;
  __shared__ float red[8];
  // the k range grows with the row block (TRANS=0): launch the long tiles first so they do not form the tail
  const int by = TRANS ? blockIdx.y : gridDim.y - 1 - blockIdx.y;
  const int m0 = by * T::BM, n0 = blockIdx.x * T::BN;
  LoadLinvA<TRANS> a{Li, ldi, N};
  LoadMatB b{B, ldb, N, nrhs};
  float acc[T::TM][T::TN];
  const int kb = TRANS ? m0 : 0;
  const int ke = TRANS ? N : min(N, m0 + T::BM);
  simt_gemm_tile<T>(sm, a, b, m0, n0, kb, ke, acc);
  float local = 0.0f;
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= N) continue;
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (n < nrhs) { out[(size_t)m * ldo + n] = acc[i][jj]; local = fmaf(acc[i][jj], acc[i][jj], local); }
    }
  }
  if (ssq != nullptr) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int w = 0; w < 8; ++w) t += red[w];
      atomicAdd(ssq, t);
    }
  }
}

// 64 x 64 tiles with the (triangular) K range of the tile split over a thread-block cluster (simt_gemm.cuh): the horses-sized
// products (N = 328, D = 1024) are 96 tiles walking up to 328 columns each — latency-bound at 23-26 us per launch
template <int TRANS>
__global__ void __launch_bounds__(256) trmm_split_kernel(const float* __restrict__ Li, int ldi, int N, const float* __restrict__ B,
                                                         int ldb, int nrhs, float* __restrict__ out, int ldo,
                                                         float* __restrict__ ssq, int splits) {
  using T = TileSmall;
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  __shared__ float wsum[8];
  const int by = TRANS ? blockIdx.y : gridDim.y - 1 - blockIdx.y;
  const int m0 = by * T::BM, n0 = blockIdx.x * T::BN;
  LoadLinvA<TRANS> a{Li, ldi, N};
  LoadMatB b{B, ldb, N, nrhs};
  float acc[T::TM][T::TN];
  const int kb = TRANS ? m0 : 0;
  const int ke = TRANS ? N : min(N, m0 + T::BM);
  int sb, se;
  simt_split_range<T>(ke - kb, splits, sb, se);
  simt_gemm_tile<T>(sm, a, b, m0, n0, kb + sb, kb + se, acc);
  float local = 0.0f;
  simt_cluster_reduce<T>(red, acc, splits, [&](int r, int c, float x) {
    const int m = m0 + r, n = n0 + c;
    if (m < N && n < nrhs) { out[(size_t)m * ldo + n] = x; local = fmaf(x, x, local); }
  });
  if (ssq != nullptr) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int w = 0; w < 8; ++w) t += wsum[w];
      atomicAdd(ssq, t);
    }
  }
}

__global__ void logdet_xsq_kernel(const float* __restrict__ L, int ldl, int N, const float* __restrict__ X, int nx, GpAccum* acc) {
  float s = 0.0f, x2 = 0.0f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += logf(__ldg(L + (size_t)i * ldl + i));
  for (int i = threadIdx.x; i < nx; i += blockDim.x) { const float v = __ldg(X + i); x2 = fmaf(v, v, x2); }
  __shared__ float r1[32], r2[32];
  s = warp_sum(s); x2 = warp_sum(x2);
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s; r2[threadIdx.x >> 5] = x2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, b = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += r1[w]; b += r2[w]; }
    acc->logdet = 2.0f * a;      // gplvm.py:226-228
    acc->xsq = b;
  }
}

// ---- K6: fused gradient -------------------------------------------------------------------------------
// Tile (I >= J) of G = 0.5 (D * Linv^T Linv - A A^T), multiplied by K_f on the fly, reduced into
// dX (atomics), tr G, sum W, sum W r^2.   
struct LoadLinvT_A {   // a(m,k) = (D/2) * Linv[k, m]
  static constexpr bool k_contig = false;
  const float* Li; int ldi; int N; float scale;
  __device__ __forceinline__ float operator()(int m, int k) const { return (m < N && k < N) ? scale * __ldg(Li + (size_t)k * ldi + m) : 0.0f; }
};
struct LoadLinvT_B {   // b(k,n) = Linv[k, n]
  static constexpr bool k_contig = false;
  const float* Li; int ldi; int N;
  __device__ __forceinline__ float operator()(int k, int n) const { return (n < N && k < N) ? __ldg(Li + (size_t)k * ldi + n) : 0.0f; }
};
struct LoadA_A {       // a(m,k) = -0.5 * Avec[m, k]
  static constexpr bool k_contig = true;
  const float* A; int ld; int N, D;
  __device__ __forceinline__ float operator()(int m, int k) const { return (m < N && k < D) ? -0.5f * __ldg(A + (size_t)m * ld + k) : 0.0f; }
};
struct LoadA_B {       // b(k,n) = Avec[n, k]
  static constexpr bool k_contig = true;
  const float* A; int ld; int N, D;
  __device__ __forceinline__ float operator()(int k, int n) const { return (n < N && k < D) ? __ldg(A + (size_t)n * ld + k) : 0.0f; }
};

// Contract one lower tile (I >= J) of a SYMMETRIC matrix G with K_f on the fly:
//   dXacc_i += -(2/gamma^2) sum_j G_ij Kf_ij (x_i - x_j),  trG, sum G o Kf, sum G o Kf o r^2  
template <class T>
__device__ __forceinline__ void gp_contract_tile(float (&gacc)[T::TM][T::TN], int m0, int n0, bool diag_tile,
                                                 const float* __restrict__ X, int N, int Q, const float* __restrict__ hyp,
                                                 float* __restrict__ dXacc, GpAccum* accum, float (*rowacc)[MAXQ],
                                                 float (*colacc)[MAXQ], float (*red)[8]) {
  const float alpha = __ldg(hyp), inv_g2 = 1.0f / (__ldg(hyp + 1) * __ldg(hyp + 1));
  float trg = 0.0f, sw = 0.0f, swr = 0.0f;
  if (Q <= 2) {
    // fast path (every example uses Q = 2): per-thread row / column partial sums in registers, reduced with
    // shuffles over the 16 threads that share a row, shared-memory atomics only once per warp and column
    constexpr int QF = 2;
    float racc[T::TM][QF], cacc[T::TN][QF];
#pragma unroll
    for (int i = 0; i < T::TM; ++i) { racc[i][0] = racc[i][1] = 0.0f; }
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) { cacc[jj][0] = cacc[jj][1] = 0.0f; }
    float xn[T::TN][QF];
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      xn[jj][0] = n < N ? __ldg(X + (size_t)n * Q) : 0.0f;
      xn[jj][1] = (n < N && Q > 1) ? __ldg(X + (size_t)n * Q + 1) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < T::TM; ++i) {
      const int m = simt_row<T>(m0, i);
      const float xm0 = m < N ? __ldg(X + (size_t)m * Q) : 0.0f;
      const float xm1 = (m < N && Q > 1) ? __ldg(X + (size_t)m * Q + 1) : 0.0f;
#pragma unroll
      for (int jj = 0; jj < T::TN; ++jj) {
        const int n = simt_col<T>(n0, jj);
        if (m >= N || n >= N || (diag_tile && n > m)) continue;
        const float g = gacc[i][jj];
        if (m == n) { trg += g; sw += g * alpha; continue; }
        const float d0 = xm0 - xn[jj][0], d1 = xm1 - xn[jj][1];
        const float r2 = fmaf(d0, d0, d1 * d1);
        const float w = g * alpha * __expf(-0.5f * r2 * inv_g2);
        sw += 2.0f * w; swr += 2.0f * w * r2;
        racc[i][0] = fmaf(w, d0, racc[i][0]); racc[i][1] = fmaf(w, d1, racc[i][1]);
        cacc[jj][0] = fmaf(-w, d0, cacc[jj][0]); cacc[jj][1] = fmaf(-w, d1, cacc[jj][1]);
      }
    }
    const int tx = threadIdx.x % (T::BN / T::TN), lane = threadIdx.x & 31;
    static_assert(T::BN / T::TN == 16, "row reduction assumes 16 threads per tile row");
#pragma unroll
    for (int i = 0; i < T::TM; ++i) {
#pragma unroll
      for (int q = 0; q < QF; ++q) {
        float v = racc[i][q];
        v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8);
        if (tx == 0 && q < Q) rowacc[simt_row<T>(0, i)][q] = v;        // rows are owned by one 16-thread group
      }
    }
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
#pragma unroll
      for (int q = 0; q < QF; ++q) {
        float v = cacc[jj][q];
        v += __shfl_xor_sync(0xffffffffu, v, 16);                       // the two tile rows held by this warp
        if (lane < 16 && q < Q) atomicAdd(&colacc[simt_col<T>(0, jj)][q], v);
      }
    }
  } else {
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= N) continue;
    float xm[MAXQ];
    for (int q = 0; q < Q; ++q) xm[q] = __ldg(X + (size_t)m * Q + q);
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (n >= N || (diag_tile && n > m)) continue;
      const float g = gacc[i][jj];
      if (m == n) { trg += g; sw += g * alpha; continue; }   // Kf_ii = alpha, r = 0: no dX contribution
      float r2 = 0.0f, df[MAXQ];
      for (int q = 0; q < Q; ++q) { df[q] = xm[q] - __ldg(X + (size_t)n * Q + q); r2 = fmaf(df[q], df[q], r2); }
      const float w = g * alpha * __expf(-0.5f * r2 * inv_g2);
      sw += 2.0f * w; swr += 2.0f * w * r2;                  // (i,j) and (j,i)
      for (int q = 0; q < Q; ++q) {
        atomicAdd(&rowacc[m - m0][q], w * df[q]);
        atomicAdd(&colacc[n - n0][q], -w * df[q]);
      }
    }
  }
  }
  __syncthreads();
  // dX_i = -(2/gamma^2) sum_j W_ij (x_i - x_j); each unordered pair was visited once -> both rows updated
  const float sc = -2.0f * inv_g2;
  for (int t = threadIdx.x; t < T::BM * Q; t += 256) {
    const int r = t / Q, q = t % Q;
    if (m0 + r < N) atomicAdd(dXacc + (size_t)(m0 + r) * Q + q, sc * rowacc[r][q]);
    if (n0 + r < N) atomicAdd(dXacc + (size_t)(n0 + r) * Q + q, sc * colacc[r][q]);
  }
  trg = warp_sum(trg); sw = warp_sum(sw); swr = warp_sum(swr);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = trg; red[1][threadIdx.x >> 5] = sw; red[2][threadIdx.x >> 5] = swr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
    atomicAdd(&accum->trG, a); atomicAdd(&accum->sumW, b); atomicAdd(&accum->sumWr2, c);
  }
}

__global__ void __launch_bounds__(256) gp_grad_kernel(const float* __restrict__ Li, int ldi, const float* __restrict__ Av, int lda,
                                                      const float* __restrict__ X, int N, int Q, int D,
                                                      const float* __restrict__ hyp, float* __restrict__ dXacc, GpAccum* accum) {
  using T = TileLarge;
  if (blockIdx.x > blockIdx.y) return;
  __shared__ typename T::Smem sm;
  __shared__ float rowacc[T::BM][MAXQ];
  __shared__ float colacc[T::BN][MAXQ];
  __shared__ float red[3][8];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  const bool diag_tile = blockIdx.x == blockIdx.y;
  for (int t = threadIdx.x; t < T::BM * MAXQ; t += 256) { (&rowacc[0][0])[t] = 0.0f; (&colacc[0][0])[t] = 0.0f; }

  float gacc[T::TM][T::TN];   // G tile = (D/2) Linv^T Linv - (1/2) A A^T, one accumulator for both products
  {
    LoadLinvT_A a{Li, ldi, N, 0.5f * (float)D};
    LoadLinvT_B b{Li, ldi, N};
    simt_gemm_tile<T>(sm, a, b, m0, n0, m0 /* Linv[k,i] = 0 for k < i and i >= m0 */, N, gacc);
  }
  __syncthreads();
  {
    LoadA_A a{Av, lda, N, D};
    LoadA_B b{Av, lda, N, D};
    simt_gemm_tile<T, LoadA_A, LoadA_B, false>(sm, a, b, m0, n0, 0, D, gacc);
  }
  gp_contract_tile<T>(gacc, m0, n0, diag_tile, X, N, Q, hyp, dXacc, accum, rowacc, colacc, red);
}

// Small N (the reference's N = 328: a 3 x 3 grid of 128-row tiles, six useful CTAs walking K = N + D serially, 236 us):
// G = (D/2) Linv^T Linv - (1/2) A A^T is formed by 64 x 64 tiles whose K ranges are split over a thread-block cluster
// (simt_gemm.cuh) and written to memory (lower tiles); gp_contract_mem_kernel then contracts it with K_f.
__global__ void __launch_bounds__(256) gp_g_split_kernel(const float* __restrict__ Li, int ldi, const float* __restrict__ Av, int lda,
                                                         int N, int D, float* __restrict__ G, int splits) {
  using T = TileSmall;
  if (blockIdx.x > blockIdx.y) return;      // the CTAs of a cluster share (x, y): they leave together
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  float acc[T::TM][T::TN];
  int kb, ke;
  {
    simt_split_range<T>(N - m0, splits, kb, ke);          // Linv[k,i] = 0 for k < i and i >= m0
    LoadLinvT_A a{Li, ldi, N, 0.5f * (float)D};
    LoadLinvT_B b{Li, ldi, N};
    simt_gemm_tile<T>(sm, a, b, m0, n0, m0 + kb, m0 + ke, acc);
  }
  __syncthreads();
  {
    simt_split_range<T>(D, splits, kb, ke);
    LoadA_A a{Av, lda, N, D};
    LoadA_B b{Av, lda, N, D};
    simt_gemm_tile<T, LoadA_A, LoadA_B, false>(sm, a, b, m0, n0, kb, ke, acc);
  }
  simt_cluster_reduce<T>(red, acc, splits, [&](int r, int c, float x) {
    if (m0 + r < N && n0 + c < N) G[(size_t)(m0 + r) * N + n0 + c] = x;
  });
}

// scalars + prior term of dX
__global__ void gp_finalize_kernel(const GpAccum* __restrict__ acc, const float* __restrict__ hyp, const float* __restrict__ opt,
                                   int flags, int N, int Q, int D, float x_prior_w, int include_const, int want_grad,
                                   const float* __restrict__ X, const float* __restrict__ dXacc, float* __restrict__ dX,
                                   float* __restrict__ out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (want_grad && dX != nullptr)
    for (int i = tid; i < N * Q; i += gridDim.x * blockDim.x) dX[i] = dXacc[i] + 2.0f * x_prior_w * __ldg(X + i);
  if (tid == 0) {
    const float alpha = hyp[0], gamma = hyp[1];
    float loss = 0.5f * (acc->ssq + (float)D * acc->logdet);
    if (include_const) loss += 0.5f * (float)((double)D * (double)N * 1.8378770664093453);   // D N log(2 pi), gplvm.py:238
    loss += x_prior_w * acc->xsq;                                                          // gprbm.py:242
    out[0] = loss; out[1] = acc->logdet; out[2] = acc->ssq;
    out[3] = (flags & 1) ? (acc->sumW / alpha) * sigmoidf_(opt[0]) : 0.0f;
    out[4] = (flags & 2) ? (acc->sumWr2 / (gamma * gamma * gamma)) * sigmoidf_(opt[1]) : 0.0f;
    out[5] = (flags & 4) ? acc->trG * sigmoidf_(opt[2]) : 0.0f;
    out[6] = alpha; out[7] = hyp[2];
  }
}

// predictive variance: var[m] = alpha - sum_k V1[k,m]^2 ; flags the first negative entry (gplvm.py:283-286)
__global__ void pred_var_kernel(const float* __restrict__ V1, int N, int M, const float* __restrict__ hyp, float* __restrict__ var,
                                int* __restrict__ info) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = 0.0f;
  for (int k = 0; k < N; ++k) { const float v = __ldg(V1 + (size_t)k * M + m); s = fmaf(v, v, s); }
  const float pv = __ldg(hyp) - s;
  var[m] = pv;
  if (pv < 0.0f && info) atomicCAS(info, 0, m + 1);
}
// mean[m,d] = sum_k V1[k,m] S[k,d] + y_mean[d]
struct LoadV1T {
  static constexpr bool k_contig = false;
  const float* V1; int N, M;
  __device__ __forceinline__ float operator()(int m, int k) const { return (m < M && k < N) ? __ldg(V1 + (size_t)k * M + m) : 0.0f; }
};
__global__ void __launch_bounds__(256) pred_mean_kernel(const float* __restrict__ V1, const float* __restrict__ S, int N, int M, int D,
                                                        const float* __restrict__ y_mean, float* __restrict__ mean) {
  using T = TileSmall;
  __shared__ typename T::Smem sm;
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  LoadV1T a{V1, N, M};
  LoadMatB b{S, D, N, D};
  float acc[T::TM][T::TN];
  simt_gemm_tile<T>(sm, a, b, m0, n0, 0, N, acc);
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= M) continue;
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (n < D) mean[(size_t)m * D + n] = acc[i][jj] + (y_mean ? __ldg(y_mean + n) : 0.0f);
    }
  }
}

// ---- host-side building blocks -----------------------------------------------------------------------
int run_gram(db_handle_t h, cudaStream_t st, const float* X, const float* Y, int N, int M, int Q, const float* hyp,
             int add_noise, float* out, int ldo) {
  const int rows = ((long)((M + 255) / 256) * ((N + 127) / 128) >= 2L * h->num_sms) ? 128 : 32;
  dim3 grid((M + 255) / 256, (N + rows - 1) / rows);
  const size_t smem = sizeof(float) * (256 + rows) * Q;
  switch (Q) {
    case 1: gram_kernel<1><<<grid, 256, smem, st>>>(X, Y, N, M, Q, hyp, add_noise, out, ldo, rows); break;
    case 2: gram_kernel<2><<<grid, 256, smem, st>>>(X, Y, N, M, Q, hyp, add_noise, out, ldo, rows); break;
    case 3: gram_kernel<3><<<grid, 256, smem, st>>>(X, Y, N, M, Q, hyp, add_noise, out, ldo, rows); break;
    case 4: gram_kernel<4><<<grid, 256, smem, st>>>(X, Y, N, M, Q, hyp, add_noise, out, ldo, rows); break;
    default: gram_kernel<0><<<grid, 256, smem, st>>>(X, Y, N, M, Q, hyp, add_noise, out, ldo, rows); break;
  }
  return check_launch(h, "gram_rbf");
}

// lo_buf (optional, same shape / leading dimension as A): scratch for the tf32 remainders of the panels; when given
// and the operands are TMA-addressable, the wide K = 512 trailing updates run on the tensor cores (3xTF32).
// Scratch of the 3 x fp16 products (gemm_f16x3.cu): scales[0..2] a-priori (L panels, L^-1, K^-1), [3] Y, [4] A from their
// measured maxima; fp16 hi/lo pairs of the current Cholesky panel, of U = L^-T, of K^-1, of Y^T and of A.
struct F16Scratch {
  float* scales; unsigned int* bits;
  __half *l_hi, *l_lo;      // [N,N]: the Cholesky panels (strictly lower 512-blocks, scale[0]); diagonal blocks: Linv_jj (scale[1])
  __half *u_hi, *u_lo;      // [N,N]: U = L^-T (scale[1])
  __half *wt_hi, *wt_lo;    // [N,OB]: block-row products of the triangular inverse (scale[5])
  __half *k_hi, *k_lo, *y_hi, *y_lo, *a_hi, *a_lo;   // K^-1 [N,N] (scale[2]), Y^T [D,N] (scale[3]), A [N,D] (scale[4])
};
inline size_t f16_scratch_bytes(int N, int D) {
  const size_t nn = (size_t)N * N, nd = (size_t)N * D, no = (size_t)N * OB;
  return 256 + 2 * sizeof(__half) * (no + 3 * nn + 2 * nd) + 12 * 256;
}
inline F16Scratch f16_scratch(void* base, int N, int D) {
  F16Scratch f;
  uint8_t* p = static_cast<uint8_t*>(base);
  auto take = [&](size_t halves) { __half* q = reinterpret_cast<__half*>(p); p += (halves * sizeof(__half) + 255) & ~size_t(255); return q; };
  f.scales = reinterpret_cast<float*>(p); f.bits = reinterpret_cast<unsigned int*>(p + 64); p += 256;
  const size_t nn = (size_t)N * N, nd = (size_t)N * D, no = (size_t)N * OB;
  f.l_hi = take(nn); f.l_lo = take(nn); f.u_hi = take(nn); f.u_lo = take(nn); f.wt_hi = take(no); f.wt_lo = take(no);
  f.k_hi = take(nn); f.k_lo = take(nn); f.y_hi = take(nd); f.y_lo = take(nd); f.a_hi = take(nd); f.a_lo = take(nd);
  return f;
}
// the fp16 form needs 16-byte row pitches of the fp16 matrices (N, D multiples of 8); DB_GP_F16=0 keeps 3xTF32 everywhere
inline bool gp_use_f16(int N, int D) {
  static const bool off = [] { const char* e = getenv("DB_GP_F16"); return e && e[0] == '0'; }();
  return !off && (N % 8) == 0 && (D % 8) == 0;
}

// Triangular inverse riding along the factorisation on the handle's two auxiliary streams (3 x fp16 chain only; see
// run_trinv_f16_side_block below). reserve(blk): SMs the launches of outer block blk leave to those streams.
struct TrinvOverlap {
  static constexpr int kMaxBlocks = 44;
  static constexpr int kMinReserve = 16;
  cudaStream_t diag, prod;  // lane of the diagonal-block inverses, lane of the block-row products
  cudaEvent_t* ev;          // [0] fork, [1 + blk] "outer block blk is final", [1 + kMaxBlocks + blk] "Linv_blk is ready", [94] join
  float* Li; float* U;
  int mode;                 // 1: the lanes also run under the trailing updates (whose grid is capped), 2: only under the steps
  int trail_reserve;        // SMs a trailing update leaves free in mode 1
  int num_sms, N;
  int pending;              // first slice not issued yet (slices wait while the step launches need every SM)
  int err;                  // first non-zero return code of a side product launch (checked by the caller after the join)
  // The step launches of outer block blk want one strip CTA per 64 rows below the diagonal plus about as many in-panel update
  // CTAs; what is left belongs to the lanes. Below kMinReserve nothing is reserved and the slices are deferred.
  int reserve(int blk) const {
    // strip + update CTAs per strip CTA; measured at N = 5000 (ms per evaluation): 1.5 / 1.75 / 2 / 2.5 / 3 -> 2.44 / 2.34 / 2.30 /
    // 2.36 / 2.74 (N = 6144: 3.35 / - / 3.19 / 3.62 / -)
    static const float ratio = [] { const char* e = getenv("DB_GP_UPD_RATIO"); return e ? (float)atof(e) : 2.0f; }();
    const int strips0 = (N - blk * OB + NB - 1) / NB;
    const int r = num_sms - (int)(ratio * (float)strips0 + 0.5f);
    if (r < kMinReserve) return 0;
    return r > num_sms - 8 ? num_sms - 8 : r;
  }
  cudaEvent_t block_final(int blk) const { return ev[1 + blk]; }
  cudaEvent_t linv_ready(int blk) const { return ev[1 + kMaxBlocks + blk]; }
  cudaEvent_t join() const { return ev[94]; }
};
static_assert(2 * TrinvOverlap::kMaxBlocks + 1 < 94 && db_context_s::kSideEvents > 94, "event pool layout");
// issues the slices that can start now; returns true when any was issued (the lanes are then busy beside the next launches)
bool run_trinv_f16_side_blocks(db_handle_t h, TrinvOverlap& ov, cudaStream_t main, float* L, int blk, const F16Scratch& f);

int run_potrf(db_handle_t h, cudaStream_t st, float* A, int lda, int N, int* info, float* lo_buf = nullptr,
              const F16Scratch* f16 = nullptr, TrinvOverlap* ov = nullptr) {
  cudaMemsetAsync(info, 0, sizeof(int), st);
  static const bool legacy = [] { const char* e = getenv("DB_POTRF_LEGACY"); return e && e[0] == '1'; }();
  static const bool pdl = [] { const char* e = getenv("DB_POTRF_PDL"); return !(e && e[0] == '0'); }();
  static std::once_flag smem_once;
  std::call_once(smem_once, [] { cudaFuncSetAttribute(potrf_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEP_SMEM); });
  // Tail merge (3 x fp16 chain): once at most tail_rows rows remain, the outer blocks stop handing their trailing update to a
  // GEMM launch (~20 us of launch latency for ~1 us of work there); the update CTAs of the step launches apply every panel
  // to ALL remaining columns instead. Measured at N = 5000 (ms per evaluation, DB_POTRF_TAIL_ROWS = 0 / 1000 / 1500 / 2000 /
  // 3000): 2.273 / 2.257 / 2.269 / 2.349 / 2.722 with the side lanes, 2.711 / - / 2.684 on one stream.
  static const int tail_rows = [] { const char* e = getenv("DB_POTRF_TAIL_ROWS"); return e ? atoi(e) : 1000; }();
  for (int b0 = 0; b0 < N; b0 += OB) {
    const int b1 = b0 + OB < N ? b0 + OB : N;
    const bool merged = f16 != nullptr && tail_rows > 0 && b0 > 0 && N - b0 <= tail_rows;
    const bool prev_merged = merged && b0 - OB > 0 && N - (b0 - OB) <= tail_rows;
    const int upd_end = merged ? N : b1;          // columns the update CTAs of this block's launches cover
    static const bool skip_steps = [] { const char* e = getenv("DB_POTRF_SKIP_STEPS"); return e && e[0] == '1'; }();   // timing diagnostics only
    for (int j = b0; j < b1 && !skip_steps; j += NB) {
      if (legacy) {
        potrf_colblock_kernel<<<(N - j + NB - 1) / NB, NB, 0, st>>>(A, lda, N, j, info);
        const int j2 = j + NB;
        if (j2 < b1) {     // narrow update: remaining columns of this outer block, all rows below
          dim3 grid((b1 - j2 + TileLarge::BN - 1) / TileLarge::BN, (N - j2 + TileLarge::BM - 1) / TileLarge::BM);
          syrk_block_kernel<<<grid, 256, 0, st>>>(A, lda, N, j2, j2, b1, j, j2);
        }
        continue;
      }
      const int strips = (N - j + NB - 1) / NB;
      const int jprev = j > b0 ? j - NB : (prev_merged ? j - NB : -1);
      int updates = 0;                 // update CTAs: each loops over 128x64 tiles of the columns [j + 64, upd_end)
      if (jprev >= 0 && j + NB < upd_end) {
        int tiles = 0;
        for (int cb = j + NB; cb < upd_end; cb += NB) tiles += (N - cb + 2 * NB - 1) / (2 * NB);
        const int free_sms = h->num_sms - strips - (ov ? ov->reserve(b0 / OB) : 0);
        updates = merged ? (tiles + 2) / 3 : tiles;
        const int cap = free_sms > 8 ? free_sms : 8;
        if (updates > cap) updates = cap;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(strips + updates); cfg.blockDim = dim3(STEP_THREADS); cfg.dynamicSmemBytes = STEP_SMEM; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = (pdl && strips + updates <= 2 * h->num_sms) ? 1 : 0;   // measured: slower on larger grids
      cfg.attrs = attr; cfg.numAttrs = 1;
      long long* dbg = nullptr;
#ifdef DB_POTRF_TIMING
      static long long* dbg_buf = nullptr;
      static int dbg_step = 0;
      if (getenv("DB_POTRF_DEBUG")) {
        if (!dbg_buf) { cudaMalloc(&dbg_buf, 8 * 128 * sizeof(long long)); }
        if (j == 0) {
          if (dbg_step > 0) {   // dump the previous factorisation
            static long long hbuf[8 * 128];
            cudaStreamSynchronize(st);
            cudaMemcpy(hbuf, dbg_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
            for (int i = 0; i < dbg_step && i < 128; ++i) {
              const long long* d = hbuf + 8 * i;
              fprintf(stderr, "step %3d: load %6lld upd %6lld factor %6lld solve-end %6lld total %6lld clk | since prev end %7.2f us\n", i,
                      d[1] - d[0], d[2] - d[1], d[3] - d[2], d[4] - d[2], d[5] - d[0], i ? (d[6] - hbuf[8 * (i - 1) + 6]) * 1e-3 : 0.0);
            }
          }
          dbg_step = 0;
          cudaMemsetAsync(dbg_buf, 0, 8 * 128 * sizeof(long long), st);
        }
        if (dbg_step < 128) dbg = dbg_buf + 8 * dbg_step;
        ++dbg_step;
      }
#endif
      cudaLaunchKernelEx(&cfg, potrf_step_kernel, A, lda, N, j, jprev, upd_end, strips, info, dbg);
    }
    static const bool skip_wide = [] { const char* e = getenv("DB_POTRF_SKIP_WIDE"); return e && e[0] == '1'; }();   // timing diagnostics only
    if (b1 < N && !skip_wide) {        // wide update of everything right of the outer block, K = b1 - b0
      const int rows = N - b1;
      const float* P = A + (size_t)b1 * lda + b0;
      bool done = false;
      if (lo_buf != nullptr && h->cc_major >= 10) {
        // the remainders are kept for every panel: the tensor-core triangular inverse reads them again (run_trinv_tc)
        float* Plo = lo_buf + (size_t)b1 * lda + b0;
        if (f16 == nullptr) tf32_split_lo(h, st, P, rows, b1 - b0, lda, Plo);
        if (f16 != nullptr) {
          // the panel as a scaled fp16 pair (|L_ij| <= sqrt(alpha + s), scales[0]), kept in the N x N pair for the triangular
          // inverse; trailing update as 3 x fp16
          const F16Operand P16 = F16Operand{f16->l_hi, f16->l_lo, f16->scales}.at((size_t)b1 * lda + b0);
          f16_split(h, st, P, rows, b1 - b0, lda, P16, lda);
          int cap = 0;
          if (ov && ov->mode == 1) {       // block b0 / OB is final and converted: its part of the inverse starts now
            if (run_trinv_f16_side_blocks(h, *ov, st, A, b0 / OB, *f16)) cap = h->num_sms - ov->trail_reserve;
          }
          int rc = -1000;
          if (merged) rc = 0;    // the step launches of this and the following blocks carry the update
          else if (rows >= 128)       // (the 3xTF32 form kept the SIMT kernel below 512 rows: 71 us for the 392-row tail at N = 5000)
            rc = f16_gemm_launch(h, st, rows, rows, b1 - b0, -1.0f, P16, lda, P16, lda, 1.0f, A + (size_t)b1 * lda + b1, lda,
                                 /*lower_only=*/1, /*tri_k=*/0, nullptr, 0, cap);
          if (rc != 0 && rc != -1000) return set_err(h, DB_ERR_CUDA, "f16x3 trailing update failed (%d)", rc);
          done = rc == 0;
        }
        if (!done && f16 == nullptr && rows >= 512 && tf32_gemm_supported(rows, rows, b1 - b0, P, Plo, lda, P, Plo, lda)) {
          const int rc = tf32_gemm_launch(h, st, rows, rows, b1 - b0, -1.0f, P, Plo, lda, P, Plo, lda, 1.0f,
                                          A + (size_t)b1 * lda + b1, lda, /*lower_only=*/1, /*tri_k=*/0);
          if (rc != 0 && rc != -1000) return set_err(h, DB_ERR_CUDA, "tf32 trailing update failed (%d)", rc);
          done = rc == 0;
        }
      }
      if (!done) {
        const int t = (rows + TileLarge::BM - 1) / TileLarge::BM;
        syrk_block_kernel<<<dim3(t, t), 256, 0, st>>>(A, lda, N, b1, b1, N, b0, b1);
      }
    }
    if (ov && f16 != nullptr && (ov->mode != 1 || b1 >= N)) {     // mode 2, and the last block (no trailing update) in either mode
      run_trinv_f16_side_blocks(h, *ov, st, A, b0 / OB, *f16);
    }
  }
  return check_launch(h, "potrf_lower");
}

int run_trinv(db_handle_t h, cudaStream_t st, const float* L, int ldl, int N, float* Li, float* Tmp) {
  cudaMemsetAsync(Li, 0, sizeof(float) * (size_t)N * N, st);
  const int nblk = (N + NB - 1) / NB;
  trinv_diag_kernel<<<nblk, 64, 0, st>>>(L, ldl, N, Li, N);
  for (int s = NB; s < N; s *= 2) {
    const int pairs = (N + 2 * s - 1) / (2 * s);
    if (false && s >= 1024) {  // measured: the large tile gives too few CTAs at these level sizes (slower)
      dim3 grid((s + TileLarge::BN - 1) / TileLarge::BN, (s + TileLarge::BM - 1) / TileLarge::BM, pairs);
      trinv_level_kernel<0, TileLarge><<<grid, 256, 0, st>>>(L, ldl, N, s, Li, N, Tmp, N);
      trinv_level_kernel<1, TileLarge><<<grid, 256, 0, st>>>(L, ldl, N, s, Li, N, Tmp, N);
    } else {
      dim3 grid((s + TileSmall::BN - 1) / TileSmall::BN, (s + TileSmall::BM - 1) / TileSmall::BM, pairs);
      const int splits = simt_splits((long)grid.x * grid.y * pairs, s, h->num_sms, 32);
      if (splits > 1) {
        const int tiles_y = (int)grid.y;
        dim3 g2(grid.x, grid.y * pairs, splits);
        simt_launch_split(trinv_level_split_kernel<0>, g2, 256, st, L, ldl, N, s, Li, N, Tmp, N, tiles_y, splits);
        simt_launch_split(trinv_level_split_kernel<1>, g2, 256, st, L, ldl, N, s, Li, N, Tmp, N, tiles_y, splits);
      } else {
        trinv_level_kernel<0, TileSmall><<<grid, 256, 0, st>>>(L, ldl, N, s, Li, N, Tmp, N);
        trinv_level_kernel<1, TileSmall><<<grid, 256, 0, st>>>(L, ldl, N, s, Li, N, Tmp, N);
      }
    }
  }
  return check_launch(h, "trinv_lower");
}

// ---- tensor-core triangular inverse (N >= 512, N % 4 == 0) ---------------------------------------------------------
// Produces U = Linv^T (upper triangular, what the K^-1 = U U^T product reads) and its tf32 remainder, block row by block
// row of Linv with 512-wide blocks B_j (P_j = 512 j rows precede block j):
//   Linv_jj           : SIMT recursive doubling inside every diagonal block (levels 64, 128, 256), written to Li
//   WT  = U[:P_j,:P_j] L[B_j,:P_j]^T               (3xTF32 GEMM; X upper triangular -> K loop clipped per tile)
//   U[:P_j, B_j] = -WT Linv_jj^T                    (3xTF32 GEMM, remainder written by the epilogue)
// L and the remainders of its panels (left behind by run_potrf in lo_buf) are the only inputs; the diagonal blocks of the
// L buffer are overwritten with the remainders of Linv_jj (L_jj is dead once Linv_jj exists).
// grid (16, 16, nblk): 32x32 tiles of diagonal block z
// f16 form (li_hi != nullptr): instead of the tf32 remainders, Linv_jj and U_jj are written again as scaled fp16 pairs (scale
// *sc16 for both: |L^-1| <= 1/sqrt(s)) at the same positions of the N x N fp16 matrices.
__global__ void __launch_bounds__(256) trinv_diag_finish_kernel(const float* __restrict__ Li, int N, float* __restrict__ LiLo,
                                                                float* __restrict__ U, float* __restrict__ Ulo,
                                                                __half* __restrict__ li_hi = nullptr, __half* __restrict__ li_lo = nullptr,
                                                                __half* __restrict__ u_hi = nullptr, __half* __restrict__ u_lo = nullptr,
                                                                const float* __restrict__ sc16 = nullptr, int first_block = 0) {
  __shared__ float tile[32][33];
  const float sc = li_hi != nullptr ? __ldg(sc16) : 0.0f;
  const int b0 = (blockIdx.z + first_block) * OB;
  const int nb = min(OB, N - b0);
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  if (r0 >= nb || c0 >= nb) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.0f;
    if (r < nb && c < nb) {
      const size_t idx = (size_t)(b0 + r) * N + b0 + c;
      v = c <= r ? Li[idx] : 0.0f;
      if (li_hi != nullptr) { __half hh, ll; split_f16(v * sc, hh, ll); li_hi[idx] = hh; li_lo[idx] = ll; }
      else LiLo[idx] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;           // element (c, r) of U_jj = element (r, c) of Linv_jj
    if (c < nb && r < nb) {
      const float v = tile[tx][ty + 8 * i];
      const size_t idx = (size_t)(b0 + c) * N + b0 + r;
      U[idx] = v;
      if (u_hi != nullptr) { __half hh, ll; split_f16(v * sc, hh, ll); u_hi[idx] = hh; u_lo[idx] = ll; }
      else Ulo[idx] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    }
  }
}

int run_trinv_tc(db_handle_t h, cudaStream_t st, float* L, const float* Llo, int N, float* Li, float* U, float* Ulo, float* WT,
                 float* WTlo) {
  const int nblk = (N + OB - 1) / OB;
  cudaMemsetAsync(U, 0, sizeof(float) * (size_t)N * N, st);
  cudaMemsetAsync(Ulo, 0, sizeof(float) * (size_t)N * N, st);
  for (int j = 0; j < nblk; ++j) {      // the upper part of the diagonal blocks of Li is read by the GEMMs: zero it
    const int b0 = j * OB, nb = std::min(OB, N - b0);
    cudaMemset2DAsync(Li + (size_t)b0 * N + b0, sizeof(float) * (size_t)N, 0, sizeof(float) * (size_t)nb, nb, st);
  }
  trinv_diag_kernel<<<(N + NB - 1) / NB, 64, 0, st>>>(L, N, N, Li, N);
  for (int s = NB; s < OB && s < N; s *= 2) {
    const int pairs = (N + 2 * s - 1) / (2 * s);
    dim3 grid((s + TileSmall::BN - 1) / TileSmall::BN, (s + TileSmall::BM - 1) / TileSmall::BM, pairs);
    // scratch T of the SIMT levels: the lower part of U's diagonal blocks (overwritten by trinv_diag_finish_kernel)
    trinv_level_kernel<0, TileSmall><<<grid, 256, 0, st>>>(L, N, N, s, Li, N, U, N);
    trinv_level_kernel<1, TileSmall><<<grid, 256, 0, st>>>(L, N, N, s, Li, N, U, N);
  }
  trinv_diag_finish_kernel<<<dim3(OB / 32, OB / 32, nblk), 256, 0, st>>>(Li, N, L, U, Ulo);
  for (int j = 1; j < nblk; ++j) {
    const int P = j * OB, nb = std::min(OB, N - P);
    int rc = tf32_gemm_launch(h, st, P, nb, P, 1.0f, U, Ulo, N, L + (size_t)P * N, Llo + (size_t)P * N, N, 0.0f, WT, OB,
                              /*lower_only=*/0, /*tri_k=*/2, WTlo);
    if (rc) return set_err(h, DB_ERR_CUDA, "tf32 triangular inverse (block row product) failed (%d)", rc);
    rc = tf32_gemm_launch(h, st, P, nb, nb, -1.0f, WT, WTlo, OB, Li + (size_t)P * N + P, L + (size_t)P * N + P, N, 0.0f,
                          U + P, N, 0, 0, Ulo + P);
    if (rc) return set_err(h, DB_ERR_CUDA, "tf32 triangular inverse (block column) failed (%d)", rc);
  }
  return check_launch(h, "trinv_tc");
}

// The same block-row scheme on the 3 x fp16 GEMM: operands are the scaled fp16 pairs run_potrf left in F16Scratch (panels of
// L), the ones trinv_diag_finish_kernel writes (Linv_jj into the diagonal blocks of the L pair, U_jj), and the ones the GEMM
// epilogues emit (W^T, the off-diagonal blocks of U). U is also written in fp32 (exact diagonal of K^-1, tr K^-1).
int run_trinv_tc_f16(db_handle_t h, cudaStream_t st, float* L, int N, float* Li, float* U, const F16Scratch& f) {
  const int nblk = (N + OB - 1) / OB;
  cudaMemsetAsync(U, 0, sizeof(float) * (size_t)N * N, st);
  for (int j = 0; j < nblk; ++j) {
    const int b0 = j * OB, nb = std::min(OB, N - b0);
    cudaMemset2DAsync(Li + (size_t)b0 * N + b0, sizeof(float) * (size_t)N, 0, sizeof(float) * (size_t)nb, nb, st);
  }
  trinv_diag_kernel<<<(N + NB - 1) / NB, 64, 0, st>>>(L, N, N, Li, N);
  for (int s = NB; s < OB && s < N; s *= 2) {
    const int pairs = (N + 2 * s - 1) / (2 * s);
    dim3 grid((s + TileSmall::BN - 1) / TileSmall::BN, (s + TileSmall::BM - 1) / TileSmall::BM, pairs);
    trinv_level_kernel<0, TileSmall><<<grid, 256, 0, st>>>(L, N, N, s, Li, N, U, N);
    trinv_level_kernel<1, TileSmall><<<grid, 256, 0, st>>>(L, N, N, s, Li, N, U, N);
  }
  trinv_diag_finish_kernel<<<dim3(OB / 32, OB / 32, nblk), 256, 0, st>>>(Li, N, nullptr, U, nullptr, f.l_hi, f.l_lo, f.u_hi, f.u_lo,
                                                                         f.scales + 1);
  const F16Operand U16{f.u_hi, f.u_lo, f.scales + 1}, L16{f.l_hi, f.l_lo, f.scales}, Lid16{f.l_hi, f.l_lo, f.scales + 1};
  const F16Operand WT16{f.wt_hi, f.wt_lo, f.scales + 5};
  for (int j = 1; j < nblk; ++j) {
    const int P = j * OB, nb = std::min(OB, N - P);
    int rc = f16_gemm_launch(h, st, P, nb, P, 1.0f, U16, N, L16.at((size_t)P * N), N, 0.0f, nullptr, OB, /*lower_only=*/0, /*tri_k=*/2,
                             &WT16, OB);
    if (rc) return set_err(h, DB_ERR_CUDA, "f16x3 triangular inverse (block row product) failed (%d)", rc);
    const F16Operand Uout = U16.at(P);
    rc = f16_gemm_launch(h, st, P, nb, nb, -1.0f, WT16, OB, Lid16.at((size_t)P * N + P), N, 0.0f, U + P, N, 0, 0, &Uout, N);
    if (rc) return set_err(h, DB_ERR_CUDA, "f16x3 triangular inverse (block column) failed (%d)", rc);
  }
  return check_launch(h, "trinv_tc_f16");
}

// ---- the same inverse riding along the factorisation (round 2) ------------------------------------------------------
// Block row j of the scheme above needs L_jj (final after the column-block steps of outer block j) and the converted panels
// 0..j-1, not the rest of the factorisation, and the Cholesky is a latency chain that leaves most of the GPU idle (SM
// throughput 15 %). So the inverse is issued on the handle's two auxiliary streams, one slice per outer block, behind an
// event the factorisation records when that block is final (run_potrf, TrinvOverlap):
//   diagonal lane, slice j: Linv_jj (64x64 solves + SIMT levels on that block alone), finish (fp16 pairs, U_jj) -> event
//   product lane,  slice j: U[:P_j, B_j] = -WT_j Linv_jj^T (j >= 1), then WT_{j+1} = U[:P_{j+1},:P_{j+1}] L[B_{j+1},:P_{j+1}]^T
//                           (the long product of the NEXT block row: all its inputs exist already)
// so that after the last step only Linv_jj of the last block and one K = 512 product remain. The streams share the SMs by
// construction, not by priority: the step launches of outer block b leave reserve(b) SMs free (their in-panel update CTAs are
// capped at one per strip CTA) and the side products run on at most that many CTAs (persistent grid cap) — every kernel
// involved owns its SM through its shared-memory footprint, so an over-subscribed launch would simply queue behind the other
// stream and stretch the critical path.
void run_trinv_f16_side_prologue(db_handle_t h, const TrinvOverlap& ov, cudaStream_t main, int N) {
  cudaEventRecord(ov.ev[0], main);
  cudaStreamWaitEvent(ov.diag, ov.ev[0], 0);
  cudaStreamWaitEvent(ov.prod, ov.ev[0], 0);
  const int nblk = (N + OB - 1) / OB;
  cudaMemsetAsync(ov.U, 0, sizeof(float) * (size_t)N * N, ov.diag);
  for (int j = 0; j < nblk; ++j) {
    const int b0 = j * OB, nb = std::min(OB, N - b0);
    cudaMemset2DAsync(ov.Li + (size_t)b0 * N + b0, sizeof(float) * (size_t)N, 0, sizeof(float) * (size_t)nb, nb, ov.diag);
  }
}

// slice blk, behind the event of outer block ev_blk >= blk (a deferred slice starts with a later block's event)
static void run_trinv_f16_side_block(db_handle_t h, TrinvOverlap& ov, float* L, int blk, int ev_blk, const F16Scratch& f) {
  const int N = ov.N;
  const int nblk = (N + OB - 1) / OB;
  const int b0 = blk * OB, nb = std::min(OB, N - b0);
  const size_t d = (size_t)b0 * N + b0;
  {
    cudaStream_t st = ov.diag;
    cudaStreamWaitEvent(st, ov.block_final(ev_blk), 0);
    trinv_diag_kernel<<<(nb + NB - 1) / NB, 64, 0, st>>>(L + d, N, nb, ov.Li + d, N);
    for (int s = NB; s < OB && s < nb; s *= 2) {
      const int pairs = (nb + 2 * s - 1) / (2 * s);
      dim3 grid((s + TileSmall::BN - 1) / TileSmall::BN, (s + TileSmall::BM - 1) / TileSmall::BM, pairs);
      const int splits = simt_splits((long)grid.x * grid.y * pairs, s, h->num_sms, 32);
      if (splits > 1) {          // one 512-block is 1-16 tiles per level: split K over a cluster (as run_trinv does at small N)
        const int tiles_y = (int)grid.y;
        dim3 g2(grid.x, grid.y * pairs, splits);
        simt_launch_split(trinv_level_split_kernel<0>, g2, 256, st, (const float*)(L + d), N, nb, s, ov.Li + d, N, ov.U + d, N, tiles_y, splits);
        simt_launch_split(trinv_level_split_kernel<1>, g2, 256, st, (const float*)(L + d), N, nb, s, ov.Li + d, N, ov.U + d, N, tiles_y, splits);
      } else {
        trinv_level_kernel<0, TileSmall><<<grid, 256, 0, st>>>(L + d, N, nb, s, ov.Li + d, N, ov.U + d, N);
        trinv_level_kernel<1, TileSmall><<<grid, 256, 0, st>>>(L + d, N, nb, s, ov.Li + d, N, ov.U + d, N);
      }
    }
    trinv_diag_finish_kernel<<<dim3(OB / 32, OB / 32, 1), 256, 0, st>>>(ov.Li, N, nullptr, ov.U, nullptr, f.l_hi, f.l_lo, f.u_hi, f.u_lo,
                                                                         f.scales + 1, blk);
    cudaEventRecord(ov.linv_ready(blk), st);
  }
  cudaStream_t st = ov.prod;
  cudaStreamWaitEvent(st, ov.linv_ready(blk), 0);       // (behind block_final(ev_blk): the converted panels up to ev_blk exist as well)
  const F16Operand U16{f.u_hi, f.u_lo, f.scales + 1}, L16{f.l_hi, f.l_lo, f.scales}, Lid16{f.l_hi, f.l_lo, f.scales + 1};
  const F16Operand WT16{f.wt_hi, f.wt_lo, f.scales + 5};
  // the launches of outer block ev_blk + 1 run beside these products; after the last block the GPU is theirs
  const int cap = ev_blk + 1 < nblk ? ov.reserve(ev_blk + 1) : 0;
  int first = 1;       // the first launch behind the event wait is a plain one (gemm_f16x3.cuh, allow_pdl)
  if (blk >= 1) {
    const F16Operand Uout = U16.at(b0);
    const int rc = f16_gemm_launch(h, st, b0, nb, nb, -1.0f, WT16, OB, Lid16.at(d), N, 0.0f, ov.U + b0, N, 0, 0, &Uout, N, cap, /*allow_pdl=*/0);
    if (rc && !ov.err) ov.err = rc;
    first = 0;
  }
  if (blk + 1 < nblk) {
    const int P = b0 + OB, nb1 = std::min(OB, N - P);
    const int rc = f16_gemm_launch(h, st, P, nb1, P, 1.0f, U16, N, L16.at((size_t)P * N), N, 0.0f, nullptr, OB, /*lower_only=*/0, /*tri_k=*/2,
                                   &WT16, OB, cap, /*allow_pdl=*/first ? 0 : 1);
    if (rc && !ov.err) ov.err = rc;
  }
}

bool run_trinv_f16_side_blocks(db_handle_t h, TrinvOverlap& ov, cudaStream_t main, float* L, int blk, const F16Scratch& f) {
  const int nblk = (ov.N + OB - 1) / OB;
  if (blk + 1 < nblk && ov.reserve(blk + 1) == 0) return false;       // no room beside the next block's launches: later
  cudaEventRecord(ov.block_final(blk), main);
  for (int j = ov.pending; j <= blk; ++j) run_trinv_f16_side_block(h, ov, L, j, blk, f);
  ov.pending = blk + 1;
  return true;
}

// every diagonal-lane slice is behind an event the product lane has waited for, so joining the product lane joins both
void run_trinv_f16_side_join(const TrinvOverlap& ov, cudaStream_t main) {
  cudaEventRecord(ov.join(), ov.prod);
  cudaStreamWaitEvent(main, ov.join(), 0);
}

// 0: one stream (as before); 1 / 2: TrinvOverlap::mode. DB_GP_OVERLAP overrides.
inline int gp_overlap_mode(db_handle_t h, cudaStream_t st, int N) {
  static const int env = [] { const char* e = getenv("DB_GP_OVERLAP"); return e ? atoi(e) : 1; }();
  const int nblk = (N + OB - 1) / OB;
  if (env <= 0 || nblk < 3 || nblk > TrinvOverlap::kMaxBlocks) return 0;
  // The lanes open with the first outer block whose step launches leave SMs free (TrinvOverlap::reserve). When that is late
  // in the factorisation (N = 8192: block 8 of 16) most of the inverse is still ahead when the Cholesky ends, the slices issued
  // until then run on the few SMs reserved at their issue time, and the evaluation is slower than on one stream (measured
  // 6.84 against 6.59 ms; N = 6144: 3.43 against 3.86 ms): one stream then.
  TrinvOverlap probe{}; probe.num_sms = h->num_sms; probe.N = N;
  int first = 1;
  while (first < nblk && probe.reserve(first) == 0) ++first;
  if (first - 1 > nblk / 3) return 0;
  return side_lane(h, st) ? env : 0;
}
inline TrinvOverlap gp_overlap(db_handle_t h, float* Li, float* U, int mode, int N) {
  static const int trail = [] { const char* e = getenv("DB_GP_TRAIL_RESERVE"); return e ? atoi(e) : TrinvOverlap::kMinReserve; }();
  return TrinvOverlap{h->side_stream, h->side_stream2, h->side_events, Li, U, mode, trail, h->num_sms, N, 0, 0};
}

template <int TRANS>
int run_trmm(db_handle_t h, cudaStream_t st, const float* Li, int N, const float* B, int ldb, int nrhs, float* out, int ldo,
             float* ssq) {
  // 128 x 128 tiles only when they fill the SMs (N = 328, D = 1024 is 24 of them: 62 us per product)
  if (nrhs >= 256 && ((nrhs + 127) / 128) * ((N + 127) / 128) >= h->num_sms) {
    using T = TileLarge;
    dim3 grid((nrhs + T::BN - 1) / T::BN, (N + T::BM - 1) / T::BM);
    trmm_kernel<TRANS, T><<<grid, 256, 0, st>>>(Li, N, N, B, ldb, nrhs, out, ldo, ssq);
  } else {
    using T = TileSmall;
    dim3 grid((nrhs + T::BN - 1) / T::BN, (N + T::BM - 1) / T::BM);
    const int splits = simt_splits((long)grid.x * grid.y, N, h->num_sms);
    if (splits > 1) {
      grid.z = splits;
      simt_launch_split(trmm_split_kernel<TRANS>, grid, 256, st, Li, N, N, B, ldb, nrhs, out, ldo, ssq, splits);
    } else {
      trmm_kernel<TRANS, T><<<grid, 256, 0, st>>>(Li, N, N, B, ldb, nrhs, out, ldo, ssq);
    }
  }
  return check_launch(h, "trmm");
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct GplvmWorkspace {
  float *K, *Li, *Tmp, *S, *Av, *B1, *B2, *B3, *dXacc, *hyp, *Ulo, *WT, *WTlo;
  GpAccum* acc;
  void* f16;              // F16Scratch region (3 x fp16 products)
  size_t bytes;
};
GplvmWorkspace gplvm_workspace(void* ws, int N, int Q, int D) {
  GplvmWorkspace w;
  uint8_t* p = static_cast<uint8_t*>(ws);
  uint8_t* p0 = p;
  const size_t nn = align256(sizeof(float) * (size_t)N * N), nd = align256(sizeof(float) * (size_t)N * D);
  w.K = reinterpret_cast<float*>(p); p += nn;
  w.Li = reinterpret_cast<float*>(p); p += nn;
  w.Tmp = reinterpret_cast<float*>(p); p += nn;
  w.S = reinterpret_cast<float*>(p); p += nd;
  w.Av = reinterpret_cast<float*>(p); p += nd;
  w.B1 = reinterpret_cast<float*>(p); p += nd;     // tensor-core path: Y^T, its tf32 remainder, A^T
  w.B2 = reinterpret_cast<float*>(p); p += nd;
  w.B3 = reinterpret_cast<float*>(p); p += nd;
  w.dXacc = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)N * Q);
  w.hyp = reinterpret_cast<float*>(p); p += 256;
  w.acc = reinterpret_cast<GpAccum*>(p); p += 256;
  w.Ulo = reinterpret_cast<float*>(p); p += nn;       // tensor-core path: tf32 remainder of Linv^T
  w.WT = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)N * OB);     // block-row products of run_trinv_tc
  w.WTlo = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)N * OB);
  w.f16 = p; p += align256(f16_scratch_bytes(N, D));
  w.bytes = (size_t)(p - p0);
  return w;
}

inline bool gpdbn_tensor_factor() {     // DB_GPDBN_SIMT=1: keep the GPDBN factorisation on the fp32 SIMT kernels (A/B timing, diagnostics)
  static const bool simt = [] { const char* e = getenv("DB_GPDBN_SIMT"); return e && e[0] == '1'; }();
  return !simt;
}
inline bool gp_use_tensor_path(db_handle_t h, int N, int D) {
  static const bool simt_only = [] { const char* e = getenv("DB_GP_SIMT"); return e && e[0] == '1'; }();   // diagnostics
  if (simt_only) return false;
  return h->cc_major >= 10 && N >= 512 && (N % 4) == 0 && (D % 4) == 0 && D >= 32;
}



// ---- tensor-core form of the LML + gradient evaluation (N >= 512, N % 4 == 0, D % 4 == 0) ---------------------------
// K^-1 = Linv^T Linv as ONE 3xTF32 GEMM over the transposed factor (K range clipped to the non-zero part), A = K^-1 Y,
// -0.5 A A^T, then a streaming contraction of G = (D/2) K^-1 - 0.5 A A^T with K_f. |L^-1 Y|^2 = sum Y o A.
__global__ void __launch_bounds__(256) mirror_lower_kernel(float* __restrict__ A, float* __restrict__ A2, int N) {
  __shared__ float tile[32][33];
  __shared__ float tile2[32][33];
  if (blockIdx.x > blockIdx.y) return;                       // tile (by, bx) of the lower triangle -> (bx, by)
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < N && c < N) ? A[(size_t)r * N + c] : 0.0f;
    if (A2) tile2[ty + 8 * i][tx] = (r < N && c < N) ? A2[(size_t)r * N + c] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;           // writes element (c, r) = (r, c)^T
    if (c < N && r < N && c < r) {
      A[(size_t)c * N + r] = tile[tx][ty + 8 * i];
      if (A2) A2[(size_t)c * N + r] = tile2[tx][ty + 8 * i];
    }
  }
}
// The same mirror for the 3 x fp16 chain: the full symmetric matrix is also written as the scaled fp16 pair the next product
// reads (a separate conversion pass over the mirrored matrix cost 31 us at N = 5000). The diagonal must be final already.
// 64 x 64 tiles, 16-byte loads / stores of A and 8-byte stores of the pair (N % 4 == 0 on this path); the first version moved
// 32 x 32 tiles with 4-byte loads and 2-byte stores: 52 us for the 148 MB it touches.
__device__ __forceinline__ void store_pair4(__half* __restrict__ hi, __half* __restrict__ lo, size_t idx, float4 v, float sc) {
  __half h0, l0, h1, l1, h2, l2, h3, l3;
  split_f16(v.x * sc, h0, l0); split_f16(v.y * sc, h1, l1); split_f16(v.z * sc, h2, l2); split_f16(v.w * sc, h3, l3);
  __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3), la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&ha), *reinterpret_cast<uint32_t*>(&hb));
  *reinterpret_cast<uint2*>(lo + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&la), *reinterpret_cast<uint32_t*>(&lb));
}
__global__ void __launch_bounds__(256) mirror_lower_split_kernel(float* __restrict__ A, int N, __half* __restrict__ hi,
                                                                 __half* __restrict__ lo, const float* __restrict__ scale) {
  __shared__ float tile[64][65];
  if (blockIdx.x > blockIdx.y) return;                       // tile (by, bx) of the lower triangle -> (bx, by)
  const bool diag = blockIdx.x == blockIdx.y;
  const float sc = __ldg(scale);
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;     // sixteen float4 per tile row, sixteen rows per pass
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = rr + 16 * i, gr = r0 + r, gc = c0 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < N && gc < N) {
      v = *reinterpret_cast<const float4*>(A + (size_t)gr * N + gc);
      if (!diag) store_pair4(hi, lo, (size_t)gr * N + gc, v, sc);      // strictly lower block: its pair as it is
    }
    tile[r][4 * q] = v.x; tile[r][4 * q + 1] = v.y; tile[r][4 * q + 2] = v.z; tile[r][4 * q + 3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = rr + 16 * i;
    if (diag) {                                              // symmetric fill of the diagonal tile, whole rows
      const int gr = r0 + r, gc = c0 + 4 * q;
      if (gr < N && gc < N) {
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = (4 * q + k <= r) ? tile[r][4 * q + k] : tile[4 * q + k][r];
        const float4 v = make_float4(e[0], e[1], e[2], e[3]);
        *reinterpret_cast<float4*>(A + (size_t)gr * N + gc) = v;
        store_pair4(hi, lo, (size_t)gr * N + gc, v, sc);
      }
    } else {                                                 // row c0 + r of the transposed block: columns r0 + 4q ..
      const int gr = c0 + r, gc = r0 + 4 * q;
      if (gr < N && gc < N) {
        const float4 v = make_float4(tile[4 * q][r], tile[4 * q + 1][r], tile[4 * q + 2][r], tile[4 * q + 3][r]);
        *reinterpret_cast<float4*>(A + (size_t)gr * N + gc) = v;
        store_pair4(hi, lo, (size_t)gr * N + gc, v, sc);
      }
    }
  }
}
// A = (K^-1 Y) arrives transposed from the product ([D, N]); this transposition also forms the two reductions that read A
// next to Y — sum Y o A (dots[1]) and |A|_F^2 (dots[0]) — in fp64 (per-thread fp32 partials of four terms): two separate
// dot_sum passes over N x D cost 2 x 17.5 us at N = 5000, D = 784.
__global__ void __launch_bounds__(256) transpose_dots_kernel(const float* __restrict__ in, int R, int C, int ld_in, float* __restrict__ out,
                                                             int ld_out, const float* __restrict__ Y, double* __restrict__ dots) {
  __shared__ float tile[32][33];
  __shared__ double red[2][8];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < R && c < C) ? __ldg(in + (size_t)r * ld_in + c) : 0.0f;
  }
  __syncthreads();
  float aa = 0.0f, ya = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < C && r < R) {
      const float v = tile[tx][ty + 8 * i];
      out[(size_t)c * ld_out + r] = v;
      aa = fmaf(v, v, aa);
      ya = fmaf(v, __ldg(Y + (size_t)c * ld_out + r), ya);
    }
  }
  double a = (double)aa, b = (double)ya;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (tx == 0) { red[0][ty] = a; red[1][ty] = b; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(dots + threadIdx.x, t);
  }
}
// tr G = 0.5 (D |Linv|_F^2 - |A|_F^2) from two fp32 round-to-nearest reductions (scratch[0], scratch[1]) replaces the
// trace summed from the tensor-core tiles: sums of squares are monotone, so the truncating accumulator biases them
// one way (~3e-6), and tr G cancels ~50x. The diagonal also enters sum G o Kf (Kf_ii = alpha).
// scratch (fp64): [0] |Linv|_F^2 = tr K^-1, [1] |A|_F^2, [2] sum Y o A, [3..8] off-diagonal contractions accumulated by
// gp_contract_mem_kernel: K^-1 o K_f, K^-1 o K_f r^2, AA o K_f, AA o K_f r^2 (AA = -0.5 A A^T), sum K_f, sum K_f r^2.
// The hyper-parameter gradients contract G = (D/2) K^-1 + AA with K_f: two terms far larger than their sum, each carrying the
// one-sided error of the truncating tensor-core datapath (measured at N = 5000 against fp64: d_alpha off by 3.3e-4, d_gamma by
// 1.5e-4). With K = K_f + s I and K A = Y the plain K_f contractions have closed forms,
//   sum_ij (K^-1)_ij (K_f)_ij = N - s tr K^-1,      sum_ij (A A^T)_ij (K_f)_ij = sum Y o A - s |A|_F^2,
// so sum G o K_f (d_alpha) is formed from the fp64 scalars alone. The r^2-weighted contraction (d_gamma) has no closed form; its
// dominant error is a UNIFORM ADDITIVE shift of the A A^T tiles (A ~ Y - smooth(Y) with Y centred binary: most of the D products
// of an entry are small and of one sign, and each loses its low bits toward zero when aligned inside the MMA; measured shift
// -1.5e-6 per entry, sum AA o K_f off by +6.1 and sum AA o K_f r^2 by +8.0 = 6.1 x sum K_f r^2 / sum K_f). The closed form
// calibrates that shift, delta = (measured - exact) / sum K_f, and delta sum K_f r^2 is removed from the weighted contraction.
// The K^-1 tiles show no such pattern (errors ~1e-5 of d_gamma) and are used as they are.
__global__ void gp_fix_trace_kernel(GpAccum* acc, const double* __restrict__ scratch, const float* __restrict__ hyp, int N, int D, int debug) {
  const double alpha = hyp[0], s = hyp[2], half_d = 0.5 * (double)D;
  const double trKinv = scratch[0], a2 = scratch[1], ya = scratch[2];
  acc->trG = (float)(0.5 * ((double)D * trKinv - a2));                       // the cancelling difference itself in fp64
  const double exactA = -0.5 * (ya - (s + alpha) * a2);                      // off-diagonal part of sum (-0.5 A A^T) o K_f
  double shift = scratch[7] > 0.0 ? (scratch[5] - exactA) / scratch[7] : 0.0;
  if (!(fabs(shift * scratch[7]) <= 1e-3 * fabs(exactA))) shift = 0.0;       // the calibration only ever moves parts per million
  const double sumW = half_d * ((double)N - s * trKinv) - 0.5 * (ya - s * a2);
  const double sumWr2 = half_d * scratch[4] + (scratch[6] - shift * scratch[8]);
  if (debug)
    printf("gp_fix_trace: trKinv %.9g |A|^2 %.9g YoA %.9g | K o Kf: meas %.9g exact %.9g | AA o Kf: meas %.9g exact %.9g shift %.3e | Kr2 %.9g Ar2 %.9g -> %.9g | "
           "sumW %.9g -> %.9g, sumWr2 %.9g -> %.9g\n",
           trKinv, a2, ya, scratch[3], ((double)N - s * trKinv) - alpha * trKinv, scratch[5], exactA, shift, scratch[4], scratch[6],
           scratch[6] - shift * scratch[8], (double)acc->sumW, sumW, (double)acc->sumWr2, sumWr2);
  acc->sumW = (float)sumW;
  acc->sumWr2 = (float)sumWr2;
}
__global__ void gp_store_ssq_kernel(GpAccum* acc, const double* __restrict__ scratch) { acc->ssq = (float)scratch[2]; }
// Kinv[i,i] = sum_k LinvT[i,k]^2 in fp32 round-to-nearest (one warp per row): the tensor-core value of this monotone
// sum is biased low by the truncating accumulator, and A = K^-1 Y inherits a diagonal bias one-to-one.
// trace (optional): tr K^-1 = |Linv|_F^2 = the sum of these diagonal entries, accumulated in fp64 (per-lane fp32 partials of
// at most N / 32 terms, as in dot_sum_kernel) — a separate pass over the N x N factor cost 35 us at N = 5000.
__global__ void kinv_diag_kernel(const float* __restrict__ LinvT, int N, float* __restrict__ Kinv, float* __restrict__ Kinv_lo,
                                 double* __restrict__ trace = nullptr) {
  const int i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  if (i >= N) return;
  const float* row = LinvT + (size_t)i * N;
  float s = 0.0f;
  for (int k = i + lane; k < N; k += 32) { const float v = __ldg(row + k); s = fmaf(v, v, s); }
  double sd = (double)s;
  s = warp_sum(s);
  if (trace != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
  }
  if (lane == 0) {
    Kinv[(size_t)i * N + i] = s;
    if (Kinv_lo) Kinv_lo[(size_t)i * N + i] = s - __uint_as_float(__float_as_uint(s) & 0xffffe000u);
    if (trace != nullptr) atomicAdd(trace, sd);
  }
}
// Reductions that feed the hyper-parameter gradients (tr K^-1, |A|_F^2, sum Y o A) are accumulated in DOUBLE across threads:
// ~10^4 warp partials added with fp32 atomics random-walk to ~3e-6 relative, which the ~100x cancelling
// tr G = 0.5 (D tr K^-1 - |A|^2) turns into 3e-4 (measured at N = 5000 against the fp64 oracle). Per-thread partial sums stay
// fp32 (a few fused multiply-adds each). They ride on passes that exist anyway: kinv_diag_kernel, transpose_dots_kernel.
// G tile = (D/2) Kinv + AA (AA = -0.5 A A^T, lower part valid; Kinv == nullptr: AA is G already) read from memory,
// contracted with K_f
template <class T>
__global__ void __launch_bounds__(256) gp_contract_mem_kernel(const float* __restrict__ Kinv, const float* __restrict__ AA, int N, int Q,
                                                              int D, const float* __restrict__ X, const float* __restrict__ hyp,
                                                              float* __restrict__ dXacc, GpAccum* accum, double* __restrict__ cal) {
  if (blockIdx.x > blockIdx.y) return;
  __shared__ float rowacc[T::BM][MAXQ];
  __shared__ float colacc[T::BN][MAXQ];
  __shared__ float red[3][8];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  const bool diag_tile = blockIdx.x == blockIdx.y;
  for (int t = threadIdx.x; t < T::BM * MAXQ; t += 256) { (&rowacc[0][0])[t] = 0.0f; (&colacc[0][0])[t] = 0.0f; }
  const float half_d = 0.5f * (float)D;
  const bool calibrate = cal != nullptr && Kinv != nullptr;
  const float alpha = __ldg(hyp), nhalf_inv_g2 = -0.5f / (__ldg(hyp + 1) * __ldg(hyp + 1));
  float cK = 0.0f, cKr = 0.0f, cA = 0.0f, cAr = 0.0f, cF = 0.0f, cFr = 0.0f;   // strictly-lower contractions (see gp_fix_trace_kernel)
  float gacc[T::TM][T::TN];
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      float g = 0.0f;
      if (m < N && n < N && n <= m) {
        g = __ldg(AA + (size_t)m * N + n);
        if (Kinv != nullptr) {
          const float kv = __ldg(Kinv + (size_t)m * N + n);
          if (calibrate && n < m) {
            float r2 = 0.0f;
            for (int q = 0; q < Q; ++q) { const float d = __ldg(X + (size_t)m * Q + q) - __ldg(X + (size_t)n * Q + q); r2 = fmaf(d, d, r2); }
            const float kf = alpha * __expf(nhalf_inv_g2 * r2);
            cK = fmaf(kv, kf, cK); cKr = fmaf(kv * kf, r2, cKr); cA = fmaf(g, kf, cA); cAr = fmaf(g * kf, r2, cAr);
            cF += kf; cFr = fmaf(kf, r2, cFr);
          }
          g = fmaf(half_d, kv, g);
        }
      }
      gacc[i][jj] = g;
    }
  }
  if (calibrate) {     // per-thread fp32 partials (<= TM*TN terms), everything across threads in fp64
    double v[6] = {2.0 * (double)cK, 2.0 * (double)cKr, 2.0 * (double)cA, 2.0 * (double)cAr, 2.0 * (double)cF, 2.0 * (double)cFr};
#pragma unroll
    for (int e = 0; e < 6; ++e) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
      if ((threadIdx.x & 31) == 0) atomicAdd(cal + e, v[e]);
    }
  }
  __syncthreads();
  gp_contract_tile<T>(gacc, m0, n0, diag_tile, X, N, Q, hyp, dXacc, accum, rowacc, colacc, red);
}

// The tensor-core chain's contraction for Q <= 2 (every example) in ONE pass over the tile: gp_contract_mem_kernel first walks
// its 8 x 8 entries per thread for the calibration sums (K_f evaluated per entry with run-time-Q loads of X) and then again
// inside gp_contract_tile (K_f a second time) — 170 us at N = 5000 with issue slots 27 % busy and DRAM at 7 %. Here the rows'
// and columns' latents sit in registers, every entry costs two loads and one exponential, and the calibration sums ride along.
template <class T>
__global__ void __launch_bounds__(256) gp_contract_fused_kernel(const float* __restrict__ Kinv, const float* __restrict__ AA, int N, int Q,
                                                                int D, const float* __restrict__ X, const float* __restrict__ hyp,
                                                                float* __restrict__ dXacc, GpAccum* accum, double* __restrict__ cal) {
  if (blockIdx.x > blockIdx.y) return;
  __shared__ float rowacc[T::BM][2];
  __shared__ float colacc[T::BN][2];
  __shared__ float red[3][8];
  static_assert(T::BM == T::BN, "square tiles");
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  for (int t = threadIdx.x; t < T::BM * 2; t += 256) { (&rowacc[0][0])[t] = 0.0f; (&colacc[0][0])[t] = 0.0f; }
  const float half_d = 0.5f * (float)D;
  const float alpha = __ldg(hyp), inv_g2 = 1.0f / (__ldg(hyp + 1) * __ldg(hyp + 1));
  float cK = 0.0f, cKr = 0.0f, cA = 0.0f, cAr = 0.0f, cF = 0.0f, cFr = 0.0f;   // strictly-lower contractions (see gp_fix_trace_kernel)
  float trg = 0.0f, sw = 0.0f, swr = 0.0f;
  float racc[T::TM][2], cacc[T::TN][2], xn[T::TN][2];
#pragma unroll
  for (int i = 0; i < T::TM; ++i) { racc[i][0] = racc[i][1] = 0.0f; }
#pragma unroll
  for (int jj = 0; jj < T::TN; ++jj) {
    cacc[jj][0] = cacc[jj][1] = 0.0f;
    const int n = simt_col<T>(n0, jj);
    xn[jj][0] = n < N ? __ldg(X + (size_t)n * Q) : 0.0f;
    xn[jj][1] = (n < N && Q > 1) ? __ldg(X + (size_t)n * Q + 1) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= N) continue;
    const float xm0 = __ldg(X + (size_t)m * Q);
    const float xm1 = Q > 1 ? __ldg(X + (size_t)m * Q + 1) : 0.0f;
    float aa[T::TN], kv[T::TN];
    static_assert(T::TN % 4 == 0, "a thread's columns come in aligned groups of four");
#pragma unroll
    for (int g4 = 0; g4 < T::TN / 4; ++g4) {                // the row's loads first, 16 bytes each (N % 4 == 0 on this path)
      const int n = simt_col<T>(n0, 4 * g4);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), k = a;
      if (n < N && n <= m) {
        a = __ldg(reinterpret_cast<const float4*>(AA + (size_t)m * N + n));
        k = __ldg(reinterpret_cast<const float4*>(Kinv + (size_t)m * N + n));
      }
      aa[4 * g4] = a.x; aa[4 * g4 + 1] = a.y; aa[4 * g4 + 2] = a.z; aa[4 * g4 + 3] = a.w;
      kv[4 * g4] = k.x; kv[4 * g4 + 1] = k.y; kv[4 * g4 + 2] = k.z; kv[4 * g4 + 3] = k.w;
    }
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (n >= N || n > m) continue;
      const float g = fmaf(half_d, kv[jj], aa[jj]);
      if (m == n) { trg += g; sw += g * alpha; continue; }   // Kf_ii = alpha, r = 0: no dX contribution
      const float d0 = xm0 - xn[jj][0], d1 = xm1 - xn[jj][1];
      const float r2 = fmaf(d0, d0, d1 * d1);
      const float kf = alpha * __expf(-0.5f * r2 * inv_g2);
      const float kfr = kf * r2;
      cK = fmaf(kv[jj], kf, cK); cKr = fmaf(kv[jj], kfr, cKr); cA = fmaf(aa[jj], kf, cA); cAr = fmaf(aa[jj], kfr, cAr);
      cF += kf; cFr += kfr;
      const float w = g * kf;
      sw += 2.0f * w; swr += 2.0f * w * r2;                  // (i,j) and (j,i)
      racc[i][0] = fmaf(w, d0, racc[i][0]); racc[i][1] = fmaf(w, d1, racc[i][1]);
      cacc[jj][0] = fmaf(-w, d0, cacc[jj][0]); cacc[jj][1] = fmaf(-w, d1, cacc[jj][1]);
    }
  }

  {     // per-thread fp32 partials (<= TM*TN terms), everything across threads in fp64
    double v[6] = {2.0 * (double)cK, 2.0 * (double)cKr, 2.0 * (double)cA, 2.0 * (double)cAr, 2.0 * (double)cF, 2.0 * (double)cFr};
#pragma unroll
    for (int e = 0; e < 6; ++e) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
      if ((threadIdx.x & 31) == 0) atomicAdd(cal + e, v[e]);
    }
  }
  const int tx = threadIdx.x % (T::BN / T::TN), lane = threadIdx.x & 31;
  static_assert(T::BN / T::TN == 16, "row reduction assumes 16 threads per tile row");
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v = racc[i][q];
      v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8);
      if (tx == 0) rowacc[simt_row<T>(0, i)][q] = v;        // rows are owned by one 16-thread group
    }
  }
#pragma unroll
  for (int jj = 0; jj < T::TN; ++jj) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v = cacc[jj][q];
      v += __shfl_xor_sync(0xffffffffu, v, 16);             // the two tile rows held by this warp
      if (lane < 16) atomicAdd(&colacc[simt_col<T>(0, jj)][q], v);
    }
  }
  __syncthreads();
  const float sc = -2.0f * inv_g2;
  for (int t = threadIdx.x; t < T::BM * Q; t += 256) {
    const int r = t / Q, q = t % Q;
    if (m0 + r < N) atomicAdd(dXacc + (size_t)(m0 + r) * Q + q, sc * rowacc[r][q]);
    if (n0 + r < N) atomicAdd(dXacc + (size_t)(n0 + r) * Q + q, sc * colacc[r][q]);
  }
  trg = warp_sum(trg); sw = warp_sum(sw); swr = warp_sum(swr);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = trg; red[1][threadIdx.x >> 5] = sw; red[2][threadIdx.x >> 5] = swr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
    atomicAdd(&accum->trG, a); atomicAdd(&accum->sumW, b); atomicAdd(&accum->sumWr2, c);
  }
}

// ---- GPDBN top layer (gprbm.py), training mode x_ph = X ------------------------------------------------
// With K_Xx = K_f = K - s I (s = noise variance) the reference's quantities collapse to (after GP14):
//   pred_var = s - s^2 diag(K^-1)        (gprbm.py:208-218)
//   Hm       = H2 - s K^-1 H2            (gprbm.py:220-226)
//   gp_loss  = 0.5 (sum H2 o K^-1 H2 + D logdet K) + |X|^2      (gprbm.py:234-245)
// so the N-right-hand-side triangular solve disappears; K^-1 is formed explicitly once per step.

// Kinv = Linv^T Linv, both triangles written (lower tile pairs computed, mirrored)
__global__ void __launch_bounds__(256) kinv_kernel(const float* __restrict__ Li, int ldi, int N, float* __restrict__ Kinv) {
  using T = TileLarge;
  if (blockIdx.x > blockIdx.y) return;
  __shared__ typename T::Smem sm;
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  LoadLinvT_A a{Li, ldi, N, 1.0f};
  LoadLinvT_B b{Li, ldi, N};
  float acc[T::TM][T::TN];
  simt_gemm_tile<T>(sm, a, b, m0, n0, m0, N, acc);
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
    if (m >= N) continue;
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (n >= N) continue;
      Kinv[(size_t)m * N + n] = acc[i][jj];
      Kinv[(size_t)n * N + m] = acc[i][jj];
    }
  }
}

// small N: the same product on 64 x 64 lower tiles whose K range [m0, N) is split over a thread-block cluster
__global__ void __launch_bounds__(256) kinv_split_kernel(const float* __restrict__ Li, int ldi, int N, float* __restrict__ Kinv,
                                                         int splits) {
  using T = TileSmall;
  if (blockIdx.x > blockIdx.y) return;      // the CTAs of a cluster share (x, y)
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  LoadLinvT_A a{Li, ldi, N, 1.0f};
  LoadLinvT_B b{Li, ldi, N};
  float acc[T::TM][T::TN];
  int kb, ke;
  simt_split_range<T>(N - m0, splits, kb, ke);
  simt_gemm_tile<T>(sm, a, b, m0, n0, m0 + kb, m0 + ke, acc);
  simt_cluster_reduce<T>(red, acc, splits, [&](int r, int c, float x) {
    const int m = m0 + r, n = n0 + c;
    if (m < N && n < N) { Kinv[(size_t)m * N + n] = x; Kinv[(size_t)n * N + m] = x; }
  });
}

__global__ void gpdbn_pred_var_kernel(const float* __restrict__ Kinv, int N, const float* __restrict__ hyp, float* __restrict__ pv,
                                      int* __restrict__ info) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float s = __ldg(hyp + 2);
  const float v = s - s * s * __ldg(Kinv + (size_t)i * N + i);
  pv[i] = v;
  if (!(v >= 0.0f)) atomicCAS(info + 1, 0, i + 1);          // tf.assert_non_negative(pred_var), gprbm.py:215
}

// Hm = H2 - s B ; ssq += sum H2 o B
__global__ void gpdbn_apply_kernel(const float* __restrict__ H2, const float* __restrict__ B, size_t n, const float* __restrict__ hyp,
                                   float* __restrict__ Hm, GpAccum* acc) {
  const float s = __ldg(hyp + 2);
  float local = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float h = __ldg(H2 + i), b = __ldg(B + i);
    Hm[i] = fmaf(-s, b, h);
    local = fmaf(h, b, local);
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) atomicAdd(&acc->ssq, local);
}
__global__ void gpdbn_scalars_kernel(const GpAccum* __restrict__ acc, const float* __restrict__ hyp, int D, float* __restrict__ out) {
  out[0] = 0.5f * (acc->ssq + (float)D * acc->logdet) + acc->xsq;      // gprbm.py:242
  out[1] = acc->logdet; out[2] = acc->ssq; out[3] = acc->xsq;
  out[4] = hyp[0]; out[5] = hyp[1]; out[6] = hyp[2]; out[7] = 0.0f;
}
// g_H2 = g_Hm - s P + B ; pad[0] += sum g_Hm o B
__global__ void gpdbn_backward_h_kernel(const float* __restrict__ gHm, const float* __restrict__ P, const float* __restrict__ B,
                                        size_t n, const float* __restrict__ hyp, float* __restrict__ gH2, GpAccum* acc) {
  const float s = __ldg(hyp + 2);
  float local = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = __ldg(gHm + i), b = __ldg(B + i);
    gH2[i] = fmaf(-s, __ldg(P + i), g) + b;
    local = fmaf(g, b, local);
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) atomicAdd(&acc->pad[0], local);
}
// pad[1] = sum_i g_pv_i (1 - 2 s Kinv_ii)
__global__ void gpdbn_gpv_scalar_kernel(const float* __restrict__ gpv, const float* __restrict__ Kinv, int N,
                                        const float* __restrict__ hyp, GpAccum* acc) {
  const float s = __ldg(hyp + 2);
  float local = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    local += __ldg(gpv + i) * (1.0f - 2.0f * s * __ldg(Kinv + (size_t)i * N + i));
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) atomicAdd(&acc->pad[1], local);
}

// G = s^2 Kinv diag(g_pv) Kinv + 0.5 (E B^T + B E^T) + (D/2) Kinv, E = s P - 0.5 B   (symmetric part of dLoss/dK)
struct LoadKinvScaledA {   // a(m,k) = s^2 g_k Kinv[m,k]
  static constexpr bool k_contig = true;
  const float* Kinv; const float* gpv; int N; float s2;
  __device__ __forceinline__ float operator()(int m, int k) const {
    return (m < N && k < N) ? s2 * __ldg(gpv + k) * __ldg(Kinv + (size_t)m * N + k) : 0.0f;
  }
};
struct LoadKinvB {         // b(k,n) = Kinv[n,k]
  static constexpr bool k_contig = true;
  const float* Kinv; int N;
  __device__ __forceinline__ float operator()(int k, int n) const { return (n < N && k < N) ? __ldg(Kinv + (size_t)n * N + k) : 0.0f; }
};
struct LoadEB_A {          // a(m,k) = k < D ? 0.5 E[m,k] : 0.5 B[m,k-D]
  static constexpr bool k_contig = true;
  const float* P; const float* B; int N, D; float s;
  __device__ __forceinline__ float operator()(int m, int k) const {
    if (m >= N || k >= 2 * D) return 0.0f;
    if (k < D) return 0.5f * (s * __ldg(P + (size_t)m * D + k) - 0.5f * __ldg(B + (size_t)m * D + k));
    return 0.5f * __ldg(B + (size_t)m * D + (k - D));
  }
};
struct LoadEB_B {          // b(k,n) = k < D ? B[n,k] : E[n,k-D]
  static constexpr bool k_contig = true;
  const float* P; const float* B; int N, D; float s;
  __device__ __forceinline__ float operator()(int k, int n) const {
    if (n >= N || k >= 2 * D) return 0.0f;
    if (k < D) return __ldg(B + (size_t)n * D + k);
    return s * __ldg(P + (size_t)n * D + (k - D)) - 0.5f * __ldg(B + (size_t)n * D + (k - D));
  }
};
__global__ void __launch_bounds__(256) gpdbn_grad_kernel(const float* __restrict__ Kinv, const float* __restrict__ gpv,
                                                         const float* __restrict__ P, const float* __restrict__ B,
                                                         const float* __restrict__ X, int N, int Q, int D,
                                                         const float* __restrict__ hyp, float* __restrict__ dXacc, GpAccum* accum) {
  using T = TileLarge;
  if (blockIdx.x > blockIdx.y) return;
  __shared__ typename T::Smem sm;
  __shared__ float rowacc[T::BM][MAXQ];
  __shared__ float colacc[T::BN][MAXQ];
  __shared__ float red[3][8];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  const bool diag_tile = blockIdx.x == blockIdx.y;
  for (int t = threadIdx.x; t < T::BM * MAXQ; t += 256) { (&rowacc[0][0])[t] = 0.0f; (&colacc[0][0])[t] = 0.0f; }
  const float s = __ldg(hyp + 2);
  float gacc[T::TM][T::TN];
  {
    LoadKinvScaledA a{Kinv, gpv, N, s * s};
    LoadKinvB b{Kinv, N};
    simt_gemm_tile<T>(sm, a, b, m0, n0, 0, N, gacc);
  }
  __syncthreads();
  {
    LoadEB_A a{P, B, N, D, s};
    LoadEB_B b{P, B, N, D, s};
    simt_gemm_tile<T, LoadEB_A, LoadEB_B, false>(sm, a, b, m0, n0, 0, 2 * D, gacc);
  }
  const float half_d = 0.5f * (float)D;
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      if (m < N && n < N) gacc[i][jj] = fmaf(half_d, __ldg(Kinv + (size_t)m * N + n), gacc[i][jj]);
    }
  }
  gp_contract_tile<T>(gacc, m0, n0, diag_tile, X, N, Q, hyp, dXacc, accum, rowacc, colacc, red);
}

// small N: G = s^2 Kinv diag(g_pv) Kinv + (E B^T + B E^T)/2-type product + (D/2) Kinv formed by 64 x 64 lower tiles with a
// cluster-split K and written to memory; gp_contract_mem_kernel<TileSmall>(nullptr, G, ...) contracts it
__global__ void __launch_bounds__(256) gpdbn_g_split_kernel(const float* __restrict__ Kinv, const float* __restrict__ gpv,
                                                            const float* __restrict__ P, const float* __restrict__ B, int N, int D,
                                                            const float* __restrict__ hyp, float* __restrict__ G, int splits) {
  using T = TileSmall;
  if (blockIdx.x > blockIdx.y) return;
  __shared__ typename T::Smem sm;
  __shared__ __align__(16) float red[T::BM * T::BN];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  const float s = __ldg(hyp + 2);
  float gacc[T::TM][T::TN];
  int kb, ke;
  {
    simt_split_range<T>(N, splits, kb, ke);
    LoadKinvScaledA a{Kinv, gpv, N, s * s};
    LoadKinvB b{Kinv, N};
    simt_gemm_tile<T>(sm, a, b, m0, n0, kb, ke, gacc);
  }
  __syncthreads();
  {
    simt_split_range<T>(2 * D, splits, kb, ke);
    LoadEB_A a{P, B, N, D, s};
    LoadEB_B b{P, B, N, D, s};
    simt_gemm_tile<T, LoadEB_A, LoadEB_B, false>(sm, a, b, m0, n0, kb, ke, gacc);
  }
  const float half_d = 0.5f * (float)D;
  simt_cluster_reduce<T>(red, gacc, splits, [&](int r, int c, float x) {
    const int m = m0 + r, n = n0 + c;
    if (m < N && n < N) G[(size_t)m * N + n] = fmaf(half_d, __ldg(Kinv + (size_t)m * N + n), x);
  });
}
// Tensor-core form of the same gradient: Xs = s^2 K^-1 diag(g_pv) (and its tf32 remainder) is materialised, G1 = Xs K^-1
// runs as one 3xTF32 GEMM (lower part), and the contraction kernel reads the G1 tile from memory instead of forming it.
__global__ void __launch_bounds__(256) kinv_scale_cols_kernel(const float4* __restrict__ Kinv, const float* __restrict__ gpv, int N,
                                                              const float* __restrict__ hyp, float4* __restrict__ Xs,
                                                              float4* __restrict__ Xs_lo) {
  const float s2 = __ldg(hyp + 2) * __ldg(hyp + 2);
  const size_t n4 = (size_t)N * N / 4, row4 = (size_t)N / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % row4) * 4;
    const float4 k = __ldg(Kinv + i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gpv + c));
    float4 v = make_float4(s2 * g.x * k.x, s2 * g.y * k.y, s2 * g.z * k.z, s2 * g.w * k.w);
    Xs[i] = v;
    Xs_lo[i] = make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                           v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
  }
}
// 3 x fp16 form: Xs directly as the scaled fp16 pair the product reads (|Xs_ij| <= s^2 max|g_pv| / s: *scale from that bound)
__global__ void __launch_bounds__(256) kinv_scale_cols_f16_kernel(const float4* __restrict__ Kinv, const float* __restrict__ gpv, int N,
                                                                  const float* __restrict__ hyp, const float* __restrict__ scale,
                                                                  __half* __restrict__ hi, __half* __restrict__ lo) {
  const float f = __ldg(hyp + 2) * __ldg(hyp + 2) * __ldg(scale);
  const size_t n4 = (size_t)N * N / 4, row4 = (size_t)N / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % row4) * 4;
    const float4 k = __ldg(Kinv + i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gpv + c));
    __half h0, l0, h1, l1, h2, l2, h3, l3;
    split_f16(f * g.x * k.x, h0, l0); split_f16(f * g.y * k.y, h1, l1); split_f16(f * g.z * k.z, h2, l2); split_f16(f * g.w * k.w, h3, l3);
    __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3), la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
    *reinterpret_cast<uint2*>(hi + 4 * i) = make_uint2(*reinterpret_cast<uint32_t*>(&ha), *reinterpret_cast<uint32_t*>(&hb));
    *reinterpret_cast<uint2*>(lo + 4 * i) = make_uint2(*reinterpret_cast<uint32_t*>(&la), *reinterpret_cast<uint32_t*>(&lb));
  }
}
__global__ void __launch_bounds__(256) gpdbn_grad_mem_kernel(const float* __restrict__ G1, const float* __restrict__ Kinv,
                                                             const float* __restrict__ P, const float* __restrict__ B,
                                                             const float* __restrict__ X, int N, int Q, int D,
                                                             const float* __restrict__ hyp, float* __restrict__ dXacc, GpAccum* accum) {
  using T = TileLarge;
  if (blockIdx.x > blockIdx.y) return;
  __shared__ typename T::Smem sm;
  __shared__ float rowacc[T::BM][MAXQ];
  __shared__ float colacc[T::BN][MAXQ];
  __shared__ float red[3][8];
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  const bool diag_tile = blockIdx.x == blockIdx.y;
  for (int t = threadIdx.x; t < T::BM * MAXQ; t += 256) { (&rowacc[0][0])[t] = 0.0f; (&colacc[0][0])[t] = 0.0f; }
  const float s = __ldg(hyp + 2);
  const float half_d = 0.5f * (float)D;
  float gacc[T::TM][T::TN];
#pragma unroll
  for (int i = 0; i < T::TM; ++i) {
    const int m = simt_row<T>(m0, i);
#pragma unroll
    for (int jj = 0; jj < T::TN; ++jj) {
      const int n = simt_col<T>(n0, jj);
      float v = 0.0f;
      if (m < N && n < N) {
        const size_t lo_idx = n <= m ? (size_t)m * N + n : (size_t)n * N + m;      // G1 is symmetric, lower part stored
        v = fmaf(half_d, __ldg(Kinv + (size_t)m * N + n), __ldg(G1 + lo_idx));
      }
      gacc[i][jj] = v;
    }
  }
  __syncthreads();
  {
    LoadEB_A a{P, B, N, D, s};
    LoadEB_B b{P, B, N, D, s};
    simt_gemm_tile<T, LoadEB_A, LoadEB_B, false>(sm, a, b, m0, n0, 0, 2 * D, gacc);
  }
  gp_contract_tile<T>(gacc, m0, n0, diag_tile, X, N, Q, hyp, dXacc, accum, rowacc, colacc, red);
}
__global__ void gpdbn_finalize_kernel(const GpAccum* __restrict__ acc, const float* __restrict__ hyp, const float* __restrict__ opt,
                                      int flags, int N, int Q, float x_prior_w, const float* __restrict__ X,
                                      const float* __restrict__ dXacc, float* __restrict__ dX, float* __restrict__ dhyp) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = tid; i < N * Q; i += gridDim.x * blockDim.x) dX[i] = dXacc[i] + 2.0f * x_prior_w * __ldg(X + i);
  if (tid == 0) {
    const float alpha = hyp[0], gamma = hyp[1];
    dhyp[0] = (flags & 1) ? (acc->sumW / alpha) * sigmoidf_(opt[0]) : 0.0f;
    dhyp[1] = (flags & 2) ? (acc->sumWr2 / (gamma * gamma * gamma)) * sigmoidf_(opt[1]) : 0.0f;
    // noise: trace of dLoss/dK plus the explicit dependence of Hm and pred_var on s
    dhyp[2] = (flags & 4) ? (acc->trG - acc->pad[0] + acc->pad[1]) * sigmoidf_(opt[2]) : 0.0f;
  }
}

struct GpdbnWorkspace {
  float *K, *Li, *Kinv, *B, *P, *dXacc, *hyp, *U, *Ulo, *WT, *WTlo;
  GpAccum* acc;
  void* f16;              // F16Scratch region (3 x fp16 factorisation, triangular inverse and K^-1 product)
  size_t bytes;
};
GpdbnWorkspace gpdbn_workspace(void* ws, int N, int Q, int D) {
  GpdbnWorkspace w;
  uint8_t* p = static_cast<uint8_t*>(ws);
  uint8_t* p0 = p;
  const size_t nn = align256(sizeof(float) * (size_t)N * N), nd = align256(sizeof(float) * (size_t)N * D);
  w.K = reinterpret_cast<float*>(p); p += nn;
  w.Li = reinterpret_cast<float*>(p); p += nn;
  w.Kinv = reinterpret_cast<float*>(p); p += nn;      // doubles as the scratch of the triangular inversion
  w.B = reinterpret_cast<float*>(p); p += nd;
  w.P = reinterpret_cast<float*>(p); p += nd;
  w.dXacc = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)N * Q);
  w.hyp = reinterpret_cast<float*>(p); p += 256;
  w.acc = reinterpret_cast<GpAccum*>(p); p += 256;
  // tensor-core factorisation (N >= 512, N % 4 == 0): Linv^T, its tf32 remainder and the block-row products
  w.U = reinterpret_cast<float*>(p); p += nn;
  w.Ulo = reinterpret_cast<float*>(p); p += nn;
  w.WT = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)N * OB);
  w.WTlo = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)N * OB);
  w.f16 = p; p += align256(f16_scratch_bytes(N, 8));
  w.bytes = (size_t)(p - p0);
  return w;
}

}  // namespace
}  // namespace db

using namespace db;

extern "C" int db_gp_effective_hyp(db_handle_t h, db_stream_t stream, const float* hyp_opt, int flags, float fixed_noise,
                                   float jitter, float* hyp_out) {
  DB_REQUIRE(h, hyp_opt && hyp_out, "null operand");
  effective_hyp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyp_opt, flags, fixed_noise, jitter, hyp_out);
  return check_launch(h, "gp_effective_hyp");
}

extern "C" int db_gram_rbf(db_handle_t h, db_stream_t stream, const float* X, const float* Y, int N, int M, int Q,
                           const float* hyp, int add_noise, float* out, int ldo) {
  DB_REQUIRE(h, X && hyp && out && N > 0 && M > 0 && Q > 0 && ldo >= M, "null operand or bad shape");
  DB_REQUIRE(h, Y != nullptr || M == N, "K(X) needs M == N");
  DB_REQUIRE(h, Q <= 40, "latent dimensionality above 40 is not supported");
  return run_gram(h, (cudaStream_t)stream, X, Y, N, M, Q, hyp, add_noise, out, ldo);
}

extern "C" int db_potrf_lower(db_handle_t h, db_stream_t stream, float* A, int lda, int N, int* info) {
  DB_REQUIRE(h, A && info && N > 0 && lda >= N, "null operand or bad shape");
  return run_potrf(h, (cudaStream_t)stream, A, lda, N, info);
}

extern "C" int db_trsm_lower(db_handle_t h, db_stream_t stream, const float* L, int ldl, int N, float* B, int ldb, int nrhs,
                             int trans) {
  DB_REQUIRE(h, L && B && N > 0 && nrhs > 0 && ldl >= N && ldb >= nrhs, "null operand or bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = (N + NB - 1) / NB;
  for (int bi = 0; bi < nblk; ++bi) {
    const int blk = trans ? nblk - 1 - bi : bi;
    const int i0 = blk * NB, nb = N - i0 < NB ? N - i0 : NB;
    trsm_diag_kernel<<<(nrhs + 255) / 256, 256, 0, st>>>(L, ldl, N, i0, B, ldb, nrhs, trans);
    const int r0 = trans ? 0 : i0 + nb;
    const int nrows = trans ? i0 : N - r0;
    if (nrows > 0) {
      dim3 grid((nrhs + TileSmall::BN - 1) / TileSmall::BN, (nrows + TileSmall::BM - 1) / TileSmall::BM);
      trsm_update_kernel<<<grid, 256, 0, st>>>(L, ldl, i0, nb, r0, nrows, B, ldb, nrhs, trans);
    }
  }
  return check_launch(h, "trsm_lower");
}

extern "C" size_t db_gplvm_workspace_bytes(int N, int Q, int D) {
  if (N <= 0 || Q <= 0 || D <= 0) return 0;
  return gplvm_workspace(nullptr, N, Q, D).bytes;
}

extern "C" int db_gplvm_lml_grad(db_handle_t h, db_stream_t stream, const float* X, const float* Y, int N, int Q, int D,
                                 const float* hyp_opt, int flags, float fixed_noise, float jitter, float x_prior_weight,
                                 int include_const, int want_grad, void* workspace, float* out_scalars, float* dX,
                                 float* L_out, int* info) {
  DB_REQUIRE(h, X && Y && hyp_opt && workspace && out_scalars && info, "null operand");
  DB_REQUIRE(h, N > 1 && Q > 0 && D > 0, "bad shape");
  DB_REQUIRE(h, Q <= MAXQ, "latent dimensionality above 8 is not supported by the fused gradient");
  DB_REQUIRE(h, !want_grad || dX, "dX required when want_grad");
  cudaStream_t st = (cudaStream_t)stream;
  GplvmWorkspace w = gplvm_workspace(workspace, N, Q, D);
  int rc;
  effective_hyp_kernel<<<1, 1, 0, st>>>(hyp_opt, flags, fixed_noise, jitter, w.hyp);
  cudaMemsetAsync(w.acc, 0, sizeof(GpAccum), st);
  if ((rc = run_gram(h, st, X, nullptr, N, N, Q, w.hyp, 1, w.K, N))) return rc;
  const bool tc = gp_use_tensor_path(h, N, D);
  const bool f16 = tc && gp_use_f16(N, D);
  F16Scratch fs{};
  if (f16) { fs = f16_scratch(w.f16, N, D); f16_gp_bounds(st, w.hyp, N, fs.scales); }
  const int overlap = f16 ? gp_overlap_mode(h, st, N) : 0;
  TrinvOverlap ov{};
  if (overlap) {
    // the triangular inverse and the conversion of Y ride along the factorisation on the auxiliary stream
    ov = gp_overlap(h, w.Li, w.Tmp, overlap, N);
    run_trinv_f16_side_prologue(h, ov, st, N);
    const F16Operand Y16{fs.y_hi, fs.y_lo, fs.scales + 3};
    transpose_f32(ov.prod, Y, N, D, D, w.B1, N);                                   // B1 = Y^T [D,N]
    f16_scale_from_absmax(h, ov.prod, Y, N, D, D, fs.bits, fs.scales + 3);
    f16_split(h, ov.prod, w.B1, D, N, N, Y16, N);
  }
  if ((rc = run_potrf(h, st, w.K, N, N, info, tc ? w.Li : nullptr, f16 ? &fs : nullptr, overlap ? &ov : nullptr))) {
    if (overlap) run_trinv_f16_side_join(ov, st);
    return rc;
  }
  if (L_out) cudaMemcpyAsync(L_out, w.K, sizeof(float) * (size_t)N * N, cudaMemcpyDeviceToDevice, st);
  logdet_xsq_kernel<<<1, 1024, 0, st>>>(w.K, N, N, X, N * Q, w.acc);
  if (!tc && (rc = run_trinv(h, st, w.K, N, N, w.Li, w.Tmp))) return rc;
  if (tc) {
    const int tiles32 = (N + 31) / 32;
    double* frob = reinterpret_cast<double*>(reinterpret_cast<char*>(w.acc) + 64);   // nine fp64 scratch sums behind the accumulators
    cudaMemsetAsync(frob, 0, 9 * sizeof(double), st);
    // Tmp = Linv^T (upper triangular), Ulo = its tf32 remainder (3 x fp16 form: its scaled fp16 pair in the scratch)
    if (overlap) {
      run_trinv_f16_side_join(ov, st);
      if (ov.err) return set_err(h, DB_ERR_CUDA, "f16x3 triangular inverse (side lanes) failed (%d)", ov.err);
    }
    else if (f16) { if ((rc = run_trinv_tc_f16(h, st, w.K, N, w.Li, w.Tmp, fs))) return rc; }
    else if ((rc = run_trinv_tc(h, st, w.K, w.Li, N, w.Li, w.Tmp, w.Ulo, w.WT, w.WTlo))) return rc;
    int trc;
    if (f16) {
      // the three big products as 3 x fp16 (gemm_f16x3.cu): operands converted once into scaled fp16 pairs
      F16Operand U16{fs.u_hi, fs.u_lo, fs.scales + 1}, K16{fs.k_hi, fs.k_lo, fs.scales + 2};
      F16Operand Y16{fs.y_hi, fs.y_lo, fs.scales + 3}, A16{fs.a_hi, fs.a_lo, fs.scales + 4};
      trc = f16_gemm_launch(h, st, N, N, N, 1.0f, U16, N, U16, N, 0.0f, w.Li, N, 1, 1);        // Li = K^-1 (lower)
      if (trc) return set_err(h, DB_ERR_CUDA, "f16x3 K^-1 product failed (%d)", trc);
      kinv_diag_kernel<<<(N + 7) / 8, 256, 0, st>>>(w.Tmp, N, w.Li, nullptr, frob);     // + |Linv|_F^2 = tr K^-1
      mirror_lower_split_kernel<<<dim3((N + 63) / 64, (N + 63) / 64), 256, 0, st>>>(w.Li, N, K16.hi, K16.lo, K16.scale);   // |K^-1_ij| <= 1 / s
      if (!overlap) {
        transpose_f32(st, Y, N, D, D, w.B1, N);                                    // B1 = Y^T [D,N]
        f16_scale_from_absmax(h, st, Y, N, D, D, fs.bits, fs.scales + 3);
        f16_split(h, st, w.B1, D, N, N, Y16, N);
      }
      trc = f16_gemm_launch(h, st, D, N, N, 1.0f, Y16, N, K16, N, 0.0f, w.B3, N, 0, 0);        // B3 = (K^-1 Y)^T
      if (trc) return set_err(h, DB_ERR_CUDA, "f16x3 K^-1 Y product failed (%d)", trc);
      // Av = A = K^-1 Y [N,D]; frob[1] = |A|_F^2, frob[2] = sum Y o A = |L^-1 Y|^2
      transpose_dots_kernel<<<dim3((N + 31) / 32, (D + 31) / 32), 256, 0, st>>>(w.B3, D, N, N, w.Av, D, Y, frob + 1);
      gp_store_ssq_kernel<<<1, 1, 0, st>>>(w.acc, frob);
      if (want_grad) {
        f16_scale_from_absmax(h, st, w.Av, N, D, D, fs.bits + 1, fs.scales + 4);
        f16_split(h, st, w.Av, N, D, D, A16, D);
        trc = f16_gemm_launch(h, st, N, N, D, -0.5f, A16, D, A16, D, 0.0f, w.Tmp, N, 1, 0);     // Tmp = -0.5 A A^T (lower)
        if (trc) return set_err(h, DB_ERR_CUDA, "f16x3 A A^T product failed (%d)", trc);
      }
    } else {
    // Li = K^-1 (lower), K buffer (L is dead) = its tf32 remainder, written by the same epilogue
    trc = tf32_gemm_launch(h, st, N, N, N, 1.0f, w.Tmp, w.Ulo, N, w.Tmp, w.Ulo, N, 0.0f, w.Li, N, 1, 1, w.K);
    if (trc) return set_err(h, DB_ERR_CUDA, "tf32 K^-1 product failed (%d)", trc);
    mirror_lower_kernel<<<dim3(tiles32, tiles32), 256, 0, st>>>(w.Li, w.K, N);
    kinv_diag_kernel<<<(N + 7) / 8, 256, 0, st>>>(w.Tmp, N, w.Li, w.K, frob);        // + |Linv|_F^2 = tr K^-1
    transpose_f32(st, Y, N, D, D, w.B1, N);                                      // B1 = Y^T [D,N]
    tf32_split_lo(h, st, w.B1, D, N, N, w.B2);
    trc = tf32_gemm_launch(h, st, D, N, N, 1.0f, w.B1, w.B2, N, w.Li, w.K, N, 0.0f, w.B3, N, 0, 0);          // B3 = (K^-1 Y)^T
    if (trc) return set_err(h, DB_ERR_CUDA, "tf32 K^-1 Y product failed (%d)", trc);
    transpose_dots_kernel<<<dim3((N + 31) / 32, (D + 31) / 32), 256, 0, st>>>(w.B3, D, N, N, w.Av, D, Y, frob + 1);   // as above
    gp_store_ssq_kernel<<<1, 1, 0, st>>>(w.acc, frob);
    if (want_grad) {
      tf32_split_lo(h, st, w.Av, N, D, D, w.S);                                  // S buffer = remainder of A
      trc = tf32_gemm_launch(h, st, N, N, D, -0.5f, w.Av, w.S, D, w.Av, w.S, D, 0.0f, w.Tmp, N, 1, 0);       // Tmp = -0.5 A A^T (lower)
      if (trc) return set_err(h, DB_ERR_CUDA, "tf32 A A^T product failed (%d)", trc);
    }
    }
    if (want_grad) {
      cudaMemsetAsync(w.dXacc, 0, sizeof(float) * (size_t)N * Q, st);
      const int tiles = (N + TileLarge::BM - 1) / TileLarge::BM;
      if (Q <= 2 && (N % 4) == 0) gp_contract_fused_kernel<TileLarge><<<dim3(tiles, tiles), 256, 0, st>>>(w.Li, w.Tmp, N, Q, D, X, w.hyp, w.dXacc, w.acc, frob + 3);
      else gp_contract_mem_kernel<TileLarge><<<dim3(tiles, tiles), 256, 0, st>>>(w.Li, w.Tmp, N, Q, D, X, w.hyp, w.dXacc, w.acc, frob + 3);
      static const int dbg = [] { const char* e = getenv("DB_GP_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
      gp_fix_trace_kernel<<<1, 1, 0, st>>>(w.acc, frob, w.hyp, N, D, dbg);
    }
  } else {
  if ((rc = run_trmm<0>(h, st, w.Li, N, Y, D, D, w.S, D, &w.acc->ssq))) return rc;
  if (want_grad) {
    if ((rc = run_trmm<1>(h, st, w.Li, N, w.S, D, D, w.Av, D, nullptr))) return rc;
    cudaMemsetAsync(w.dXacc, 0, sizeof(float) * (size_t)N * Q, st);
    const int tiles = (N + TileLarge::BM - 1) / TileLarge::BM;
    const int tiles_s = (N + TileSmall::BM - 1) / TileSmall::BM;
    const int splits = simt_splits((long)tiles_s * (tiles_s + 1) / 2, D + N / 2, h->num_sms);
    if (splits > 1) {      // few tiles: cluster split-K into w.Tmp (free once Linv is formed), then the contraction from memory
      simt_launch_split(gp_g_split_kernel, dim3(tiles_s, tiles_s, splits), 256, st, (const float*)w.Li, N, (const float*)w.Av, D, N, D,
                        w.Tmp, splits);
      gp_contract_mem_kernel<TileSmall><<<dim3(tiles_s, tiles_s), 256, 0, st>>>(nullptr, w.Tmp, N, Q, D, X, w.hyp, w.dXacc, w.acc, nullptr);
    } else {
      gp_grad_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(w.Li, N, w.Av, D, X, N, Q, D, w.hyp, w.dXacc, w.acc);
    }
  }
  }
  gp_finalize_kernel<<<(N * Q + 255) / 256, 256, 0, st>>>(w.acc, w.hyp, hyp_opt, flags, N, Q, D, x_prior_weight, include_const,
                                                       want_grad, X, w.dXacc, dX, out_scalars);
  return check_launch(h, "gplvm_lml_grad");
}

extern "C" int db_gp_predict(db_handle_t h, db_stream_t stream, const float* X, const float* L, const float* S,
                             const float* y_mean, const float* xstar, int N, int Q, int D, int M, const float* hyp,
                             float* workspace, float* mean, float* var, int* info) {
  DB_REQUIRE(h, X && L && xstar && hyp && workspace && N > 0 && Q > 0 && M > 0, "null operand or bad shape");
  DB_REQUIRE(h, mean == nullptr || (S != nullptr && D > 0), "mean needs S = L^-1 Y");
  DB_REQUIRE(h, mean || var, "no output requested");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (info) cudaMemsetAsync(info, 0, sizeof(int), st);
  if ((rc = run_gram(h, st, X, xstar, N, M, Q, hyp, 0, workspace, M))) return rc;   // K_Xx [N,M]
  if ((rc = db_trsm_lower(h, stream, L, N, N, workspace, M, M, 0))) return rc;      // V1 = L^-1 K_Xx
  if (var) pred_var_kernel<<<(M + 127) / 128, 128, 0, st>>>(workspace, N, M, hyp, var, info);
  if (mean) {
    dim3 grid((D + TileSmall::BN - 1) / TileSmall::BN, (M + TileSmall::BM - 1) / TileSmall::BM);
    pred_mean_kernel<<<grid, 256, 0, st>>>(workspace, S, N, M, D, y_mean, mean);
  }
  return check_launch(h, "gp_predict");
}

// ---- GPDBN top layer -------------------------------------------------------------------------------------
extern "C" size_t db_gpdbn_workspace_bytes(int N, int Q, int D) {
  if (N <= 0 || Q <= 0 || D <= 0) return 0;
  return gpdbn_workspace(nullptr, N, Q, D).bytes;
}

extern "C" int db_gpdbn_factor(db_handle_t h, db_stream_t stream, const float* X, int N, int Q, int D, const float* hyp_opt,
                               int flags, float fixed_noise, float jitter, void* workspace, float* pred_var, float* L_out,
                               int* info) {
  DB_REQUIRE(h, X && hyp_opt && workspace && pred_var && info, "null operand");
  DB_REQUIRE(h, N > 1 && Q > 0 && Q <= MAXQ && D > 0, "bad shape (Q <= 8)");
  cudaStream_t st = (cudaStream_t)stream;
  GpdbnWorkspace w = gpdbn_workspace(workspace, N, Q, D);
  int rc;
  effective_hyp_kernel<<<1, 1, 0, st>>>(hyp_opt, flags, fixed_noise, jitter, w.hyp);
  cudaMemsetAsync(w.acc, 0, sizeof(GpAccum), st);
  cudaMemsetAsync(info, 0, 2 * sizeof(int), st);
  if ((rc = run_gram(h, st, X, nullptr, N, N, Q, w.hyp, 1, w.K, N))) return rc;
  const bool tc = gp_use_tensor_path(h, N, 32) && gpdbn_tensor_factor();
  F16Scratch fs{};
  const bool f16 = tc && gp_use_f16(N, 8);
  if (f16) { fs = f16_scratch(w.f16, N, 8); f16_gp_bounds(st, w.hyp, N, fs.scales); }
  const int overlap = f16 ? gp_overlap_mode(h, st, N) : 0;
  TrinvOverlap ov{};
  if (overlap) {
    ov = gp_overlap(h, w.Li, w.U, overlap, N);
    run_trinv_f16_side_prologue(h, ov, st, N);
  }
  if ((rc = run_potrf(h, st, w.K, N, N, info, tc ? w.Li : nullptr, f16 ? &fs : nullptr, overlap ? &ov : nullptr))) {      // clears info[0] itself
    if (overlap) run_trinv_f16_side_join(ov, st);
    return rc;
  }
  if (L_out) cudaMemcpyAsync(L_out, w.K, sizeof(float) * (size_t)N * N, cudaMemcpyDeviceToDevice, st);
  logdet_xsq_kernel<<<1, 1024, 0, st>>>(w.K, N, N, X, N * Q, w.acc);
  if (tc) {
    // Linv^T and K^-1 = Linv^T Linv on the tensor cores, diagonal of K^-1 in exact fp32 (it carries pred_var)
    const int tiles32 = (N + 31) / 32;
    if (f16) {
      if (overlap) {
        run_trinv_f16_side_join(ov, st);
        if (ov.err) return set_err(h, DB_ERR_CUDA, "f16x3 triangular inverse (side lanes) failed (%d)", ov.err);
      }
      else if ((rc = run_trinv_tc_f16(h, st, w.K, N, w.Li, w.U, fs))) return rc;
      const F16Operand U16{fs.u_hi, fs.u_lo, fs.scales + 1};
      const int trc = f16_gemm_launch(h, st, N, N, N, 1.0f, U16, N, U16, N, 0.0f, w.Kinv, N, 1, 1);
      if (trc) return set_err(h, DB_ERR_CUDA, "f16x3 K^-1 product failed (%d)", trc);
      kinv_diag_kernel<<<(N + 7) / 8, 256, 0, st>>>(w.U, N, w.Kinv, nullptr);
      // mirrored, and again as the fp16 pair the backward product (db_gpdbn_backward_x, 3 x fp16) reads
      mirror_lower_split_kernel<<<dim3((N + 63) / 64, (N + 63) / 64), 256, 0, st>>>(w.Kinv, N, fs.k_hi, fs.k_lo, fs.scales + 2);
    } else {
    if ((rc = run_trinv_tc(h, st, w.K, w.Li, N, w.Li, w.U, w.Ulo, w.WT, w.WTlo))) return rc;
    // the K buffer (L is dead) keeps the tf32 remainder of K^-1 for the backward GEMM (db_gpdbn_backward_x)
    const int trc = tf32_gemm_launch(h, st, N, N, N, 1.0f, w.U, w.Ulo, N, w.U, w.Ulo, N, 0.0f, w.Kinv, N, 1, 1, w.K);
    if (trc) return set_err(h, DB_ERR_CUDA, "tf32 K^-1 product failed (%d)", trc);
    mirror_lower_kernel<<<dim3(tiles32, tiles32), 256, 0, st>>>(w.Kinv, w.K, N);
    kinv_diag_kernel<<<(N + 7) / 8, 256, 0, st>>>(w.U, N, w.Kinv, w.K);
    }
  } else {
    if ((rc = run_trinv(h, st, w.K, N, N, w.Li, w.Kinv))) return rc;
    const int tiles = (N + TileLarge::BM - 1) / TileLarge::BM;
    const int tiles_s = (N + TileSmall::BM - 1) / TileSmall::BM;
    const int splits = simt_splits((long)tiles_s * (tiles_s + 1) / 2, N, h->num_sms);
    if (splits > 1) simt_launch_split(kinv_split_kernel, dim3(tiles_s, tiles_s, splits), 256, st, (const float*)w.Li, N, N, w.Kinv, splits);
    else kinv_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(w.Li, N, N, w.Kinv);
  }
  gpdbn_pred_var_kernel<<<(N + 255) / 256, 256, 0, st>>>(w.Kinv, N, w.hyp, pred_var, info);
  return check_launch(h, "gpdbn_factor");
}

extern "C" int db_gpdbn_apply(db_handle_t h, db_stream_t stream, const float* H2, int N, int Q, int D, void* workspace,
                              float* Hm, float* scalars) {
  DB_REQUIRE(h, H2 && workspace && Hm && scalars && N > 1 && D > 0, "null operand or bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  GpdbnWorkspace w = gpdbn_workspace(workspace, N, Q, D);
  int rc;
  cudaMemsetAsync(&w.acc->ssq, 0, sizeof(float), st);
  if ((rc = run_gemm_f32(h, st, 0, 0, N, D, N, 1.0f, w.Kinv, N, H2, D, 0.0f, w.B, D))) return rc;
  const size_t n = (size_t)N * D;
  gpdbn_apply_kernel<<<(int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256), 256, 0, st>>>(H2, w.B, n, w.hyp, Hm, w.acc);
  gpdbn_scalars_kernel<<<1, 1, 0, st>>>(w.acc, w.hyp, D, scalars);
  return check_launch(h, "gpdbn_apply");
}

extern "C" int db_gpdbn_backward_h(db_handle_t h, db_stream_t stream, const float* g_Hm, int N, int Q, int D, void* workspace,
                                   float* g_H2) {
  DB_REQUIRE(h, g_Hm && workspace && g_H2 && N > 1 && D > 0, "null operand or bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  GpdbnWorkspace w = gpdbn_workspace(workspace, N, Q, D);
  int rc;
  cudaMemsetAsync(&w.acc->pad[0], 0, 2 * sizeof(float), st);
  if ((rc = run_gemm_f32(h, st, 0, 0, N, D, N, 1.0f, w.Kinv, N, g_Hm, D, 0.0f, w.P, D))) return rc;
  const size_t n = (size_t)N * D;
  gpdbn_backward_h_kernel<<<(int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256), 256, 0, st>>>(g_Hm, w.P, w.B, n, w.hyp, g_H2,
                                                                                                   w.acc);
  return check_launch(h, "gpdbn_backward_h");
}

extern "C" int db_gpdbn_backward_x(db_handle_t h, db_stream_t stream, const float* X, const float* g_pv, int N, int Q, int D,
                                   const float* hyp_opt, int flags, float x_prior_weight, void* workspace, float* dX,
                                   float* dhyp) {
  DB_REQUIRE(h, X && g_pv && hyp_opt && workspace && dX && dhyp, "null operand");
  DB_REQUIRE(h, N > 1 && Q > 0 && Q <= MAXQ && D > 0, "bad shape (Q <= 8)");
  cudaStream_t st = (cudaStream_t)stream;
  GpdbnWorkspace w = gpdbn_workspace(workspace, N, Q, D);
  cudaMemsetAsync(w.dXacc, 0, sizeof(float) * (size_t)N * Q, st);
  cudaMemsetAsync(&w.acc->trG, 0, 3 * sizeof(float), st);      // trG, sumW, sumWr2
  gpdbn_gpv_scalar_kernel<<<(N + 255) / 256 > 64 ? 64 : (N + 255) / 256, 256, 0, st>>>(g_pv, w.Kinv, N, w.hyp, w.acc);
  const int tiles = (N + TileLarge::BM - 1) / TileLarge::BM;
  bool done = false;
  if (gp_use_tensor_path(h, N, 32) && gpdbn_tensor_factor() && gp_use_f16(N, 8) && (reinterpret_cast<uintptr_t>(g_pv) & 15) == 0) {
    // same decision as db_gpdbn_factor: the scratch holds K^-1 as a scaled fp16 pair (scale [2]); Xs goes where U was (dead
    // since the K^-1 product), scale [6] from its bound s max|g_pv|. 688 us as 3xTF32 at N = 5000.
    const F16Scratch fs = f16_scratch(w.f16, N, 8);
    const F16Operand K16{fs.k_hi, fs.k_lo, fs.scales + 2}, Xs16{fs.u_hi, fs.u_lo, fs.scales + 6};
    f16_scale_from_absmax(h, st, g_pv, 1, N, N, fs.bits + 2, fs.scales + 6, w.hyp + 2);
    kinv_scale_cols_f16_kernel<<<h->num_sms * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(w.Kinv), g_pv, N, w.hyp, fs.scales + 6,
                                                              fs.u_hi, fs.u_lo);
    const int trc = f16_gemm_launch(h, st, N, N, N, 1.0f, Xs16, N, K16, N, 0.0f, w.Ulo, N, 1, 0);
    if (trc != 0 && trc != -1000) return set_err(h, DB_ERR_CUDA, "f16x3 gradient product failed (%d)", trc);
    if (trc == 0) {
      gpdbn_grad_mem_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(w.Ulo, w.Kinv, w.P, w.B, X, N, Q, D, w.hyp, w.dXacc, w.acc);
      done = true;
    }
  }
  if (!done && gp_use_tensor_path(h, N, 32) && gpdbn_tensor_factor() && !gp_use_f16(N, 8) && (reinterpret_cast<uintptr_t>(g_pv) & 15) == 0) {
    // same decision as db_gpdbn_factor: the K buffer holds the remainder of K^-1
    kinv_scale_cols_kernel<<<h->num_sms * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(w.Kinv), g_pv, N, w.hyp,
                                                          reinterpret_cast<float4*>(w.Li), reinterpret_cast<float4*>(w.U));
    const int trc = tf32_gemm_launch(h, st, N, N, N, 1.0f, w.Li, w.U, N, w.Kinv, w.K, N, 0.0f, w.Ulo, N, 1, 0);
    if (trc != 0 && trc != -1000) return set_err(h, DB_ERR_CUDA, "tf32 gradient product failed (%d)", trc);
    if (trc == 0) {
      gpdbn_grad_mem_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(w.Ulo, w.Kinv, w.P, w.B, X, N, Q, D, w.hyp, w.dXacc, w.acc);
      done = true;
    }
  }
  if (!done) {
    const int tiles_s = (N + TileSmall::BM - 1) / TileSmall::BM;
    const int splits = simt_splits((long)tiles_s * (tiles_s + 1) / 2, N + 2 * D, h->num_sms);
    if (splits > 1) {       // few tiles (the horses runs, N = 328): cluster split-K into w.Ulo, contraction from memory
      simt_launch_split(gpdbn_g_split_kernel, dim3(tiles_s, tiles_s, splits), 256, st, (const float*)w.Kinv, g_pv, (const float*)w.P,
                        (const float*)w.B, N, D, (const float*)w.hyp, w.Ulo, splits);
      gp_contract_mem_kernel<TileSmall><<<dim3(tiles_s, tiles_s), 256, 0, st>>>(nullptr, w.Ulo, N, Q, D, X, w.hyp, w.dXacc, w.acc, nullptr);
    } else {
      gpdbn_grad_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(w.Kinv, g_pv, w.P, w.B, X, N, Q, D, w.hyp, w.dXacc, w.acc);
    }
  }
  gpdbn_finalize_kernel<<<(N * Q + 255) / 256, 256, 0, st>>>(w.acc, w.hyp, hyp_opt, flags, N, Q, x_prior_weight, X, w.dXacc, dX,
                                                            dhyp);
  return check_launch(h, "gpdbn_backward_x");
}
