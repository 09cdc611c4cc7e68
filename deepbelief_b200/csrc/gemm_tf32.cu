// 3xTF32 tensor-core GEMM for the GP chain (gplvm.py:212 tf.cholesky trailing updates, :235 / :248 triangular
// solves expressed through L^-1, and the K^-1 / A A^T products of the analytic gradient):
//
//   C[M, N] = alpha * X[M, K] * Y[N, K]^T + beta * C        fp32 in / fp32 out, both operands K-major (row-major)
//
// Every operand is used as x = hi + lo with hi = tf32(x) (the tensor core reads the top 19 bits of the fp32 word
// in shared memory: no conversion pass) and lo = x - hi precomputed once per operand (db_tf32_split_lo); the three
// products hi*hi + lo*hi + hi*lo accumulate into one TMEM tile, which restores fp32-level accuracy (~2^-21).
// The tensor core adds into the fp32 accumulator with truncation, so K is cut into chunks of 128 columns for the
// hi*hi product (its own accumulator; the corrections use a second one); the partial tiles are summed by the
// epilogue threads in registers, in fp32 round-to-nearest (deterministic).
//
// Same persistent, warp-specialised structure as gemm_tc.cu: warp 0 TMA producer (4 tensor maps, 128B swizzle,
// 128B swizzle, 3-stage ring of 64 KB), warp 1 tcgen05.mma kind::tf32 128x128x8, warp 2 TMEM allocator (2 main + 2
// correction accumulator stages), warps 4-7 epilogue.
// Output orientation: TMEM lanes = rows of Y (= columns of C), so a warp writes 32 consecutive floats of a row of C.
#include "context.h"
#include "ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>

namespace db {
namespace {

constexpr int BM = 128;       // rows of Y per tile (TMEM lanes)   -> columns of C
constexpr int BN = 128;       // rows of X per tile (TMEM columns) -> rows of C (128 partial sums per epilogue thread)
constexpr int BK = 32;        // fp32 elements = one 128-byte swizzle row
constexpr int UK = 8;         // tf32 MMA K
constexpr int STAGES = 3;         // 3 x 64 KB in flight; 12 MMAs (768 cycles) per barrier round trip (BK = 16 / 6 stages
                                  // kept the tensor pipe only 48-50 % busy: the single issuing thread was handshake-bound)
constexpr int A_BYTES = BM * BK * 4;    // 16 KB
constexpr int B_BYTES = BN * BK * 4;    // 16 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // hi + lo of both operands: 64 KB
constexpr int ACC_STAGES = 2;     // main accumulators (hi*hi), flushed every K_CHUNK columns
constexpr int CORR_STAGES = 2;    // correction accumulators (lo*hi + hi*lo), flushed once per tile
constexpr int TMEM_COLS = (ACC_STAGES + CORR_STAGES) * BN;
constexpr int THREADS = 256;
// The tensor core adds into its fp32 accumulator with truncation toward zero: every add shrinks the partial sum by
// ~half an ulp, a systematic relative bias of ~T * 2^-24 after T adds (measured: 2.4e-6 at T = 48). The hi*hi products
// therefore get their own accumulator, flushed every 128 columns (T = 16); the two correction products are 2^-11
// smaller, so their accumulator can run over the whole K range of a tile. Measured on tr G of the N=1500 GPLVM test (a
// 89x cancelling difference, scripts/diag_trace.py): chunk 32 / 64 / 128 / 512 / 2048 -> 2.5e-4 / 1.8e-4 / 1.3e-4 /
// 4.4e-4 / 8.4e-4 relative (one shared accumulator at 128: 5.0e-4).
constexpr int K_CHUNK = 128;
// Each truncating add loses on average ~0.7 ulp-halves of the running sum; the epilogue multiplies every hi*hi chunk by
// 1 + kShrinkPerAdd * adds to remove the expected shrink. Calibrated on random data (scripts/check_tf32.py: max error vs
// fp64 at K = 4096 is 7.9e-7 / 4.9e-7 / 3.8e-7 / 6.0e-7 for 0 / 2e-8 / 4e-8 / 6e-8; scripts/diag_trace.py: tr G error
// 2.6e-4 / 1.5e-4 / 1.5e-4 / 2.6e-5).
constexpr float kShrinkPerAdd = 4e-8f;     // 16 adds per chunk -> 1 + 5 ulp
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;

struct Tf32Params {
  int M, N, K;            // C is [M, N]; X [M, K], Y [N, K]
  float alpha, beta;
  float* C; int ldc;
  float* C_lo;            // optional: tf32 remainder of the written values (same layout), so C can be the next operand
  int lower_only;         // 1: only elements with col <= row are produced (tiles above the diagonal are skipped)
  int kb_chunk;           // k-blocks (of BK columns) accumulated in TMEM before the partial tile is added to C
  int tri_k;              // 1: X[i,k] = 0 for k < i and Y[j,k] = 0 for k < j (transposed triangular factors): k starts at max(i0, j0)
                          // 2: only X is upper triangular (X[i,k] = 0 for k < i): k starts at the first row of the tile
  int passes;             // 3 (debug: DB_TF32_PASSES=1 issues only the hi*hi product)
  float debias;           // 1 + expected relative shrink of one K chunk of the truncating hi*hi accumulator
};

// K-major operand tile written by TMA with SWIZZLE_128B: rows are 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) { return ptx::umma_desc_k_sw128(smem_addr); }

__global__ void __launch_bounds__(THREADS, 1)
tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap tm_y_hi, const __grid_constant__ CUtensorMap tm_y_lo,
                   const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo, const Tf32Params p) {
#if defined(DB_SM100)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint64_t* cfull_bar = tempty_bar + ACC_STAGES;
  uint64_t* cempty_bar = cfull_bar + CORR_STAGES;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(cempty_bar + CORR_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.N + BM - 1) / BM;     // along columns of C
  const int tiles_n = (p.M + BN - 1) / BN;     // along rows of C
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int kb_per_chunk = p.kb_chunk;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tm_y_hi); ptx::tma_prefetch_desc(&tm_y_lo);
    ptx::tma_prefetch_desc(&tm_x_hi); ptx::tma_prefetch_desc(&tm_x_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < ACC_STAGES; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 128); }
    for (int a = 0; a < CORR_STAGES; ++a) { ptx::mbar_init(&cfull_bar[a], 1); ptx::mbar_init(&cempty_bar[a], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(tmem_base_smem, TMEM_COLS); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  // programmatic dependent launch (as in gemm_tc.cu): the prologue above overlaps the tail of the previous kernel of the GP
  // chain; no global access before that kernel has completed
  ptx::grid_launch_dependents();
  ptx::grid_dependency_wait();

  // tile t -> (column block of C, row block of C); row-block-major so consecutive CTAs share the X rows in L2.
  // lower_only on a square tile grid enumerates the T (T + 1) / 2 lower tiles only (row block i, column block <= i, i
  // ascending): the CTAs get equal tile counts, and with tri_k the long-K tiles (small i) are dealt first. Measured on
  // the K = 512 trailing update with the rectangular enumeration + skipping: SMs active 51 % of the kernel.
  const bool tri_enum = p.lower_only && tiles_m == tiles_n;
  const int total_tiles = tri_enum ? tiles_n * (tiles_n + 1) / 2 : num_tiles;
  auto tile_coords = [&](int t, int& m0, int& n0) {
    if (tri_enum) {
      int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
      while (i * (i + 1) / 2 > t) --i;
      while ((i + 1) * (i + 2) / 2 <= t) ++i;
      n0 = i * BN; m0 = (t - i * (i + 1) / 2) * BM;
    } else {
      m0 = (t % tiles_m) * BM; n0 = (t / tiles_m) * BN;
    }
  };
  auto skipped_at = [&](int m0, int n0) { return p.lower_only && m0 > n0 + BN - 1; };   // min col > max row
  auto kb0_at = [&](int m0, int n0) { return p.tri_k == 1 ? max(m0, n0) / BK : (p.tri_k == 2 ? n0 / BK : 0); };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0;
        tile_coords(t, m0, n0);
        if (skipped_at(m0, n0)) continue;
        for (int kb = kb0_at(m0, n0); kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          ptx::tma_load_2d(st, &tm_y_hi, &full_bar[stage], kb * BK, m0);
          ptx::tma_load_2d(st + A_BYTES, &tm_y_lo, &full_bar[stage], kb * BK, m0);
          ptx::tma_load_2d(st + 2 * A_BYTES, &tm_x_hi, &full_bar[stage], kb * BK, n0);
          ptx::tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tm_x_lo, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(/*TF32*/ 2, BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int cacc = 0; uint32_t cacc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0;
        tile_coords(t, m0, n0);
        if (skipped_at(m0, n0)) continue;
        const int kb_first = kb0_at(m0, n0);
        if (kb_first >= num_kb) continue;
        const uint32_t tmem_c = tmem_base + (uint32_t)((ACC_STAGES + cacc) * BN);
        if (p.passes == 3) { ptx::mbar_wait(&cempty_bar[cacc], cacc_phase ^ 1); ptx::tc_fence_after(); }
        for (int kb_begin = kb_first; kb_begin < num_kb; kb_begin += kb_per_chunk) {
          const int kb_end = min(num_kb, kb_begin + kb_per_chunk);
          ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
          for (int kb = kb_begin; kb < kb_end; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * STAGE_BYTES);
            const uint64_t d_yhi = umma_desc_k_sw64(sa), d_ylo = umma_desc_k_sw64(sa + A_BYTES);
            const uint64_t d_xhi = umma_desc_k_sw64(sa + 2 * A_BYTES), d_xlo = umma_desc_k_sw64(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t koff = (uint64_t)((k * UK * 4) >> 4);
              ptx::mma_tf32_ss(tmem_d, d_yhi + koff, d_xhi + koff, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
              if (p.passes == 3) {
                ptx::mma_tf32_ss(tmem_c, d_ylo + koff, d_xhi + koff, idesc, (kb != kb_first || k != 0) ? 1u : 0u);
                ptx::mma_tf32_ss(tmem_c, d_yhi + koff, d_xlo + koff, idesc, 1u);
              }
            }
            ptx::mma_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit(&tfull_bar[acc]);
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (p.passes == 3) {
          ptx::mma_commit(&cfull_bar[cacc]);
          if (++cacc == CORR_STAGES) { cacc = 0; cacc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // Epilogue: the K-chunk partial tiles are summed in REGISTERS in fp32 round-to-nearest (thread = one column of C,
    // 128 rows), so the truncating tensor-core accumulator never sees more than K_CHUNK columns; one write at the end.
    const int quarter = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    int cacc = 0; uint32_t cacc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int m0, n0;
      tile_coords(t, m0, n0);
      if (skipped_at(m0, n0)) continue;
      const int col = m0 + quarter * 32 + lane;                // column of C handled by this thread
      const int row0 = n0;
      const bool col_ok = col < p.N;
      float sum[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) sum[j] = 0.0f;
      const int kb_first = kb0_at(m0, n0);
      for (int kb_begin = kb_first; kb_begin < num_kb; kb_begin += kb_per_chunk) {
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr0 + (uint32_t)(chunk * 32), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[chunk * 32 + j] = fmaf(__uint_as_float(r[j]), p.debias, sum[chunk * 32 + j]);
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty_bar[acc]);
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      if (p.passes == 3 && kb_first < num_kb) {
        ptx::mbar_wait(&cfull_bar[cacc], cacc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((ACC_STAGES + cacc) * BN);
#pragma unroll
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr0 + (uint32_t)(chunk * 32), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[chunk * 32 + j] += __uint_as_float(r[j]);
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&cempty_bar[cacc]);
        if (++cacc == CORR_STAGES) { cacc = 0; cacc_phase ^= 1; }
      }
      // rows [jlo, jhi) of this thread's column are written: inside C, and at or below the diagonal when lower_only.
      // One pointer walks down the column; all loads of a read-modify-write group are issued before its first store.
      const int jlo = p.lower_only ? max(0, col - row0) : 0;
      const int jhi = col_ok ? min(BN, p.M - row0) : 0;
      const unsigned span = jhi > jlo ? (unsigned)(jhi - jlo) : 0u;
      float* cp = p.C + (size_t)row0 * p.ldc + col;
      float* lp = p.C_lo ? p.C_lo + (size_t)row0 * p.ldc + col : nullptr;
#pragma unroll
      for (int g = 0; g < BN / 16; ++g) {
        float oldv[16];
        if (p.beta != 0.0f) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            oldv[j] = (unsigned)(g * 16 + j - jlo) < span ? __ldcg(cp + (size_t)(g * 16 + j) * p.ldc) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if ((unsigned)(g * 16 + j - jlo) < span) {
            float v = p.alpha * sum[g * 16 + j];
            if (p.beta != 0.0f) v = fmaf(p.beta, oldv[j], v);
            cp[(size_t)(g * 16 + j) * p.ldc] = v;
            if (lp != nullptr) lp[(size_t)(g * 16 + j) * p.ldc] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, TMEM_COLS); }
#endif
}

__global__ void tf32_split_lo_kernel(const float* __restrict__ x, size_t rows, size_t cols, size_t ld, float* __restrict__ lo) {
  const size_t total = rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    const float v = __ldg(x + r * ld + c);
    lo[r * ld + c] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
}

// [R, C] row-major (ld_in) -> [C, R] row-major (ld_out), optionally also the tf32 remainder of the transposed copy
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ in, int R, int C, int ld_in,
                                                            float* __restrict__ out, int ld_out) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < R && c < C) ? __ldg(in + (size_t)r * ld_in + c) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < C && r < R) out[(size_t)c * ld_out + r] = tile[tx][ty + 8 * i];
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

bool make_tmap_f32(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  auto enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool tf32_gemm_supported(int M, int N, int K, const float* X, const float* X_lo, int ldx, const float* Y, const float* Y_lo, int ldy) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(X_lo) | reinterpret_cast<uintptr_t>(Y) |
                      reinterpret_cast<uintptr_t>(Y_lo);
  return M > 0 && N > 0 && K > 0 && (a & 15) == 0 && (ldx % 4 == 0) && (ldy % 4 == 0) && ldx >= K && ldy >= K;
}

// C[M,N] = alpha X Y^T + beta C; returns 0, or a negative code when the shape cannot use TMA (caller falls back to SIMT)
int tf32_gemm_launch(db_handle_t h, cudaStream_t st, int M, int N, int K, float alpha, const float* X, const float* X_lo, int ldx,
                     const float* Y, const float* Y_lo, int ldy, float beta, float* C, int ldc, int lower_only, int tri_k, float* C_lo) {
  if (!tf32_gemm_supported(M, N, K, X, X_lo, ldx, Y, Y_lo, ldy)) return -1000;
  CUtensorMap tyh, tyl, txh, txl;
  if (!make_tmap_f32(&tyh, Y, N, K, ldy, BM) || !make_tmap_f32(&tyl, Y_lo, N, K, ldy, BM) ||
      !make_tmap_f32(&txh, X, M, K, ldx, BN) || !make_tmap_f32(&txl, X_lo, M, K, ldx, BN)) return -1001;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tf32x3_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return (int)attr_err;
  Tf32Params p;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C = C; p.ldc = ldc; p.C_lo = C_lo; p.lower_only = lower_only;
  int kchunk = K_CHUNK;
  { const char* e = getenv("DB_TF32_KCHUNK"); if (e) kchunk = atoi(e) > 0 ? atoi(e) : K_CHUNK; }
  p.kb_chunk = kchunk / BK < 1 ? 1 : kchunk / BK;
  p.tri_k = tri_k;
  { const char* e = getenv("DB_TF32_PASSES"); p.passes = (e && e[0] == '1') ? 1 : 3; }
  {
    // every accumulator add truncates toward zero: expected relative loss per add ~ kShrinkPerAdd of the running sum
    float per_add = kShrinkPerAdd;
    const char* e = getenv("DB_TF32_DEBIAS"); if (e) per_add = (float)atof(e);
    const int adds = p.kb_chunk * (BK / UK);
    p.debias = 1.0f + per_add * (float)adds;
  }
  const int tm = (N + BM - 1) / BM, tn = (M + BN - 1) / BN;
  const int tiles = (lower_only && tm == tn) ? tn * (tn + 1) / 2 : tm * tn;
  const int ctas = tc_ctas(h);
  const int grid = tiles < ctas ? tiles : ctas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = getenv("DB_TF32_PDL"); return (e && e[0] == '0') ? 0 : 1; }();   // diagnostics
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, tf32x3_gemm_kernel, tyh, tyl, txh, txl, p);
}

void tf32_split_lo(db_handle_t h, cudaStream_t st, const float* x, size_t rows, size_t cols, size_t ld, float* lo) {
  const size_t total = rows * cols;
  size_t blocks = (total + 255) / 256, cap = (size_t)h->num_sms * 8;
  tf32_split_lo_kernel<<<(int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks)), 256, 0, st>>>(x, rows, cols, ld, lo);
}

void transpose_f32(cudaStream_t st, const float* in, int R, int C, int ld_in, float* out, int ld_out) {
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  transpose_f32_kernel<<<grid, 256, 0, st>>>(in, R, C, ld_in, out, ld_out);
}

}  // namespace db

using namespace db;

extern "C" int db_tf32_split_lo(db_handle_t h, db_stream_t stream, const float* x, int rows, int cols, int ld, float* lo) {
  DB_REQUIRE(h, x && lo && rows > 0 && cols > 0 && ld >= cols, "null operand or bad shape");
  tf32_split_lo(h, (cudaStream_t)stream, x, rows, cols, ld, lo);
  return check_launch(h, "tf32_split_lo");
}

extern "C" int db_gemm_tf32x3_nt(db_handle_t h, db_stream_t stream, int M, int N, int K, float alpha, const float* X,
                                 const float* X_lo, int ldx, const float* Y, const float* Y_lo, int ldy, float beta, float* C,
                                 int ldc, int lower_only) {
  DB_REQUIRE(h, X && X_lo && Y && Y_lo && C && M > 0 && N > 0 && K > 0 && ldc >= N, "null operand or bad shape");
  if (h->cc_major < 10) return set_err(h, DB_ERR_UNSUPPORTED, "tcgen05 path needs sm_100a (device is sm_%d%d)", h->cc_major, h->cc_minor);
  const int rc = tf32_gemm_launch(h, (cudaStream_t)stream, M, N, K, alpha, X, X_lo, ldx, Y, Y_lo, ldy, beta, C, ldc, lower_only, 0, nullptr);
  if (rc == -1000) return set_err(h, DB_ERR_UNSUPPORTED, "3xTF32 GEMM needs 16-byte aligned operands and leading dimensions that are multiples of 4");
  if (rc) return set_err(h, DB_ERR_CUDA, "tf32x3 GEMM launch failed (%d)", rc);
  return DB_OK;
}
