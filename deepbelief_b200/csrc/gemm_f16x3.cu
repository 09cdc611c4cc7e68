// 3 x fp16 tensor-core GEMM for the big products of the GP chain (round 2): the same product as gemm_tf32.cu,
//
//   C[M, N] = alpha * X[M, K] * Y[N, K]^T + beta * C        fp32 result, both operands K-major,
//
// with every operand given as a power-of-two scaled fp16 pair, x * 2^e = hi + lo (round-to-nearest split, >= 22 significant
// bits — tf32(x) + tf32(x - tf32(x)) keeps 21 and truncates), and the three products hi*hi + lo*hi + hi*lo. A kind::f16 MMA
// does twice the FLOP of a kind::tf32 one in the same 64 cycles while reading the same bytes of shared memory, and the
// 128 x 128 tf32 tile is shared-memory-bandwidth bound (tensor pipe 48-50 % active, profiles/r01_gp_tensor_chain_summary.txt):
// per k-step of 8 the three MMAs read 24 KB in 192 cycles (128 B/clk, the whole LDS bandwidth) while TMA refills 64 KB per
// 768 cycles on top. Same persistent warp-specialised structure: warp 0 TMA producer (4 tensor maps, 128B swizzle, 3-stage
// ring of 64 KB = 64-column k-blocks), warp 1 tcgen05.mma kind::f16 128x128x16, warp 2 TMEM allocator (2 main + 2 correction
// accumulator stages), warps 4-7 epilogue; the hi*hi products are flushed to fp32 registers every K_CHUNK columns (the
// truncating accumulator, see gemm_tf32.cu), the 2^-11 smaller corrections once per tile.
// Scales are device scalars (2^e as float) chosen by the producer of each operand: a-priori bounds from the
// hyper-parameters (|L| <= sqrt(alpha + s), |L^-1| <= 1/sqrt(s), |K^-1| <= 1/s) or a measured maximum (Y, A).
// Entries far below the bound lose relative, not absolute, precision: the error of an entry is <= 2^-39 of the bound.
#include "gemm_f16x3.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>

namespace db {
namespace {

constexpr int BM = 128;       // rows of Y per tile (TMEM lanes)   -> columns of C
constexpr int BN = 128;       // rows of X per tile (TMEM columns) -> rows of C (128 partial sums per epilogue thread)
constexpr int BK = 64;        // fp16 elements = one 128-byte swizzle row
constexpr int UK = 16;        // fp16 MMA K
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;    // 16 KB
constexpr int B_BYTES = BN * BK * 2;    // 16 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // hi + lo of both operands: 64 KB
constexpr int ACC_STAGES = 2;     // main accumulators (hi*hi), flushed every K_CHUNK columns
constexpr int CORR_STAGES = 2;    // correction accumulators (lo*hi + hi*lo), flushed once per tile
constexpr int TMEM_COLS = (ACC_STAGES + CORR_STAGES) * BN;
constexpr int THREADS = 256;
constexpr int K_CHUNK = 256;      // 16 truncating adds per chunk, as the tf32 kernel's 128 columns
constexpr float kShrinkPerAdd = 4e-8f;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;

struct F16Params {
  int M, N, K;            // C is [M, N]; X [M, K], Y [N, K]
  float alpha, beta;
  const float* sx; const float* sy;     // device scalars: the power-of-two scales X and Y were written with
  float* C; int ldc;                    // fp32 result (may be null when only the fp16 pair is wanted)
  __half* c16_hi; __half* c16_lo; int ldc16; const float* sc;   // optional: the result again as a scaled fp16 pair (next operand)
  int lower_only, kb_chunk, tri_k, passes;
  float debias;
};

__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) { return ptx::umma_desc_k_sw128(smem_addr); }

__global__ void __launch_bounds__(THREADS, 1)
f16x3_gemm_kernel(const __grid_constant__ CUtensorMap tm_y_hi, const __grid_constant__ CUtensorMap tm_y_lo,
                   const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo, const F16Params p) {
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
      constexpr uint32_t idesc = ptx::umma_idesc(/*F16*/ 0, BM, BN);
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
              const uint64_t koff = (uint64_t)((k * UK * 2) >> 4);
              ptx::mma_f16_ss(tmem_d, d_yhi + koff, d_xhi + koff, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
              if (p.passes == 3) {
                ptx::mma_f16_ss(tmem_c, d_ylo + koff, d_xhi + koff, idesc, (kb != kb_first || k != 0) ? 1u : 0u);
                ptx::mma_f16_ss(tmem_c, d_yhi + koff, d_xlo + koff, idesc, 1u);
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
      const float alpha_eff = p.alpha / (__ldg(p.sx) * __ldg(p.sy));   // undoes the two power-of-two operand scales
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
      float* cp = p.C ? p.C + (size_t)row0 * p.ldc + col : nullptr;
      __half* hp = p.c16_hi ? p.c16_hi + (size_t)row0 * p.ldc16 + col : nullptr;
      __half* lp = p.c16_hi ? p.c16_lo + (size_t)row0 * p.ldc16 + col : nullptr;
      const float sc16 = p.c16_hi ? __ldg(p.sc) : 0.0f;
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
            float v = alpha_eff * sum[g * 16 + j];
            if (p.beta != 0.0f) v = fmaf(p.beta, oldv[j], v);
            if (cp != nullptr) cp[(size_t)(g * 16 + j) * p.ldc] = v;
            if (hp != nullptr) {
              __half hh, ll;
              split_f16(v * sc16, hh, ll);
              hp[(size_t)(g * 16 + j) * p.ldc16] = hh; lp[(size_t)(g * 16 + j) * p.ldc16] = ll;
            }
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


// x * scale -> fp16 hi + lo (round to nearest), [rows, cols] block, leading dimensions ld (fp32) and ld16.
// Vector form (cols, ld, ld16 multiples of 4, 16-byte aligned x / 8-byte aligned hi, lo): one float4 in, two 8-byte stores out.
__global__ void __launch_bounds__(256) f16_split_kernel(const float* __restrict__ x, size_t rows, size_t cols, size_t ld,
                                                        const float* __restrict__ scale, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, size_t ld16, int vec) {
  const float sc = __ldg(scale);
  if (vec) {
    const size_t c4n = cols / 4, total = rows * c4n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      const size_t r = i / c4n, c = (i - r * c4n) * 4;
      const float4 v = __ldcs(reinterpret_cast<const float4*>(x + r * ld + c));
      __half h0, l0, h1, l1, h2, l2, h3, l3;
      split_f16(v.x * sc, h0, l0); split_f16(v.y * sc, h1, l1); split_f16(v.z * sc, h2, l2); split_f16(v.w * sc, h3, l3);
      __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3), la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
      *reinterpret_cast<uint2*>(hi + r * ld16 + c) = make_uint2(*reinterpret_cast<uint32_t*>(&ha), *reinterpret_cast<uint32_t*>(&hb));
      *reinterpret_cast<uint2*>(lo + r * ld16 + c) = make_uint2(*reinterpret_cast<uint32_t*>(&la), *reinterpret_cast<uint32_t*>(&lb));
    }
    return;
  }
  const size_t total = rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    __half h, l;
    split_f16(__ldg(x + r * ld + c) * sc, h, l);
    hi[r * ld16 + c] = h; lo[r * ld16 + c] = l;
  }
}
__global__ void __launch_bounds__(256) f16_absmax_kernel(const float* __restrict__ x, size_t rows, size_t cols, size_t ld,
                                                         unsigned int* __restrict__ bits, int vec) {
  float m = 0.0f;
  if (vec) {
    const size_t c4n = cols / 4, total = rows * c4n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      const size_t r = i / c4n, c = (i - r * c4n) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    const size_t total = rows * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      const size_t r = i / cols, c = i - r * cols;
      m = fmaxf(m, fabsf(__ldg(x + r * ld + c)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(bits, __float_as_uint(m));
}
// scale = 2^(14 - ceil(log2(bound))): the bound lands in (2^13, 2^14]
__device__ __forceinline__ float scale_for_bound(float b) {
  if (!(b > 0.0f) || !isfinite(b)) return 1.0f;
  int q; const float f = frexpf(b, &q);            // b = f 2^q, f in [0.5, 1)
  if (f == 0.5f) --q;                               // exact power of two: b = 2^(q-1)
  return ldexpf(1.0f, max(-100, min(100, 14 - q)));
}
__global__ void f16_scale_from_bits_kernel(const unsigned int* __restrict__ bits, float* __restrict__ scale, const float* __restrict__ mul) {
  *scale = scale_for_bound(__uint_as_float(*bits) * (mul != nullptr ? fabsf(*mul) : 1.0f));
}
// scales[0] L panels, [1] L^-1 / U, [2] K^-1 from the effective hyper-parameters (alpha, gamma, s)
__global__ void f16_gp_bounds_kernel(const float* __restrict__ hyp, int N, float* __restrict__ scales) {
  const float alpha = hyp[0], s = hyp[2];
  scales[0] = scale_for_bound(sqrtf(alpha + s));                    // |L_ij| <= sqrt(K_ii)
  scales[1] = scale_for_bound(rsqrtf(s));                           // |L^-1_ij| <= |L^-1|_2 = lambda_min(K)^-1/2 <= s^-1/2
  scales[2] = scale_for_bound(1.0f / s);                            // |K^-1_ij| <= 1 / s
  scales[5] = scale_for_bound(sqrtf((float)N * (alpha + s) / s));   // |L_21 L_11^-1|_2 <= sqrt(lambda_max / lambda_min), lambda_max <= tr K
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn16() {
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
bool make_tmap_h(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  auto enc = encode_fn16();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int ew_blocks(size_t total, int num_sms) {
  size_t blocks = (total + 255) / 256, cap = (size_t)num_sms * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace

bool f16_gemm_supported(int M, int N, int K, const F16Operand& X, int ldx, const F16Operand& Y, int ldy) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(X.hi) | reinterpret_cast<uintptr_t>(X.lo) | reinterpret_cast<uintptr_t>(Y.hi) |
                      reinterpret_cast<uintptr_t>(Y.lo);
  return M > 0 && N > 0 && K > 0 && X.hi && X.lo && Y.hi && Y.lo && X.scale && Y.scale && (a & 15) == 0 && (ldx % 8 == 0) &&
         (ldy % 8 == 0) && ldx >= K && ldy >= K;
}

int f16_gemm_launch(db_handle_t h, cudaStream_t st, int M, int N, int K, float alpha, const F16Operand& X, int ldx,
                    const F16Operand& Y, int ldy, float beta, float* C, int ldc, int lower_only, int tri_k, const F16Operand* C16,
                    int ldc16, int max_ctas, int allow_pdl) {
  if (!f16_gemm_supported(M, N, K, X, ldx, Y, ldy)) return -1000;
  if (C == nullptr && (C16 == nullptr || beta != 0.0f)) return -1000;
  CUtensorMap tyh, tyl, txh, txl;
  if (!make_tmap_h(&tyh, Y.hi, N, K, ldy, BM) || !make_tmap_h(&tyl, Y.lo, N, K, ldy, BM) ||
      !make_tmap_h(&txh, X.hi, M, K, ldx, BN) || !make_tmap_h(&txl, X.lo, M, K, ldx, BN)) return -1001;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(f16x3_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return (int)attr_err;
  F16Params p;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.sx = X.scale; p.sy = Y.scale; p.C = C; p.ldc = ldc;
  p.c16_hi = C16 ? C16->hi : nullptr; p.c16_lo = C16 ? C16->lo : nullptr; p.ldc16 = ldc16; p.sc = C16 ? C16->scale : nullptr;
  p.lower_only = lower_only;
  int kchunk = K_CHUNK;
  { const char* e = getenv("DB_F16_KCHUNK"); if (e) kchunk = atoi(e) > 0 ? atoi(e) : K_CHUNK; }
  p.kb_chunk = kchunk / BK < 1 ? 1 : kchunk / BK;
  p.tri_k = tri_k;
  p.passes = 3;
  {
    float per_add = kShrinkPerAdd;
    const char* e = getenv("DB_F16_DEBIAS"); if (e) per_add = (float)atof(e);
    p.debias = 1.0f + per_add * (float)(p.kb_chunk * (BK / UK));
  }
  const int tm = (N + BM - 1) / BM, tn = (M + BN - 1) / BN;
  const int tiles = (lower_only && tm == tn) ? tn * (tn + 1) / 2 : tm * tn;
  int ctas = tc_ctas(h);
  if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
  const int grid = tiles < ctas ? tiles : ctas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  static const int pdl = [] { const char* e = getenv("DB_F16_PDL"); return (e && e[0] == '0') ? 0 : 1; }();   // diagnostics
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (pdl && allow_pdl) ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, f16x3_gemm_kernel, tyh, tyl, txh, txl, p);
}

void f16_split(db_handle_t h, cudaStream_t st, const float* x, size_t rows, size_t cols, size_t ld, const F16Operand& out, size_t ld16) {
  const int vec = (cols % 4 == 0) && (ld % 4 == 0) && (ld16 % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(out.hi) | reinterpret_cast<uintptr_t>(out.lo)) & 7) == 0;
  f16_split_kernel<<<ew_blocks(rows * cols / (vec ? 4 : 1), h->num_sms), 256, 0, st>>>(x, rows, cols, ld, out.scale, out.hi, out.lo, ld16, vec);
}
void f16_scale_from_absmax(db_handle_t h, cudaStream_t st, const float* x, size_t rows, size_t cols, size_t ld, unsigned int* bits_scratch,
                           float* scale_out, const float* bound_mul) {
  cudaMemsetAsync(bits_scratch, 0, sizeof(unsigned int), st);
  const int vec = (cols % 4 == 0) && (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  f16_absmax_kernel<<<ew_blocks(rows * cols / (vec ? 4 : 1), h->num_sms), 256, 0, st>>>(x, rows, cols, ld, bits_scratch, vec);
  f16_scale_from_bits_kernel<<<1, 1, 0, st>>>(bits_scratch, scale_out, bound_mul);
}
void f16_gp_bounds(cudaStream_t st, const float* hyp, int N, float* scales) { f16_gp_bounds_kernel<<<1, 1, 0, st>>>(hyp, N, scales); }

}  // namespace db

using namespace db;

// scratch: [256 B: two scale floats, two absmax words] [X hi | X lo : M*Kp halves each] [Y hi | Y lo : N*Kp halves each], Kp = K rounded to 8
extern "C" size_t db_gemm_f16x3_scratch_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const size_t Kp = ((size_t)K + 7) / 8 * 8;
  return 256 + 2 * ((size_t)M + (size_t)N) * Kp * sizeof(__half) + 64;
}

extern "C" int db_gemm_f16x3_nt(db_handle_t h, db_stream_t stream, int M, int N, int K, float alpha, const float* X, int ldx,
                                const float* Y, int ldy, float beta, float* C, int ldc, int lower_only, void* scratch) {
  DB_REQUIRE(h, X && Y && C && scratch && M > 0 && N > 0 && K > 0 && ldc >= N && ldx >= K && ldy >= K, "null operand or bad shape");
  if (h->cc_major < 10) return set_err(h, DB_ERR_UNSUPPORTED, "tcgen05 path needs sm_100a (device is sm_%d%d)", h->cc_major, h->cc_minor);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Kp = ((size_t)K + 7) / 8 * 8;
  uint8_t* p = static_cast<uint8_t*>(scratch);
  float* scales = reinterpret_cast<float*>(p);
  unsigned int* bits = reinterpret_cast<unsigned int*>(p + 64);
  __half* base = reinterpret_cast<__half*>(p + 256);
  F16Operand Xo{base, base + (size_t)M * Kp, scales};
  F16Operand Yo{base + 2 * (size_t)M * Kp, base + 2 * (size_t)M * Kp + (size_t)N * Kp, scales + 1};
  f16_scale_from_absmax(h, st, X, M, K, ldx, bits, scales);
  f16_scale_from_absmax(h, st, Y, N, K, ldy, bits + 1, scales + 1);
  f16_split(h, st, X, M, K, ldx, Xo, Kp);
  f16_split(h, st, Y, N, K, ldy, Yo, Kp);
  const int rc = f16_gemm_launch(h, st, M, N, K, alpha, Xo, (int)Kp, Yo, (int)Kp, beta, C, ldc, lower_only, 0, nullptr, 0);
  if (rc == -1000) return set_err(h, DB_ERR_UNSUPPORTED, "3 x fp16 GEMM needs a 16-byte aligned scratch buffer");
  if (rc) return set_err(h, DB_ERR_CUDA, "f16x3 GEMM launch failed (%d)", rc);
  return check_launch(h, "gemm_f16x3_nt");
}
