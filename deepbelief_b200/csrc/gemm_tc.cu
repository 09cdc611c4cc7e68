// K1/K2 tensor-core form: persistent, warp-specialised tcgen05 GEMM for the binary-RBM CD chain.
//
//   D[Ma, Nb] = A[Ma, K] * B[Nb, K]^T      (fp16 operands, fp32 accumulate in TMEM)
//
// Replaces the TF ops behind rbm.py:129-131 (v@W+c), rbm.py:143-147 (W@h^T+b), rbm.py:160/174
// (sigmoid), rbm.py:202-208/230-236 (uniform draw + strict compare) and the two dW products TF
// autodiff derives from rbm.py:388-415 — one launch each, nothing materialised in between.
//
// Precision: one operand of every CD GEMM of a binary-binary RBM is exactly {0,1} (SURVEY.md §7
// hard part 1), so only the real operand (W, or the probabilities in dW) is split: x*2^s = hi + lo
// in fp16 (>= 22 significant bits), two kind::f16 MMA passes accumulate into the same TMEM tile.
// That is fp32-level accuracy at 2 fp16 passes = the cost of ONE TF32 pass (fp16 runs at 2x TF32).
//
// Accumulation: the tensor core adds into its fp32 TMEM accumulator with truncation toward zero. The hi products
// (an 11-bit fp16 value times {0,1}) land on or above the accumulator's last bit and add EXACTLY; the lo products (the
// 2^-11 remainder) always fall below it, and every lo MMA shaves ~half an ulp off the running sum: measured against fp64
// on the pre-activation of a 4096 / 8192-deep layer, a mean shrink of 4e-6 / 8e-6 and probabilities off by up to
// 1.3e-5 / 3.5e-5 relative, outside the 1e-5 the path must hold. The two passes therefore accumulate into SEPARATE TMEM
// tiles (H: columns 0-255, L: columns 256-511): in L the same truncation is relative to a 2^-11 times smaller sum. The
// epilogue adds H + L once, in fp32 round-to-nearest, while it drains both into registers (128 sums per thread);
// the TMEM is released as soon as the drain is done, and the epilogue math of tile t runs under the MMAs of tile t+1.
// (Promoting one shared accumulator into registers every 512 / 1024 / 2048 columns instead was measured at 2e-6 / 4e-6 /
// 9e-6 and +15 % / +7 % / +1 % time: the 20 us of epilogue math per tile delays the promotions of the next tile.)
// The dW product runs over its whole K = 2 x batch range the same way (no split-K partial tiles in HBM).
//
// Roles (384 threads, 1 CTA/SM, persistent over tiles; setmaxnreg moves registers from warps 0-3 to the epilogue):
//   warp 0    : TMA producer   (A_hi, A_lo, B tiles -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1    : MMA issuer     (one elected thread: tcgen05.mma, tcgen05.commit)
//   warp 2    : TMEM allocator (512 columns = the hi and the lo accumulator of one 128x256 tile)
//   warps 4-11: epilogue       (tcgen05.ld of H and L + add, release; then scale+bias -> logistic -> Philox/injected
//                               uniform -> compare -> global stores), overlapped with the next tile's MMAs.
//                               Warp w owns TMEM lanes 32 (w % 4) .. +32 and the column half (w - 4) / 4 of the tile.
#include "../../include/deepbelief_b200.h"
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>

namespace db {

namespace {

constexpr int BLOCK_M = 128;   // rows of A per tile  (TMEM lanes)
constexpr int BLOCK_N = 256;   // rows of B per tile  (TMEM columns) at full width; p.tile_n = 256 / 128 / 64 is what a launch uses
constexpr int BLOCK_K = 64;    // fp16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + B_TILE_BYTES;   // 64 KB
constexpr int TMEM_COLS = 2 * BLOCK_N;                // 512: H (hi products) | L (lo products)
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;       // warps 4-11
constexpr int REGS_EPI = 232, REGS_CTRL = 40;   // 256 x 232 + 128 x 40 = 64512 <= 65536
constexpr int GROUP_M_DEFAULT = 16;   // measured at config 5 (V=4096, H=8192, N=16384): 8 / 16 / 32 / 64 -> 32.58 / 32.43 / 33.75 / 35.55 ms per step
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TileCoord { int m_blk, n_blk; };

__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_m, int tiles_n, int group_m) {
  // groups of group_m row-blocks sweep all column-blocks: CTAs running together share A and B tiles in L2
  const int group_size = group_m * tiles_n;
  const int g = t / group_size;
  const int first_m = g * group_m;
  const int gm = min(tiles_m - first_m, group_m);
  const int in_g = t - g * group_size;
  TileCoord c;
  c.m_blk = first_m + in_g % gm;
  c.n_blk = in_g / gm;
  return c;
}

// work item -> (k-split, tile, column offset and width inside the tile)
struct WorkItem { int ks, tile, n_off, bn; };
__device__ __forceinline__ WorkItem decode_work(int w, int num_tiles, int num_full, int k_splits, int tile_n) {
  WorkItem wi;
  if (k_splits > 1) { wi.ks = w / num_tiles; wi.tile = w - wi.ks * num_tiles; wi.n_off = 0; wi.bn = tile_n; return wi; }
  wi.ks = 0;
  if (w < num_full) { wi.tile = w; wi.n_off = 0; wi.bn = tile_n; return wi; }
  const int h = w - num_full;
  wi.tile = num_full + (h >> 1); wi.n_off = (h & 1) * (tile_n / 2); wi.bn = tile_n / 2;
  return wi;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// One 32-column slice of the tile, values in registers (val[j] = accumulator of column nb0 + j for row m of D):
// scale + bias -> logistic -> sample -> the requested outputs.
__device__ __forceinline__ void epilogue_slice(const TcGemmParams& p, float (&val)[32], int nb0, int m, bool m_ok, float bias_m,
                                               float acc_scale, float* __restrict__ out32, uint32_t draw_lo, uint32_t draw_hi) {
  const float inv_pT = p.pT_scale;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = fmaf(val[j], acc_scale, bias_m);
    val[j] = p.act ? __fdividef(1.0f, 1.0f + __expf(-x)) : x;
  }

  uint32_t sbits = 0;
  if (p.sample) {
    if (p.U != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int nb = nb0 + j;
        float u = (m_ok && nb < p.Nb) ? __ldg(p.U + (size_t)nb * p.ldu + m) : 2.0f;
        sbits |= (val[j] > u ? 1u : 0u) << j;
      }
    } else {
      const uint32_t grow0 = p.rng.row_offset + (uint32_t)nb0;   // multiple of 4 by construction
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        Philox4 o = philox4x32_10((uint32_t)m, (grow0 >> 2) + q, draw_lo, draw_hi, p.rng.seed_lo, p.rng.seed_hi);
        sbits |= (val[4 * q + 0] > u32_to_uniform(o.x) ? 1u : 0u) << (4 * q + 0);
        sbits |= (val[4 * q + 1] > u32_to_uniform(o.y) ? 1u : 0u) << (4 * q + 1);
        sbits |= (val[4 * q + 2] > u32_to_uniform(o.z) ? 1u : 0u) << (4 * q + 2);
        sbits |= (val[4 * q + 3] > u32_to_uniform(o.w) ? 1u : 0u) << (4 * q + 3);
      }
    }
  }

  // ---- [Nb, Ma] outputs: for fixed j the 32 lanes write 32 consecutive units ----
  if (out32 != nullptr && m_ok) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (nb0 + j < p.Nb) out32[(size_t)(nb0 + j) * p.ld32 + m] = val[j];
  }
  // ---- fused K2 tail: val = gradient tile -> optimizer step on theta, split-weight cache rewritten ----
  if (p.opt.kind >= 0) {
    const TcOptArgs& o = p.opt;
    const float sc = __ldg(o.scale_dev);
    const OptRule rule{o.kind, o.lr, o.lr_t, o.momentum, o.beta1, o.beta2, o.eps, o.weight_decay};
    float mx = 0.0f;
    if (m_ok) {
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 4) {
        float th[4], a[4], b[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {          // all loads of the group first
          const bool ok = nb0 + j0 + jj < p.Nb;
          const size_t idx = (size_t)(nb0 + j0 + jj) * p.ld32 + m;
          th[jj] = ok ? o.theta[idx] : 0.0f;
          a[jj] = (ok && o.kind != DB_OPT_SGD) ? o.st[idx] : 0.0f;
          b[jj] = (ok && o.kind == DB_OPT_ADAM) ? o.st2[idx] : 0.0f;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          if (nb0 + j0 + jj >= p.Nb) { val[j0 + jj] = 0.0f; continue; }
          const size_t idx = (size_t)(nb0 + j0 + jj) * p.ld32 + m;
          float s1 = a[jj], s2 = b[jj];
          const float t = opt_apply(rule, th[jj], val[j0 + jj], s1, s2);
          if (o.kind != DB_OPT_SGD) o.st[idx] = s1;
          if (o.kind == DB_OPT_ADAM) o.st2[idx] = s2;
          o.theta[idx] = t;
          mx = fmaxf(mx, fabsf(t));
          val[j0 + jj] = t * sc;
          __half h, l;
          split_f16(val[j0 + jj], h, l);
          const size_t nidx = (size_t)(nb0 + j0 + jj) * o.ldn + m;
          o.n_hi[nidx] = h; o.n_lo[nidx] = l;
        }
      }
      __half* dhi = o.t_hi + (size_t)m * o.ldt + nb0;
      __half* dlo = o.t_lo + (size_t)m * o.ldt + nb0;
      if (nb0 + 32 <= p.Nb) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __half h0, l0, h1, l1;
            split_f16(val[8 * g + 2 * e], h0, l0);
            split_f16(val[8 * g + 2 * e + 1], h1, l1);
            __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
            hi[e] = *reinterpret_cast<uint32_t*>(&hh);
            lo[e] = *reinterpret_cast<uint32_t*>(&ll);
          }
          reinterpret_cast<uint4*>(dhi)[g] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          reinterpret_cast<uint4*>(dlo)[g] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nb0 + j < p.Nb) {
            __half h0, l0;
            split_f16(val[j], h0, l0);
            dhi[j] = h0; dlo[j] = l0;
          }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(o.max_bits, __float_as_uint(mx));
    return;
  }
  if (p.out_s16 != nullptr && m_ok) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (nb0 + j < p.Nb) {
        float s = p.sample ? (float)((sbits >> j) & 1u) : val[j];
        p.out_s16[(size_t)(nb0 + j) * p.lds16 + m] = __float2half_rn(s);
      }
  }
  // ---- [Ma, Nb] "T" outputs: each thread owns 32 consecutive columns of its row ----
  const bool full_chunk = nb0 + 32 <= p.Nb;
  if (p.out_sT16 != nullptr && m_ok) {
    __half* dst = p.out_sT16 + (size_t)m * p.ldsT + tc_interleaved_col(nb0, p.sT_phase);
    if (full_chunk) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        uint32_t b = sbits >> (8 * g);
        v.x = pack_h2((float)(b & 1u), (float)((b >> 1) & 1u));
        v.y = pack_h2((float)((b >> 2) & 1u), (float)((b >> 3) & 1u));
        v.z = pack_h2((float)((b >> 4) & 1u), (float)((b >> 5) & 1u));
        v.w = pack_h2((float)((b >> 6) & 1u), (float)((b >> 7) & 1u));
        reinterpret_cast<uint4*>(dst)[g] = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nb0 + j < p.Nb) dst[j] = __float2half_rn((float)((sbits >> j) & 1u));
    }
  }
  if (p.out_pT_hi != nullptr && m_ok) {
    __half* dhi = p.out_pT_hi + (size_t)m * p.ldpT + tc_interleaved_col(nb0, p.pT_phase);
    __half* dlo = p.out_pT_lo + (size_t)m * p.ldpT + tc_interleaved_col(nb0, p.pT_phase);
    if (full_chunk) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __half h0, l0, h1, l1;
          split_f16(val[8 * g + 2 * e] * inv_pT, h0, l0);
          split_f16(val[8 * g + 2 * e + 1] * inv_pT, h1, l1);
          __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
          hi[e] = *reinterpret_cast<uint32_t*>(&hh);
          lo[e] = *reinterpret_cast<uint32_t*>(&ll);
        }
        reinterpret_cast<uint4*>(dhi)[g] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        reinterpret_cast<uint4*>(dlo)[g] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nb0 + j < p.Nb) {
          __half h0, l0;
          split_f16(val[j] * inv_pT, h0, l0);
          dhi[j] = h0; dlo[j] = l0;
        }
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_b_half,
               const TcGemmParams p) {
#if defined(DB_SM100)
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;         // accumulators complete -> epilogue
  uint64_t* tempty_bar = tfull_bar + 1;            // accumulators drained -> MMA
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool two_pass = p.a_lo != nullptr;

  const int tiles_m = (p.Ma + BLOCK_M - 1) / BLOCK_M;
  const int tile_n = p.tile_n;                       // 256, or 128 / 64 when wider tiles would leave most SMs without work
  const int tiles_n = (p.Nb + tile_n - 1) / tile_n;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  // work item w = (k-split ks, tile t): all tiles of one K range run before the next range starts
  const int k_splits = p.k_splits > 1 ? p.k_splits : 1;
  const int kb_per_split = (num_kb + k_splits - 1) / k_splits;
  // Tail balancing: when the last wave would occupy at most half of the CTAs, its tiles are cut into two
  // 128-column halves (UMMA N = 128), so the wave costs about half a tile time instead of a full one.
  const int tail_tiles = (k_splits == 1 && p.split_tail) ? num_tiles % (int)gridDim.x : 0;
  const int num_full = num_tiles - tail_tiles;
  const int num_work = k_splits > 1 ? num_tiles * k_splits : num_full + 2 * tail_tiles;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tm_a_hi);
    if (two_pass) ptx::tma_prefetch_desc(&tm_a_lo);
    ptx::tma_prefetch_desc(&tm_b);
    if (tail_tiles) ptx::tma_prefetch_desc(&tm_b_half);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(tfull_bar, 1); ptx::mbar_init(tempty_bar, EPI_THREADS);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  // Programmatic dependent launch: everything above (descriptor prefetch, barrier init, TMEM allocation) overlaps the tail
  // of the previous kernel of the chain; nothing below touches global memory before that kernel has completed. The next
  // kernel may start its own prologue right away (it lands on idle SMs when this launch is a small one).
  ptx::grid_launch_dependents();
  ptx::grid_dependency_wait();

  if (warp < EPI_WARP0) {
    ptx::setmaxnreg_dec<REGS_CTRL>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t a_bytes = two_pass ? 2 * A_TILE_BYTES : A_TILE_BYTES;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
          const WorkItem wi = decode_work(w, num_tiles, num_full, k_splits, tile_n);
          const int ks = wi.ks;
          const TileCoord tc = tile_coord(wi.tile, tiles_m, tiles_n, p.group_m);
          const int m0 = tc.m_blk * BLOCK_M, n0 = tc.n_blk * tile_n + wi.n_off;
          const uint32_t tx_bytes = a_bytes + (uint32_t)wi.bn * BLOCK_K * 2;
          const int kb_end = min(num_kb, (ks + 1) * kb_per_split);
          for (int kb = ks * kb_per_split; kb < kb_end; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            if (p.l2_hints) {      // the row-block group of A stays in L2 while the column blocks of B stream past it
              ptx::tma_load_2d_hint(st, &tm_a_hi, &full_bar[stage], kb * BLOCK_K, m0, ptx::kL2EvictLast);
              if (two_pass) ptx::tma_load_2d_hint(st + A_TILE_BYTES, &tm_a_lo, &full_bar[stage], kb * BLOCK_K, m0, ptx::kL2EvictLast);
              ptx::tma_load_2d_hint(st + 2 * A_TILE_BYTES, wi.bn == tile_n ? &tm_b : &tm_b_half, &full_bar[stage], kb * BLOCK_K, n0,
                                    ptx::kL2EvictFirst);
            } else {
              ptx::tma_load_2d(st, &tm_a_hi, &full_bar[stage], kb * BLOCK_K, m0);
              if (two_pass) ptx::tma_load_2d(st + A_TILE_BYTES, &tm_a_lo, &full_bar[stage], kb * BLOCK_K, m0);
              ptx::tma_load_2d(st + 2 * A_TILE_BYTES, wi.bn == tile_n ? &tm_b : &tm_b_half, &full_bar[stage], kb * BLOCK_K, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        uint32_t acc_phase = 0;
        const uint32_t tmem_h = tmem_base, tmem_l = tmem_base + (uint32_t)BLOCK_N;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
          const WorkItem wi = decode_work(w, num_tiles, num_full, k_splits, tile_n);
          const int ks = wi.ks;
          const uint32_t idesc = ptx::umma_idesc(/*F16*/ 0, BLOCK_M, (uint32_t)wi.bn);
          const int kb_begin = ks * kb_per_split, kb_end = min(num_kb, (ks + 1) * kb_per_split);
          ptx::mbar_wait(tempty_bar, acc_phase ^ 1);          // the epilogue has drained both accumulators
          ptx::tc_fence_after();
          for (int kb = kb_begin; kb < kb_end; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);          // TMA bytes have landed
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * STAGE_BYTES);
            const uint64_t d_ahi = ptx::umma_desc_k_sw128(sa);
            const uint64_t d_alo = ptx::umma_desc_k_sw128(sa + A_TILE_BYTES);
            const uint64_t d_b = ptx::umma_desc_k_sw128(sa + 2 * A_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);   // advance inside the swizzle row
              const uint32_t accumulate = (kb != kb_begin || k != 0) ? 1u : 0u;
              ptx::mma_f16_ss(tmem_h, d_ahi + koff, d_b + koff, idesc, accumulate);
              if (two_pass) ptx::mma_f16_ss(tmem_l, d_alo + koff, d_b + koff, idesc, accumulate);
            }
            ptx::mma_commit(&empty_bar[stage]);               // frees the smem slot when the MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit(tfull_bar);                         // accumulators complete -> epilogue
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    ptx::setmaxnreg_inc<REGS_EPI>();
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may read
    const int half = (warp - EPI_WARP0) >> 2;       // column half of the work item this warp owns
    uint32_t acc_phase = 0;
    const RngArgs rng = rng_resolved(p.rng);        // draw id += device clock when the step runs from a CUDA graph
    float accv[4][32];                              // 128 sums: row m, columns half * bn/2 + [0, bn/2)
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem wi = decode_work(w, num_tiles, num_full, k_splits, tile_n);
      const int ks = wi.ks;
      float* const out32 = p.out32 ? p.out32 + (size_t)ks * p.split_stride : nullptr;
      const TileCoord tc = tile_coord(wi.tile, tiles_m, tiles_n, p.group_m);
      const int m = tc.m_blk * BLOCK_M + quarter * 32 + lane;
      const bool m_ok = m < p.Ma;
      const int half_cols = wi.bn >> 1;             // 128, or 64 for a tail half-tile
      const int n_slices = half_cols >> 5;          // 4 or 2
      const int n_half0 = tc.n_blk * tile_n + wi.n_off + half * half_cols;
      const float bias_m = (p.bias != nullptr && m_ok) ? __ldg(p.bias + m) : 0.0f;
      const float acc_scale = p.acc_scale_dev ? p.acc_scale * __ldg(p.acc_scale_dev) : p.acc_scale;

      ptx::mbar_wait(tfull_bar, acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * half_cols);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s < n_slices) {                         // warp-uniform
          uint32_t rh[32], rl[32];
          ptx::tmem_ld_32x32(taddr0 + (uint32_t)(s * 32), rh);
          if (two_pass) ptx::tmem_ld_32x32(taddr0 + (uint32_t)(BLOCK_N + s * 32), rl);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) accv[s][j] = two_pass ? __uint_as_float(rh[j]) + __uint_as_float(rl[j]) : __uint_as_float(rh[j]);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(tempty_bar);                 // 256 arrivals release the accumulators to the next tile's MMAs
      acc_phase ^= 1;

#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int nb0 = n_half0 + s * 32;
        if (s < n_slices && nb0 < p.Nb) {           // warp-uniform
          float* dst32 = out32;
          if (p.peer_world > 0) {                   // this slice belongs to rank `owner`: store into its slot for this rank
            const int owner = nb0 / p.peer_rows;
            dst32 = p.peer_out[owner] + ((ptrdiff_t)p.peer_rank - (ptrdiff_t)owner) * (ptrdiff_t)p.peer_rows * (ptrdiff_t)p.ld32;
          }
          epilogue_slice(p, accv[s], nb0, m, m_ok, bias_m, acc_scale, dst32, rng.draw_lo, rng.draw_hi);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

// fp16 [rows, cols] row-major with leading dimension ld (elements); box = [box_rows, 64], 128B swizzle
bool make_tmap_f16(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  auto enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

bool tc_gemm_supported(int Ma, int Nb, int K, int lda, int ldb) {
  // TMA: 16-byte aligned row pitch; everything else (ragged Ma/Nb/K) is handled by OOB zero fill + masks
  return Ma > 0 && Nb > 0 && K > 0 && (lda % 8 == 0) && (ldb % 8 == 0) && lda >= K && ldb >= K;
}

int tc_gemm_launch(const TcGemmParams& p, int num_sms, cudaStream_t stream) {
  if (!tc_gemm_supported(p.Ma, p.Nb, p.K, p.lda, p.ldb)) return -1000;
  if ((reinterpret_cast<uintptr_t>(p.a_hi) | reinterpret_cast<uintptr_t>(p.b) |
       reinterpret_cast<uintptr_t>(p.a_lo)) & 15) return -1000;
  if (p.sample && p.U == nullptr && (p.rng.row_offset & 3u)) return -1000;
  if (p.k_splits > 1) {   // partial products are only meaningful for the plain linear fp32 output
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    if (p.act || p.sample || p.bias || p.out_s16 || p.out_sT16 || p.out_pT_hi || !p.out32 || p.k_splits > num_kb) return -1000;
    const int per = (num_kb + p.k_splits - 1) / p.k_splits;
    if ((p.k_splits - 1) * per >= num_kb) return -1000;     // an empty chunk would publish an unwritten accumulator
  }
  if (p.peer_world > 0) {
    if (p.peer_world > 8 || p.peer_rank < 0 || p.peer_rank >= p.peer_world || p.k_splits > 1 || p.opt.kind >= 0 || p.act || p.sample ||
        p.peer_rows <= 0 || (p.peer_rows % BLOCK_N) || p.peer_rows * p.peer_world != p.Nb) return -1000;
    for (int r = 0; r < p.peer_world; ++r) if (!p.peer_out[r]) return -1000;
  }
  if (p.opt.kind >= 0) {
    if (p.k_splits > 1 || p.act || p.sample || !p.opt.theta || !p.opt.n_hi || !p.opt.n_lo || !p.opt.t_hi || !p.opt.t_lo ||
        !p.opt.scale_dev || !p.opt.max_bits || (p.opt.ldt % 8) || p.opt.ldn < p.Ma ||
        ((reinterpret_cast<uintptr_t>(p.opt.t_hi) | reinterpret_cast<uintptr_t>(p.opt.t_lo)) & 15) ||
        (p.opt.kind != DB_OPT_SGD && !p.opt.st) || (p.opt.kind == DB_OPT_ADAM && !p.opt.st2)) return -1000;
  }
  if (p.out_sT16 && ((p.ldsT % 8) || (reinterpret_cast<uintptr_t>(p.out_sT16) & 15))) return -1000;
  if (p.out_pT_hi && ((p.ldpT % 8) || (reinterpret_cast<uintptr_t>(p.out_pT_hi) & 15) ||
                      (reinterpret_cast<uintptr_t>(p.out_pT_lo) & 15))) return -1000;

  CUtensorMap tm_a_hi, tm_a_lo, tm_b, tm_b_half;
  if (!make_tmap_f16(&tm_a_hi, p.a_hi, p.Ma, p.K, p.lda, BLOCK_M)) return -1001;
  if (!make_tmap_f16(&tm_a_lo, p.a_lo ? p.a_lo : p.a_hi, p.Ma, p.K, p.lda, BLOCK_M)) return -1001;
  // Tile width: 128 x 256 tiles unless they would leave most of the SMs idle (the reference's own layer sizes: 328 rows x
  // 500 units is 8 such tiles) — then 128 or 64 batch rows per tile, so more SMs share the work and each epilogue is shorter.
  const int tiles_m_host = (p.Ma + BLOCK_M - 1) / BLOCK_M;
  int tile_n = BLOCK_N;
  while (tile_n > 64 && tiles_m_host * ((p.Nb + tile_n - 1) / tile_n) * (p.k_splits > 1 ? p.k_splits : 1) < num_sms / 2) tile_n >>= 1;
  if (!make_tmap_f16(&tm_b, p.b, p.Nb, p.K, p.ldb, tile_n)) return -1001;
  if (!make_tmap_f16(&tm_b_half, p.b, p.Nb, p.K, p.ldb, tile_n / 2)) return -1001;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return (int)attr_err;

  const int base_tiles = tiles_m_host * ((p.Nb + tile_n - 1) / tile_n);
  const int tiles = base_tiles * (p.k_splits > 1 ? p.k_splits : 1);
  const int grid = tiles < num_sms ? tiles : num_sms;
  TcGemmParams q = p;
  const int rem = base_tiles % grid;
  {
    static const int env_gm = [] { const char* e = getenv("DB_TC_GROUP_M"); return e ? atoi(e) : 0; }();   // diagnostics
    q.group_m = env_gm > 0 ? env_gm : GROUP_M_DEFAULT;
  }
  q.tile_n = tile_n;
  {
    static const int env_hints = [] { const char* e = getenv("DB_TC_L2_HINTS"); return e ? atoi(e) : 0; }();   // diagnostics
    q.l2_hints = env_hints;
  }
  q.split_tail = (tile_n >= 128 && p.k_splits <= 1 && base_tiles > grid && rem > 0 && 2 * rem <= grid) ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = getenv("DB_TC_PDL"); return (e && e[0] == '0') ? 0 : 1; }();   // diagnostics
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, tc_gemm_kernel, tm_a_hi, tm_a_lo, tm_b, tm_b_half, q);
}

}  // namespace db
