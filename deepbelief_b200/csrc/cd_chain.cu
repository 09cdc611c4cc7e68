// Fused CD-k driver for the binary-binary RBM on the tcgen05 path (rbm.py:628-646 chain, rbm.py:664
// cd_loss and the gradient TF autodiff derives from it — closed form in SURVEY.md §8 R9).
//
// Data layout in HBM (per rank; N = local batch rows, NP = N rounded up to 64):
//   weight cache : W  as fp16 hi/lo [V,H]  and  W^T as fp16 hi/lo [H,V], pre-scaled by 2^e (e chosen
//                  from max|W| on the device), refreshed once per optimizer step, read 2k+1 times.
//   V0b, Vb      : [N,V] fp16 binary states (data / current chain visibles)  -> B operand of "up"
//   Hb           : [N,H] fp16 binary chain hiddens (or the caller's PCD state) -> B operand of "down"
//   VT           : [V, 2*NP] fp16, data and model visibles, 64-column interleaved -> B operand of dW
//   PT_hi, PT_lo : [H, 2*NP] fp16, +p0 and -pk scaled by 2^14, interleaved alike   -> A operand of dW
// Every GEMM is  D[units, batch] = REAL_split[units, K] * BINARY[batch, K]^T  (gemm_tc.cu), so the
// chain is 2k+1 launches + 1 dW launch + 3 small memory-bound kernels.
#include "context.h"
#include "gemm_tc.cuh"

namespace db {
namespace {

constexpr float kProbScale = 16384.0f;   // 2^14: p in (0,1) -> fp16 normal range, hi+lo >= 22 bits

struct WeightCacheHeader {
  float scale;       // 2^e
  float inv_scale;   // 2^-e
  unsigned int max_bits;       // bit pattern of max |W| the scale was chosen from
  unsigned int new_max_bits;   // max |W_new| collected by the fused dW epilogue (db_cd_tc_step_update)
  unsigned int resplit;        // set by rescale_after_update_kernel: the scale moved, the cache must be re-split
  unsigned int pad[59];
};
static_assert(sizeof(WeightCacheHeader) == 256, "header is one 256-byte slot");

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
// Row pitch (elements) of the INTERNAL fp16 matrices: TMA needs 16-byte row pitches, so widths that are not multiples of 8
// (the reference's own 500-unit layer) are padded; the padding is never read (the tensor maps carry the true width and
// zero-fill beyond it) and the fp32 tensors the caller sees keep their natural layout.
inline int pad8(int x) { return (x + 7) & ~7; }

struct WeightCacheView {
  WeightCacheHeader* hdr;
  __half *w_hi, *w_lo;     // [V,H]
  __half *wt_hi, *wt_lo;   // [H,V]
};
inline WeightCacheView weight_cache_view(void* cache, int V, int H) {
  WeightCacheView v;
  uint8_t* p = static_cast<uint8_t*>(cache);
  const size_t mat = align256((size_t)V * pad8(H) * sizeof(__half)), mat_t = align256((size_t)H * pad8(V) * sizeof(__half));
  v.hdr = reinterpret_cast<WeightCacheHeader*>(p); p += 256;
  v.w_hi = reinterpret_cast<__half*>(p); p += mat;       // [V, pad8(H)]
  v.w_lo = reinterpret_cast<__half*>(p); p += mat;
  v.wt_hi = reinterpret_cast<__half*>(p); p += mat_t;    // [H, pad8(V)]
  v.wt_lo = reinterpret_cast<__half*>(p);
  return v;
}

// The dW GEMM runs over its whole K = 2 NP range in one work item per tile (gemm_tc.cu promotes the TMEM accumulator into
// fp32 registers every 512 columns, so depth costs no accuracy and no partial tiles go through HBM). Only when the
// [H, V] result has too few 128 x 256 tiles to occupy the GPU is the K range cut into up to 8 independent ranges of
// >= 4096 columns whose partial products dw_reduce_kernel sums in a fixed order.
constexpr int kDwChunk = 4096;
constexpr int kDwMaxSplits = 8;
constexpr int kDwEnoughTiles = 96;
constexpr int kFuseMinK = 16384;
inline int dw_splits(int ldT, int V, int H) {
  const int tiles = ((H + 127) / 128) * ((V + 255) / 256);
  if (tiles >= kDwEnoughTiles) return 1;
  int s = (ldT + kDwChunk - 1) / kDwChunk;
  return s < 1 ? 1 : (s > kDwMaxSplits ? kDwMaxSplits : s);
}

struct ChainWorkspace {
  __half *v0b, *vb, *hb, *vt, *pt_hi, *pt_lo;
  float* dw_part;      // [splits, V*H] partial dW products (splits > 1 only)
  int NP, splits;
  size_t bytes;
};
inline ChainWorkspace chain_workspace(void* ws, int N, int V, int H) {
  ChainWorkspace w;
  w.NP = (N + 63) / 64 * 64;
  uint8_t* p = static_cast<uint8_t*>(ws);
  uint8_t* p0 = p;
  w.v0b = reinterpret_cast<__half*>(p); p += align256((size_t)N * pad8(V) * 2);
  w.vb = reinterpret_cast<__half*>(p); p += align256((size_t)N * pad8(V) * 2);
  w.hb = reinterpret_cast<__half*>(p); p += align256((size_t)N * pad8(H) * 2);
  w.vt = reinterpret_cast<__half*>(p); p += align256((size_t)V * 2 * w.NP * 2);
  w.pt_hi = reinterpret_cast<__half*>(p); p += align256((size_t)H * 2 * w.NP * 2);
  w.pt_lo = reinterpret_cast<__half*>(p); p += align256((size_t)H * 2 * w.NP * 2);
  w.splits = dw_splits(2 * w.NP, V, H);
  w.dw_part = reinterpret_cast<float*>(p);
  if (w.splits > 1) p += align256((size_t)w.splits * V * H * sizeof(float));
  w.bytes = (size_t)(p - p0);
  return w;
}

// ---- max |W| and the power-of-two scale ----------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ out_bits) {
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order as uints
}
__device__ __forceinline__ int scale_exponent(float m) {
  int e = 0;
  if (m > 0.0f && isfinite(m)) {
    int q; frexpf(m, &q);           // m = f * 2^q, f in [0.5, 1)  ->  m * 2^(14-q) in [2^13, 2^14)
    e = 14 - q;
    e = max(-100, min(100, e));
  }
  return e;
}
__global__ void choose_scale_kernel(WeightCacheHeader* hdr) {
  const int e = scale_exponent(__uint_as_float(hdr->max_bits));
  hdr->scale = ldexpf(1.0f, e);
  hdr->inv_scale = ldexpf(1.0f, -e);
  hdr->new_max_bits = 0u;
  hdr->resplit = 0u;
}
// After a fused update: the epilogue wrote the cache with the scale in the header and collected max |W_new|. The scale
// rule is that of db_cd_tc_refresh_weights (a function of max |W| alone), so whenever the new maximum keeps the
// exponent the cache is already what a refresh would produce; when it crosses a power of two the cache is re-split
// by the guarded cvt_transpose launch that follows (values up to 4x over the old maximum are still finite fp16, and
// anything beyond is overwritten by that re-split).
__global__ void rescale_after_update_kernel(WeightCacheHeader* hdr) {
  const float m = __uint_as_float(hdr->new_max_bits);
  const float sc = ldexpf(1.0f, scale_exponent(m));
  hdr->resplit = (sc != hdr->scale) ? 1u : 0u;
  hdr->scale = sc;
  hdr->inv_scale = 1.0f / sc;
  hdr->max_bits = hdr->new_max_bits;
  hdr->new_max_bits = 0u;
}

// ---- fp32 [R,C] -> fp16 row-major copy and/or fp16 transposed copy, optional hi/lo split ----------
// t_interleave_phase >= 0: the transposed copy is written at tc_interleaved_col(r, phase) (CD "T" layout)
__device__ __forceinline__ float load_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_as_float(const uint8_t* p) { return (float)__ldg(p); }
template <bool SPLIT, typename SrcT = float>
__global__ void __launch_bounds__(256) cvt_transpose_kernel(const SrcT* __restrict__ src, int R, int C,
                                                            const float* __restrict__ scale_dev, __half* __restrict__ d_hi,
                                                            __half* __restrict__ d_lo, int ldd, __half* __restrict__ t_hi,
                                                            __half* __restrict__ t_lo, int ldt, int t_interleave_phase,
                                                            const unsigned int* __restrict__ guard) {
  if (guard != nullptr && __ldg(guard) == 0u) return;      // conditional re-split (rescale_after_update_kernel)
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float s = scale_dev ? __ldg(scale_dev) : 1.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = (r < R && c < C) ? load_as_float(src + (size_t)r * C + c) * s : 0.0f;
    tile[ty + 8 * i][tx] = v;
    if (d_hi != nullptr && r < R && c < C) {
      if (SPLIT) { __half hi, lo; split_f16(v, hi, lo); d_hi[(size_t)r * ldd + c] = hi; d_lo[(size_t)r * ldd + c] = lo; }
      else d_hi[(size_t)r * ldd + c] = __float2half_rn(v);
    }
  }
  __syncthreads();
  if (t_hi != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, r = r0 + tx;
      if (c < C && r < R) {
        const float v = tile[tx][ty + 8 * i];
        const int col = t_interleave_phase >= 0 ? tc_interleaved_col(r, t_interleave_phase) : r;
        if (SPLIT) { __half hi, lo; split_f16(v, hi, lo); t_hi[(size_t)c * ldt + col] = hi; t_lo[(size_t)c * ldt + col] = lo; }
        else t_hi[(size_t)c * ldt + col] = __float2half_rn(v);
      }
    }
  }
}

// ---- optimizer step on W + split-weight cache rewrite in ONE pass (the data-parallel step, after the gradient all-reduce) ----
// 64 x 64 tile of W [V, H] per CTA: g, W and the optimizer slots are read once (16-byte accesses along H), the rule is applied, W / slots
// / the [V, H] fp16 split are written in place, and the [H, V] split goes through a shared-memory transpose. The cache is
// written with the scale in the header and max |W_new| is collected, exactly as the fused dW epilogue does (gemm_tc.cu);
// rescale_after_update_kernel + the guarded re-split then handle a moved scale. The last CTAs update b | c.
// Adam step size with the step number taken from the device clock (CUDA-graph replay): lr sqrt(1 - b2^t) / (1 - b1^t), t = step + clock[1]
__device__ __forceinline__ float adam_lr_from_clock(float lr, float beta1, float beta2, int step, const unsigned long long* clock) {
  const double t = (double)step + (double)__ldg(clock + 1);
  return (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
}
__global__ void adam_lr_kernel(float lr, float beta1, float beta2, int step, const unsigned long long* clock, float* out) {
  *out = adam_lr_from_clock(lr, beta1, beta2, step, clock);
}
struct ApplyArgs {
  const float* lr_t_dev;      // != nullptr: Adam's bias-corrected step size, formed on the device by adam_lr_kernel (graph replay)
  OptRule rule;
  float* theta; const float* grad; float* st; float* st2;     // packed W | b | c and slots
  int V, H, ldw, ldwt;        // ldw / ldwt: row pitches of the [V, H] and [H, V] fp16 cache matrices
  int vec_ok;                 // 16-byte accesses allowed (H % 4 == 0 and theta / grad / slot pointers 16-byte aligned)
  __half *w_hi, *w_lo, *wt_hi, *wt_lo;
  const float* scale_dev; unsigned int* max_bits;
};
constexpr int kApplyTile = 64;
__global__ void __launch_bounds__(256) apply_update_kernel(ApplyArgs a) {
  __shared__ float tile[kApplyTile][kApplyTile + 1];        // scaled new weights, [v][h]
  if (a.lr_t_dev != nullptr) a.rule.lr_t = __ldg(a.lr_t_dev);
  const int tiles_h = (a.H + kApplyTile - 1) / kApplyTile, tiles_v = (a.V + kApplyTile - 1) / kApplyTile;
  if ((int)blockIdx.x >= tiles_h * tiles_v) {               // bias part: V + H elements behind W
    const size_t nW = (size_t)a.V * a.H;
    const int i = ((int)blockIdx.x - tiles_h * tiles_v) * 256 + threadIdx.x;
    if (i < a.V + a.H) {
      const size_t idx = nW + i;
      float s1 = a.rule.kind != 0 ? a.st[idx] : 0.0f, s2 = a.rule.kind == 2 ? a.st2[idx] : 0.0f;
      a.theta[idx] = opt_apply(a.rule, a.theta[idx], __ldg(a.grad + idx), s1, s2);
      if (a.rule.kind != 0) a.st[idx] = s1;
      if (a.rule.kind == 2) a.st2[idx] = s2;
    }
    return;
  }
  const int r0 = ((int)blockIdx.x / tiles_h) * kApplyTile, c0 = ((int)blockIdx.x % tiles_h) * kApplyTile;
  const float sc = __ldg(a.scale_dev);
  const bool full = a.vec_ok && r0 + kApplyTile <= a.V && c0 + kApplyTile <= a.H;
  float mx = 0.0f;
  if (full) {
    // phase 1: thread = 4 consecutive h of one row v (16-byte accesses, 256 B per 16 threads), 16 rows per pass
    const int c4 = (threadIdx.x & 15) * 4, rr = threadIdx.x >> 4;
    float4 g[4], th[4], m1[4], m2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                           // all loads first
      const size_t idx = (size_t)(r0 + rr + 16 * i) * a.H + c0 + c4;
      g[i] = __ldcs(reinterpret_cast<const float4*>(a.grad + idx));
      th[i] = *reinterpret_cast<const float4*>(a.theta + idx);
      m1[i] = a.rule.kind != 0 ? *reinterpret_cast<const float4*>(a.st + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      m2[i] = a.rule.kind == 2 ? *reinterpret_cast<const float4*>(a.st2 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rr + 16 * i;
      const size_t idx = (size_t)(r0 + r) * a.H + c0 + c4;
      float4 t;
      t.x = opt_apply(a.rule, th[i].x, g[i].x, m1[i].x, m2[i].x);
      t.y = opt_apply(a.rule, th[i].y, g[i].y, m1[i].y, m2[i].y);
      t.z = opt_apply(a.rule, th[i].z, g[i].z, m1[i].z, m2[i].z);
      t.w = opt_apply(a.rule, th[i].w, g[i].w, m1[i].w, m2[i].w);
      *reinterpret_cast<float4*>(a.theta + idx) = t;
      if (a.rule.kind != 0) *reinterpret_cast<float4*>(a.st + idx) = m1[i];
      if (a.rule.kind == 2) *reinterpret_cast<float4*>(a.st2 + idx) = m2[i];
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
      const float v0 = t.x * sc, v1 = t.y * sc, v2 = t.z * sc, v3 = t.w * sc;
      tile[r][c4] = v0; tile[r][c4 + 1] = v1; tile[r][c4 + 2] = v2; tile[r][c4 + 3] = v3;
      __half h0, l0, h1, l1, h2, l2, h3, l3;
      split_f16(v0, h0, l0); split_f16(v1, h1, l1); split_f16(v2, h2, l2); split_f16(v3, h3, l3);
      __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3), la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
      const size_t widx = (size_t)(r0 + r) * a.ldw + c0 + c4;
      *reinterpret_cast<uint2*>(a.w_hi + widx) = make_uint2(*reinterpret_cast<uint32_t*>(&ha), *reinterpret_cast<uint32_t*>(&hb));
      *reinterpret_cast<uint2*>(a.w_lo + widx) = make_uint2(*reinterpret_cast<uint32_t*>(&la), *reinterpret_cast<uint32_t*>(&lb));
    }
    __syncthreads();
    // phase 2: thread = 16 consecutive v of one column h -> 32 contiguous bytes of W^T hi and of W^T lo
    const int c = threadIdx.x >> 2, rq = (threadIdx.x & 3) * 16;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      __half h0, l0, h1, l1;
      split_f16(tile[rq + 2 * e][c], h0, l0);
      split_f16(tile[rq + 2 * e + 1][c], h1, l1);
      __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
      hi[e] = *reinterpret_cast<uint32_t*>(&hh); lo[e] = *reinterpret_cast<uint32_t*>(&ll);
    }
    const size_t tidx = (size_t)(c0 + c) * a.ldwt + r0 + rq;
    reinterpret_cast<uint4*>(a.wt_hi + tidx)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    reinterpret_cast<uint4*>(a.wt_hi + tidx)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    reinterpret_cast<uint4*>(a.wt_lo + tidx)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    reinterpret_cast<uint4*>(a.wt_lo + tidx)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  } else {
    // ragged edge tiles: one element at a time
    for (int e = threadIdx.x; e < kApplyTile * kApplyTile; e += 256) {
      const int r = e / kApplyTile, c = e % kApplyTile;
      float v = 0.0f;
      if (r0 + r < a.V && c0 + c < a.H) {
        const size_t idx = (size_t)(r0 + r) * a.H + c0 + c;
        float s1 = a.rule.kind != 0 ? a.st[idx] : 0.0f, s2 = a.rule.kind == 2 ? a.st2[idx] : 0.0f;
        const float t = opt_apply(a.rule, a.theta[idx], __ldg(a.grad + idx), s1, s2);
        if (a.rule.kind != 0) a.st[idx] = s1;
        if (a.rule.kind == 2) a.st2[idx] = s2;
        a.theta[idx] = t;
        mx = fmaxf(mx, fabsf(t));
        v = t * sc;
        __half hi, lo;
        split_f16(v, hi, lo);
        a.w_hi[(size_t)(r0 + r) * a.ldw + c0 + c] = hi; a.w_lo[(size_t)(r0 + r) * a.ldw + c0 + c] = lo;
      }
      tile[r][c] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kApplyTile * kApplyTile; e += 256) {
      const int c = e / kApplyTile, r = e % kApplyTile;
      if (r0 + r < a.V && c0 + c < a.H) {
        __half hi, lo;
        split_f16(tile[r][c], hi, lo);
        a.wt_hi[(size_t)(c0 + c) * a.ldwt + r0 + r] = hi; a.wt_lo[(size_t)(c0 + c) * a.ldwt + r0 + r] = lo;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(a.max_bits, __float_as_uint(mx));
}

// ---- data-parallel update on this rank's rows of W, gradient reduce-scattered by the dW epilogues of all ranks ----------------
// slots [world][rows_per * H]: slot s = rank s's partial gradient for MY rows (written by its dW epilogue over NVLink,
// gemm_tc.cu peer_out). Summed in rank order (deterministic), the rule applied with this rank's slice of the optimizer
// slots, and the new rows stored into EVERY rank's parameter vector (peer-mapped pointers; the all-gather by push). b | c
// are updated identically on every rank from the all-reduced bias gradients in `grad`.
struct ShardArgs {
  OptRule rule; const float* lr_t_dev;
  int V, H, world, rank, rows_per;
  const float* slots; float* grad; float* st; float* st2;
  float* peer_params[8];
};
__global__ void __launch_bounds__(256) apply_shard_kernel(ShardArgs a) {
  if (a.lr_t_dev != nullptr) a.rule.lr_t = __ldg(a.lr_t_dev);
  const size_t shard = (size_t)a.rows_per * a.H, n4 = shard / 4, base = (size_t)a.rank * shard;
  const size_t nW = (size_t)a.V * a.H;
  float* theta = a.peer_params[a.rank];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 g = __ldcs(reinterpret_cast<const float4*>(a.slots) + i);
    for (int s = 1; s < a.world; ++s) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(a.slots + (size_t)s * shard) + i);
      g.x += b.x; g.y += b.y; g.z += b.z; g.w += b.w;
    }
    const size_t idx = base + 4 * i;
    float4 th = *reinterpret_cast<const float4*>(theta + idx);
    float4 m1 = a.rule.kind != 0 ? *reinterpret_cast<const float4*>(a.st + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 m2 = a.rule.kind == 2 ? *reinterpret_cast<const float4*>(a.st2 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 t;
    t.x = opt_apply(a.rule, th.x, g.x, m1.x, m2.x);
    t.y = opt_apply(a.rule, th.y, g.y, m1.y, m2.y);
    t.z = opt_apply(a.rule, th.z, g.z, m1.z, m2.z);
    t.w = opt_apply(a.rule, th.w, g.w, m1.w, m2.w);
    if (a.rule.kind != 0) *reinterpret_cast<float4*>(a.st + idx) = m1;
    if (a.rule.kind == 2) *reinterpret_cast<float4*>(a.st2 + idx) = m2;
    *reinterpret_cast<float4*>(a.grad + idx) = g;
    for (int r = 0; r < a.world; ++r) *reinterpret_cast<float4*>(a.peer_params[r] + idx) = t;
  }
  // biases: V + H elements behind W, the same on every rank
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)(a.V + a.H); i += (size_t)gridDim.x * blockDim.x) {
    const size_t idx = nW + i;
    float s1 = a.rule.kind != 0 ? a.st[idx] : 0.0f, s2 = a.rule.kind == 2 ? a.st2[idx] : 0.0f;
    theta[idx] = opt_apply(a.rule, theta[idx], __ldg(a.grad + idx), s1, s2);
    if (a.rule.kind != 0) a.st[idx] = s1;
    if (a.rule.kind == 2) a.st2[idx] = s2;
  }
}

// ---- bias gradients from the interleaved T buffers: one warp per unit ------------------------------
// gb[v] = -inv_n * (sum_n v0[n,v] - sum_n vk[n,v])          (sign from the phase of each 64-block)
// gc[h] = -inv_n * 2^-14 * sum_cols (PT_hi + PT_lo)[h,:]     (phase-1 entries are stored negated)
// upd.kind >= 0: the same thread then applies the optimizer rule to its bias (b | c are contiguous behind W in the packed
// parameter vector, so unit indexes both; same expressions as optimizer_kernel, layer_simt.cu)
struct BiasUpdate {
  int kind;                    // -1: gradients only
  float lr, lr_t, momentum, beta1, beta2, eps, weight_decay;
  float* theta; float* st; float* st2;      // each already offset to the first bias
};
__device__ __forceinline__ void bias_update(const BiasUpdate& u, int unit, float g) {
  if (u.kind < 0) return;
  const OptRule rule{u.kind, u.lr, u.lr_t, u.momentum, u.beta1, u.beta2, u.eps, u.weight_decay};
  float s1 = u.kind != DB_OPT_SGD ? u.st[unit] : 0.0f, s2 = u.kind == DB_OPT_ADAM ? u.st2[unit] : 0.0f;
  u.theta[unit] = opt_apply(rule, u.theta[unit], g, s1, s2);
  if (u.kind != DB_OPT_SGD) u.st[unit] = s1;
  if (u.kind == DB_OPT_ADAM) u.st2[unit] = s2;
}
__global__ void tc_bias_grads_kernel(const __half* __restrict__ vt, const __half* __restrict__ pt_hi,
                                     const __half* __restrict__ pt_lo, int V, int H, int ld, float inv_n,
                                     float* __restrict__ gb, float* __restrict__ gc, const BiasUpdate upd) {
  const int unit = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (unit >= V + H) return;
  float s = 0.0f;
  if (unit < V) {
    const __half* row = vt + (size_t)unit * ld;
    for (int c = lane * 8; c < ld; c += 256) {
      const uint4 q = *reinterpret_cast<const uint4*>(row + c);
      const __half2* h2 = reinterpret_cast<const __half2*>(&q);
      float t = 0.0f;
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h2[e]); t += f.x + f.y; }
      s += ((c >> 6) & 1) ? -t : t;
    }
    s = warp_sum(s);
    if (lane == 0) { const float g = -inv_n * s; gb[unit] = g; bias_update(upd, unit, g); }
  } else {
    const int hrow = unit - V;
    const __half* rh = pt_hi + (size_t)hrow * ld;
    const __half* rl = pt_lo + (size_t)hrow * ld;
    float shi = 0.0f, slo = 0.0f;
    for (int c = lane * 8; c < ld; c += 256) {
      const uint4 qh = *reinterpret_cast<const uint4*>(rh + c);
      const uint4 ql = *reinterpret_cast<const uint4*>(rl + c);
      const __half2* a = reinterpret_cast<const __half2*>(&qh);
      const __half2* b = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = __half22float2(a[e]), fb = __half22float2(b[e]);
        shi += fa.x + fa.y; slo += fb.x + fb.y;
      }
    }
    s = warp_sum(shi) + warp_sum(slo);
    if (lane == 0) { const float g = -inv_n * (1.0f / kProbScale) * s; gc[hrow] = g; bias_update(upd, unit, g); }
  }
}

// gW = sum over the K-chunk partials (fixed order: deterministic)
__global__ void dw_reduce_kernel(const float4* __restrict__ part, size_t n4, size_t stride4, int splits,
                                 float4* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = __ldcs(part + i);
    for (int s = 1; s < splits; ++s) {
      const float4 b = __ldcs(part + (size_t)s * stride4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
  }
}

// [rows, cols] fp16 with row pitch ld -> dense fp32
__global__ void half_to_float_kernel(const __half* __restrict__ src, size_t rows, size_t cols, size_t ld, float* __restrict__ dst) {
  const size_t n = rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    dst[i] = __half2float(src[r * ld + c]);
  }
}

inline int ew_grid(size_t n, int num_sms) {
  size_t blocks = (n + 255) / 256, cap = (size_t)num_sms * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// N a multiple of 8 (Philox draws are keyed by 4-row groups, the interleaved buffers by 64-row blocks); any V, H >= 8
inline bool tc_shape_ok(int N, int V, int H) { return N >= 8 && (N % 8 == 0) && V >= 8 && H >= 8; }


// dW[v,h] = -(1/N) (v0^T p0 - vk^T pk) for rows v in [v_begin, v_begin + v_count): one GEMM over the interleaved
// K = 2*NP axis (K-chunked, see dw_splits)
struct PeerPush { float* const* slots; int world, rank; };
int launch_dw(db_handle_t h, cudaStream_t st, const db_cd_tc_desc_t* d, const ChainWorkspace& ws, float* gW, int v_begin,
              int v_count, const TcOptArgs* fused = nullptr, const PeerPush* push = nullptr) {
  const int V = d->V, H = d->H;
  const int ldT = 2 * ws.NP;
  TcGemmParams p{};
  p.opt.kind = -1;
  if (fused) p.opt = *fused;
  if (push) {
    p.peer_world = push->world; p.peer_rank = push->rank; p.peer_rows = V / push->world;
    for (int r = 0; r < push->world && r < 8; ++r) p.peer_out[r] = push->slots[r];
  }
  p.Ma = H; p.Nb = v_count; p.K = ldT;
  p.a_hi = ws.pt_hi; p.a_lo = ws.pt_lo; p.lda = ldT; p.b = ws.vt + (size_t)v_begin * ldT; p.ldb = ldT;
  p.acc_scale = -d->inv_global_n / kProbScale; p.act = 0; p.sample = 0;
  p.ld32 = H;
  const bool split = ws.splits > 1 && ((size_t)V * H) % 4 == 0 && ((size_t)v_begin * H) % 4 == 0 && ((size_t)v_count * H) % 4 == 0;
  if (split) { p.out32 = ws.dw_part + (size_t)v_begin * H; p.k_splits = ws.splits; p.split_stride = (size_t)V * H; }
  else p.out32 = gW + (size_t)v_begin * H;
  int rc = tc_gemm_launch(p, tc_ctas(h), st);
  if (rc) return set_err(h, DB_ERR_CUDA, "cd_tc dW launch failed (%d)", rc);
  if (split) {
    const size_t n4 = (size_t)v_count * H / 4, stride4 = (size_t)V * H / 4;
    dw_reduce_kernel<<<ew_grid(n4, h->num_sms), 256, 0, st>>>(reinterpret_cast<const float4*>(ws.dw_part + (size_t)v_begin * H), n4,
                                                              stride4, ws.splits, reinterpret_cast<float4*>(gW + (size_t)v_begin * H));
  }
  return DB_OK;
}

}  // namespace
}  // namespace db

using namespace db;

extern "C" size_t db_cd_tc_workspace_bytes(const db_cd_tc_desc_t* d) {
  if (!d || !tc_shape_ok(d->N, d->V, d->H)) return 0;
  return chain_workspace(nullptr, d->N, d->V, d->H).bytes;
}

extern "C" size_t db_cd_tc_weight_cache_bytes(int V, int H) {
  return 256 + 2 * align256((size_t)V * pad8(H) * sizeof(__half)) + 2 * align256((size_t)H * pad8(V) * sizeof(__half));
}

extern "C" int db_cd_tc_refresh_weights(db_handle_t h, db_stream_t stream, const float* W, int V, int H, void* cache) {
  DB_REQUIRE(h, W && cache && V > 0 && H > 0, "null operand or empty shape");
  if (h->cc_major < 10) return set_err(h, DB_ERR_UNSUPPORTED, "tcgen05 path needs sm_100a (device is sm_%d%d)", h->cc_major, h->cc_minor);
  cudaStream_t st = (cudaStream_t)stream;
  WeightCacheView wc = weight_cache_view(cache, V, H);
  cudaMemsetAsync(&wc.hdr->max_bits, 0, sizeof(unsigned int), st);
  const size_t n = (size_t)V * H;
  absmax_kernel<<<ew_grid(n, h->num_sms), 256, 0, st>>>(W, n, &wc.hdr->max_bits);
  choose_scale_kernel<<<1, 1, 0, st>>>(wc.hdr);
  dim3 grid((H + 31) / 32, (V + 31) / 32);
  cvt_transpose_kernel<true><<<grid, 256, 0, st>>>(W, V, H, &wc.hdr->scale, wc.w_hi, wc.w_lo, pad8(H), wc.wt_hi, wc.wt_lo, pad8(V), -1, nullptr);
  return check_launch(h, "cd_tc_refresh_weights");
}

// opt != nullptr: the fused update (db_cd_tc_step_update) — params, opt_state and the weight cache are rewritten by the
// dW epilogue and the bias kernel; otherwise gradients only (db_cd_tc_step).
static int cd_tc_step_impl(db_handle_t h, db_stream_t stream, const db_cd_tc_desc_t* d, const void* v0,
                           float* params, void* weight_cache, void* pcd_state, const float* injected,
                           const db_rng_t* rng, void* workspace, float* grad, float* vk_out, float* hk_out,
                           float* p0_out, float* pk_out, const db_opt_t* opt, float* opt_state, const PeerPush* push = nullptr) {
  DB_REQUIRE(h, d && v0 && params && weight_cache && workspace && grad, "null operand");
  DB_REQUIRE(h, d->k >= 1, "k must be > 0");   // rbm.py:364
  DB_REQUIRE(h, injected || rng, "sampling needs injected draws or an rng");
  DB_REQUIRE(h, !d->pcd || pcd_state, "pcd=1 needs the persistent chain state");
  if (h->cc_major < 10) return set_err(h, DB_ERR_UNSUPPORTED, "tcgen05 path needs sm_100a (device is sm_%d%d)", h->cc_major, h->cc_minor);
  const int N = d->N, V = d->V, H = d->H, k = d->k;
  if (!tc_shape_ok(N, V, H)) return set_err(h, DB_ERR_UNSUPPORTED, "tensor path needs N, V, H multiples of 8 (got %d, %d, %d)", N, V, H);
  if (!injected && (rng->row_offset & 3u)) return set_err(h, DB_ERR_UNSUPPORTED, "row_offset must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;

  WeightCacheView wc = weight_cache_view(weight_cache, V, H);
  ChainWorkspace ws = chain_workspace(workspace, N, V, H);
  const int ldT = 2 * ws.NP;
  const float* b = params + (size_t)V * H;
  const float* c = b + V;
  float* gW = grad; float* gb = grad + (size_t)V * H; float* gc = gb + V;
  __half* hstate = d->pcd ? static_cast<__half*>(pcd_state) : ws.hb;

  RngArgs r; to_rng(rng, r);
  const uint64_t draw0 = rng ? rng->draw : 0;
  auto set_draw = [&](TcGemmParams& p, uint64_t i) {
    p.rng = r; const uint64_t dr = draw0 + i; p.rng.draw_lo = (uint32_t)dr; p.rng.draw_hi = (uint32_t)(dr >> 32);
  };
  const size_t NH = (size_t)N * H, NV = (size_t)N * V;
  const int Vp = pad8(V), Hp = pad8(H);

  if (N != ws.NP) {   // padded columns of the interleaved buffers must not contribute to dW
    cudaMemsetAsync(ws.vt, 0, (size_t)V * ldT * 2, st);
    cudaMemsetAsync(ws.pt_hi, 0, (size_t)H * ldT * 2, st);
    cudaMemsetAsync(ws.pt_lo, 0, (size_t)H * ldT * 2, st);
  }
  {  // data batch -> fp16 [N,V] and interleaved phase 0 of VT
    dim3 grid((V + 31) / 32, (N + 31) / 32);
    if (d->v0_u8)
      cvt_transpose_kernel<false, uint8_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(v0), N, V, nullptr, ws.v0b, nullptr, Vp,
                                                                 ws.vt, nullptr, ldT, 0, nullptr);
    else
      cvt_transpose_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(v0), N, V, nullptr, ws.v0b, nullptr, Vp, ws.vt,
                                                        nullptr, ldT, 0, nullptr);
  }

  auto up = [&](const __half* vin, int phase /*-1: middle of chain*/, bool write_sample, const float* U, uint64_t draw_i,
                float* p_out) -> int {
    TcGemmParams p{};
    p.opt.kind = -1;
    p.Ma = H; p.Nb = N; p.K = V;
    p.a_hi = wc.wt_hi; p.a_lo = wc.wt_lo; p.lda = Vp; p.b = vin; p.ldb = Vp;
    p.acc_scale = 1.0f; p.acc_scale_dev = &wc.hdr->inv_scale; p.bias = c; p.act = 1;
    p.sample = write_sample ? 1 : 0; p.U = U; p.ldu = H; set_draw(p, draw_i);
    p.out32 = p_out; p.ld32 = H;
    p.out_s16 = write_sample ? hstate : nullptr; p.lds16 = Hp;
    if (phase >= 0) {
      p.out_pT_hi = ws.pt_hi; p.out_pT_lo = ws.pt_lo; p.ldpT = ldT; p.pT_phase = phase;
      p.pT_scale = phase == 0 ? kProbScale : -kProbScale;
    }
    return tc_gemm_launch(p, tc_ctas(h), st);
  };
  auto down = [&](bool last, const float* U, uint64_t draw_i) -> int {
    TcGemmParams p{};
    p.opt.kind = -1;
    p.Ma = V; p.Nb = N; p.K = H;
    p.a_hi = wc.w_hi; p.a_lo = wc.w_lo; p.lda = Hp; p.b = hstate; p.ldb = Hp;
    p.acc_scale = 1.0f; p.acc_scale_dev = &wc.hdr->inv_scale; p.bias = b; p.act = 1;
    p.sample = 1; p.U = U; p.ldu = V; set_draw(p, draw_i);
    p.out_s16 = ws.vb; p.lds16 = Vp;
    if (last) { p.out_sT16 = ws.vt; p.ldsT = ldT; p.sT_phase = 1; }
    return tc_gemm_launch(p, tc_ctas(h), st);
  };

  int rc = up(ws.v0b, 0, !d->pcd, injected, 0, p0_out);
  if (rc) return set_err(h, rc == -1000 ? DB_ERR_UNSUPPORTED : DB_ERR_CUDA, "cd_tc up(v0) launch failed (%d)", rc);
  for (int s = 1; s <= k; ++s) {
    const float* Uv = injected ? injected + NH + (size_t)(s - 1) * (NV + NH) : nullptr;
    const float* Uh = injected ? Uv + NV : nullptr;
    rc = down(s == k, Uv, (uint64_t)(2 * s - 1));
    if (rc) return set_err(h, DB_ERR_CUDA, "cd_tc down launch failed (%d)", rc);
    rc = up(ws.vb, s == k ? 1 : -1, true, Uh, (uint64_t)(2 * s), s == k ? pk_out : nullptr);
    if (rc) return set_err(h, DB_ERR_CUDA, "cd_tc up launch failed (%d)", rc);
  }
  BiasUpdate bu{};
  bu.kind = -1;
  if (opt) {
    // north_star K2: dW accumulation and the optimizer update in ONE kernel. The epilogue of the dW GEMM applies the rule
    // to W and rewrites the fp16 split-weight cache (W and W^T); the bias kernel does the same for b | c.
    const size_t n_total = (size_t)V * H + V + H, nW = (size_t)V * H;
    float lr_t = opt->lr;
    if (opt->kind == DB_OPT_ADAM) {    // TF AdamOptimizer: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) [TF, not vendored]; as db_optimizer_step
      const double b1t = pow((double)opt->beta1, opt->step), b2t = pow((double)opt->beta2, opt->step);
      lr_t = (float)((double)opt->lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    }
    TcOptArgs fo{};
    fo.kind = opt->kind; fo.lr = opt->lr; fo.lr_t = lr_t; fo.momentum = opt->momentum; fo.beta1 = opt->beta1; fo.beta2 = opt->beta2;
    fo.eps = opt->eps; fo.weight_decay = opt->weight_decay;
    fo.theta = params;
    fo.st = opt->kind != DB_OPT_SGD ? opt_state : nullptr;
    fo.st2 = opt->kind == DB_OPT_ADAM ? opt_state + n_total : nullptr;
    fo.n_hi = wc.w_hi; fo.n_lo = wc.w_lo; fo.ldn = Hp; fo.t_hi = wc.wt_hi; fo.t_lo = wc.wt_lo; fo.ldt = Vp;
    fo.scale_dev = &wc.hdr->scale; fo.max_bits = &wc.hdr->new_max_bits;
    bu.kind = opt->kind; bu.lr = opt->lr; bu.lr_t = lr_t; bu.momentum = opt->momentum; bu.beta1 = opt->beta1; bu.beta2 = opt->beta2;
    bu.eps = opt->eps; bu.weight_decay = opt->weight_decay;
    bu.theta = params + nW; bu.st = fo.st ? fo.st + nW : nullptr; bu.st2 = fo.st2 ? fo.st2 + nW : nullptr;
    rc = launch_dw(h, st, d, ws, gW, 0, V, &fo);
    if (rc) return rc;
  } else {
    rc = launch_dw(h, st, d, ws, gW, 0, V, nullptr, push);
    if (rc) return rc;
  }
  tc_bias_grads_kernel<<<(V + H + 7) / 8, 256, 0, st>>>(ws.vt, ws.pt_hi, ws.pt_lo, V, H, ldT, d->inv_global_n, gb, gc, bu);
  if (opt) {
    rescale_after_update_kernel<<<1, 1, 0, st>>>(wc.hdr);
    dim3 grid((H + 31) / 32, (V + 31) / 32);
    cvt_transpose_kernel<true><<<grid, 256, 0, st>>>(params, V, H, &wc.hdr->scale, wc.w_hi, wc.w_lo, Hp, wc.wt_hi, wc.wt_lo, Vp, -1,
                                                     &wc.hdr->resplit);
  }
  if (vk_out) half_to_float_kernel<<<ew_grid(NV, h->num_sms), 256, 0, st>>>(ws.vb, N, V, Vp, vk_out);
  if (hk_out) half_to_float_kernel<<<ew_grid(NH, h->num_sms), 256, 0, st>>>(hstate, N, H, Hp, hk_out);
  return check_launch(h, "cd_tc_step");
}

extern "C" int db_cd_tc_step(db_handle_t h, db_stream_t stream, const db_cd_tc_desc_t* d, const void* v0,
                             const float* params, const void* weight_cache, void* pcd_state, const float* injected,
                             const db_rng_t* rng, void* workspace, float* grad, float* vk_out, float* hk_out,
                             float* p0_out, float* pk_out) {
  return cd_tc_step_impl(h, stream, d, v0, const_cast<float*>(params), const_cast<void*>(weight_cache), pcd_state, injected, rng,
                         workspace, grad, vk_out, hk_out, p0_out, pk_out, nullptr, nullptr);
}

extern "C" int db_cd_tc_update_fusable(const db_cd_tc_desc_t* d) {
  if (!d || !tc_shape_ok(d->N, d->V, d->H)) return 0;
  // the epilogue moves ~1.2 MB per 128 x 256 tile through HBM; it hides under the tile's MMAs only when K = 2 NP is deep
  // (measured at V = 4096, H = 8192: N = 16384 -0.7 ms per step, N = 2048 +0.75 ms against the separate one-pass update)
  const int ldT = 2 * ((d->N + 63) / 64 * 64);
  if (dw_splits(ldT, d->V, d->H) != 1) return 0;
  return ldT >= kFuseMinK ? 2 : 1;
}

extern "C" int db_cd_tc_step_update(db_handle_t h, db_stream_t stream, const db_cd_tc_desc_t* d, const void* v0,
                                    float* params, void* weight_cache, void* pcd_state, const float* injected,
                                    const db_rng_t* rng, void* workspace, float* grad, const db_opt_t* opt,
                                    float* opt_state) {
  DB_REQUIRE(h, d && opt, "null operand");
  DB_REQUIRE(h, opt->kind == DB_OPT_SGD || opt->kind == DB_OPT_MOMENTUM || opt->kind == DB_OPT_ADAM, "unknown optimizer kind");
  DB_REQUIRE(h, opt->kind == DB_OPT_SGD || opt_state, "the optimizer rule needs its slot buffer");
  if (opt->clock) return set_err(h, DB_ERR_UNSUPPORTED, "the fused update takes the step number by value (no device clock)");
  if (!db_cd_tc_update_fusable(d))
    return set_err(h, DB_ERR_UNSUPPORTED, "dW of this shape runs as K-split partial products; use db_cd_tc_step + db_optimizer_step");
  return cd_tc_step_impl(h, stream, d, v0, params, weight_cache, pcd_state, injected, rng, workspace, grad, nullptr, nullptr,
                         nullptr, nullptr, opt, opt_state);
}

extern "C" int db_cd_tc_apply_update(db_handle_t h, db_stream_t stream, int V, int H, float* params, const float* grad,
                                     const db_opt_t* opt, float* opt_state, void* weight_cache) {
  DB_REQUIRE(h, params && grad && opt && weight_cache && V > 0 && H > 0, "null operand or empty shape");
  DB_REQUIRE(h, opt->kind == DB_OPT_SGD || opt->kind == DB_OPT_MOMENTUM || opt->kind == DB_OPT_ADAM, "unknown optimizer kind");
  DB_REQUIRE(h, opt->kind == DB_OPT_SGD || opt_state, "the optimizer rule needs its slot buffer");
  cudaStream_t st = (cudaStream_t)stream;
  WeightCacheView wc = weight_cache_view(weight_cache, V, H);
  const size_t n_total = (size_t)V * H + V + H;
  float lr_t = opt->lr;
  if (opt->kind == DB_OPT_ADAM) {
    const double b1t = pow((double)opt->beta1, opt->step), b2t = pow((double)opt->beta2, opt->step);
    lr_t = (float)((double)opt->lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  }
  ApplyArgs a{};
  a.lr_t_dev = nullptr;
  if (opt->clock && opt->kind == DB_OPT_ADAM) {      // fp64 pow is slow on this part: once, not per CTA
    float* slot = reinterpret_cast<float*>(&wc.hdr->pad[0]);
    adam_lr_kernel<<<1, 1, 0, st>>>(opt->lr, opt->beta1, opt->beta2, opt->step, (const unsigned long long*)opt->clock, slot);
    a.lr_t_dev = slot;
  }
  a.rule = OptRule{opt->kind, opt->lr, lr_t, opt->momentum, opt->beta1, opt->beta2, opt->eps, opt->weight_decay};
  a.theta = params; a.grad = grad;
  a.st = opt->kind != DB_OPT_SGD ? opt_state : nullptr;
  a.st2 = opt->kind == DB_OPT_ADAM ? opt_state + n_total : nullptr;
  a.V = V; a.H = H; a.ldw = pad8(H); a.ldwt = pad8(V);
  a.vec_ok = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(a.st) |
                               reinterpret_cast<uintptr_t>(a.st2)) & 15) == 0;
  a.w_hi = wc.w_hi; a.w_lo = wc.w_lo; a.wt_hi = wc.wt_hi; a.wt_lo = wc.wt_lo;
  a.scale_dev = &wc.hdr->scale; a.max_bits = &wc.hdr->new_max_bits;
  const int tiles = ((H + kApplyTile - 1) / kApplyTile) * ((V + kApplyTile - 1) / kApplyTile), bias_blocks = (V + H + 255) / 256;
  apply_update_kernel<<<tiles + bias_blocks, 256, 0, st>>>(a);
  rescale_after_update_kernel<<<1, 1, 0, st>>>(wc.hdr);
  dim3 grid((H + 31) / 32, (V + 31) / 32);
  cvt_transpose_kernel<true><<<grid, 256, 0, st>>>(params, V, H, &wc.hdr->scale, wc.w_hi, wc.w_lo, pad8(H), wc.wt_hi, wc.wt_lo, pad8(V), -1,
                                                   &wc.hdr->resplit);
  return check_launch(h, "cd_tc_apply_update");
}

extern "C" int db_cd_tc_push_ok(const db_cd_tc_desc_t* d, int world) {
  if (!d || world < 2 || world > 8 || !tc_shape_ok(d->N, d->V, d->H)) return 0;
  if (d->V % (world * 256) != 0) return 0;                       // shard boundaries on tile boundaries
  if (d->H % 4 != 0 || ((size_t)d->V * d->H + d->V + d->H) % 4 != 0) return 0;   // 16-byte accesses of the sharded update
  return dw_splits(2 * ((d->N + 63) / 64 * 64), d->V, d->H) == 1 ? 1 : 0;
}

extern "C" int db_cd_tc_step_push(db_handle_t h, db_stream_t stream, const db_cd_tc_desc_t* d, const void* v0,
                                  const float* params, const void* weight_cache, void* pcd_state, const float* injected,
                                  const db_rng_t* rng, void* workspace, float* grad, float* const* peer_slots, int world,
                                  int rank) {
  DB_REQUIRE(h, d && peer_slots && rank >= 0 && rank < world, "null operand or bad rank");
  if (!db_cd_tc_push_ok(d, world))
    return set_err(h, DB_ERR_UNSUPPORTED, "peer push needs 2..8 ranks, V a multiple of 256 x ranks and an un-split dW product");
  PeerPush push{peer_slots, world, rank};
  return cd_tc_step_impl(h, stream, d, v0, const_cast<float*>(params), const_cast<void*>(weight_cache), pcd_state, injected, rng,
                         workspace, grad, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &push);
}

extern "C" int db_cd_tc_apply_update_sharded(db_handle_t h, db_stream_t stream, int V, int H, int world, int rank,
                                             float* const* peer_params, const float* slots, float* grad, const db_opt_t* opt,
                                             float* opt_state) {
  DB_REQUIRE(h, peer_params && slots && grad && opt && V > 0 && H > 0, "null operand or empty shape");
  DB_REQUIRE(h, world >= 2 && world <= 8 && rank >= 0 && rank < world && V % world == 0 && H % 4 == 0, "bad rank layout");
  DB_REQUIRE(h, opt->kind == DB_OPT_SGD || opt->kind == DB_OPT_MOMENTUM || opt->kind == DB_OPT_ADAM, "unknown optimizer kind");
  DB_REQUIRE(h, opt->kind == DB_OPT_SGD || opt_state, "the optimizer rule needs its slot buffer");
  if (opt->clock) return set_err(h, DB_ERR_UNSUPPORTED, "the sharded update takes the step number by value (no device clock)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_total = (size_t)V * H + V + H;
  float lr_t = opt->lr;
  if (opt->kind == DB_OPT_ADAM) {
    const double b1t = pow((double)opt->beta1, opt->step), b2t = pow((double)opt->beta2, opt->step);
    lr_t = (float)((double)opt->lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  }
  ShardArgs a{};
  a.rule = OptRule{opt->kind, opt->lr, lr_t, opt->momentum, opt->beta1, opt->beta2, opt->eps, opt->weight_decay};
  a.lr_t_dev = nullptr;
  a.V = V; a.H = H; a.world = world; a.rank = rank; a.rows_per = V / world;
  a.slots = slots; a.grad = grad;
  a.st = opt->kind != DB_OPT_SGD ? opt_state : nullptr;
  a.st2 = opt->kind == DB_OPT_ADAM ? opt_state + n_total : nullptr;
  for (int r = 0; r < world; ++r) { DB_REQUIRE(h, peer_params[r], "null peer pointer"); a.peer_params[r] = peer_params[r]; }
  const size_t n4 = (size_t)a.rows_per * H / 4;
  apply_shard_kernel<<<ew_grid(n4, h->num_sms), 256, 0, st>>>(a);
  return check_launch(h, "cd_tc_apply_update_sharded");
}
