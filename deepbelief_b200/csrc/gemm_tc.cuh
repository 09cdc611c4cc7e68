// Interface of the tcgen05 split-fp16 GEMM (gemm_tc.cu) used by the CD chain driver (cd_chain.cu).
#pragma once
#include "common.cuh"

namespace db {

// D[Ma, Nb] = A[Ma, K] * B[Nb, K]^T, both operands K-major fp16.
//   A = the REAL operand, pre-scaled by a power of two and split into hi + lo (two MMA passes, each into
//       its own TMEM accumulator, summed once by the epilogue); a_lo == nullptr -> single pass.
//   B = the EXACT operand (binary {0,1} states, exactly representable in fp16).
// The epilogue sees D "transposed": thread <-> row m of D (a unit), registers <-> columns nb (batch
// rows), so [Nb, Ma] row-major outputs (the layout every reference tensor uses, rbm.py:123-127) are
// written with lane-contiguous stores and the [Ma, Nb] "T" outputs with per-thread vector stores.
// Fused tail of the dW launch (north_star K2; rbm.py:664-670 cd_loss -> optimizer.minimize): the epilogue turns the
// accumulated gradient tile straight into the parameter update and rewrites the split-weight cache, so W, the
// optimizer slots and the four fp16 cache matrices make ONE pass through HBM per step, hidden under the MMAs.
// Indexing is that of out32: element (nb, m) lives at nb * ld32 + m  ([V, H] row-major for the dW launch).
struct TcOptArgs {
  int kind;                   // -1: no fused update; else DB_OPT_SGD / DB_OPT_MOMENTUM / DB_OPT_ADAM
  float lr, lr_t, momentum, beta1, beta2, eps, weight_decay;   // lr_t: bias-corrected Adam step size (host, fp64)
  float* theta;               // parameters, updated in place
  float* st; float* st2;      // optimizer slots (Momentum: st; Adam: st = m, st2 = v)
  __half* n_hi; __half* n_lo; // [Nb, Ma] fp16 split of theta_new * scale   (W cache), leading dimension ldn
  int ldn;
  __half* t_hi; __half* t_lo; // [Ma, Nb] fp16 split of theta_new * scale   (W^T cache), leading dimension ldt
  int ldt;
  const float* scale_dev;     // power-of-two scale the cache is written with
  unsigned int* max_bits;     // atomicMax of the bit pattern of |theta_new| (the next scale decision)
};

struct TcGemmParams {
  int Ma, Nb, K;
  const __half* a_hi; const __half* a_lo; int lda;   // [Ma, K]
  const __half* b; int ldb;                          // [Nb, K]
  float acc_scale;            // applied to the accumulator first (undoes the power-of-two pre-scale)
  const float* acc_scale_dev; // optional device scalar multiplied into acc_scale (weight-cache 1/scale)
  const float* bias;          // [Ma] or nullptr
  int act;                    // 0 linear, 1 logistic
  int sample;                 // 0 none, 1 Bernoulli (p > u, strict: rbm.py:207)
  const float* U; int ldu;    // injected uniforms [Nb, Ma], nullptr -> Philox
  RngArgs rng;
  float* out32; int ld32;             // [Nb, Ma] fp32 value after act (probabilities / linear result)
  // split-K (out32 only): the K axis is cut into k_splits contiguous chunks, chunk s writes its partial
  // product to out32 + s * split_stride; the caller sums the partials in fp32 (round-to-nearest). The
  // tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of
  // accumulation steps — chunking bounds it (measured: 1.2e-4 relative at K = 32768 unsplit).
  int k_splits; size_t split_stride;
  int group_m;                // set by tc_gemm_launch: row-blocks per rasterisation group
  int l2_hints;               // set by tc_gemm_launch: TMA loads carry L2 eviction priorities (A evict_last, B evict_first)
  int tile_n;                 // set by tc_gemm_launch: batch rows (TMEM columns) per tile: 256, 128 or 64
  int split_tail;             // set by tc_gemm_launch: cut the tiles of a sparse last wave into 128-column halves
  __half* out_s16; int lds16;         // [Nb, Ma] fp16 sample (sample==1) or value
  // "T" outputs [Ma, 2*Nb_pad] hold the two ends of the CD chain INTERLEAVED in blocks of 64 columns:
  // column(nb, phase) = (nb / 64) * 128 + phase * 64 + nb % 64   (phase 0 = data end, 1 = model end),
  // so the dW GEMM alternates +v0 p0 and -vk pk blocks along K and its fp32 partial sums stay small.
  __half* out_sT16; int ldsT; int sT_phase;    // fp16 sample, row-wise
  __half* out_pT_hi; __half* out_pT_lo; int ldpT; int pT_phase; float pT_scale;   // split of value*pT_scale
  TcOptArgs opt;              // opt.kind < 0 (set by every caller but the fused dW launch): unused
  // Data-parallel dW (reduce-scatter by PUSH over NVLink): with peer_world > 0 the fp32 result rows [r * peer_rows,
  // (r + 1) * peer_rows) of the [Nb, Ma] output are not written to out32 but stored straight into rank r's slot buffer
  // peer_out[r] (a peer-mapped pointer; this rank's own buffer for r == peer_rank) at slot peer_rank:
  //   peer_out[r][(peer_rank * peer_rows + (nb - r * peer_rows)) * ld32 + m].
  // peer_rows is a multiple of the tile width, so a 32-row slice never straddles two owners.
  float* peer_out[8]; int peer_world, peer_rank, peer_rows;
};

__host__ __device__ __forceinline__ int tc_interleaved_col(int nb, int phase) {
  return (nb >> 6) * 128 + phase * 64 + (nb & 63);
}

// returns cudaError_t as int (0 ok); -1000 = shape not supported by the tensor path
int tc_gemm_launch(const TcGemmParams& p, int num_sms, cudaStream_t stream);
bool tc_gemm_supported(int Ma, int Nb, int K, int lda, int ldb);

}  // namespace db
