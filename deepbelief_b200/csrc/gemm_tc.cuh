// Interface of the tcgen05 split-fp16 GEMM (gemm_tc.cu) used by the CD chain driver (cd_chain.cu).
#pragma once
#include "common.cuh"

namespace db {

// D[Ma, Nb] = A[Ma, K] * B[Nb, K]^T, both operands K-major fp16.
//   A = the REAL operand, pre-scaled by a power of two and split into hi + lo (two MMA passes into
//       one TMEM accumulator); a_lo == nullptr -> single pass.
//   B = the EXACT operand (binary {0,1} states, exactly representable in fp16).
// The epilogue sees D "transposed": thread <-> row m of D (a unit), registers <-> columns nb (batch
// rows), so [Nb, Ma] row-major outputs (the layout every reference tensor uses, rbm.py:123-127) are
// written with lane-contiguous stores and the [Ma, Nb] "T" outputs with per-thread vector stores.
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
  int kb_chunk;               // k-blocks (64 columns) per TMEM accumulation chunk before the fp32 promotion; 0 = default (8)
  int split_tail;             // set by tc_gemm_launch: cut the tiles of a sparse last wave into 128-column halves
  __half* out_s16; int lds16;         // [Nb, Ma] fp16 sample (sample==1) or value
  // "T" outputs [Ma, 2*Nb_pad] hold the two ends of the CD chain INTERLEAVED in blocks of 64 columns:
  // column(nb, phase) = (nb / 64) * 128 + phase * 64 + nb % 64   (phase 0 = data end, 1 = model end),
  // so the dW GEMM alternates +v0 p0 and -vk pk blocks along K and its fp32 partial sums stay small.
  __half* out_sT16; int ldsT; int sT_phase;    // fp16 sample, row-wise
  __half* out_pT_hi; __half* out_pT_lo; int ldpT; int pT_phase; float pT_scale;   // split of value*pT_scale
};

__host__ __device__ __forceinline__ int tc_interleaved_col(int nb, int phase) {
  return (nb >> 6) * 128 + phase * 64 + (nb & 63);
}

// returns cudaError_t as int (0 ok); -1000 = shape not supported by the tensor path
int tc_gemm_launch(const TcGemmParams& p, int num_sms, cudaStream_t stream);
bool tc_gemm_supported(int Ma, int Nb, int K, int lda, int ldb);

}  // namespace db
