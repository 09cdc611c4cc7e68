// Internal interface of the 3 x fp16 tcgen05 GEMM (gemm_f16x3.cu) used by the big products of the GP chain (gp.cu).
#pragma once
#include "context.h"

namespace db {

// A matrix as a power-of-two scaled fp16 pair: x * (*scale) = hi + lo, same indexing (leading dimension) for hi and lo.
struct F16Operand {
  __half* hi; __half* lo;
  const float* scale;     // device scalar, 2^e
  F16Operand at(size_t elem_offset) const { return F16Operand{hi + elem_offset, lo + elem_offset, scale}; }
};

// C[M,N] = alpha X[M,K] Y[N,K]^T + beta C (fp32 result). lower_only / tri_k as tf32_gemm_launch. Leading dimensions in
// elements, multiples of 8; 16-byte aligned operands. Returns 0, -1000 when TMA cannot address the operands, else a CUDA error.
// C16 (optional): the written values again as a scaled fp16 pair (scale *C16->scale, leading dimension ldc16), so that a
// product that becomes an operand needs no conversion pass; C may then be null (beta must be 0).
// max_ctas > 0 caps the persistent grid (a launch that shares the GPU with another stream's kernels, gp.cu). allow_pdl = 0:
// launched without the programmatic-dependent-launch attribute — REQUIRED for the first launch behind a cross-stream event
// wait: captured into a CUDA graph, a dependent launch did not honour the event edge (measured: N = 8192 replays differed
// by 1e-2 from the eager run until that launch was made a plain one).
int f16_gemm_launch(db_handle_t h, cudaStream_t st, int M, int N, int K, float alpha, const F16Operand& X, int ldx,
                    const F16Operand& Y, int ldy, float beta, float* C, int ldc, int lower_only, int tri_k,
                    const F16Operand* C16 = nullptr, int ldc16 = 0, int max_ctas = 0, int allow_pdl = 1);
bool f16_gemm_supported(int M, int N, int K, const F16Operand& X, int ldx, const F16Operand& Y, int ldy);
// out = split(x * (*out.scale)) over a [rows, cols] block (x: leading dimension ld, out: ld16)
void f16_split(db_handle_t h, cudaStream_t st, const float* x, size_t rows, size_t cols, size_t ld, const F16Operand& out, size_t ld16);
// *scale_out = 2^(14 - ceil(log2 (max|x| * m))) (bits_scratch: one device word; m = |*bound_mul| on the device, 1 when null: the
// scale of a matrix that is bounded by max|x| times a device scalar)
void f16_scale_from_absmax(db_handle_t h, cudaStream_t st, const float* x, size_t rows, size_t cols, size_t ld, unsigned int* bits_scratch,
                           float* scale_out, const float* bound_mul = nullptr);
// a-priori scales from the effective hyper-parameters: [0] Cholesky panels, [1] L^-1 (U), [2] K^-1, [5] L_21 L_11^-1 products
void f16_gp_bounds(cudaStream_t st, const float* hyp, int N, float* scales);

}  // namespace db
