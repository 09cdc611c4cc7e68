// Internal interface of the 3xTF32 tcgen05 GEMM (gemm_tf32.cu) used by the GP chain (gp.cu).
#pragma once
#include "context.h"

namespace db {

// C[M,N] = alpha X[M,K] Y[N,K]^T + beta C (fp32, K-major operands with their tf32 remainders). lower_only: only
// col <= row; tri_k: both operands are transposed lower-triangular factors (zero for k < row index), so each tile
// starts its K loop at max(row0, col0); tri_k = 2: only X is upper triangular (K loop from row0). C_lo (optional): the
// tf32 remainder of every written value, same layout as C. Returns 0, -1000 when TMA cannot address the operands, else a CUDA error.
int tf32_gemm_launch(db_handle_t h, cudaStream_t st, int M, int N, int K, float alpha, const float* X, const float* X_lo, int ldx,
                     const float* Y, const float* Y_lo, int ldy, float beta, float* C, int ldc, int lower_only, int tri_k, float* C_lo = nullptr);
bool tf32_gemm_supported(int M, int N, int K, const float* X, const float* X_lo, int ldx, const float* Y, const float* Y_lo, int ldy);
// lo = x - tf32(x) over a [rows, cols] block with leading dimension ld (same layout for lo)
void tf32_split_lo(db_handle_t h, cudaStream_t st, const float* x, size_t rows, size_t cols, size_t ld, float* lo);
// out[C,R] = in[R,C]^T
void transpose_f32(cudaStream_t st, const float* in, int R, int C, int ld_in, float* out, int ld_out);

}  // namespace db
