// Tensor-core GEMM with fp32 accuracy (3xTF32) on blocked + split operands - see tgemm.cu.
#pragma once
#include "common.cuh"

namespace basq {

// Matrix X [rows, kdim] stored as two fp32 arrays (x = hi + lo, both tf32-representable) in the
// layout [row tile of 128][K chunk of 4 (kdim padded to 32)][128 rows][4]; padding is zero.
struct BlkOperand {
  DevBuf hi, lo;
  int rows = 0, kdim = 0;
  int RT = 0;  // row tiles
  int KC = 0;  // K chunks of 4 elements (multiple of 8)
  int alloc(basq_ctx* ctx, int rows, int kdim);
};

// op <- src (fp64, row-major, leading dimension ld).  transposed = false: src is [rows, kdim];
// transposed = true: src is [kdim, rows] (the operand is src^T).
int blk_from_f64(basq_ctx* ctx, const double* src, int64_t ld, bool transposed, BlkOperand* op);

// out = alpha * A B^T  ([A.rows, B.rows], fp64, row-major with leading dimension ldo), or its
// transpose ([B.rows, A.rows]) when `transposed`; `accumulate` adds to out instead of overwriting it.
int tgemm(basq_ctx* ctx, const BlkOperand& A, const BlkOperand& B, double alpha, double* out, int64_t ldo,
          bool transposed, bool accumulate = false);

}  // namespace basq
