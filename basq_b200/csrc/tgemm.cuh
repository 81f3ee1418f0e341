// Tensor-core GEMM with fp32 accuracy (fp16 hi / lo split, three products) on blocked operands - see tgemm.cu.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace basq {

// Matrix X [rows, kdim] stored as two fp16 arrays of the row-scaled matrix (x * rscale = hi + lo; rscale a power
// of two that brings the row's largest magnitude into [2^13, 2^14)) in the layout
// [row tile of 128][K chunk of 8 (kdim padded to 64)][128 rows][8]; padding is zero.  rinv[row] = 1 / rscale.
struct BlkOperand {
  DevBuf hi, lo, rinv;
  int rows = 0, kdim = 0;
  int RT = 0;  // row tiles
  int KC = 0;  // K chunks of 8 elements (multiple of 8: one 16 KB piece per row tile and 64-wide K block)
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
