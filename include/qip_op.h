/*
 * qip_op.h -- plain-C descriptor of one RustQIP gate ("MatrixOp").
 *
 * Mirrors `enum MatrixOp<P>` of the reference
 * (qip-iterators/src/iterators/ops.rs:11-20):
 *     Matrix(indices, data)             -> QIP_OP_MATRIX
 *     SparseMatrix(indices, rows)       -> QIP_OP_SPARSE  (rows flattened to CSR)
 *     Swap(m, a_indices ++ b_indices)   -> QIP_OP_SWAP
 *     Control(nc, controls ++ inner.indices, Box<inner>) -> QIP_OP_CONTROL
 *
 * Conventions (all pinned by the reference's tests):
 *   - qubit q lives at index bit (n-1-q): qubit 0 is the MSB
 *     (qip-iterators/src/matrix_ops.rs:12-21).
 *   - indices[0] is the MSB of the 2^k sub-index; dense data is row-major,
 *     out_sub[row] = sum_col data[row*2^k + col] * in_sub[col]
 *     (qip-iterators/src/utils.rs:5-8, iterators/qubit_iterators.rs:23-31).
 *   - amplitudes are interleaved (re, im) pairs of float or double
 *     (num_complex::Complex<P>, P in {f32,f64}: qip/src/types.rs:6-13).
 *
 * All pointers are HOST pointers borrowed for the duration of a call.
 * This header is shared by the product library (include/qipb200.h) and by the
 * CPU oracle (oracle/qip_oracle.h) so that tests hand the same bytes to both.
 */
#ifndef QIP_OP_H
#define QIP_OP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qip_prec { QIP_F32 = 0, QIP_F64 = 1 } qip_prec;

typedef enum qip_op_kind {
  QIP_OP_MATRIX = 0,
  QIP_OP_SPARSE = 1,
  QIP_OP_SWAP = 2,
  QIP_OP_CONTROL = 3
} qip_op_kind;

typedef struct qip_op {
  int32_t kind;           /* qip_op_kind */
  uint32_t n_indices;     /* k = len(indices); SWAP: 2*m; CONTROL: nc + inner k */
  uint32_t n_control;     /* CONTROL only: nc (leading entries of indices)      */
  uint32_t reserved;      /* must be 0 */
  uint64_t n_entries;     /* MATRIX: complex values in `dense` (must be 4^k, the
                             reference checks dat.len(): state_ops/matrix_ops.rs:14-23);
                             SPARSE: number of rows (must be 2^k, :35-46); else 0 */
  const uint64_t *indices;   /* reference qubit numbers, order significant      */
  const void *dense;         /* MATRIX: 4^k complex<prec>, row-major            */
  const uint64_t *sp_rowptr; /* SPARSE: 2^k + 1 row offsets into sp_col/sp_val  */
  const uint64_t *sp_col;    /* SPARSE: column of each stored entry             */
  const void *sp_val;        /* SPARSE: complex<prec> value of each entry       */
  const struct qip_op *inner;/* CONTROL: the controlled op (may itself be CONTROL) */
} qip_op;

#ifdef __cplusplus
}
#endif
#endif /* QIP_OP_H */
