/*
 * qip_oracle.c -- CPU ORACLE (plain C restatement of the reference hot path).
 * TEST INFRASTRUCTURE ONLY: see qip_oracle.h for who may load this and for the
 * parity-pinning statement.  Build: `make -C oracle` (gcc -O2 -ffp-contract=off
 * -fopenmp; no -ffast-math, no FMA contraction, so the per-element arithmetic
 * order equals the reference's).
 */
#include "qip_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- qip-iterators/src/utils.rs ------------------------------------------------ */

uint64_t qo_get_flat_index(uint64_t nindices, uint64_t i, uint64_t j) { /* utils.rs:5-8 */
  return i * ((uint64_t)1 << nindices) + j;
}

uint64_t qo_flip_bits(uint64_t n, uint64_t num) { /* utils.rs:22-25: reverse_bits >> (64-n) */
  uint64_t r = 0;
  for (uint64_t i = 0; i < n; ++i) r |= ((num >> i) & 1) << (n - 1 - i);
  return r;
}

uint64_t qo_set_bit(uint64_t num, uint64_t bit_index, int value) { /* utils.rs:37-44 */
  uint64_t v = (uint64_t)1 << bit_index;
  return value ? (num | v) : (num & ~v);
}

int qo_get_bit(uint64_t num, uint64_t bit_index) { /* utils.rs:55-57 */
  return (int)((num >> bit_index) & 1);
}

/* ---- qip-iterators/src/matrix_ops.rs:12-30 ------------------------------------- */

uint64_t qo_full_to_sub(uint64_t n, const uint64_t *mat_indices, uint64_t k, uint64_t full_index) {
  uint64_t acc = 0;
  for (uint64_t j = 0; j < k; ++j) {
    int bit = qo_get_bit(full_index, n - 1 - mat_indices[j]);
    acc = qo_set_bit(acc, k - 1 - j, bit);
  }
  return acc;
}

uint64_t qo_sub_to_full(uint64_t n, const uint64_t *mat_indices, uint64_t k, uint64_t sub_index,
                        uint64_t base) {
  uint64_t acc = base;
  for (uint64_t j = 0; j < k; ++j) {
    int bit = qo_get_bit(sub_index, k - 1 - j);
    acc = qo_set_bit(acc, n - 1 - mat_indices[j], bit);
  }
  return acc;
}

/* ---- qip/src/utils.rs:21-60 ----------------------------------------------------- */

uint64_t qo_entwine_bits(uint64_t n, uint64_t selector, uint64_t off_bits, uint64_t on_bits) {
  uint64_t result = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if ((selector & 1) == 0) {
      result |= (off_bits & 1) << i;
      off_bits >>= 1;
    } else {
      result |= (on_bits & 1) << i;
      on_bits >>= 1;
    }
    selector >>= 1;
  }
  return result;
}

uint64_t qo_extract_bits(uint64_t num, const uint64_t *indices, uint64_t n_indices) {
  uint64_t acc = 0;
  for (uint64_t i = 0; i < n_indices; ++i) acc |= ((num >> indices[i]) & 1) << i;
  return acc;
}

/* ---- op flattening ---------------------------------------------------------------
 * Control(nc, idx, inner) dispatches to inner.sum_for_control_iterator, and a
 * nested Control adds its nc to the running count (iterators/ops.rs:111-115,
 * :147-153).  The bit positions always come from the OUTERMOST op's indices
 * (matrix_ops.rs:106,137: op.indices()).                                        */
typedef struct flat_op {
  uint64_t k;        /* len(outer indices)                        */
  const uint64_t *indices;
  uint64_t thr;      /* 2^k - 2^kop, 0 when there are no controls */
  uint64_t kop;      /* index count of the innermost non-control op */
  int base_kind;
  const void *dense;
  const uint64_t *sp_rowptr, *sp_col;
  const void *sp_val;
} flat_op;

static int flatten_op(const qip_op *op, flat_op *f) {
  if (!op || !op->indices) return -1;
  memset(f, 0, sizeof(*f));
  f->k = op->n_indices;
  f->indices = op->indices;
  uint64_t nc = 0;
  const qip_op *cur = op;
  uint64_t kop = op->n_indices;
  int depth = 0;
  while (cur->kind == QIP_OP_CONTROL) {
    if (!cur->inner || cur->n_control > cur->n_indices || ++depth > 64) return -1;
    nc += cur->n_control;
    kop = cur->n_indices - cur->n_control; /* ops.rs:112,150 */
    cur = cur->inner;
  }
  if (nc + kop != f->k || f->k == 0 || f->k > 63) return -1;
  f->kop = kop;
  f->thr = nc ? (((uint64_t)1 << f->k) - ((uint64_t)1 << kop)) : 0; /* qubit_iterators.rs:130-131 */
  f->base_kind = cur->kind;
  switch (cur->kind) {
    case QIP_OP_MATRIX:
      if (!cur->dense) return -1;
      f->dense = cur->dense;
      break;
    case QIP_OP_SPARSE:
      if (!cur->sp_rowptr || !cur->sp_col || !cur->sp_val) return -1;
      f->sp_rowptr = cur->sp_rowptr;
      f->sp_col = cur->sp_col;
      f->sp_val = cur->sp_val;
      break;
    case QIP_OP_SWAP:
      break;
    default:
      return -1;
  }
  return 0;
}

uint64_t qo_row_entries(const qip_op *op, int prec, uint64_t row, uint64_t *cols, double *vals,
                        uint64_t cap) {
  flat_op f;
  if (flatten_op(op, &f) != 0) return 0;
  uint64_t cnt = 0;
#define EMIT(c_, re_, im_)            \
  do {                                \
    if (cnt < cap) {                  \
      cols[cnt] = (c_);               \
      vals[2 * cnt] = (re_);          \
      vals[2 * cnt + 1] = (im_);      \
    }                                 \
    ++cnt;                            \
  } while (0)
  if (row < f.thr) {
    EMIT(row, 1.0, 0.0);
    return cnt;
  }
  uint64_t r = row - f.thr;
  if (f.base_kind == QIP_OP_MATRIX) {
    uint64_t side = (uint64_t)1 << f.kop;
    for (uint64_t c = 0; c < side; ++c) {
      double re, im;
      if (prec == QIP_F32) {
        re = ((const float *)f.dense)[2 * (r * side + c)];
        im = ((const float *)f.dense)[2 * (r * side + c) + 1];
      } else {
        re = ((const double *)f.dense)[2 * (r * side + c)];
        im = ((const double *)f.dense)[2 * (r * side + c) + 1];
      }
      if (re == 0.0 && im == 0.0) continue;
      EMIT(c + f.thr, re, im);
    }
  } else if (f.base_kind == QIP_OP_SPARSE) {
    for (uint64_t e = f.sp_rowptr[r]; e < f.sp_rowptr[r + 1]; ++e) {
      double re, im;
      if (prec == QIP_F32) {
        re = ((const float *)f.sp_val)[2 * e];
        im = ((const float *)f.sp_val)[2 * e + 1];
      } else {
        re = ((const double *)f.sp_val)[2 * e];
        im = ((const double *)f.sp_val)[2 * e + 1];
      }
      EMIT(f.sp_col[e] + f.thr, re, im);
    }
  } else {
    uint64_t half = f.kop >> 1;
    uint64_t lower_mask = ~(~(uint64_t)0 << half);
    EMIT((((r & lower_mask) << half) + (r >> half)) + f.thr, 1.0, 0.0);
  }
#undef EMIT
  return cnt;
}

#define T double
#define SFX f64
#define SQRT sqrt
#include "qip_oracle_impl.inc"
#undef T
#undef SFX
#undef SQRT

#define T float
#define SFX f32
#define SQRT sqrtf
#include "qip_oracle_impl.inc"
#undef T
#undef SFX
#undef SQRT

int qo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void qo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
