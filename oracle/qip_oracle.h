/*
 * qip_oracle.h -- CPU ORACLE for the RustQIP gate-application hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * algorithm; only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * (`cpu_baseline`, `--impl reference`) may load it.  The product library
 * (rustqip_b200/libqipb200.so) never links, loads or calls anything here.
 *
 * Parity status: the reference is Rust and cannot be compiled in this image
 * (no cargo/rustc, crates not vendored), so there is no oracle/_ref.  The
 * restatement is pinned against every known-answer test the reference holds
 * for this path (tests/test_oracle_reference_kat.py ports them 1:1), i.e.
 * index maps, row/col order, MSB conventions, X/SWAP/CNOT/identity matrices,
 * measurement collapse.  For GENERAL COMPLEX gate values (H, T, Rz, Haar
 * blocks) the reference itself holds no numeric test: there parity is
 * "unpinned by the reference's own tests" and rests on this restatement of
 * the cited lines plus kron/unitarity identities.
 *
 * Every function cites the reference lines it restates
 * (paths relative to /root/reference).
 */
#ifndef QIP_ORACLE_H
#define QIP_ORACLE_H

#include "../include/qip_op.h"

#ifdef __cplusplus
extern "C" {
#endif

/* qip-iterators/src/utils.rs:5-57 */
uint64_t qo_get_flat_index(uint64_t nindices, uint64_t i, uint64_t j);
uint64_t qo_flip_bits(uint64_t n, uint64_t num);
uint64_t qo_set_bit(uint64_t num, uint64_t bit_index, int value);
int qo_get_bit(uint64_t num, uint64_t bit_index);

/* qip-iterators/src/matrix_ops.rs:12-30 */
uint64_t qo_full_to_sub(uint64_t n, const uint64_t *mat_indices, uint64_t k, uint64_t full_index);
uint64_t qo_sub_to_full(uint64_t n, const uint64_t *mat_indices, uint64_t k, uint64_t sub_index,
                        uint64_t base);

/* qip/src/utils.rs:21-60 */
uint64_t qo_entwine_bits(uint64_t n, uint64_t selector, uint64_t off_bits, uint64_t on_bits);
uint64_t qo_extract_bits(uint64_t num, const uint64_t *indices, uint64_t n_indices);

/* Non-zero (col,val) entries of one op row, in the order the reference's row
 * iterators yield them (iterators/ops.rs:100-156, qubit_iterators.rs:8-219).
 * Writes at most `cap` entries; returns the number of entries of the row.
 * vals: interleaved (re,im) doubles (f32 ops are widened exactly). */
uint64_t qo_row_entries(const qip_op *op, int prec, uint64_t row, uint64_t *cols, double *vals,
                        uint64_t cap);

/* apply_op (accumulate != 0, matrix_ops.rs:98-123) and apply_op_overwrite
 * (accumulate == 0, matrix_ops.rs:127-152): one evaluation of
 * apply_op_row_indices (matrix_ops.rs:62-94) per output element, OpenMP over
 * output rows (the analogue of par_iter_mut, matrix_ops.rs:122,151).
 * Arithmetic: ascending non-zero columns from a zero accumulator, 4-mul/2-add
 * complex product (num-complex 0.4 `Mul`), no FMA contraction.
 * Lengths are in amplitudes. Returns 0, or -1 for a malformed op. */
int qo_apply_op_f64(uint64_t n, const qip_op *op, const double *input, uint64_t input_len,
                    double *output, uint64_t output_len, uint64_t input_offset,
                    uint64_t output_offset, int accumulate);
int qo_apply_op_f32(uint64_t n, const qip_op *op, const float *input, uint64_t input_len,
                    float *output, uint64_t output_len, uint64_t input_offset,
                    uint64_t output_offset, int accumulate);

/* apply_ops (matrix_ops.rs:158-219): [] = copy of the overlapping window, [op] = apply_op, several ops = the
 * multi-op row iterator (iterators/iterator_mapper.rs:8-31 + qubit_multi_iterator.rs:13-79), accumulated into
 * `output`.  Restated AS IT IS, including SURVEY.md quirk Q5 (row bits are peeled low-first per op while columns are
 * composed first-op-high: for ops that are not all alike the result is not their tensor product).  Pinned by the six
 * MultiOpIterator known-answer tests of the reference (qubit_multi_iterator.rs:82-205, ported in
 * tests/test_oracle_reference_kat.py) through qo_multi_op_iterator_*.  At most QO_MAX_MULTI_OPS ops. */
#define QO_MAX_MULTI_OPS 16
int qo_apply_ops_f64(uint64_t n, const qip_op *ops, uint64_t n_ops, const double *input, uint64_t input_len,
                     double *output, uint64_t output_len, uint64_t input_offset, uint64_t output_offset);
int qo_apply_ops_f32(uint64_t n, const qip_op *ops, uint64_t n_ops, const float *input, uint64_t input_len,
                     float *output, uint64_t output_len, uint64_t input_offset, uint64_t output_offset);
/* MultiOpIterator::new(ns, lists).collect(): list i = lens[i] entries (cols[i][e], vals[i][2e..2e+1]). */
uint64_t qo_multi_op_iterator_f64(const uint64_t *ns, const uint64_t *const *cols, const double *const *vals,
                                  const uint64_t *lens, uint64_t n_lists, uint64_t *out_cols, double *out_vals,
                                  uint64_t cap);
uint64_t qo_multi_op_iterator_f32(const uint64_t *ns, const uint64_t *const *cols, const float *const *vals,
                                  const uint64_t *lens, uint64_t n_lists, uint64_t *out_cols, float *out_vals,
                                  uint64_t cap);

/* qip/src/state_ops/measurement_ops.rs:11-13 */
double qo_prob_magnitude_f64(const double *input, uint64_t len);
float qo_prob_magnitude_f32(const float *input, uint64_t len);
/* measurement_ops.rs:44-112 (input_offset < 0 == None) */
double qo_measure_prob_f64(uint64_t n, uint64_t measured, const uint64_t *indices, uint64_t n_indices,
                           const double *input, uint64_t input_len, uint64_t input_offset);
float qo_measure_prob_f32(uint64_t n, uint64_t measured, const uint64_t *indices, uint64_t n_indices,
                          const float *input, uint64_t input_len, uint64_t input_offset);
/* measurement_ops.rs:115-127: out has 2^n_indices entries */
void qo_measure_probs_f64(uint64_t n, const uint64_t *indices, uint64_t n_indices, const double *input,
                          uint64_t input_len, uint64_t input_offset, double *out);
void qo_measure_probs_f32(uint64_t n, const uint64_t *indices, uint64_t n_indices, const float *input,
                          uint64_t input_len, uint64_t input_offset, float *out);
/* measurement_ops.rs:153-176 with the random draw r in [0,1) supplied by the caller */
uint64_t qo_soft_measure_f64(uint64_t n, const uint64_t *indices, uint64_t n_indices, const double *input,
                             uint64_t input_len, uint64_t input_offset, double r);
uint64_t qo_soft_measure_f32(uint64_t n, const uint64_t *indices, uint64_t n_indices, const float *input,
                             uint64_t input_len, uint64_t input_offset, double r);
/* measurement_ops.rs:220-269 */
void qo_measure_state_f64(uint64_t n, const uint64_t *indices, uint64_t n_indices, uint64_t measured,
                          double measured_prob, const double *input, uint64_t input_len, double *output,
                          uint64_t output_len, uint64_t input_offset, uint64_t output_offset);
void qo_measure_state_f32(uint64_t n, const uint64_t *indices, uint64_t n_indices, uint64_t measured,
                          float measured_prob, const float *input, uint64_t input_len, float *output,
                          uint64_t output_len, uint64_t input_offset, uint64_t output_offset);

int qo_max_threads(void);
void qo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
