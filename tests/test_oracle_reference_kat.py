"""Pins the CPU oracle against every known-answer test the reference holds for the
gate-application path.  Each test names the reference test it ports.
No GPU needed."""
import math

import numpy as np
import pytest

from oracle import qip_oracle as qo
from rustqip_b200.ops import (MatrixOp, Representation, from_reals, make_control_op,
                              make_sparse_matrix_op)
from rustqip_b200 import matrix_ops as host_mo


def kron_helper(before, mat, after):
    """ndarray_kron_helper, qip-iterators/src/matrix_ops.rs:257-269."""
    eye = np.eye(2)
    for _ in range(before):
        mat = np.kron(eye, mat)
    for _ in range(after):
        mat = np.kron(mat, eye)
    return mat


# ---- qip-iterators/src/matrix_ops.rs:272-374 (8 kron-equality cases) -------------------
@pytest.mark.parametrize("name,n,indices,data,before,after", [
    ("test_ident", 3, [0], [1, 0, 0, 1], 0, 2),
    ("test_flip", 3, [0], [0, 1, 1, 0], 0, 2),
    ("test_flip_mid", 3, [1], [0, 1, 1, 0], 1, 1),
    ("test_flip_end", 3, [2], [0, 1, 1, 0], 2, 0),
    ("test_flip_mid_twobody", 4, [1, 2], [1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1], 1, 1),
    ("test_counting", 3, [0], [1, 2, 3, 4], 0, 2),
])
def test_matrix_ops_kron(name, n, indices, data, before, after):
    op = MatrixOp.new_matrix(indices, data)
    side = 1 << len(indices)
    arr = np.array(data, dtype=np.float64).reshape(side, side)
    mat = qo.make_op_matrix(n, op)
    assert np.array_equal(mat, kron_helper(before, arr, after).astype(np.complex128)), name


def test_counting_order():  # matrix_ops.rs:351-361
    data = list(range(16))
    op = MatrixOp.new_matrix([0, 1], data)
    assert np.array_equal(qo.make_op_matrix(2, op).real, np.array(data, dtype=float).reshape(4, 4))


def test_counting_order_flipped():  # matrix_ops.rs:364-374
    data = list(range(16))
    op = MatrixOp.new_matrix([1, 0], data)
    assert not np.array_equal(qo.make_op_matrix(2, op).real, np.array(data, dtype=float).reshape(4, 4))


# ---- qip-iterators/src/iterators/qubit_iterators.rs:290-379 -------------------------------
def _row_matrix(op, k):
    m = np.zeros((1 << k, 1 << k))
    for r in range(1 << k):
        for c, _ in qo.row_entries(op, r):
            m[r, c] = 1.0
    return m


def test_mat_iterator():  # :290-307
    op = MatrixOp.new_matrix([0], from_reals([0.0, 1.0, 1.0, 0.0]))
    assert np.array_equal(_row_matrix(op, 1), [[0, 1], [1, 0]])


def test_sparse_mat_iterator():  # :310-328
    op = MatrixOp.new_sparse([0], [[(1, 1.0)], [(0, 1.0)]])
    assert np.array_equal(_row_matrix(op, 1), [[0, 1], [1, 0]])


def test_swap_iterator():  # :331-352
    op = MatrixOp.new_swap([0], [1])
    assert np.array_equal(_row_matrix(op, 2), [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])


def test_c_iterator():  # :355-379
    op = MatrixOp.new_control([0], [1], MatrixOp.new_matrix([1], from_reals([0.0, 1.0, 1.0, 0.0])))
    assert np.array_equal(_row_matrix(op, 2), [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])


# ---- qip-iterators/src/iterators/qubit_multi_iterator.rs:82-205 (MultiOpIterator) -------------------
def test_multi_iter_trivial():  # :87-98
    assert qo.multi_op_iterator([1, 1], [[(1, 1.0)], [(0, 1.0)]]) == [(2, 1 + 0j)]


def test_multi_iter_nontrivial():  # :100-112
    assert qo.multi_op_iterator([1, 1], [[(0, 1.0), (1, 1.0)], [(0, 1.0)]]) == [(0, 1 + 0j), (2, 1 + 0j)]


def test_multi_iter_nontrivial_other():  # :114-126
    assert qo.multi_op_iterator([1, 1], [[(0, 1.0)], [(0, 1.0), (1, 1.0)]]) == [(0, 1 + 0j), (1, 1 + 0j)]


def _multi_iter_matrix(n, lists_of_row, ns):
    m = np.zeros((1 << n, 1 << n))
    for i in range(1 << n):
        for col, _ in qo.multi_op_iterator(ns, lists_of_row(i)):
            m[i, col] = 1.0
    return m


def test_multi_iter_mat_iterator():  # :128-149
    assert np.array_equal(_multi_iter_matrix(1, lambda i: [[(1 - i, 1.0)]], [1]), [[0, 1], [1, 0]])


def test_multi_iter_double_mat_identity():  # :151-177
    m = _multi_iter_matrix(2, lambda i: [[((i & 2) >> 1, 1.0)], [(i & 1, 1.0)]], [1, 1])
    assert np.array_equal(m, np.eye(4))


def test_multi_iter_double_mat_swap():  # :179-205
    m = _multi_iter_matrix(2, lambda i: [[((~i & 2) >> 1, 1.0)], [(~i & 1, 1.0)]], [1, 1])
    assert np.array_equal(m, np.eye(4)[::-1])


def test_apply_ops_empty_and_single():  # matrix_ops.rs:167-183
    rng = np.random.default_rng(5)
    n = 4
    psi = (rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex128)
    out = np.zeros(10, dtype=np.complex128)
    qo.apply_ops(n, [], psi[2:14].copy(), out, input_offset=2, output_offset=5)  # window [5, 14)
    assert np.array_equal(out[:9], psi[5:14]) and out[9] == 0
    op = MatrixOp.new_matrix([1], from_reals([0.0, 1.0, 1.0, 0.0]))
    a = np.zeros(16, dtype=np.complex128)
    b = np.zeros(16, dtype=np.complex128)
    qo.apply_ops(n, [op], psi, a)
    qo.apply_op(n, op, psi, b)
    assert np.array_equal(a, b)


def test_apply_ops_identical_single_qubit_ops():
    """What the reference's benches do (state_bench.rs:226-236): the same 1-qubit gate on several qubits.  Op i takes
    its ROW bit from the qubit of op (last - i) and writes its COLUMN bit at its own qubit (quirk Q5), so the result is
    the tensor product applied to the input with those qubits in reversed order."""
    rng = np.random.default_rng(6)
    n = 5
    psi = (rng.standard_normal(32) + 1j * rng.standard_normal(32)).astype(np.complex128)
    h = [x / np.sqrt(2) for x in (1.0, 1.0, 1.0, -1.0)]
    qs = (0, 2, 3)
    ops = [MatrixOp.new_matrix([q], from_reals(h)) for q in qs]
    out = np.zeros(32, dtype=np.complex128)
    qo.apply_ops(n, ops, psi, out)
    rev = np.zeros_like(psi)
    for i in range(32):
        j = i
        for a, b in zip(qs, qs[::-1]):
            j = (j & ~(1 << (n - 1 - a))) | (((i >> (n - 1 - b)) & 1) << (n - 1 - a))
        rev[j] = psi[i]
    assert np.allclose(out, qo.run_pipeline(n, ops, state=rev), atol=1e-14)
    assert not np.allclose(out, qo.run_pipeline(n, ops, state=psi), atol=1e-3)


def test_apply_ops_quirk_q5():
    """SURVEY.md Q5, restated as the reference computes it: row bits are peeled low-first per op
    (iterator_mapper.rs:16-25), columns composed first-op-high (qubit_multi_iterator.rs:48-52): [X(q0), I(q1)] on
    two qubits is a 4-cycle, not X (x) I."""
    x = MatrixOp.new_matrix([0], from_reals([0.0, 1.0, 1.0, 0.0]))
    i2 = MatrixOp.new_matrix([1], from_reals([1.0, 0.0, 0.0, 1.0]))
    m = np.zeros((4, 4))
    for col in range(4):
        e = np.zeros(4, dtype=np.complex128)
        e[col] = 1
        out = np.zeros(4, dtype=np.complex128)
        qo.apply_ops(2, [x, i2], e, out)
        m[:, col] = out.real
    # row r = (b1 b0): op 0 (X) sees b0, op 1 (I) sees b1; column = (X-col << 1) | I-col = ((1 - b0) << 1) | b1
    expect = np.zeros((4, 4))
    for r in range(4):
        b1, b0 = r >> 1, r & 1
        expect[r, ((1 - b0) << 1) | b1] = 1
    assert np.array_equal(m, expect)
    assert not np.array_equal(m, np.kron([[0, 1], [1, 0]], np.eye(2)))
    p = np.linalg.matrix_power(m, 4)
    assert np.array_equal(p, np.eye(4)) and not np.array_equal(np.linalg.matrix_power(m, 2), np.eye(4))


# ---- qip/src/state_ops/matrix_ops.rs:265-377 ---------------------------------------------
def test_get_bit_set_bit():  # :265-275
    assert not qo.get_bit(1, 1)
    assert qo.get_bit(1, 0)
    assert qo.set_bit(1, 0, True) == 1
    assert qo.set_bit(1, 1, True) == 3


def test_get_index_simple_condition_swap():  # :277-304
    op = MatrixOp.new_matrix([0, 1, 2], [])
    assert op.num_indices() == 3 and [host_mo.get_index(op, i) for i in range(3)] == [0, 1, 2]
    cop = make_control_op([0, 1], MatrixOp.new_matrix([2, 3], []))
    assert cop.num_indices() == 4 and [host_mo.get_index(cop, i) for i in range(4)] == [0, 1, 2, 3]
    sop = MatrixOp.new_swap([0, 1], [2, 3])
    assert sop.num_indices() == 4 and [host_mo.get_index(sop, i) for i in range(4)] == [0, 1, 2, 3]


def test_apply_identity():  # :307-314
    op = MatrixOp.new_matrix([0], from_reals([1.0, 0.0, 0.0, 1.0]))
    inp, out = from_reals([1.0, 0.0]), from_reals([0.0, 0.0])
    qo.apply_op(1, op, inp, out)
    assert np.array_equal(inp, out)


def test_apply_swap_mat():  # :317-325
    op = MatrixOp.new_matrix([0], from_reals([0.0, 1.0, 1.0, 0.0]))
    inp, out = from_reals([1.0, 0.0]), from_reals([0.0, 0.0])
    qo.apply_op(1, op, inp, out)
    assert np.array_equal(inp[::-1], out)


def test_apply_swap_mat_first():  # :328-344 (qubit 0 is the MSB)
    inp = from_reals([1.0, 0.0, 0.0, 0.0])
    out = from_reals([0.0] * 4)
    qo.apply_op(2, MatrixOp.new_matrix([0], from_reals([0.0, 1.0, 1.0, 0.0])), inp, out)
    assert np.array_equal(out, from_reals([0.0, 0.0, 1.0, 0.0]))
    out = from_reals([0.0] * 4)
    qo.apply_op(2, MatrixOp.new_matrix([1], from_reals([0.0, 1.0, 1.0, 0.0])), inp, out)
    assert np.array_equal(out, from_reals([0.0, 1.0, 0.0, 0.0]))


def test_make_sparse_mat():  # :347-377
    one = 1 + 0j
    expected = [[(1, one)], [(0, one)], [(3, one)], [(2, one)]]
    op1 = make_sparse_matrix_op([0, 1], expected, Representation.BigEndian)
    op2 = make_sparse_matrix_op([0, 1], [[(2, one)], [(3, one)], [(0, one)], [(1, one)]],
                                Representation.LittleEndian)
    assert op1.rows == expected and op2.rows == expected


# ---- doctests: qip-iterators/src/utils.rs:14-20,30-35,49-53; qip/src/utils.rs:12-20,49-53 ----
def test_utils_doctests():
    assert qo.flip_bits(3, 0b100) == 0b001
    assert qo.flip_bits(3, 0b010) == 0b010
    assert qo.flip_bits(4, 0b1010) == 0b0101
    assert qo.set_bit(0, 1, True) == 2 and qo.set_bit(1, 1, True) == 3 and qo.set_bit(1, 0, False) == 0
    assert qo.get_bit(2, 1)
    assert qo.entwine_bits(3, 0b010, 0b01, 0b1) == 0b011
    assert qo.extract_bits(0b1010, [3, 0]) == 0b01
    assert qo.get_flat_index(2, 1, 3) == 7
    # the host mirror's copies agree with the oracle
    for full in range(32):
        assert host_mo.full_to_sub(5, [3, 0, 4], full) == qo.full_to_sub(5, [3, 0, 4], full)
        assert host_mo.sub_to_full(5, [3, 0, 4], full & 7, full) == qo.sub_to_full(5, [3, 0, 4], full & 7, full)


# ---- qip/src/state_ops/measurement_ops.rs:24-43,136-152 (doctests), :291-335 (tests) -------
def test_measure_prob_doctest():
    inp = from_reals([0.0, 0.0, 1.0, 0.0])
    assert qo.measure_prob(2, 0, [0], inp) == 0.0
    assert qo.measure_prob(2, 1, [0], inp) == 1.0
    assert qo.measure_prob(2, 1, [0, 1], inp) == 1.0
    assert qo.measure_prob(2, 2, [1, 0], inp) == 1.0


def test_soft_measure_doctest():
    inp = from_reals([0.0, 0.0, 1.0, 0.0])
    for r in (1e-12, 0.3, 0.999):  # r == 0.0 exactly selects index 0 in the reference too
        assert qo.soft_measure(2, [0], inp, r) == 1
        assert qo.soft_measure(2, [1], inp, r) == 0
        assert qo.soft_measure(2, [0, 1], inp, r) == 0b01
        assert qo.soft_measure(2, [1, 0], inp, r) == 0b10


@pytest.mark.parametrize("m,expected", [(0, [math.sqrt(0.5), math.sqrt(0.5), 0, 0]),
                                        (1, [0, 0, math.sqrt(0.5), math.sqrt(0.5)])])
def test_measure_state(m, expected):  # :291-326
    inp = from_reals([0.5, 0.5, 0.5, 0.5])
    p = qo.measure_prob(2, m, [0], inp)
    assert abs(p - 0.5) < np.finfo(np.float64).eps
    out = inp.copy()
    qo.measure_state(2, [0], m, p, inp, out)
    assert np.allclose(out, from_reals(expected), atol=1e-10, rtol=0)


def test_measure_probs():  # :329-335
    inp = from_reals([0.5, 0.5, 0.5, 0.5])
    assert list(qo.measure_probs(2, [1], inp)) == [0.5, 0.5]


# ---- structural identities that pin general complex gates (unpinned by the reference) -------
def test_general_complex_gate_matches_kron():
    rng = np.random.default_rng(7)
    for n, idx in [(4, [2]), (4, [0, 3]), (5, [4, 1, 2])]:
        k = len(idx)
        u = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
        mat = qo.make_op_matrix(n, MatrixOp.new_matrix(idx, u.reshape(-1)))
        # expected via explicit index arithmetic (row-major, idx[0] = MSB, qubit q <-> bit n-1-q)
        exp = np.zeros_like(mat)
        for r in range(1 << n):
            for c in range(1 << n):
                if all(((r ^ c) >> (n - 1 - q)) & 1 == 0 for q in range(n) if q not in idx):
                    rs = sum(((r >> (n - 1 - q)) & 1) << (k - 1 - j) for j, q in enumerate(idx))
                    cs = sum(((c >> (n - 1 - q)) & 1) << (k - 1 - j) for j, q in enumerate(idx))
                    exp[r, c] = u[rs, cs]
        assert np.allclose(mat, exp, atol=1e-15)


def test_offsets_zero_outside_window():
    """matrix_ops.rs:79-89: partners outside [input_offset, input_offset+len) read as zero."""
    n = 3
    op = MatrixOp.new_matrix([0], [0, 1, 1, 0])  # X on the MSB: partner = i ^ 4
    full = (np.arange(8) + 1).astype(np.complex128)
    out = np.zeros(4, dtype=np.complex128)
    qo.apply_op_overwrite(n, op, full[4:], out, input_offset=4, output_offset=0)
    assert np.array_equal(out, full[4:])          # rows 0..3 read partners 4..7
    out = np.zeros(4, dtype=np.complex128)
    qo.apply_op_overwrite(n, op, full[:4], out, input_offset=0, output_offset=0)
    assert np.array_equal(out, np.zeros(4))       # partners 4..7 are outside the window
