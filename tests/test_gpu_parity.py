"""GPU parity: every CUDA path against the CPU oracle on the same seeded inputs.

* the host-buffer drop-ins (qipb200_apply_op / _overwrite == qip_iterators::matrix_ops)
  must be BIT-IDENTICAL to the oracle: the row kernel issues the reference's exact
  arithmetic sequence with non-contracted IEEE ops;
* the device-resident state path (in-place kernels, FMA allowed) must agree within
  rel-tol 1e-10 (f64) / 1e-5 (f32) on amplitudes, and exactly (==, +-0 equal) for
  the X / CNOT / Toffoli-X / SWAP index permutations (BASELINE.json north_star).
"""
import math

import numpy as np
import pytest

from oracle import qip_oracle as qo
from rustqip_b200 import circuits, gates
from rustqip_b200 import matrix_ops as mo
from rustqip_b200.errors import CircuitError
from rustqip_b200.ops import MatrixOp, make_control_op, make_matrix_op, make_swap_op
from rustqip_b200.state import State

pytestmark = pytest.mark.gpu

TOL = {np.complex128: 1e-10, np.complex64: 1e-5}


def rand_state(n, dtype, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    return np.ascontiguousarray(v.astype(dtype))


def rand_unitary(k, rng):
    a = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def assert_close(got, want, dtype):
    tol = TOL[np.dtype(dtype).type]
    scale = max(np.max(np.abs(want)), 1e-300)
    err = np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))) / scale
    assert err <= tol, "max rel err %.3e > %.1e" % (err, tol)
    l2 = np.linalg.norm(got.astype(np.complex128) - want.astype(np.complex128)) / max(
        np.linalg.norm(want.astype(np.complex128)), 1e-300)
    assert l2 <= tol


def oracle_apply(n, op, psi):
    out = np.zeros_like(psi)
    qo.apply_op_overwrite(n, op, psi, out)
    return out


def gpu_state_apply(ctx, n, ops, psi, fusion=False):
    with State(n, psi.dtype, ctx) as st:
        st.upload(psi)
        if len(ops) == 1:
            st.apply_op(ops[0])
        else:
            st.apply_schedule(ops, fusion=fusion)
        return st.download()


def op_zoo(n, rng):
    """(name, op, exact?) covering every MatrixOp kind and kernel class."""
    zoo = []
    u1 = rand_unitary(1, rng)
    for q in range(n):
        zoo.append(("dense1_q%d" % q, make_matrix_op([q], u1.reshape(-1)), False))
        zoo.append(("x_q%d" % q, gates.x(q), True))
        zoo.append(("t_q%d" % q, gates.t(q), False))
    zoo.append(("h_q0", gates.h(0), False))
    zoo.append(("y_q1", gates.y(1), False))
    zoo.append(("z_q2", gates.z(2), False))
    zoo.append(("s_q0", gates.s(n - 1), False))
    zoo.append(("rz", gates.rz(1, 0.7123), False))
    for (a, b) in [(0, 1), (1, 0), (0, n - 1), (n - 1, 0), (n - 2, n - 1), (n // 2, 1)]:
        if a == b:
            continue
        zoo.append(("cnot_%d_%d" % (a, b), gates.cnot(a, b), True))
        zoo.append(("cz_%d_%d" % (a, b), gates.cz(a, b), False))
        zoo.append(("cphase_%d_%d" % (a, b), gates.cphase(a, b, math.pi / 8), False))
        zoo.append(("swap_%d_%d" % (a, b), make_swap_op([a], [b]), True))
        zoo.append(("dense2_%d_%d" % (a, b), make_matrix_op([a, b], rand_unitary(2, rng).reshape(-1)), False))
        zoo.append(("cdense1_%d_%d" % (a, b), make_control_op([a], make_matrix_op([b], u1.reshape(-1))), False))
    if n >= 4:
        zoo.append(("toffoli", gates.toffoli(0, n - 1, 2), True))
        zoo.append(("toffoli2", gates.toffoli(3, 1, 0), True))
        zoo.append(("dense3", make_matrix_op([n - 1, 0, 2], rand_unitary(3, rng).reshape(-1)), False))
        zoo.append(("dense3b", make_matrix_op([1, 2, 3], rand_unitary(3, rng).reshape(-1)), False))
        zoo.append(("dense4", make_matrix_op([3, 0, 2, 1], rand_unitary(4, rng).reshape(-1)), False))
        zoo.append(("cdense2", make_control_op([2], make_matrix_op([0, 3], rand_unitary(2, rng).reshape(-1))), False))
        zoo.append(("ccdense1", make_control_op([3, 0], make_matrix_op([1], u1.reshape(-1))), False))
        zoo.append(("swap2", make_swap_op([0, 1], [n - 1, n - 2]), True))
        zoo.append(("fredkin", make_control_op([1], make_swap_op([0], [3])), True))
        zoo.append(("diag2", make_matrix_op([2, 0], np.diag(np.exp(1j * rng.standard_normal(4))).reshape(-1)), False))
        zoo.append(("nested_ctrl", MatrixOp.new_control([0], [1, 2], MatrixOp.new_control([1], [2], gates.x(2))), True))
        perm4 = np.eye(4)[[2, 0, 3, 1]]
        zoo.append(("perm_dense2", make_matrix_op([1, n - 1], perm4.reshape(-1)), True))
    if n >= 6:
        zoo.append(("dense5", make_matrix_op([5, 0, 2, 4, 1], rand_unitary(5, rng).reshape(-1)), False))
        zoo.append(("diag5", make_matrix_op([0, 1, 2, 3, 5], np.diag(np.exp(1j * rng.standard_normal(32))).reshape(-1)), False))
    # sparse ops: X as sparse, a sparse 2-qubit op with duplicate-free rows, controlled sparse
    zoo.append(("sparse_x", MatrixOp.new_sparse([1], [[(1, 1.0)], [(0, 1.0)]]), True))
    zoo.append(("sparse2", MatrixOp.new_sparse([0, n - 1], [[(0, 0.6), (3, 0.8j)], [(1, 1.0)], [(2, -1.0)],
                                                          [(0, 0.8j), (3, 0.6)]]), False))
    zoo.append(("csparse", MatrixOp.new_control([n - 1], [0], MatrixOp.new_sparse([0], [[(1, 1j)], [(0, -1j)]])), False))
    return zoo


# ---------------------------------------------------------------------------------------
# 1. host-buffer drop-ins: bit-identical
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_dropin_apply_op_overwrite_bit_identical(ctx, dtype):
    n = 7
    rng = np.random.default_rng(11)
    psi = rand_state(n, dtype, 1)
    for name, op, _ in op_zoo(n, rng):
        want = np.zeros_like(psi)
        qo.apply_op_overwrite(n, op, psi, want)
        got = np.full_like(psi, 7.0)
        mo.apply_op_overwrite(n, op, psi, got, ctx=ctx)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)) or np.array_equal(got, want), name


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_dropin_apply_op_accumulates(ctx, dtype):
    n = 6
    rng = np.random.default_rng(12)
    psi = rand_state(n, dtype, 2)
    base = rand_state(n, dtype, 3)
    for name, op, _ in op_zoo(n, rng)[::3]:
        want = base.copy()
        qo.apply_op(n, op, psi, want)
        got = base.copy()
        mo.apply_op(n, op, psi, got, ctx=ctx)
        assert np.array_equal(got, want), name


def test_dropin_offsets_and_ragged_windows(ctx):
    """matrix_ops.rs:74-89: shard windows; partners outside the input window read as zero."""
    n = 6
    rng = np.random.default_rng(13)
    full = rand_state(n, np.complex128, 4)
    ops = [gates.h(0), gates.cnot(0, 5), make_matrix_op([1, 4], rand_unitary(2, rng).reshape(-1)),
           make_swap_op([0], [3]), gates.toffoli(0, 1, 2)]
    windows = [(0, 64, 0, 64), (32, 32, 0, 32), (0, 32, 32, 32), (16, 40, 8, 50), (5, 3, 60, 4), (0, 64, 10, 0)]
    for op in ops:
        for (io, il, oo, ol) in windows:
            inp = np.ascontiguousarray(full[io:io + il])
            want = np.zeros(ol, dtype=np.complex128)
            got = np.zeros(ol, dtype=np.complex128)
            qo.apply_op_overwrite(n, op, inp, want, io, oo)
            mo.apply_op_overwrite(n, op, inp, got, io, oo, ctx=ctx)
            assert np.array_equal(got, want), (repr(op), io, il, oo, ol)
    # two shards of the input summed with the accumulating variant == full application
    op = make_matrix_op([0, 3], rand_unitary(2, rng).reshape(-1))
    want = oracle_apply(n, op, full)
    got = np.zeros_like(full)
    mo.apply_op(n, op, np.ascontiguousarray(full[:32]), got, 0, 0, ctx=ctx)
    mo.apply_op(n, op, np.ascontiguousarray(full[32:]), got, 32, 0, ctx=ctx)
    assert np.allclose(got, want, atol=1e-15)


def test_dropin_apply_ops(ctx):
    n = 5
    psi = rand_state(n, np.complex128, 5)
    # [] : overlap copy (matrix_ops.rs:170-183)
    out = np.zeros(20, dtype=np.complex128)
    mo.apply_ops(n, [], np.ascontiguousarray(psi[4:28]), out, 4, 10, ctx=ctx)
    assert np.array_equal(out[:18], psi[10:28]) and np.all(out[18:] == 0)
    # [op] : apply_op
    op = gates.h(2)
    want = np.zeros_like(psi)
    qo.apply_op(n, op, psi, want)
    got = np.zeros_like(psi)
    mo.apply_ops(n, [op], psi, got, ctx=ctx)
    assert np.array_equal(got, want)
    # several ops: the reference's multi-op row iterator (quirk Q5 included), accumulated: bit-identical to the oracle
    ops = [gates.x(0), gates.x(3), gates.h(1)]
    want = psi.copy()
    qo.apply_ops(n, ops, psi, want)
    got = psi.copy()
    mo.apply_ops(n, ops, psi, got, ctx=ctx)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_dropin_apply_ops_multi_op_iterator_bit_identical(ctx, dtype):
    """matrix_ops.rs:184-217 + iterator_mapper.rs:8-31 + qubit_multi_iterator.rs:38-78 on the device (k_multi_gather):
    every op kind in the list, offsets and ragged windows, accumulation into a non-zero output."""
    rng = np.random.default_rng(77)
    n = 9
    psi = rand_state(n, dtype, 8)
    sparse = MatrixOp.new_sparse([6, 7], [[(0, 1.0), (3, 0.5j)], [(1, -1.0)], [(2, 1j), (3, 2.0)], [(0, 0.25)]])
    lists = [
        [gates.h(0), gates.t(4)],
        [gates.cnot(1, 5), make_matrix_op([2, 3], rand_unitary(2, rng).reshape(-1)), gates.x(8)],
        [make_swap_op([0, 1], [4, 5]), sparse, gates.toffoli(2, 3, 8)],
        [make_control_op([7], make_control_op([1], gates.h(3))), make_matrix_op([0, 5, 6], rand_unitary(3, rng).reshape(-1))],
        [gates.h(q) for q in (8, 2, 5, 0)],
        [gates.x(q) for q in range(8)],
    ]
    for ops in lists:
        want = rand_state(n, dtype, 9)
        got = want.copy()
        qo.apply_ops(n, ops, psi, want)
        mo.apply_ops(n, ops, psi, got, ctx=ctx)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), [repr(o) for o in ops]
    # windows: input [40, 400), output [100, 500)
    ops = lists[1]
    want = np.zeros(400, dtype=dtype)
    got = np.zeros(400, dtype=dtype)
    qo.apply_ops(n, ops, np.ascontiguousarray(psi[40:400]), want, 40, 100)
    mo.apply_ops(n, ops, np.ascontiguousarray(psi[40:400]), got, 40, 100, ctx=ctx)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8)) and np.any(got != 0)
    with pytest.raises(Exception):
        mo.apply_ops(n, [gates.x(0)] * 9, psi, got, ctx=ctx)  # more than 8 ops in one sweep: UNSUPPORTED, not a wrong answer


# ---------------------------------------------------------------------------------------
# 2. device-resident state: every op kind x every bit position x both precisions
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 11])
def test_state_single_ops(ctx, dtype, n):
    rng = np.random.default_rng(100 + n)
    psi = rand_state(n, dtype, 6)
    if n < 3:
        zoo = [("x", gates.x(0), True), ("h", gates.h(n - 1), False), ("t", gates.t(0), False)]
        if n == 2:
            zoo += [("cnot", gates.cnot(0, 1), True), ("cnot2", gates.cnot(1, 0), True),
                    ("swap", make_swap_op([0], [1]), True),
                    ("d2", make_matrix_op([1, 0], rand_unitary(2, rng).reshape(-1)), False)]
    else:
        zoo = op_zoo(n, rng) if n <= 7 else op_zoo(n, rng)[::2]
    with State(n, dtype, ctx) as st:
        for name, op, exact in zoo:
            want = oracle_apply(n, op, psi)
            st.upload(psi)
            st.apply_op(op)
            got = st.download()
            if exact:
                assert np.array_equal(got, want), name  # == treats +-0 as equal (quirk Q7)
            else:
                assert_close(got, want, dtype)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_state_every_target_bit_n16(ctx, dtype):
    """One generic dense 1-qubit gate, X, T and CNOT on every bit position of a 16-qubit state."""
    n = 16
    rng = np.random.default_rng(21)
    u1 = rand_unitary(1, rng)
    psi = rand_state(n, dtype, 7)
    with State(n, dtype, ctx) as st:
        for q in range(n):
            for op, exact in [(make_matrix_op([q], u1.reshape(-1)), False), (gates.x(q), True),
                              (gates.t(q), False), (gates.cnot((q + 5) % n, q), True),
                              (gates.cphase(q, (q + 1) % n, 0.3), False)]:
                want = oracle_apply(n, op, psi)
                st.upload(psi)
                st.apply_op(op)
                got = st.download()
                if exact:
                    assert np.array_equal(got, want), (q, repr(op))
                else:
                    assert_close(got, want, dtype)


def test_state_schedule_matches_pipeline(ctx):
    """apply_schedule == the reference's per-entry fold (builder.rs:423-514), fused or not."""
    for n, depth, dtype in [(10, 6, np.complex128), (12, 4, np.complex64), (9, 8, np.complex128)]:
        ops = circuits.random_circuit(n, depth, 0xABC0 + n, "H,T,CNOT")
        ops += circuits.random_circuit(n, 2, 0xABD0 + n, "H,CZ,CNOT")
        want = qo.run_pipeline(n, ops, 0, dtype)
        for fusion in (False, True):
            with State(n, dtype, ctx) as st:
                st.set_basis(0)
                st.apply_schedule(ops, fusion=fusion)
                got = st.download()
            assert_close(got, want, dtype)


def test_permutation_circuit_is_exact(ctx):
    """A circuit of X/CNOT/Toffoli/SWAP only must come out value-exact, fused or not."""
    n = 12
    rng = np.random.default_rng(5)
    ops = []
    for _ in range(60):
        a, b, c = rng.choice(n, 3, replace=False)
        kind = rng.integers(4)
        ops.append([gates.x(int(a)), gates.cnot(int(a), int(b)), gates.toffoli(int(a), int(b), int(c)),
                    make_swap_op([int(a)], [int(b)])][kind])
    psi = rand_state(n, np.complex128, 8)
    want = qo.run_pipeline(n, ops, state=psi)
    for fusion in (False, True):
        got = gpu_state_apply(ctx, n, ops, psi, fusion=fusion)
        assert np.array_equal(got, want)


def test_qft_small(ctx):
    n = 9
    ops = circuits.qft(n)
    psi = circuits.random_state(n, 0x5EED0003, np.complex64)
    want = qo.run_pipeline(n, ops, state=psi, dtype=np.complex64)
    got = gpu_state_apply(ctx, n, ops, psi, fusion=True)
    assert_close(got, want, np.complex64)
    # QFT of a basis state |x> has flat magnitude 2^(-n/2)
    with State(n, np.complex128, ctx) as st:
        st.set_basis(37)
        st.apply_schedule(ops)
        out = st.download()
    assert np.allclose(np.abs(out), 2.0 ** (-n / 2), atol=1e-12)


def test_dense4_blocks_small(ctx):
    n = 10
    ops = circuits.config4(n, blocks=6)
    want = qo.run_pipeline(n, ops, 0, np.complex128)
    with State(n, np.complex128, ctx) as st:
        st.set_basis(0)
        st.apply_schedule(ops)
        got = st.download()
        assert abs(st.norm2() - 1.0) < 1e-12
    assert_close(got, want, np.complex128)


def test_calculate_state_entry(ctx):
    """qipb200_calculate_state == LocalBuilder::calculate_state_with_init for unitary pipelines."""
    import ctypes as C
    from rustqip_b200 import _lib
    from rustqip_b200._abi import QIP_F64, marshal_ops
    n = 8
    ops = circuits.random_circuit(n, 5, 99)
    arr, keep = marshal_ops(ops, QIP_F64)
    out = np.zeros(1 << n, dtype=np.complex128)
    for flags in (_lib.SCHED_DEFAULT, _lib.SCHED_NO_FUSION):
        st = _lib.lib().qipb200_calculate_state(ctx.handle, QIP_F64, n, 5, arr, len(ops), flags, out.ctypes.data)
        _lib.check(st, ctx.handle)
        assert_close(out, qo.run_pipeline(n, ops, 5), np.complex128)


# ---------------------------------------------------------------------------------------
# 3. errors never abort; edge cases
# ---------------------------------------------------------------------------------------
def test_error_codes(ctx):
    with State(3, np.complex128, ctx) as st:
        st.set_basis(0)
        for bad in [MatrixOp.new_matrix([3], [1, 0, 0, 1]),            # index out of range
                    MatrixOp.new_matrix([0, 0], np.eye(4).reshape(-1)),  # repeated index
                    MatrixOp.new_matrix([0], [1, 0, 0]),                # wrong length
                    MatrixOp.new_swap([0], [1, 2]),                     # unequal halves
                    MatrixOp.new_matrix([], [])]:
            with pytest.raises(CircuitError):
                st.apply_op(bad)
        with pytest.raises(CircuitError):
            st.set_basis(8)
        # the state is untouched by the failed calls
        assert st.download()[0] == 1.0


def test_identity_and_empty_schedule(ctx):
    n = 5
    psi = rand_state(n, np.complex128, 9)
    with State(n, np.complex128, ctx) as st:
        st.upload(psi)
        st.apply_schedule([])
        st.apply_op(make_matrix_op([2], [1, 0, 0, 1]))
        st.apply_op(make_matrix_op([4, 0], np.eye(4).reshape(-1)))
        assert np.array_equal(st.download(), psi)
        assert abs(st.norm2() - 1.0) < 1e-12


# ---------------------------------------------------------------------------------------
# 4. measurement kernels vs oracle (qip/src/state_ops/measurement_ops.rs)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_measurement(ctx, dtype):
    n = 10
    psi = rand_state(n, dtype, 10)
    tol = 1e-12 if dtype == np.complex128 else 1e-6
    with State(n, dtype, ctx) as st:
        st.upload(psi)
        for indices in ([0], [9], [3, 7], [7, 3], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9], [5, 0, 9]):
            want = qo.measure_probs(n, indices, psi).astype(np.float64)
            got = st.measure_probs(indices)
            assert np.allclose(got, want, atol=tol), indices
            assert abs(st.measure_prob(1, indices) - want[1]) < tol
        for r in (1e-9, 0.25, 0.5, 0.75, 0.999999):
            for indices in ([0], [2, 8], [9, 0, 4]):
                assert st.soft_measure(indices, r) == qo.soft_measure(n, indices, psi, r)
        indices, m = [1, 6], 2
        p = qo.measure_prob(n, m, indices, psi)
        want = np.zeros_like(psi)
        qo.measure_state(n, indices, m, p, psi, want)
        st.collapse(indices, m, p)
        got = st.download()
        assert np.allclose(got, want, rtol=1e-6 if dtype == np.complex64 else 1e-14, atol=0)
        assert abs(st.norm2() - 1.0) < (1e-5 if dtype == np.complex64 else 1e-12)


def test_measure_state_reference_kat(ctx):
    """measurement_ops.rs:291-335 on the device."""
    with State(2, np.complex128, ctx) as st:
        for m, expect in [(0, [math.sqrt(.5), math.sqrt(.5), 0, 0]), (1, [0, 0, math.sqrt(.5), math.sqrt(.5)])]:
            st.upload(np.array([.5, .5, .5, .5], dtype=np.complex128))
            p = st.measure_prob(m, [0])
            assert abs(p - 0.5) < 1e-15
            st.collapse([0], m, p)
            assert np.allclose(st.download(), expect, atol=1e-10)
        st.upload(np.array([.5, .5, .5, .5], dtype=np.complex128))
        assert list(st.measure_probs([1])) == [0.5, 0.5]


# ---------------------------------------------------------------------------------------
# 5. fused tile passes with several tiles / high tile bits (n > T)
# ---------------------------------------------------------------------------------------
def _mixed_circuit(n, count, seed):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(count):
        q = [int(x) for x in rng.choice(n, 4, replace=False)]
        kind = int(rng.integers(14))
        ops.append([
            lambda: gates.h(q[0]), lambda: gates.t(q[0]), lambda: gates.x(q[0]), lambda: gates.cnot(q[0], q[1]),
            lambda: gates.cz(q[0], q[1]), lambda: gates.cphase(q[0], q[1], 0.37), lambda: gates.rz(q[0], 1.1),
            lambda: make_swap_op([q[0]], [q[1]]), lambda: gates.toffoli(q[0], q[1], q[2]),
            lambda: make_matrix_op([q[0], q[1]], rand_unitary(2, rng).reshape(-1)),
            lambda: make_matrix_op([q[2], q[0], q[1]], rand_unitary(3, rng).reshape(-1)),
            lambda: make_control_op([q[0]], make_matrix_op([q[1]], rand_unitary(1, rng).reshape(-1))),
            lambda: make_control_op([q[3]], make_swap_op([q[0]], [q[1]])),
            lambda: make_matrix_op([q[1], q[3]], np.diag(np.exp(1j * rng.standard_normal(4))).reshape(-1)),
        ][kind]())
    return ops


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [13, 15, 17])
def test_fused_multi_tile(ctx, dtype, n):
    ops = _mixed_circuit(n, 150, 4000 + n)
    psi = rand_state(n, dtype, 12)
    want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    l0 = ctx.kernel_launches()
    got = gpu_state_apply(ctx, n, ops, psi, fusion=True)
    fused_launches = ctx.kernel_launches() - l0
    assert_close(got, want, dtype)
    assert fused_launches < len(ops) // 2  # it really fused
    got_unfused = gpu_state_apply(ctx, n, ops, psi, fusion=False)
    assert_close(got_unfused, want, dtype)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_fused_condition_table_overflow(ctx, dtype):
    """More distinct controls-outside-the-tile conditions in one pass (175) than the 63 slots of the per-CTA
    condition word: the rest must take the in-record test.  Tile = index bits 0-4 and 15-21, controls on 5-14."""
    import itertools
    n = 22
    rng = np.random.default_rng(31)
    ops = []
    for k in (1, 2, 3):
        for ctrls in itertools.combinations(range(7, 17), k):
            tgt = n - 1 - int(rng.integers(3))
            inner = [gates.x(tgt), gates.h(tgt), gates.t(tgt), gates.mat([tgt], rand_unitary(1, rng).reshape(-1))][int(rng.integers(4))]
            ops.append(make_control_op(list(ctrls), inner))
    psi = rand_state(n, dtype, 14)
    want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    l0 = ctx.kernel_launches()
    got = gpu_state_apply(ctx, n, ops, psi, fusion=True)
    assert ctx.kernel_launches() - l0 <= 3
    assert_close(got, want, dtype)


def test_fused_random_circuit_n20(ctx):
    n = 20
    ops = circuits.random_circuit(n, 12, 0x5EED0002, "H,T,CNOT") + circuits.random_circuit(n, 6, 0x5EED0005, "H,CZ,CNOT")
    want = qo.run_pipeline(n, ops, 0, np.complex128)
    with State(n, np.complex128, ctx) as st:
        st.set_basis(0)
        st.apply_schedule(ops, fusion=True)
        got = st.download()
        assert abs(st.norm2() - 1.0) < 1e-11
    assert_close(got, want, np.complex128)


def test_fused_qft_f32_n18(ctx):
    n = 18
    ops = circuits.qft(n)
    psi = circuits.random_state(n, 0x5EED0003, np.complex64)
    want = qo.run_pipeline(n, ops, state=psi, dtype=np.complex64)
    got = gpu_state_apply(ctx, n, ops, psi, fusion=True)
    assert_close(got, want, np.complex64)


def test_fused_permutation_exact_multi_tile(ctx):
    n = 16
    rng = np.random.default_rng(17)
    ops = []
    for _ in range(200):
        a, b, c = [int(x) for x in rng.choice(n, 3, replace=False)]
        ops.append([gates.x(a), gates.cnot(a, b), gates.toffoli(a, b, c), make_swap_op([a], [b])][int(rng.integers(4))])
    for dtype in (np.complex128, np.complex64):
        psi = rand_state(n, dtype, 13)
        want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
        got = gpu_state_apply(ctx, n, ops, psi, fusion=True)
        assert np.array_equal(got, want)
