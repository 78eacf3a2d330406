"""Host logic of the fusion planner (rustqip_b200/csrc/planner.cpp), validated without a GPU:
the planner's serialised passes are executed by a CPU emulator of the tile kernel
(tests/native/plan_emulator.cpp) and compared with the oracle's per-entry fold."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import qip_oracle as qo
from rustqip_b200 import circuits, gates
from rustqip_b200._abi import QipOp, marshal_ops, prec_of
from rustqip_b200.ops import MatrixOp, make_control_op, make_matrix_op, make_swap_op

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "native", "_build", "libplan_emul.so")
SRCS = [os.path.join(HERE, "native", "plan_emulator.cpp"),
        os.path.join(ROOT, "rustqip_b200", "csrc", "planner.cpp"),
        os.path.join(ROOT, "rustqip_b200", "csrc", "opcompile.cpp"),
        os.path.join(ROOT, "rustqip_b200", "csrc", "jit_codegen.cpp")]
HDRS = [os.path.join(ROOT, "rustqip_b200", "csrc", "tile.cuh"),
        os.path.join(ROOT, "rustqip_b200", "csrc", "jit_codegen.h"),
        os.path.join(ROOT, "rustqip_b200", "csrc", "opcompile.h"),
        os.path.join(ROOT, "include", "qip_op.h")]


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    stale = not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRCS + HDRS)
    if stale:
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO] + SRCS + ["-ldl"])
    L = C.CDLL(SO)
    L.emul_schedule.restype = C.c_int
    L.emul_schedule.argtypes = [C.c_int, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_void_p, C.c_uint32,
                                C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_char_p, C.c_size_t]
    L.emul_schedule_blocked.restype = C.c_int
    L.emul_schedule_blocked.argtypes = [C.c_int, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_void_p, C.c_uint32,
                                        C.c_uint32, C.c_uint64, C.c_void_p, C.c_char_p, C.c_size_t]
    L.emul_sharded_uniform.restype = C.c_int
    L.emul_sharded_uniform.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_void_p, C.c_uint32,
                                       C.c_uint32, C.c_void_p, C.c_char_p, C.c_size_t]
    L.emul_schedule_rotate.restype = C.c_int
    L.emul_schedule_rotate.argtypes = [C.c_int, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
    L.emul_plan_stats.restype = C.c_int
    L.emul_plan_stats.argtypes = [C.c_int, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_uint32, C.c_uint32,
                                  C.c_int, C.c_uint32, C.c_void_p]
    return L


def run_emul(L, n, ops, psi, dtype=np.complex128, T=0, Lo=0, fuse=True, max_k=0):
    prec = prec_of(dtype)
    arr, keep = marshal_ops(ops, prec)
    st = np.ascontiguousarray(psi.astype(np.complex128))
    stats = np.zeros(64, dtype=np.uint64)
    err = C.create_string_buffer(256)
    rc = L.emul_schedule(prec, n, arr, len(ops), st.ctypes.data, T, Lo, int(fuse), max_k, stats.ctypes.data, err, 256)
    assert rc == 0, err.value
    return st, stats


def rand_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return v / np.linalg.norm(v)


def rand_unitary(k, rng):
    a = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def mixed_circuit(n, count, seed):
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


@pytest.mark.parametrize("n,T,Lo", [(8, 5, 2), (9, 6, 3), (10, 7, 2), (10, 10, 4), (7, 3, 1), (11, 8, 5)])
@pytest.mark.parametrize("fuse", [True, False])
def test_planner_matches_oracle_mixed(emul, n, T, Lo, fuse):
    ops = mixed_circuit(n, 120, 1000 + n + T)
    psi = rand_state(n, 5)
    want = qo.run_pipeline(n, ops, state=psi)
    got, stats = run_emul(emul, n, ops, psi, T=T, Lo=Lo, fuse=fuse)
    assert np.max(np.abs(got - want)) < 1e-12
    assert stats[0] > 0  # passes were actually produced


@pytest.mark.parametrize("knobs", [
    {"QIPB200_SEED_SEARCH": "1"}, {"QIPB200_KEEP_REAL": "1"}, {"QIPB200_NO_HAD": "1"}, {"QIPB200_NO_LOOKBACK": "1"},
    {"QIPB200_NO_FILL": "1"}, {"QIPB200_NO_PEEPHOLE": "1"}, {"QIPB200_NO_PHASEN": "1"}, {"QIPB200_X_MOVES": "1"},
    {"QIPB200_COMPOSE": "3"}, {"QIPB200_SEED_SEARCH": "1", "QIPB200_KEEP_REAL": "1", "QIPB200_NO_FILL": "1"},
    {"QIPB200_PLAN_LOOKAHEAD": "2"}, {"QIPB200_PLAN_LOOKAHEAD": "1", "QIPB200_SEED_SEARCH": "1"},
])
def test_planner_knobs_keep_parity(emul, monkeypatch, knobs):
    """Every planner switch (read from the environment at plan time) must leave the amplitudes alone."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    for n, T, Lo, seed in [(10, 7, 3, 5), (9, 6, 2, 6)]:
        ops = mixed_circuit(n, 100, 2000 + seed) + circuits.random_circuit(n, 6, 90 + seed, "H,T,CNOT") + circuits.qft(n)[:30]
        psi = rand_state(n, seed)
        want = qo.run_pipeline(n, ops, state=psi)
        got, stats = run_emul(emul, n, ops, psi, T=T, Lo=Lo)
        assert np.max(np.abs(got - want)) < 1e-12, knobs


def test_planner_random_and_qft(emul):
    for n, T, Lo in [(10, 6, 3), (12, 8, 3)]:
        ops = circuits.random_circuit(n, 8, 77, "H,T,CNOT") + circuits.random_circuit(n, 4, 78, "H,CZ,CNOT")
        psi = rand_state(n, 6)
        want = qo.run_pipeline(n, ops, state=psi)
        got, stats = run_emul(emul, n, ops, psi, T=T, Lo=Lo)
        assert np.max(np.abs(got - want)) < 1e-12
        assert stats[0] + stats[1] < len(ops) / 3  # far fewer sweeps than gates
    n = 10
    ops = circuits.qft(n)
    psi = rand_state(n, 7)
    got, stats = run_emul(emul, n, ops, psi, T=7, Lo=3)
    assert np.max(np.abs(got - qo.run_pipeline(n, ops, state=psi))) < 1e-12


def many_condition_circuit(n, outside_qubits, seed):
    """Gates on the three lowest index bits under every 1-, 2- and 3-subset of `outside_qubits` as controls:
    92 distinct CTA-uniform conditions for 8 qubits, more than the 63 slots of a pass's condition table."""
    import itertools
    rng = np.random.default_rng(seed)
    ops = []
    for k in (1, 2, 3):
        for ctrls in itertools.combinations(outside_qubits, k):
            tgt = n - 1 - int(rng.integers(3))
            inner = [gates.x(tgt), gates.h(tgt), gates.t(tgt), gates.mat([tgt], rand_unitary(1, rng).reshape(-1))][int(rng.integers(4))]
            ops.append(make_control_op(list(ctrls), inner))
    return ops


def test_planner_condition_table_overflow(emul):
    n, T, Lo = 16, 8, 3
    ops = many_condition_circuit(n, list(range(5, 13)), 21)  # index bits 3..10: outside the tile (bits 0-2, 11-15)
    psi = rand_state(n, 22)
    want = qo.run_pipeline(n, ops, state=psi)
    got, stats = run_emul(emul, n, ops, psi, T=T, Lo=Lo)
    assert np.max(np.abs(got - want)) < 1e-12
    assert stats[0] == 1 and stats[5] == 63 and stats[4] >= 10  # one pass, full table, the rest on the in-record path


@pytest.mark.parametrize("fuse", [True, False])
def test_planner_unnormalised_hadamards(emul, fuse):
    """Hadamards run as add/sub butterflies; the pending 1/sqrt(2)^k must end up in exactly one later gate of the
    pass (a full 2x2, a dense block) or back in the last butterfly -- whatever the rest of the pass looks like."""
    n, T, Lo = 10, 6, 3
    rng = np.random.default_rng(41)
    u2 = rand_unitary(2, rng).reshape(-1)
    cases = {
        "only H": [gates.h(q) for q in range(n)],
        "H then conditional gates only": [gates.h(q) for q in range(n)] + [gates.cnot(0, 9), gates.cz(1, 8), gates.toffoli(2, 3, 7)],
        "H then a dense block": [gates.h(9), gates.h(8), make_matrix_op([8, 9], u2), gates.h(7)],
        "H between full gates": [gates.h(9), gates.mat([9], rand_unitary(1, rng).reshape(-1)), gates.h(9), gates.h(8),
                                 gates.rz(8, 0.3), gates.h(8), gates.x(7), gates.h(7)],
        "H H cancels": [gates.h(5), gates.h(5), gates.h(4), gates.t(4), gates.h(4)],
        "controlled H is not a global scale": [make_control_op([0], gates.h(9)), gates.h(9), make_control_op([8], gates.h(9))],
    }
    for name, ops in cases.items():
        psi = rand_state(n, 42)
        want = qo.run_pipeline(n, ops, state=psi)
        for dtype, tol in ((np.complex128, 1e-12), (np.complex64, 5e-6)):
            got, _ = run_emul(emul, n, ops, psi, dtype=dtype, T=T, Lo=Lo, fuse=fuse)
            assert np.max(np.abs(got - want)) < tol, name


@pytest.mark.parametrize("n,T,Lo,remote", [(10, 6, 3, 0b11 << 8), (11, 7, 3, 0b111 << 8), (9, 6, 2, 1 << 8)])
def test_planner_epochs_with_blocked_ops(emul, n, T, Lo, remote):
    """The epoch loop of a sharded state on one address space: ops that act non-diagonally on the `remote` index
    bits are blocked until 'their' migration; everything that overtakes them must commute with them."""
    for seed in (1, 2, 3):
        ops = mixed_circuit(n, 90, 3000 + seed) + circuits.random_circuit(n, 5, 70 + seed, "H,T,CNOT") + circuits.qft(n)[:25]
        psi = rand_state(n, 30 + seed)
        want = qo.run_pipeline(n, ops, state=psi)
        arr, keep = marshal_ops(ops, prec_of(np.complex128))
        st = np.ascontiguousarray(psi.astype(np.complex128))
        stats = np.zeros(64, dtype=np.uint64)
        err = C.create_string_buffer(256)
        rc = emul.emul_schedule_blocked(prec_of(np.complex128), n, arr, len(ops), st.ctypes.data, T, Lo, remote,
                                        stats.ctypes.data, err, 256)
        assert rc == 0, (rc, err.value)
        assert np.max(np.abs(st - want)) < 1e-12
        assert stats[1] > 1  # it really went through several epochs


def test_planner_f32_data_path(emul):
    n = 9
    ops = mixed_circuit(n, 80, 99)
    psi = rand_state(n, 8)
    want = qo.run_pipeline(n, ops, state=psi)
    got, _ = run_emul(emul, n, ops, psi, dtype=np.complex64, T=6, Lo=3)
    assert np.max(np.abs(got - want)) < 5e-6  # matrices were rounded to f32


def test_planner_permutations_exact(emul):
    n = 10
    rng = np.random.default_rng(3)
    ops = []
    for _ in range(150):
        a, b, c = [int(x) for x in rng.choice(n, 3, replace=False)]
        ops.append([gates.x(a), gates.cnot(a, b), gates.toffoli(a, b, c), make_swap_op([a], [b])][int(rng.integers(4))])
    psi = rand_state(n, 9)
    got, _ = run_emul(emul, n, ops, psi, T=6, Lo=2)
    assert np.array_equal(got, qo.run_pipeline(n, ops, state=psi))


def test_plan_stats_bench_circuits(emul):
    """The bench workloads must fuse: report sweeps per circuit (also printed for DESIGN.md)."""
    out = {}
    for name, n, ops, dtype in [("n30_rand_HTCNOT_d40", 30, circuits.random_circuit(30, 40, 0x5EED0002), np.complex128),
                                ("cfg2_n28", 28, circuits.config2(), np.complex128),
                                ("cfg3_qft30_f32", 30, circuits.qft(30), np.complex64),
                                ("cfg5_local_n30", 30, circuits.random_circuit(30, 30, 0x5EED0005, "H,CZ,CNOT"), np.complex128)]:
        prec = prec_of(dtype)
        arr, keep = marshal_ops(ops, prec)
        stats = np.zeros(64, dtype=np.uint64)
        assert emul.emul_plan_stats(prec, n, arr, len(ops), 0, 0, 1, 0, stats.ctypes.data) == 0
        out[name] = (len(ops), [int(x) for x in stats[:9]])
        sweeps = int(stats[0] + stats[1])
        assert sweeps < len(ops) / 4, (name, sweeps)
    print(out)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,g,T,Lo", [(11, 1, 6, 2), (12, 2, 6, 2), (12, 3, 5, 2), (13, 2, 7, 3)])
def test_sharded_ranks_plan_alike_with_uniform_selection(emul, n, g, T, Lo, dtype):
    """Paired send (schedule.cu) pairs up tiles across GPUs: with the rank-independent selection data (op_uniform_info)
    every virtual rank of a 2^g-way sharded state must plan the same steps from ITS restriction of the ops (controls and
    diagonal bits held by the rank index differ from rank to rank), and executing what each rank emitted for itself on
    its shard must give the oracle's amplitudes -- epochs, migrations and layout restore included."""
    for seed in (1, 2):
        ops = circuits.sharded_parity_circuit(n, g, seed) + circuits.random_circuit(n, 5, 40 + seed, "H,T,CNOT") + circuits.qft(n)[:40]
        psi = rand_state(n, 30 + seed)
        want = qo.run_pipeline(n, ops, state=psi)
        arr, keep = marshal_ops(ops, prec_of(dtype))  # f32: the gate constants are rounded to f32, the emulator computes in f64
        st = np.ascontiguousarray(psi.astype(np.complex128))
        stats = np.zeros(8, dtype=np.uint64)
        err = C.create_string_buffer(256)
        rc = emul.emul_sharded_uniform(prec_of(dtype), n, g, arr, len(ops), st.ctypes.data, T, Lo, stats.ctypes.data, err, 256)
        assert rc == 0, (rc, err.value)
        assert np.max(np.abs(st - want)) < (1e-12 if dtype == np.complex128 else 2e-5)
        assert stats[0] > 0 and stats[2] > 0 and stats[3] > 0  # passes ran, qubits migrated, some ops were rank-dependent


def test_sharded_ranks_plan_differently_without_uniform_selection(emul, monkeypatch):
    """Negative control of the test above: planned from each rank's own restriction of the ops, two ranks of an 8-way
    sharded state choose different steps (return code -10) -- the reason the paired send needs op_uniform_info."""
    monkeypatch.setenv("EMUL_NO_UNIFORM", "1")
    n, g = 12, 3
    seen = set()
    for seed in (1, 2):
        ops = circuits.sharded_parity_circuit(n, g, seed) + circuits.random_circuit(n, 5, 40 + seed, "H,T,CNOT") + circuits.qft(n)[:40]
        arr, keep = marshal_ops(ops, prec_of(np.complex128))
        st = np.ascontiguousarray(rand_state(n, 30 + seed).astype(np.complex128))
        stats = np.zeros(8, dtype=np.uint64)
        err = C.create_string_buffer(256)
        seen.add(emul.emul_sharded_uniform(prec_of(np.complex128), n, g, arr, len(ops), st.ctypes.data, 5, 2, stats.ctypes.data, err, 256))
    assert -10 in seen, seen


def run_rotate(L, n, ops, st, dtype=np.complex128, T=0, Lo=0, restore=True, layout=None, unpermute=False):
    arr, keep = marshal_ops(ops, prec_of(dtype))
    stats = np.zeros(8, dtype=np.uint64)
    err = C.create_string_buffer(256)
    rc = L.emul_schedule_rotate(prec_of(dtype), n, arr, len(ops), st.ctypes.data if st is not None else None, T, Lo, int(restore),
                                layout.ctypes.data if layout is not None else None, int(unpermute), stats.ctypes.data, err, 256)
    assert rc == 0, (rc, err.value)
    return [int(x) for x in stats[:5]]  # passes, single steps, relabelling swaps, restore steps, gates in passes


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,T,Lo", [(10, 7, 3), (9, 6, 2), (12, 7, 3), (11, 6, 3), (12, 8, 2)])
def test_rotating_plan_matches_oracle(emul, n, T, Lo, dtype):
    """Qubit rotation (planner.cpp: plan_rotating, opt-in QIPB200_ROTATE): passes that end with a permutation of their
    own tile bits, every pass emitted under the layout of the moment.  (1) With the inverse permutation inside the plan the
    amplitudes are the oracle's; (2) lazily: a schedule leaves the state in a permuted layout, the next schedule starts
    from it, and the planned restore (what the API runs before a download) brings the canonical order back."""
    seed = n + T
    a = mixed_circuit(n, 120, 2000 + seed) + circuits.random_circuit(n, 12, 90 + seed, "H,T,CNOT")
    b = circuits.qft(n)[:40] + circuits.random_circuit(n, 6, 190 + seed, "H,T,CNOT")
    psi = rand_state(n, seed)
    want = qo.run_pipeline(n, a + b, state=psi)
    tol = 1e-12 if dtype == np.complex128 else 2e-5
    st = np.ascontiguousarray(psi.astype(np.complex128))
    stats = run_rotate(emul, n, a + b, st, dtype, T, Lo, restore=True)
    assert np.max(np.abs(st - want)) < tol and stats[2] > 0
    st = np.ascontiguousarray(psi.astype(np.complex128))
    layout = np.arange(n, dtype=np.uint32)
    run_rotate(emul, n, a, st, dtype, T, Lo, restore=False, layout=layout)
    assert np.any(layout != np.arange(n)) and sorted(layout) == list(range(n))  # a permutation, not the identity
    run_rotate(emul, n, b, st, dtype, T, Lo, restore=False, layout=layout, unpermute=True)
    assert np.all(layout == np.arange(n))
    assert np.max(np.abs(st - want)) < tol


def test_rotating_plan_needs_fewer_sweeps(emul):
    """The point of the rotation: with the state left in the layout of the last pass, the N=30 headline circuit and
    BASELINE configs[1] need clearly fewer HBM sweeps than with the fixed layout (plan only, no amplitudes)."""
    out = {}
    for name, n, ops in [("n30_d40", 30, circuits.random_circuit(30, 40, 0x5EED0002)), ("cfg2_n28", 28, circuits.config2())]:
        arr, keep = marshal_ops(ops, prec_of(np.complex128))
        stats = np.zeros(64, dtype=np.uint64)
        assert emul.emul_plan_stats(prec_of(np.complex128), n, arr, len(ops), 0, 0, 1, 0, stats.ctypes.data) == 0
        fixed = int(stats[0] + stats[1])
        rot = run_rotate(emul, n, ops, None, restore=False)
        lazy = rot[0] + rot[1]
        rot = run_rotate(emul, n, ops, None, restore=True)
        full = rot[0] + rot[1]
        out[name] = (fixed, lazy, full)
        assert lazy <= fixed - 4 and full <= fixed, out
    print(out)
