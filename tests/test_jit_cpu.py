"""The per-pass kernel generator (rustqip_b200/csrc/jit_codegen.cpp) validated without a GPU: the generated
source is compiled for the HOST (g++ -DQIP_JIT_HOST) and run over the whole state, pass by pass, by
tests/native/plan_emulator.cpp::emul_schedule_jit; the result must equal the oracle's per-entry fold
(oracle/qip_oracle.c == qip-iterators/src/matrix_ops.rs:127-152 applied entry by entry, builder.rs:423-514).

What this pins: thread -> amplitude maps (every group visited exactly once per super-op), swizzled addressing
(XOR / additive split), register renaming for X / CNOT / SWAP, zero / +-1 folding of gate constants, conditional
(CTA-uniform) ops, conditional-phase factor tables, the global-phase fold into the last super-op, and the layout
of the parameter block."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import qip_oracle as qo
from rustqip_b200 import circuits, gates
from rustqip_b200._abi import QipOp, marshal_ops, prec_of
from rustqip_b200.ops import make_control_op, make_matrix_op, make_swap_op

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "native", "_build", "libplan_emul.so")
CSRC = os.path.join(ROOT, "rustqip_b200", "csrc")
SRCS = [os.path.join(HERE, "native", "plan_emulator.cpp")] + [os.path.join(CSRC, f) for f in
                                                                ("planner.cpp", "opcompile.cpp", "jit_codegen.cpp")]
HDRS = [os.path.join(CSRC, f) for f in ("tile.cuh", "opcompile.h", "jit_codegen.h")] + [os.path.join(ROOT, "include", "qip_op.h")]


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    stale = not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRCS + HDRS)
    if stale:
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO] + SRCS + ["-ldl"])
    L = C.CDLL(SO)
    L.emul_schedule_jit.restype = C.c_int
    L.emul_schedule_jit.argtypes = [C.c_int, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32,
                                    C.c_char_p, C.c_void_p, C.c_char_p, C.c_size_t]
    return L


def run_jit(L, n, ops, psi, dtype):
    prec = prec_of(dtype)
    arr, keep = marshal_ops(ops, prec)
    st = np.ascontiguousarray(psi.astype(dtype))
    stats = np.zeros(8, dtype=np.uint64)
    err = C.create_string_buffer(512)
    with tempfile.TemporaryDirectory() as d:
        rc = L.emul_schedule_jit(prec, n, arr, len(ops), st.ctypes.data, 0, 0, d.encode(), stats.ctypes.data, err, 512)
    assert rc == 0, err.value
    return st, stats


def rand_state(n, seed, dtype):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return (v / np.linalg.norm(v)).astype(dtype)


def rand_unitary(k, rng):
    a = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    return np.linalg.qr(a)[0]


TOL = {np.complex128: 1e-12, np.complex64: 2e-5}


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_generated_random_circuit(emul, dtype):
    n = 15
    ops = circuits.random_circuit(n, 8, 0x5EED0002, "H,T,CNOT") + circuits.random_circuit(n, 3, 0x5EED0005, "H,CZ,CNOT")
    psi = rand_state(n, 1, dtype)
    want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    got, stats = run_jit(emul, n, ops, psi, dtype)
    assert stats[0] >= 1 and stats[1] == 0
    assert np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))) <= TOL[dtype] * np.max(np.abs(want)) * 10
    assert stats[5] > 0          # CNOTs inside a group were renamed, not computed
    assert stats[4] > 0          # and some super-ops needed only a warp-level sync


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_generated_permutations_are_exact(emul, dtype):
    """X / CNOT / Toffoli / SWAP only: register renaming must reproduce the reference's pure moves exactly."""
    n = 14
    rng = np.random.default_rng(7)
    ops = []
    for _ in range(80):
        a, b, c = [int(x) for x in rng.choice(n, 3, replace=False)]
        ops.append([gates.x(a), gates.cnot(a, b), gates.toffoli(a, b, c), make_swap_op([a], [b])][int(rng.integers(4))])
    psi = rand_state(n, 2, dtype)
    want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    got, stats = run_jit(emul, n, ops, psi, dtype)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_generated_mixed_zoo(emul, dtype):
    """Dense 2/3-qubit blocks, controlled gates with controls outside the tile, diagonal gates, conditional X."""
    n = 16
    rng = np.random.default_rng(11)
    ops = []
    for _ in range(120):
        q = [int(x) for x in rng.choice(n, 4, replace=False)]
        kind = int(rng.integers(13))
        ops.append([
            lambda: gates.h(q[0]), lambda: gates.t(q[0]), lambda: gates.x(q[0]), lambda: gates.cnot(q[0], q[1]),
            lambda: gates.cz(q[0], q[1]), lambda: gates.cphase(q[0], q[1], 0.37), lambda: gates.rz(q[0], 1.1),
            lambda: make_swap_op([q[0]], [q[1]]), lambda: gates.toffoli(q[0], q[1], q[2]), lambda: gates.s(q[0]),
            lambda: make_matrix_op([q[0], q[1]], rand_unitary(2, rng).reshape(-1)),
            lambda: make_matrix_op([q[2], q[0], q[1]], rand_unitary(3, rng).reshape(-1)),
            lambda: make_control_op([q[0]], make_matrix_op([q[1]], rand_unitary(1, rng).reshape(-1))),
        ][kind]())
    psi = rand_state(n, 3, dtype)
    want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    got, stats = run_jit(emul, n, ops, psi, dtype)
    assert stats[0] >= 1
    assert np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))) <= TOL[dtype] * np.max(np.abs(want)) * 10


def test_generated_qft_conditional_phase_tables(emul):
    """QFT: runs of controlled phases whose controls lie outside the tile (EC_PHASEN factor tables)."""
    n = 16
    ops = circuits.qft(n)
    psi = rand_state(n, 4, np.complex64)
    want = qo.run_pipeline(n, ops, state=psi, dtype=np.complex64)
    got, stats = run_jit(emul, n, ops, psi, np.complex64)
    assert np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))) <= 2e-5 * np.max(np.abs(want))


def test_nvrtc_compiles_every_pass_of_the_bench_workloads():
    """No GPU needed: the generated source of every pass of the N=30 headline circuit and of the N=30 f32 QFT
    must compile with NVRTC for sm_100a (qipb200_jit_precompile plans, generates, compiles)."""
    from rustqip_b200 import _lib
    L = _lib.lib()
    for n, dtype, ops in [(30, np.complex128, circuits.random_circuit(30, 40, 0x5EED0002, "H,T,CNOT")),
                          (30, np.complex64, circuits.qft(30))]:
        arr, keep = marshal_ops(ops, prec_of(dtype))
        out = (C.c_double * 5)()
        log = C.create_string_buffer(4000)
        rc = L.qipb200_jit_precompile(prec_of(dtype), n, arr, len(ops), out, log, 4000)
        assert rc == 0, log.value
        assert out[0] >= 7 and out[1] == out[0] and out[2] == out[0], (list(out), log.value)


_CACHE_PROBE = r"""
import ctypes as C, json, sys
import numpy as np
from rustqip_b200 import _lib, circuits
from rustqip_b200._abi import marshal_ops, prec_of
L = _lib.lib()
ops = circuits.qft(24)
arr, keep = marshal_ops(ops, prec_of(np.complex64))
out = (C.c_double * 5)()
log = C.create_string_buffer(4000)
rc = L.qipb200_jit_precompile(prec_of(np.complex64), 24, arr, len(ops), out, log, 4000)
print(json.dumps({"rc": rc, "out": list(out), "log": log.value.decode()}))
"""


def test_jit_disk_cache_between_processes(tmp_path):
    """QIPB200_JIT_CACHE_DIR: the second PROCESS planning the same circuit loads every cubin from disk instead of
    calling NVRTC; a corrupted file is ignored and rewritten."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QIPB200_JIT_CACHE_DIR=str(tmp_path), PYTHONPATH=root)

    def probe():
        r = subprocess.run([sys.executable, "-c", _CACHE_PROBE], env=env, cwd=root, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout.strip().splitlines()[-1])

    first = probe()
    assert first["rc"] == 0 and first["out"][2] == first["out"][0] >= 3, first
    assert "disk cache: 0 programs" in first["log"], first
    files = sorted(p for p in os.listdir(tmp_path) if p.endswith(".cubin"))
    assert 1 <= len(files) <= int(first["out"][0]) and not [p for p in os.listdir(tmp_path) if ".tmp." in p]
    second = probe()
    assert second["out"][2] == first["out"][2] and ("disk cache: %d programs" % len(files)) in second["log"], second
    assert second["out"][3] < first["out"][3]
    # a damaged entry must not be trusted
    victim = os.path.join(tmp_path, files[0])
    blob = bytearray(open(victim, "rb").read())
    blob[16] ^= 0xFF  # inside the header's first hash
    open(victim, "wb").write(bytes(blob))
    third = probe()
    assert third["out"][2] == first["out"][2] and ("disk cache: %d programs" % (len(files) - 1)) in third["log"], third
    assert open(victim, "rb").read() != bytes(blob)  # recompiled and rewritten


def test_jit_disk_cache_concurrent_writers(tmp_path):
    """Two processes filling the same cache directory at the same time (temporary name + rename): both succeed, no
    temporary file is left behind, and a third process finds every program on disk."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QIPB200_JIT_CACHE_DIR=str(tmp_path), PYTHONPATH=root)
    procs = [subprocess.Popen([sys.executable, "-c", _CACHE_PROBE], env=env, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for _ in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    res = [json.loads(o[0].strip().splitlines()[-1]) for o in outs]
    assert all(p.returncode == 0 for p in procs), [o[1][-500:] for o in outs]
    assert all(r["rc"] == 0 and r["out"][2] == r["out"][0] for r in res), res
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]
    r = subprocess.run([sys.executable, "-c", _CACHE_PROBE], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    third = json.loads(r.stdout.strip().splitlines()[-1])
    n_files = len([f for f in os.listdir(tmp_path) if f.endswith(".cubin")])
    assert ("disk cache: %d programs" % n_files) in third["log"], third


def test_generated_kernels_run_rotating_plans(emul, monkeypatch):
    """Passes of the rotating planner (planner.cpp: plan_rotating) end with swaps among their tile bits -- register
    renaming in the generated kernels: the host-compiled generated source of every such pass against the oracle."""
    monkeypatch.setenv("EMUL_ROTATE", "1")
    for n, seed, dtype, tol in [(14, 3, np.complex128, 1e-12), (15, 4, np.complex64, 2e-5)]:
        ops = circuits.random_circuit(n, 10, 70 + seed, "H,T,CNOT") + circuits.qft(n)[:30]
        psi = rand_state(n, seed, dtype)
        want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
        got, stats = run_jit(emul, n, ops, psi, dtype)
        assert np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))) <= tol * max(1.0, np.max(np.abs(want)))
        assert stats[0] > 0 and stats[5] > 0  # passes ran through generated code, some ops were pure renaming
