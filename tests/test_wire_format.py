"""Schedule wire format + state dump/load."""
import os

import numpy as np
import pytest

from oracle import qip_oracle as qo
from rustqip_b200 import circuits, gates, wire
from rustqip_b200.errors import CircuitError
from rustqip_b200.ops import MatrixOp, make_control_op, make_matrix_op, make_swap_op


def _zoo(n):
    rng = np.random.default_rng(2)
    u = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))[0]
    return circuits.random_circuit(n, 3, 5) + circuits.qft(n)[:20] + [
        make_matrix_op([2, 0], u.reshape(-1)), make_swap_op([0, 1], [3, 4]), gates.toffoli(0, 4, 2),
        make_control_op([1], make_swap_op([0], [3])),
        MatrixOp.new_sparse([0, n - 1], [[(0, 0.6), (3, 0.8j)], [(1, 1.0)], [(2, -1.0)], [(0, 0.8j), (3, 0.6)]]),
        MatrixOp.new_control([0], [1, 2], MatrixOp.new_control([1], [2], gates.x(2)))]


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_schedule_round_trip(tmp_path, dtype):
    n = 6
    ops = _zoo(n)
    path = os.path.join(tmp_path, "circuit.qips")
    wire.dump_schedule(path, n, ops, dtype)
    n2, dt2, ops2 = wire.load_schedule(path)
    assert n2 == n and dt2 == dtype and len(ops2) == len(ops)
    assert [repr(a) for a in ops] == [repr(b) for b in ops2]
    psi = np.ascontiguousarray((np.arange(1, 65) / 100.0).astype(dtype))
    assert np.array_equal(qo.run_pipeline(n, ops, state=psi, dtype=dtype), qo.run_pipeline(n, ops2, state=psi, dtype=dtype))


def test_bad_files_are_rejected(tmp_path):
    p = os.path.join(tmp_path, "x.qips")
    open(p, "wb").write(b"nonsense-nonsense-nonsense")
    with pytest.raises(CircuitError):
        wire.load_schedule(p)
    wire.dump_schedule(p, 3, [gates.h(0)])
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-5])
    with pytest.raises(CircuitError, match="truncated"):
        wire.load_schedule(p)


@pytest.mark.gpu
def test_schedule_and_state_files_on_device(ctx, tmp_path):
    from rustqip_b200.state import State
    n = 11
    ops = _zoo(n)
    sched = os.path.join(tmp_path, "c.qips")
    snap = os.path.join(tmp_path, "s.qipa")
    wire.dump_schedule(sched, n, ops)
    _, _, loaded = wire.load_schedule(sched)
    want = qo.run_pipeline(n, ops, 3)
    with State(n, np.complex128, ctx) as st:
        st.set_basis(3)
        st.apply_schedule(loaded[:40])
        wire.dump_state(snap, st)          # checkpoint mid-circuit ...
    with State(n, np.complex128, ctx) as st2:
        wire.load_state(snap, st2)         # ... resume in a fresh state
        st2.apply_schedule(loaded[40:])
        got = st2.download()
    assert np.max(np.abs(got - want)) < 1e-10
    with State(n, np.complex64, ctx) as st3:
        with pytest.raises(CircuitError, match="does not match"):
            wire.load_state(snap, st3)
    # the same file format through the C ABI (qipb200_state_save / _load): byte-identical files, cross-readable
    snap_c = os.path.join(tmp_path, "s_c.qipa")
    with State(n, np.complex128, ctx) as st4:
        st4.set_basis(3)
        st4.apply_schedule(loaded[:40])
        st4.save(snap_c)
    assert open(snap_c, "rb").read() == open(snap, "rb").read()
    with State(n, np.complex128, ctx) as st5:
        st5.load(snap)                     # written by the Python writer, read by the library
        st5.apply_schedule(loaded[40:])
        assert np.max(np.abs(st5.download() - want)) < 1e-10
    with State(n, np.complex64, ctx) as st6:
        with pytest.raises(CircuitError, match="does not match"):
            st6.load(snap_c)
    with State(n, np.complex128, ctx) as st7:
        open(snap_c, "ab").write(b"x")
        with pytest.raises(CircuitError, match="trailing"):
            st7.load(snap_c)


# ---- the same format through the C ABI (qipb200_schedule_parse / _serialise: no GPU needed) ------------
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_c_abi_reader_and_writer_agree_with_the_python_reference(tmp_path, dtype):
    n = 6
    ops = _zoo(n)
    path = os.path.join(tmp_path, "circuit.qips")
    wire.dump_schedule(path, n, ops, dtype)
    data = open(path, "rb").read()
    assert wire.serialise_ops(n, ops, dtype) == data           # C writer == Python writer, byte for byte
    with wire.ParsedSchedule(data) as sched:                    # C reader ...
        assert (sched.n_qubits, sched.n_ops) == (n, len(ops))
        assert sched.serialise() == data                       # ... loses nothing
        # the parsed records are valid ops for an n-qubit state (reference constructor checks, in the library)
        from rustqip_b200 import _lib
        for i in range(sched.n_ops):
            assert _lib.lib().qipb200_validate_op(None, sched.prec, n, sched.ops[i]) == 0
        # and they drive the CPU oracle to the same amplitudes as the Python op tree
        psi = np.ascontiguousarray((np.arange(1, 65) / 100.0).astype(dtype))
        want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
        got = psi.copy()
        for i in range(sched.n_ops):
            got = qo.apply_op_raw(n, sched.ops[i], got, dtype)
        assert np.array_equal(got, want)


def test_c_abi_reader_rejects_malformed_buffers(tmp_path):
    path = os.path.join(tmp_path, "c.qips")
    wire.dump_schedule(path, 5, _zoo(5))
    data = open(path, "rb").read()
    with pytest.raises(CircuitError, match="QIPS"):
        wire.ParsedSchedule(b"nonsense-nonsense-nonsense-nonsense")
    for cut in (3, 23, 30, len(data) // 2, len(data) - 1):
        with pytest.raises(CircuitError, match="truncated"):
            wire.ParsedSchedule(data[:cut])
    with pytest.raises(CircuitError, match="trailing"):
        wire.ParsedSchedule(data + b"\0")
    huge = bytearray(data)
    huge[16:24] = (2 ** 60).to_bytes(8, "little")            # announces 2^60 records
    with pytest.raises(CircuitError, match="truncated"):
        wire.ParsedSchedule(bytes(huge))
    bad_kind = bytearray(data)
    bad_kind[24] = 9
    with pytest.raises(CircuitError, match="unknown op kind"):
        wire.ParsedSchedule(bytes(bad_kind))


@pytest.mark.gpu
def test_c_abi_parsed_schedule_runs_on_device(ctx, tmp_path):
    from rustqip_b200.state import State
    n = 11
    ops = _zoo(n)
    data = wire.serialise_ops(n, ops)
    want = qo.run_pipeline(n, ops, 3)
    with wire.ParsedSchedule(data) as sched, State(n, np.complex128, ctx) as st:
        st.set_basis(3)
        st.apply_marshalled(sched.ops, sched.n_ops)
        got = st.download()
    assert np.max(np.abs(got - want)) < 1e-10


def test_c_abi_reader_survives_mutated_buffers(tmp_path):
    """Byte flips, splices and truncations of a valid schedule: the parser must answer (ok or CircuitError) for each
    of them without reading outside the buffer; whatever it accepts must survive validate + re-serialise."""
    from rustqip_b200 import _lib
    path = os.path.join(tmp_path, "c.qips")
    wire.dump_schedule(path, 5, _zoo(5))
    data = bytearray(open(path, "rb").read())
    rng = np.random.default_rng(7)
    accepted = 0
    for trial in range(1500):
        buf = bytearray(data)
        for _ in range(int(rng.integers(1, 4))):
            kind = int(rng.integers(4))
            pos = int(rng.integers(len(buf)))
            if kind == 0:
                buf[pos] ^= 1 << int(rng.integers(8))
            elif kind == 1:
                buf[pos:pos + 8] = int(rng.integers(0, 2 ** 63)).to_bytes(8, "little")
            elif kind == 2:
                del buf[pos:pos + int(rng.integers(1, 40))]
            else:
                buf = buf[:pos]
        if len(buf) == 0:
            continue
        try:
            with wire.ParsedSchedule(bytes(buf)) as sched:
                accepted += 1
                for i in range(sched.n_ops):
                    _lib.lib().qipb200_validate_op(None, sched.prec, sched.n_qubits, sched.ops[i])  # any status, no crash
                assert sched.serialise() == bytes(buf)
        except CircuitError:
            pass
    assert accepted > 0  # flips inside matrix data leave a well-formed file
