"""ctypes wrapper over oracle/libqip_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module
(see oracle/qip_oracle.h).  Nothing under rustqip_b200/ may.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rustqip_b200._abi import QIP_F32, QIP_F64, QipOp, cdtype, marshal_op, prec_of

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqip_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("qip_oracle.c", "qip_oracle_impl.inc", "qip_oracle.h",
                                             "../include/qip_op.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, p = C.c_uint64, C.c_void_p
        opp = C.POINTER(QipOp)
        for sfx in ("f64", "f32"):
            f = getattr(L, "qo_apply_op_" + sfx)
            f.restype = C.c_int
            f.argtypes = [u64, opp, p, u64, p, u64, u64, u64, C.c_int]
            f = getattr(L, "qo_apply_ops_" + sfx)
            f.restype = C.c_int
            f.argtypes = [u64, opp, u64, p, u64, p, u64, u64, u64]
            f = getattr(L, "qo_multi_op_iterator_" + sfx)
            f.restype = u64
            f.argtypes = [p, p, p, p, u64, p, p, u64]
            real = C.c_double if sfx == "f64" else C.c_float
            g = getattr(L, "qo_prob_magnitude_" + sfx)
            g.restype, g.argtypes = real, [p, u64]
            g = getattr(L, "qo_measure_prob_" + sfx)
            g.restype, g.argtypes = real, [u64, u64, p, u64, p, u64, u64]
            g = getattr(L, "qo_measure_probs_" + sfx)
            g.restype, g.argtypes = None, [u64, p, u64, p, u64, u64, p]
            g = getattr(L, "qo_soft_measure_" + sfx)
            g.restype, g.argtypes = u64, [u64, p, u64, p, u64, u64, C.c_double]
            g = getattr(L, "qo_measure_state_" + sfx)
            g.restype, g.argtypes = None, [u64, p, u64, u64, real, p, u64, p, u64, u64, u64]
        L.qo_full_to_sub.restype, L.qo_full_to_sub.argtypes = u64, [u64, p, u64, u64]
        L.qo_sub_to_full.restype, L.qo_sub_to_full.argtypes = u64, [u64, p, u64, u64, u64]
        L.qo_get_flat_index.restype, L.qo_get_flat_index.argtypes = u64, [u64, u64, u64]
        L.qo_flip_bits.restype, L.qo_flip_bits.argtypes = u64, [u64, u64]
        L.qo_set_bit.restype, L.qo_set_bit.argtypes = u64, [u64, u64, C.c_int]
        L.qo_get_bit.restype, L.qo_get_bit.argtypes = C.c_int, [u64, u64]
        L.qo_entwine_bits.restype, L.qo_entwine_bits.argtypes = u64, [u64, u64, u64, u64]
        L.qo_extract_bits.restype, L.qo_extract_bits.argtypes = u64, [u64, p, u64]
        L.qo_row_entries.restype = u64
        L.qo_row_entries.argtypes = [opp, C.c_int, u64, p, p, u64]
        L.qo_max_threads.restype = C.c_int
        L.qo_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _u64arr(v):
    return np.ascontiguousarray(np.asarray(list(v), dtype=np.uint64))


# ---- bit helpers ------------------------------------------------------------
def full_to_sub(n, mat_indices, full_index):
    a = _u64arr(mat_indices)
    return int(lib().qo_full_to_sub(n, a.ctypes.data, len(a), full_index))


def sub_to_full(n, mat_indices, sub_index, base):
    a = _u64arr(mat_indices)
    return int(lib().qo_sub_to_full(n, a.ctypes.data, len(a), sub_index, base))


def get_flat_index(nindices, i, j):
    return int(lib().qo_get_flat_index(nindices, i, j))


def flip_bits(n, num):
    return int(lib().qo_flip_bits(n, num))


def set_bit(num, bit_index, value):
    return int(lib().qo_set_bit(num, bit_index, 1 if value else 0))


def get_bit(num, bit_index):
    return bool(lib().qo_get_bit(num, bit_index))


def entwine_bits(n, selector, off_bits, on_bits):
    return int(lib().qo_entwine_bits(n, selector, off_bits, on_bits))


def extract_bits(num, indices):
    a = _u64arr(indices)
    return int(lib().qo_extract_bits(num, a.ctypes.data, len(a)))


def row_entries(op, row, dtype=np.complex128):
    prec = prec_of(dtype)
    cop, keep = marshal_op(op, prec)
    cap = 1 << min(op.num_indices(), 20)
    cols = np.zeros(cap, dtype=np.uint64)
    vals = np.zeros(cap, dtype=np.complex128)
    cnt = int(lib().qo_row_entries(C.byref(cop), prec, row, cols.ctypes.data, vals.ctypes.data, cap))
    return [(int(cols[i]), complex(vals[i])) for i in range(min(cnt, cap))]


# ---- the hot path -----------------------------------------------------------
def _apply(n, op, inp, out, input_offset, output_offset, accumulate):
    prec = prec_of(inp.dtype)
    assert inp.dtype == cdtype(prec) and out.dtype == inp.dtype
    assert inp.flags.c_contiguous and out.flags.c_contiguous
    cop, keep = marshal_op(op, prec)
    f = lib().qo_apply_op_f64 if prec == QIP_F64 else lib().qo_apply_op_f32
    rc = f(n, C.byref(cop), inp.ctypes.data, inp.shape[0], out.ctypes.data, out.shape[0],
           input_offset, output_offset, 1 if accumulate else 0)
    if rc != 0:
        raise ValueError("oracle: malformed op")


def apply_op(n, op, inp, out, input_offset=0, output_offset=0):
    """qip_iterators::matrix_ops::apply_op (accumulating), matrix_ops.rs:98-123."""
    _apply(n, op, inp, out, input_offset, output_offset, True)


def apply_op_overwrite(n, op, inp, out, input_offset=0, output_offset=0):
    """qip_iterators::matrix_ops::apply_op_overwrite, matrix_ops.rs:127-152."""
    _apply(n, op, inp, out, input_offset, output_offset, False)


def apply_ops(n, ops, inp, out, input_offset=0, output_offset=0):
    """qip_iterators::matrix_ops::apply_ops (matrix_ops.rs:158-219), accumulating into `out`."""
    from rustqip_b200._abi import marshal_ops
    prec = prec_of(inp.dtype)
    assert inp.dtype == cdtype(prec) and out.dtype == inp.dtype
    arr, keep = marshal_ops(ops, prec)
    f = lib().qo_apply_ops_f64 if prec == QIP_F64 else lib().qo_apply_ops_f32
    if f(n, arr, len(ops), inp.ctypes.data, inp.shape[0], out.ctypes.data, out.shape[0], input_offset, output_offset) != 0:
        raise ValueError("oracle: malformed op list")


def multi_op_iterator(ns, lists, dtype=np.complex128):
    """MultiOpIterator::new(ns, lists).collect() (qubit_multi_iterator.rs:13-79); lists[i] = [(col, val), ...]."""
    prec = prec_of(dtype)
    cols = [np.ascontiguousarray(np.array([c for c, _ in l], dtype=np.uint64)) for l in lists]
    vals = [np.ascontiguousarray(np.array([v for _, v in l], dtype=dtype)) for l in lists]
    cp = (C.c_void_p * len(lists))(*[c.ctypes.data for c in cols])
    vp = (C.c_void_p * len(lists))(*[v.ctypes.data for v in vals])
    lens = _u64arr([len(l) for l in lists])
    nsa = _u64arr(ns)
    cap = 1
    for l in lists:
        cap *= max(len(l), 1)
    oc = np.zeros(cap, dtype=np.uint64)
    ov = np.zeros(cap, dtype=dtype)
    f = lib().qo_multi_op_iterator_f64 if prec == QIP_F64 else lib().qo_multi_op_iterator_f32
    cnt = int(f(nsa.ctypes.data, C.cast(cp, C.c_void_p), C.cast(vp, C.c_void_p), lens.ctypes.data, len(lists),
                oc.ctypes.data, ov.ctypes.data, cap))
    return [(int(oc[i]), complex(ov[i])) for i in range(min(cnt, cap))]


def apply_op_raw(n, cop, state, dtype=np.complex128):
    """apply_op_overwrite for an already marshalled `qip_op` record (e.g. one owned by the product library's
    schedule parser): returns the new state."""
    prec = prec_of(dtype)
    state = np.ascontiguousarray(state.astype(dtype, copy=False))
    out = np.zeros_like(state)
    f = lib().qo_apply_op_f64 if prec == QIP_F64 else lib().qo_apply_op_f32
    if f(n, C.byref(cop), state.ctypes.data, state.shape[0], out.ctypes.data, out.shape[0], 0, 0, 0) != 0:
        raise ValueError("oracle: malformed op")
    return out


def run_pipeline(n, ops, init_index=0, dtype=np.complex128, state=None):
    """The unitary part of LocalBuilder::calculate_state_with_init (builder.rs:406-514):
    state = e_init; arena = 0; per entry: apply_op_overwrite(state -> arena); swap."""
    if state is None:
        state = np.zeros(1 << n, dtype=dtype)
        state[init_index] = 1
    else:
        state = np.ascontiguousarray(state.astype(dtype, copy=True))
    arena = np.zeros_like(state)
    for op in ops:
        apply_op_overwrite(n, op, state, arena)
        state, arena = arena, state
    return state


def make_op_matrix(n, op, dtype=np.complex128):
    """Test helper of the reference (matrix_ops.rs:229-255): column i = op applied to e_i
    with the accumulating apply_op; returned as out[row, col] (after reversed_axes)."""
    m = np.zeros((1 << n, 1 << n), dtype=dtype)
    for i in range(1 << n):
        inp = np.zeros(1 << n, dtype=dtype)
        out = np.zeros(1 << n, dtype=dtype)
        inp[i] = 1
        apply_op(n, op, inp, out)
        m[:, i] = out
    return m


# ---- measurement ------------------------------------------------------------
def _sfx(a):
    return "f64" if a.dtype == np.complex128 else "f32"


def prob_magnitude(inp):
    return float(getattr(lib(), "qo_prob_magnitude_" + _sfx(inp))(inp.ctypes.data, inp.shape[0]))


def measure_prob(n, measured, indices, inp, input_offset=None):
    a = _u64arr(indices)
    return float(getattr(lib(), "qo_measure_prob_" + _sfx(inp))(
        n, measured, a.ctypes.data, len(a), inp.ctypes.data, inp.shape[0], input_offset or 0))


def measure_probs(n, indices, inp, input_offset=None):
    a = _u64arr(indices)
    out = np.zeros(1 << len(a), dtype=np.float64 if inp.dtype == np.complex128 else np.float32)
    getattr(lib(), "qo_measure_probs_" + _sfx(inp))(
        n, a.ctypes.data, len(a), inp.ctypes.data, inp.shape[0], input_offset or 0, out.ctypes.data)
    return out


def soft_measure(n, indices, inp, r, input_offset=None):
    a = _u64arr(indices)
    return int(getattr(lib(), "qo_soft_measure_" + _sfx(inp))(
        n, a.ctypes.data, len(a), inp.ctypes.data, inp.shape[0], input_offset or 0, float(r)))


def measure_state(n, indices, measured, prob, inp, out, offsets=None):
    a = _u64arr(indices)
    io, oo = offsets or (0, 0)
    getattr(lib(), "qo_measure_state_" + _sfx(inp))(
        n, a.ctypes.data, len(a), measured, prob, inp.ctypes.data, inp.shape[0], out.ctypes.data,
        out.shape[0], io, oo)


def max_threads():
    return int(lib().qo_max_threads())


def set_threads(n):
    lib().qo_set_threads(int(n))
