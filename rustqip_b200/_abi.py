"""ctypes view of include/qip_op.h + marshalling of ``MatrixOp`` into it.

The same bytes are handed to libqipb200 (product) and, in tests, to the CPU
oracle, so both sides see identical inputs.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from .errors import CircuitError
from .ops import MatrixOp

QIP_F32, QIP_F64 = 0, 1
QIP_OP_MATRIX, QIP_OP_SPARSE, QIP_OP_SWAP, QIP_OP_CONTROL = 0, 1, 2, 3


class QipOp(C.Structure):
    pass


QipOp._fields_ = [
    ("kind", C.c_int32),
    ("n_indices", C.c_uint32),
    ("n_control", C.c_uint32),
    ("reserved", C.c_uint32),
    ("n_entries", C.c_uint64),
    ("indices", C.POINTER(C.c_uint64)),
    ("dense", C.c_void_p),
    ("sp_rowptr", C.POINTER(C.c_uint64)),
    ("sp_col", C.POINTER(C.c_uint64)),
    ("sp_val", C.c_void_p),
    ("inner", C.POINTER(QipOp)),
]


def prec_of(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype in (np.dtype(np.complex64), np.dtype(np.float32)):
        return QIP_F32
    if dtype in (np.dtype(np.complex128), np.dtype(np.float64)):
        return QIP_F64
    raise CircuitError("precision must be f32 or f64 (Precision trait, qip/src/types.rs:6-13)")


def cdtype(prec: int):
    return np.complex64 if prec == QIP_F32 else np.complex128


def _u64(values: Sequence[int]) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(list(values), dtype=np.uint64))


def _fill(op: MatrixOp, prec: int, out: QipOp, keep: list) -> None:
    idx = _u64(op.indices())
    keep.append(idx)
    out.n_indices = len(op.indices())
    out.n_control = 0
    out.reserved = 0
    out.n_entries = 0
    out.indices = idx.ctypes.data_as(C.POINTER(C.c_uint64))
    if op.kind == "matrix":
        out.kind = QIP_OP_MATRIX
        d = np.ascontiguousarray(np.asarray(op.data).reshape(-1).astype(cdtype(prec)))
        keep.append(d)
        out.n_entries = d.shape[0]
        out.dense = d.ctypes.data
    elif op.kind == "sparse":
        out.kind = QIP_OP_SPARSE
        rowptr = [0]
        cols: List[int] = []
        vals: List[complex] = []
        for row in op.rows:
            for c, v in row:
                cols.append(c)
                vals.append(v)
            rowptr.append(len(cols))
        rp, cl = _u64(rowptr), _u64(cols)
        vl = np.ascontiguousarray(np.asarray(vals, dtype=cdtype(prec)))
        keep.extend([rp, cl, vl])
        out.n_entries = len(op.rows)
        out.sp_rowptr = rp.ctypes.data_as(C.POINTER(C.c_uint64))
        out.sp_col = cl.ctypes.data_as(C.POINTER(C.c_uint64))
        out.sp_val = vl.ctypes.data
    elif op.kind == "swap":
        out.kind = QIP_OP_SWAP
    elif op.kind == "control":
        out.kind = QIP_OP_CONTROL
        out.n_control = op.n_control
        inner = QipOp()
        keep.append(inner)
        _fill(op.inner, prec, inner, keep)
        out.inner = C.pointer(inner)
    else:  # pragma: no cover
        raise CircuitError("unknown op kind %r" % (op.kind,))


def marshal_op(op: MatrixOp, prec: int) -> Tuple[QipOp, list]:
    """Return (struct, keepalive); the struct borrows numpy buffers held by keepalive."""
    keep: list = []
    out = QipOp()
    _fill(op, prec, out, keep)
    return out, keep


def marshal_ops(ops: Sequence[MatrixOp], prec: int):
    """Array of qip_op for a schedule; returns (array, keepalive)."""
    keep: list = []
    arr = (QipOp * max(1, len(ops)))()
    for i, op in enumerate(ops):
        _fill(op, prec, arr[i], keep)
    return arr, keep
