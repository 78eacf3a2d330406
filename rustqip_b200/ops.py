"""Host-side mirror of the reference's gate descriptor ``MatrixOp<P>``.

Reference: qip-iterators/src/iterators/ops.rs:11-91 (enum + constructors) and
qip/src/state_ops/matrix_ops.rs:12-122 (validating ``make_*_op`` constructors,
LittleEndian->BigEndian sparse re-ordering).  Pure host bookkeeping: no
amplitudes are touched here; the objects only describe a gate until
``rustqip_b200._abi`` marshals them into the C ``qip_op`` struct.
"""
from __future__ import annotations

import enum
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .errors import CircuitError


class Representation(enum.Enum):
    """qip/src/types.rs:17-23."""

    LittleEndian = 0
    BigEndian = 1


def flip_bits(n: int, num: int) -> int:
    """qip-iterators/src/utils.rs:22-25."""
    r = 0
    for i in range(n):
        r |= ((num >> i) & 1) << (n - 1 - i)
    return r


class MatrixOp:
    """``enum MatrixOp<P>`` (qip-iterators/src/iterators/ops.rs:11-20).

    kind       payload
    "matrix"   indices, data   (row-major 4^k complex, indices[0] = sub-index MSB)
    "sparse"   indices, rows   (per row a list of (col, value), stored order kept)
    "swap"     m, indices      (indices = a(m) ++ b(m))
    "control"  nc, indices, inner  (indices = controls(nc) ++ inner.indices)
    """

    __slots__ = ("kind", "_indices", "data", "rows", "n_control", "inner", "swap_n")

    def __init__(self, kind, indices, data=None, rows=None, n_control=0, inner=None, swap_n=0):
        self.kind = kind
        self._indices = [int(i) for i in indices]
        self.data = data
        self.rows = rows
        self.n_control = int(n_control)
        self.inner = inner
        self.swap_n = int(swap_n)

    # -- ops.rs:24-46 ---------------------------------------------------------
    def num_indices(self) -> int:
        if self.kind == "swap":
            return self.swap_n * 2
        return len(self._indices)

    def indices(self) -> List[int]:
        return self._indices

    # -- ops.rs:49-91 ---------------------------------------------------------
    @classmethod
    def new_matrix(cls, indices: Sequence[int], data) -> "MatrixOp":
        return cls("matrix", list(indices), data=np.asarray(data).reshape(-1))

    @classmethod
    def new_sparse(cls, indices: Sequence[int], rows) -> "MatrixOp":
        return cls("sparse", list(indices), rows=[[(int(c), complex(v)) for c, v in r] for r in rows])

    @classmethod
    def new_swap(cls, a: Sequence[int], b: Sequence[int]) -> "MatrixOp":
        a = list(a)
        return cls("swap", a + list(b), swap_n=len(a))

    @classmethod
    def new_control(cls, c: Sequence[int], r: Sequence[int], op: "MatrixOp") -> "MatrixOp":
        c = list(c)
        return cls("control", c + list(r), n_control=len(c), inner=op)

    def __repr__(self):  # ops.rs:159-181
        if self.kind == "control":
            return "C(%r)[%s]" % (self.inner, ", ".join(map(str, self._indices[: self.n_control])))
        name = {"matrix": "Matrix", "sparse": "SparseMatrix", "swap": "Swap"}[self.kind]
        return "%s[%s]" % (name, ", ".join(map(str, self._indices)))


# ---- qip/src/state_ops/matrix_ops.rs:12-122 ---------------------------------

def make_matrix_op(indices: Sequence[int], dat) -> MatrixOp:
    """state_ops/matrix_ops.rs:12-27."""
    indices = list(indices)
    n = len(indices)
    dat = np.asarray(dat).reshape(-1)
    if n == 0:
        raise CircuitError("Must supply at least one op index")
    if dat.shape[0] != 1 << (2 * n):
        raise CircuitError(
            "Matrix data has %d entries versus expected 2^2*%d" % (dat.shape[0], n))
    return MatrixOp.new_matrix(indices, dat)


def make_sparse_matrix_op(indices: Sequence[int], dat, order: Representation) -> MatrixOp:
    """state_ops/matrix_ops.rs:32-81 (incl. the LittleEndian re-ordering :62-77)."""
    indices = list(indices)
    n = len(indices)
    if n == 0:
        raise CircuitError("Must supply at least one op index")
    if len(dat) != 1 << n:
        raise CircuitError("Sparse matrix has %d rows versus expected 2^%d" % (len(dat), n))
    for row, v in enumerate(dat):
        if len(v) == 0:
            raise CircuitError("All rows of sparse matrix must have data (%d is empty)" % row)
    if order == Representation.LittleEndian:
        tagged = [(indx, [(flip_bits(n, c), v) for c, v in row]) for indx, row in enumerate(dat)]
        tagged.sort(key=lambda t: flip_bits(n, t[0]))  # stable, like sort_by_key
        dat = [row for _, row in tagged]
    return MatrixOp.new_sparse(indices, dat)


def make_swap_op(a_indices: Sequence[int], b_indices: Sequence[int]) -> MatrixOp:
    """state_ops/matrix_ops.rs:84-100."""
    a_indices, b_indices = list(a_indices), list(b_indices)
    if not a_indices or not b_indices:
        raise CircuitError("Need at least 1 swap index for a and b")
    if len(a_indices) != len(b_indices):
        raise CircuitError(
            "Swap must be performed on two sets of indices of equal length, found %d vs %d"
            % (len(a_indices), len(b_indices)))
    return MatrixOp.new_swap(a_indices, b_indices)


def make_control_op(c_indices: Sequence[int], op: MatrixOp) -> MatrixOp:
    """state_ops/matrix_ops.rs:103-122: nested controls are flattened."""
    c_indices = list(c_indices)
    if not c_indices:
        raise CircuitError("Must supply at least one control index")
    if op.kind == "control":
        return MatrixOp("control", c_indices + op.indices(), n_control=len(c_indices) + op.n_control,
                        inner=op.inner)
    return MatrixOp("control", c_indices + op.indices(), n_control=len(c_indices), inner=op)


def from_reals(data) -> np.ndarray:
    """state_ops/matrix_ops.rs:205-213."""
    return np.asarray(data, dtype=np.float64).astype(np.complex128)
