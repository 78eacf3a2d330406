"""Host logic without a GPU: the mirror's constructors, the C-ABI library's exports and its
validation / error reporting (no compute entry is called)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rustqip_b200 import _lib
from rustqip_b200.errors import B200Unavailable, CircuitError
from rustqip_b200.ops import (MatrixOp, Representation, make_control_op, make_matrix_op,
                              make_sparse_matrix_op, make_swap_op)
from rustqip_b200.state import Context, plan_exchanges, validate_op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qipb200.h")).read()
    declared = set(re.findall(r"^(?:int|void|uint64_t|size_t|const char \*|const qip_op \*)\s*\*?(qipb200_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert L.qipb200_abi_version() == 1000


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to start (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(B200Unavailable) as e:
        Context(0)
    assert "no CPU path" in str(e.value)
    # compute entries reject a NULL context instead of computing on the host
    L = _lib.lib()
    out = np.zeros(4, dtype=np.complex128)
    st = L.qipb200_calculate_state(None, 1, 2, 0, None, 0, 0, out.ctypes.data)
    assert st != 0 and b"no CPU path" in L.qipb200_last_error(None)


def test_constructor_validation_mirrors_reference():
    # qip/src/state_ops/matrix_ops.rs:12-122
    with pytest.raises(CircuitError, match="at least one op index"):
        make_matrix_op([], [])
    with pytest.raises(CircuitError, match="entries versus expected"):
        make_matrix_op([0], [1, 0, 0])
    with pytest.raises(CircuitError, match="rows versus expected"):
        make_sparse_matrix_op([0, 1], [[(0, 1)]], Representation.BigEndian)
    with pytest.raises(CircuitError, match="must have data"):
        make_sparse_matrix_op([0], [[(0, 1)], []], Representation.BigEndian)
    with pytest.raises(CircuitError, match="at least 1 swap index"):
        make_swap_op([], [1])
    with pytest.raises(CircuitError, match="equal length"):
        make_swap_op([0], [1, 2])
    with pytest.raises(CircuitError, match="at least one control"):
        make_control_op([], make_matrix_op([0], [1, 0, 0, 1]))
    # nested controls are flattened (state_ops/matrix_ops.rs:112-115)
    inner = make_control_op([1], make_matrix_op([2], [0, 1, 1, 0]))
    op = make_control_op([0], inner)
    assert op.n_control == 2 and op.indices() == [0, 1, 2] and op.inner.kind == "matrix"


@pytest.mark.parametrize("op,n,msg", [
    (MatrixOp.new_matrix([3], [1, 0, 0, 1]), 3, "out of range"),
    (MatrixOp.new_matrix([0, 0], np.eye(4).reshape(-1)), 3, "more than once"),
    (MatrixOp.new_matrix([0], [1, 0, 0]), 3, "entries versus expected"),
    (MatrixOp.new_matrix([], []), 3, "at least one op index"),
    (MatrixOp.new_swap([0], [1, 2]), 3, "equal length"),
    (MatrixOp.new_swap([], []), 3, "swap index"),
    (MatrixOp.new_sparse([0], [[(0, 1.0)], []]), 3, "must have data"),
    (MatrixOp.new_sparse([0], [[(2, 1.0)], [(0, 1.0)]]), 3, "out of range"),
    (MatrixOp.new_control([], [0], MatrixOp.new_matrix([0], [1, 0, 0, 1])), 3, "control"),
    (MatrixOp.new_control([0], [1, 2], MatrixOp.new_matrix([1], [1, 0, 0, 1])), 3, "inner op has"),
])
def test_abi_validation_errors(op, n, msg):
    with pytest.raises(CircuitError, match=msg):
        validate_op(op, n)


def test_abi_validation_accepts_good_ops():
    for op in [make_matrix_op([2, 0], np.eye(4).reshape(-1)), make_swap_op([0, 1], [2, 3]),
               make_control_op([3], make_swap_op([0], [1])),
               make_sparse_matrix_op([0, 1], [[(1, 1)], [(0, 1)], [(3, 1)], [(2, 1)]], Representation.BigEndian)]:
        validate_op(op, 4)
        validate_op(op, 4, np.complex64)


def test_plan_exchanges_counts_rank_held_targets():
    from rustqip_b200 import gates
    ops = [gates.h(0), gates.h(2), gates.t(0), gates.cnot(0, 5), gates.cnot(5, 1), gates.cz(0, 1),
           make_swap_op([0], [1]), make_matrix_op([0, 1], np.eye(4)[::-1].reshape(-1))]
    assert list(plan_exchanges(ops, 8, 4)) == [1, 0, 0, 0, 1, 0, 2, 2]
    assert list(plan_exchanges(ops, 8, 1)) == [0] * 8
