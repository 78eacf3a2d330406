"""rustqip_b200 -- B200-native (sm_100a) state-vector gate application behind
RustQIP's operator API (MatrixOp / apply_op / apply_op_overwrite and the
LocalBuilder::calculate_state_with_init schedule loop).

The compute lives in rustqip_b200/libqipb200.so (CUDA C++ behind a C ABI,
include/qipb200.h); this package is the host-side mirror of the reference
interface plus ctypes plumbing.  There is no CPU fallback.
"""
from .errors import B200Unavailable, CircuitError
from .ops import (MatrixOp, Representation, from_reals, make_control_op, make_matrix_op,
                  make_sparse_matrix_op, make_swap_op)

__all__ = [
    "B200Unavailable", "CircuitError", "MatrixOp", "Representation", "from_reals",
    "make_control_op", "make_matrix_op", "make_sparse_matrix_op", "make_swap_op",
]
