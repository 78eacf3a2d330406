"""Drop-ins for ``qip_iterators::matrix_ops`` on HOST buffers, executed on the B200.

Same names, argument order and meaning as the reference
(qip-iterators/src/matrix_ops.rs:98-219); ``input``/``output`` are numpy
complex64/complex128 vectors standing in for ``&[Complex<P>]`` /
``&mut [Complex<P>]``.  Each call copies the buffers to the device, runs the
row kernel and copies ``output`` back -- use ``rustqip_b200.State`` to keep the
amplitudes resident between gates.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._abi import cdtype, marshal_op, marshal_ops, prec_of
from .ops import MatrixOp
from .state import Context, default_context


def _bufs(inp: np.ndarray, out: np.ndarray):
    prec = prec_of(inp.dtype)
    if inp.dtype != cdtype(prec) or out.dtype != inp.dtype:
        raise TypeError("input/output must both be complex64 or both complex128")
    if not (inp.flags.c_contiguous and out.flags.c_contiguous and out.flags.writeable):
        raise TypeError("input/output must be contiguous (and output writeable)")
    return prec


def apply_op(n: int, op: MatrixOp, input: np.ndarray, output: np.ndarray, input_offset: int = 0,
             output_offset: int = 0, ctx: Optional[Context] = None) -> None:
    """matrix_ops.rs:98-123: ``output[o] += row(output_offset+o) . input``."""
    ctx = ctx or default_context()
    prec = _bufs(input, output)
    cop, keep = marshal_op(op, prec)
    _lib.check(_lib.lib().qipb200_apply_op(ctx.handle, prec, n, C.byref(cop), input.ctypes.data,
                                           input.shape[0], output.ctypes.data, output.shape[0],
                                           input_offset, output_offset), ctx.handle)


def apply_op_overwrite(n: int, op: MatrixOp, input: np.ndarray, output: np.ndarray,
                       input_offset: int = 0, output_offset: int = 0,
                       ctx: Optional[Context] = None) -> None:
    """matrix_ops.rs:127-152: same with ``=``."""
    ctx = ctx or default_context()
    prec = _bufs(input, output)
    cop, keep = marshal_op(op, prec)
    _lib.check(_lib.lib().qipb200_apply_op_overwrite(ctx.handle, prec, n, C.byref(cop),
                                                     input.ctypes.data, input.shape[0],
                                                     output.ctypes.data, output.shape[0],
                                                     input_offset, output_offset), ctx.handle)


def apply_ops(n: int, ops: Sequence[MatrixOp], input: np.ndarray, output: np.ndarray,
              input_offset: int = 0, output_offset: int = 0, ctx: Optional[Context] = None) -> None:
    """matrix_ops.rs:158-219 ([] = overlap copy, [op] = apply_op, else the reference's multi-op row iterator,
    quirk Q5 included; accumulates into `output`)."""
    ctx = ctx or default_context()
    prec = _bufs(input, output)
    arr, keep = marshal_ops(ops, prec)
    _lib.check(_lib.lib().qipb200_apply_ops(ctx.handle, prec, n, arr, len(ops), input.ctypes.data,
                                            input.shape[0], output.ctypes.data, output.shape[0],
                                            input_offset, output_offset), ctx.handle)


# host-only index helpers of the same module (matrix_ops.rs:12-35) -----------------
def full_to_sub(n: int, mat_indices: Sequence[int], full_index: int) -> int:
    k = len(mat_indices)
    acc = 0
    for j, indx in enumerate(mat_indices):
        acc |= ((full_index >> (n - 1 - indx)) & 1) << (k - 1 - j)
    return acc


def sub_to_full(n: int, mat_indices: Sequence[int], sub_index: int, base: int) -> int:
    k = len(mat_indices)
    acc = base
    for j, indx in enumerate(mat_indices):
        bit = (sub_index >> (k - 1 - j)) & 1
        pos = n - 1 - indx
        acc = (acc | (1 << pos)) if bit else (acc & ~(1 << pos))
    return acc


def get_index(op: MatrixOp, i: int) -> int:
    return op.indices()[i]
