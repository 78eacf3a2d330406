"""Gate -> MatrixOp table of ``LocalBuilder::calculate_state_with_init``.

Reference: qip/src/builder.rs:439-498.  Constants are formed in f64 and cast to
the state's precision, as the reference does with ``P::from(..)``.
"""
from __future__ import annotations

import cmath
import math
from typing import Sequence

import numpy as np

from .ops import MatrixOp, make_control_op, make_matrix_op, make_swap_op

_L, _O, _I = 1 + 0j, 0j, 1j


def x(q: int) -> MatrixOp:                      # builder.rs:440
    return make_matrix_op([q], [_O, _L, _L, _O])


def y(q: int) -> MatrixOp:                      # builder.rs:441-443
    return make_matrix_op([q], [_O, -_I, _I, _O])


def z(q: int) -> MatrixOp:                      # builder.rs:444-446
    return make_matrix_op([q], [_L, _O, _O, -_L])


def h(q: int) -> MatrixOp:                      # builder.rs:447-451
    nl = complex(math.sqrt(0.5))                # FRAC_1_SQRT_2
    return make_matrix_op([q], [nl, nl, nl, -nl])


def s(q: int) -> MatrixOp:                      # builder.rs:452
    return make_matrix_op([q], [_L, _O, _O, _I])


def t(q: int) -> MatrixOp:                      # builder.rs:453-459: from_polar(1, pi/4)
    return make_matrix_op([q], [_L, _O, _O, cmath.rect(1.0, math.pi / 4)])


def cnot(c: int, tq: int) -> MatrixOp:          # builder.rs:460-467
    return make_control_op([c], make_matrix_op([tq], [_O, _L, _L, _O]))


def cz(a: int, b: int) -> MatrixOp:             # SURVEY.md section 8d (config 5): Control(1,[a,b],Matrix([b],Z))
    return make_control_op([a], z(b))


def mat(indices: Sequence[int], data) -> MatrixOp:   # builder.rs:468-470
    return make_matrix_op(list(indices), data)


def swap(a: Sequence[int], b: Sequence[int]) -> MatrixOp:  # builder.rs:471-478
    return make_swap_op(list(a), list(b))


def rz(q: int, theta: float) -> MatrixOp:       # builder.rs:479-496: diag(e^{-i theta/2}, e^{i theta/2})
    ht = theta * 0.5
    return make_matrix_op([q], [cmath.rect(1.0, -ht), _O, _O, cmath.rect(1.0, ht)])


def cphase(c: int, tq: int, theta: float) -> MatrixOp:
    """Controlled phase used by the MatrixOp-level QFT."""
    return make_control_op([c], make_matrix_op([tq], [_L, _O, _O, cmath.rect(1.0, theta)]))


def toffoli(c0: int, c1: int, tq: int) -> MatrixOp:
    return make_control_op([c0, c1], make_matrix_op([tq], [_O, _L, _L, _O]))
