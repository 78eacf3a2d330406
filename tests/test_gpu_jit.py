"""Generated (NVRTC) pass kernels on the device: parity with the oracle, bit-equality with the interpreter kernel.

QIPB200_JIT=sync makes the scheduler compile every pass before launching it, so the generated path is the one
that runs even at the small sizes the CPU oracle reaches (by default it only engages from 22 local qubits, and
asynchronously).  Reference semantics: the per-entry fold of apply_op_overwrite
(qip-iterators/src/matrix_ops.rs:127-152, qip/src/builder.rs:423-514)."""
import numpy as np
import pytest

from oracle import qip_oracle as qo
from rustqip_b200 import circuits, gates
from rustqip_b200.ops import make_swap_op
from rustqip_b200.state import State

from test_gpu_parity import _mixed_circuit, assert_close, rand_state

pytestmark = pytest.mark.gpu


def run(ctx, n, ops, psi, monkeypatch, mode):
    monkeypatch.setenv("QIPB200_JIT", mode)
    j0 = ctx.jit_stats()["jit_passes"]
    with State(n, psi.dtype, ctx) as st:
        st.upload(psi)
        st.apply_schedule(ops, fusion=True)
        out = st.download()
    return out, ctx.jit_stats()["jit_passes"] - j0


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [13, 16, 19])
def test_generated_kernels_match_oracle_and_interpreter(ctx, monkeypatch, dtype, n):
    ops = _mixed_circuit(n, 150, 7000 + n) + circuits.random_circuit(n, 6, 0x5EED0002, "H,T,CNOT")
    psi = rand_state(n, dtype, 21)
    want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    got, njit = run(ctx, n, ops, psi, monkeypatch, "sync")
    assert njit >= 1, ctx.jit_stats()["note"]
    assert_close(got, want, dtype)
    ref, n0 = run(ctx, n, ops, psi, monkeypatch, "off")
    assert n0 == 0
    # same arithmetic in the same order on the amplitudes (zero terms dropped, +-1 as add/sub: exact for finite
    # values); only the per-CTA scalar products (global phase, conditional-phase tables) may be contracted
    # differently by the two compilers: agreement to a few ulp instead of the oracle tolerance
    scale = float(np.max(np.abs(ref)))
    assert float(np.max(np.abs(got.astype(np.complex128) - ref.astype(np.complex128)))) <= (1e-14 if dtype == np.complex128 else 2e-6) * scale


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_generated_permutations_exact(ctx, monkeypatch, dtype):
    n = 16
    rng = np.random.default_rng(17)
    ops = []
    for _ in range(200):
        a, b, c = [int(x) for x in rng.choice(n, 3, replace=False)]
        ops.append([gates.x(a), gates.cnot(a, b), gates.toffoli(a, b, c), make_swap_op([a], [b])][int(rng.integers(4))])
    psi = rand_state(n, dtype, 13)
    want = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    got, njit = run(ctx, n, ops, psi, monkeypatch, "sync")
    assert njit >= 1
    assert np.array_equal(got, want)  # register renaming: value-exact (the reference's 1*x is exact for finite x)


def test_generated_qft_f32_n20(ctx, monkeypatch):
    n = 20
    ops = circuits.qft(n)
    psi = circuits.random_state(n, 0x5EED0003, np.complex64)
    want = qo.run_pipeline(n, ops, state=psi, dtype=np.complex64)
    got, njit = run(ctx, n, ops, psi, monkeypatch, "sync")
    assert njit >= 1
    assert_close(got, want, np.complex64)


def test_async_mode_converges_to_generated_kernels(ctx, monkeypatch):
    """async: the first run may use the interpreter while NVRTC works in the background; after jit_stats(wait)
    every pass of the same schedule runs its generated kernel, and the results agree bit for bit."""
    n = 18
    ops = circuits.random_circuit(n, 10, 0x5EED0077, "H,T,CNOT")
    psi = rand_state(n, np.complex128, 3)
    first, _ = run(ctx, n, ops, psi, monkeypatch, "async")
    ctx.jit_stats(wait=True)
    t0 = ctx.jit_stats()
    second, njit = run(ctx, n, ops, psi, monkeypatch, "async")
    t1 = ctx.jit_stats()
    assert njit == t1["tile_passes"] - t0["tile_passes"] and njit >= 1
    assert float(np.max(np.abs(first - second))) <= 1e-14 * float(np.max(np.abs(first)))
