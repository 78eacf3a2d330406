"""BASELINE.json configs[0]: the README CSWAP circuit (7 qubits, f64).

CPU part: the host mirror builds the same 192-entry pipeline as LocalBuilder and the oracle
reproduces the known answer derived in SURVEY.md section 8 (Q-KA).  GPU part: the same
through calculate_state_with_init on the device, measurement included."""
import json
import math
import os

import numpy as np
import pytest

from oracle import qip_oracle as qo
from rustqip_b200.builder import readme_cswap_circuit

KA = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cswap_readme.json")))


def _expected_state():
    v = np.zeros(128, dtype=np.complex128)
    for k, (re, im) in KA["amplitudes"].items():
        v[int(k)] = complex(re, im)
    return v


def test_pipeline_and_oracle_known_answer():
    b, q, ra, rb, handle = readme_cswap_circuit()
    assert b.n() == KA["n"] and b.pipeline_depth() == KA["pipeline_depth"]
    assert len(b.unitary_ops()) == KA["unitary_entries"]
    idx = b.initial_index([(ra, 0b000), (rb, 0b001)])
    assert idx == KA["init_index"]
    st = qo.run_pipeline(7, b.unitary_ops(), idx)
    assert np.max(np.abs(st - _expected_state())) < 1e-12
    # measurement of q (qubit 0): P(0) = P(1) = 1/2 and the collapsed states of Q-KA
    assert abs(qo.measure_prob(7, 0, [0], st) - 0.5) < 1e-12
    for m, key in [(0, "post_measure_0"), (1, "post_measure_1")]:
        out = np.zeros_like(st)
        qo.measure_state(7, [0], m, qo.measure_prob(7, m, [0], st), st, out)
        want = np.zeros(128, dtype=np.complex128)
        for k, a in KA[key].items():
            want[int(k)] = a
        assert np.max(np.abs(out - want)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("fusion", [False, True])
def test_cswap_on_device(ctx, fusion):
    b, q, ra, rb, handle = readme_cswap_circuit()
    for m, key in [(0, "post_measure_0"), (1, "post_measure_1")]:
        state, meas = b.calculate_state_with_init([(ra, 0b000), (rb, 0b001)], ctx=ctx, fusion=fusion, measured=[m])
        assert meas[handle][0] == m and abs(meas[handle][1] - 0.5) < 1e-12
        want = np.zeros(128, dtype=np.complex128)
        for k, a in KA[key].items():
            want[int(k)] = a
        assert np.max(np.abs(state - want)) < 1e-10
    # drawn outcome: either branch, probability one half
    state, meas = b.calculate_state_with_init([(ra, 0b000), (rb, 0b001)], ctx=ctx, fusion=fusion,
                                              rng=np.random.default_rng(1))
    assert meas[handle][0] in (0, 1) and abs(meas[handle][1] - 0.5) < 1e-12
    assert abs(np.linalg.norm(state) - 1.0) < 1e-12


@pytest.mark.gpu
def test_oracle_vectors_on_device(ctx):
    """ORACLE-generated fixtures (tests/golden/make_golden.py) replayed on the GPU."""
    from rustqip_b200 import circuits
    from rustqip_b200.state import State
    vec = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
    cases = [("rand_htcnot_n10_f64", 10, circuits.random_circuit(10, 6, 0x5EED0002), 1e-10),
             ("rand_hczcnot_n9_f32", 9, circuits.random_circuit(9, 5, 0x5EED0005, "H,CZ,CNOT"), 1e-5),
             ("qft_n8_f32", 8, circuits.qft(8), 1e-5),
             ("dense4_n8_f64", 8, circuits.config4(8, blocks=4), 1e-10)]
    for name, n, ops, tol in cases:
        psi, want = vec[name + "_in"], vec[name + "_out"]
        for fusion in (False, True):
            with State(n, psi.dtype, ctx) as st:
                st.upload(psi)
                st.apply_schedule(ops, fusion=fusion)
                got = st.download()
            assert np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))) <= tol * max(1.0, np.max(np.abs(want))), name


@pytest.mark.gpu
def test_cpp_host_mirror_example():
    """examples/cswap_host.cpp (include/qipb200.hpp over the C ABI) runs the CSWAP on the device."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "cswap_host")
    if not os.path.exists(exe):
        pytest.skip("examples/cswap_host not built (run __graft_entry__.build())")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "P(q=0)=0.500" in p.stdout and "amp[4] = +0.500" in p.stdout and "amp[96] = -0.500" in p.stdout
