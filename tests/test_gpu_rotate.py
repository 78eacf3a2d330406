"""Qubit rotation on the device (opt-in QIPB200_ROTATE=1; planner.cpp: plan_rotating, schedule.cu: run_rotating).

The planner, the emitted passes and the generated kernels of this path are validated on the CPU
(tests/test_planner_cpu.py::test_rotating_plan_matches_oracle, tests/test_jit_cpu.py::
test_generated_kernels_run_rotating_plans); the DEVICE path below was written after the round's GPU budget was
spent and has never run on hardware -- hence the non-strict xfail marker: the suite stays green either way and the log
says which way it went."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import numpy as np
from oracle import qip_oracle as qo
from rustqip_b200 import circuits
from rustqip_b200.state import Context, State

with Context(0) as ctx:
    # (1) oracle-sized: two schedules back to back (the second starts from the layout the first left behind),
    #     measurement on the permuted layout, then the download restores the canonical order
    for n, dtype, jit, tol in [(16, np.complex128, "off", 1e-10), (18, np.complex128, "sync", 1e-10), (18, np.complex64, "sync", 1e-5)]:
        import os
        os.environ["QIPB200_JIT"] = jit
        a = circuits.random_circuit(n, 10, 11, "H,T,CNOT")
        b = circuits.qft(n)[:60] + circuits.random_circuit(n, 5, 12, "H,CZ,CNOT")
        want = qo.run_pipeline(n, a + b, 5, dtype)
        with State(n, dtype, ctx) as st:
            st.set_basis(5)
            st.apply_schedule(a)
            st.apply_schedule(b)
            probs = st.measure_probs([0, n - 1, 3])
            got = st.download()
        err = float(np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))))
        assert err <= tol, (n, jit, err)
        assert np.allclose(probs, qo.measure_probs(n, [0, n - 1, 3], want).astype(np.float64), atol=1e-5)
        print("rotate n=%d %s jit=%s: max err %.2e OK" % (n, np.dtype(dtype).name, jit, err))
    # (2) a size no host oracle reaches: rotating vs one in-place sweep per gate, compared on the device
    os.environ["QIPB200_JIT"] = "sync"
    n = 26
    ops = circuits.random_circuit(n, 20, 0x5EED0002, "H,T,CNOT")
    with State(n, np.complex128, ctx) as s1, State(n, np.complex128, ctx) as s2:
        s1.set_basis(0)
        s2.set_basis(0)
        s1.apply_schedule(ops)
        s2.apply_schedule(ops, fusion=False)
        s1.download(offset=0, length=1)  # restores the canonical layout of s1
        d = s1.max_abs_diff(s2)
        assert d <= 1e-10, d
        print("rotate n=26 vs unfused: max |diff| %.2e OK" % d)
"""


@pytest.mark.xfail(reason="opt-in path written after the round's GPU budget was spent: never run on hardware", strict=False)
def test_rotating_schedule_on_device():
    env = dict(os.environ, QIPB200_ROTATE="1", PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-c", _WORKER], cwd=ROOT, env=env, capture_output=True, text=True, timeout=420)
    sys.stdout.write(p.stdout[-3000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
