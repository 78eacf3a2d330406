"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the sharded state with
NVLink P2P qubit migration must reproduce the oracle's per-entry fold."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_state_matches_oracle(world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-4000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("standin", ["none", "0", "-1"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_paired_send_matches_oracle(world, standin):
    """QIPB200_PAIRED_SEND=1: the migration that ends an epoch is fused into the epoch's last generated tile pass and done
    in place, tile by tile, under a per-tile flag handshake with the partner's pass (one kernel = compute + NVLink
    transfer).  `standin`: which rank plays the protocol with the stand-alone kernel instead (none / rank 0 / all) --
    every mix must give the oracle's amplitudes."""
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, QIPB200_PAIRED_SEND="1", QIPB200_PAIRED_DEBUG="1")
    if standin != "none":
        env["QIPB200_PAIRED_STANDIN_RANK"] = standin
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29640 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-4000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    fused, alone = p.stderr.count("fused into the pass"), p.stderr.count("stand-in kernel")
    print("paired migrations: %d fused, %d stand-in" % (fused, alone))
    assert fused + alone > 0, p.stderr[-2000:]
    if standin == "none":
        assert fused > 0
    if standin == "-1":
        assert fused == 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_single_process_multi_device_context(world):
    """qipb200_init_multi: ONE process, one context over `world` devices (the shape a Rust B200Builder replacing
    LocalBuilder::calculate_state_with_init has, qip/src/builder.rs:400-519).  The n=17 circuit with every op kind
    on the device-held qubits, fused and unfused, f64 and f32, against the CPU oracle; measurement across devices."""
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import numpy as np
    from oracle import qip_oracle as qo
    from rustqip_b200 import circuits
    from rustqip_b200.state import Context, State

    g = (world - 1).bit_length()
    with Context(list(range(world))) as ctx:
        for n, dtype, fusion in [(13, np.complex128, False), (17, np.complex128, True), (16, np.complex64, True)]:
            ops = circuits.sharded_parity_circuit(n, g)
            want = qo.run_pipeline(n, ops, 5, dtype)
            with State(n, dtype, ctx) as st:
                st.set_basis(5)
                st.apply_schedule(ops, fusion=fusion)
                nrm = st.norm2()
                probs = st.measure_probs([0, n - 1, 3])
                got = st.download()          # the whole 2^n vector, stitched from the shards
                assert got.shape[0] == 1 << n
                m = st.soft_measure([0, 2], 0.4)
            tol = 1e-10 if dtype == np.complex128 else 1e-5
            assert float(np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128)))) <= tol
            assert abs(nrm - 1.0) < 1e-4
            assert np.allclose(probs, qo.measure_probs(n, [0, n - 1, 3], want).astype(np.float64), atol=1e-5)
            assert m == qo.soft_measure(n, [0, 2], got.astype(np.complex128), 0.4)
        # upload of an arbitrary state across the shard boundaries, one gate on a device-held qubit, partial download
        n = 14
        rng = np.random.default_rng(3)
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        from rustqip_b200 import gates
        with State(n, np.complex128, ctx) as st:
            st.upload(psi)
            st.apply_op(gates.h(0))
            want = qo.run_pipeline(n, [gates.h(0)], state=psi)
            lo, ln = (1 << (n - g)) - 100, 300     # a window straddling the first shard boundary
            part = st.download(offset=lo, length=ln)
            assert np.allclose(part, want[lo:lo + ln], atol=1e-12)
        assert ctx.launch_stats()["exchanges"] >= 1


def test_multi_device_host_example():
    """examples/multi_device_host.cpp (the C++ mirror over qipb200_init_multi) on every visible GPU."""
    exe = os.path.join(ROOT, "examples", "multi_device_host")
    if not os.path.exists(exe):
        pytest.skip("examples/multi_device_host not built (python -c 'import __graft_entry__ as g; g.build()')")
    p = subprocess.run([exe], cwd=ROOT, capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout)
    assert p.returncode == 0, p.stdout + p.stderr
