"""Worker for the multi-GPU parity test: launched by torchrun, one rank per GPU.
Compares the sharded GPU result with the CPU oracle on rank 0 and exits non-zero on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import qip_oracle as qo
    from rustqip_b200 import circuits
    from rustqip_b200.dist import gather_state, init_sharded_state
    from rustqip_b200.state import Context

    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = (world - 1).bit_length()
    ctx = Context(local_rank)
    failures = 0
    cases = [(12, np.complex128, False, "off"), (14, np.complex128, True, "off"), (15, np.complex64, True, "off"),
             (17, np.complex128, True, "off"), (18, np.complex128, True, "sync"), (19, np.complex64, True, "sync")]
    if os.environ.get("QIPB200_PAIRED_SEND"):
        # migrations fused into the last tile pass of an epoch (schedule.cu: paired send) need the generated kernels
        cases = [(18, np.complex128, True, "sync"), (19, np.complex64, True, "sync"), (20, np.complex128, True, "sync"),
                 (18, np.complex128, True, "async")]
    for n, dtype, fusion, jit in cases:
        # "sync": the generated kernels even at these sizes -- passes next to a migration run in two halves that
        # overlap the two halves of the exchange (schedule.cu / exchange_bits_split)
        os.environ["QIPB200_JIT"] = jit
        # gates on rank-held qubits (0..g-1) of every kind + random circuits touching them repeatedly
        ops = circuits.sharded_parity_circuit(n, g)
        st = init_sharded_state(n, dtype, ctx)
        st.set_basis(5)
        st.apply_schedule(ops, fusion=fusion)
        # measurement on the sharded state (collective calls, measurement_ops.rs:115-127,153-176,220-269), taken
        # BEFORE the layout is restored by the download: the histogram must follow the migrated qubits
        norm_all = st.norm2()
        probs = st.measure_probs([0, n - 1, 3])
        p1 = st.measure_prob(1, [0])
        exch = st.exchange_bytes()
        got = gather_state(st)
        draws = (1e-9, 0.3, 0.62, 0.999999)
        samples = [st.soft_measure([0, 2, n - 1], r) for r in draws]
        p10 = st.measure_probs([1, 0])
        mval = int(np.argmax(p10))          # collapse onto the likeliest outcome of qubits (1, 0): never probability 0
        mprob = st.measure_prob(mval, [1, 0])
        st.collapse([1, 0], mval, mprob)
        norm_post = st.norm2()
        post = gather_state(st)
        gathered = [None] * world
        dist.all_gather_object(gathered, (norm_all, probs.tolist(), samples, norm_post, mval, mprob))
        st.free()
        if rank == 0:
            want = qo.run_pipeline(n, ops, 5, dtype)
            tol = 1e-10 if dtype == np.complex128 else 1e-5
            # f32: the oracle (like the reference) accumulates the probabilities in f32, the device in f64
            ptol = 1e-12 if dtype == np.complex128 else 1e-4
            err = float(np.max(np.abs(got.astype(np.complex128) - want.astype(np.complex128))))
            wprobs = qo.measure_probs(n, [0, n - 1, 3], want).astype(np.float64)
            wpost = np.zeros_like(got)
            qo.measure_state(n, [1, 0], mval, mprob, got, wpost)
            checks = {
                "amplitudes": err <= tol,
                "norm": abs(norm_all - 1.0) < 1e-4,
                # every rank received the same collective results
                "ranks_agree": all(g == gathered[0] for g in gathered),
                "measure_probs": bool(np.allclose(probs, wprobs, atol=ptol)),
                "measure_prob": abs(p1 - qo.measure_prob(n, 1, [0], want)) < ptol,
                # sampling: the scan runs over OUR amplitudes (the GPU state), as the reference's would (the device
                # scan accumulates in f64 for either precision: compare with the f64 scan of the same amplitudes)
                "soft_measure": samples == [qo.soft_measure(n, [0, 2, n - 1], got.astype(np.complex128), r) for r in draws],
                "collapse": bool(np.allclose(post, wpost, rtol=1e-5 if dtype == np.complex64 else 1e-13, atol=0)),
                "norm_after_collapse": abs(norm_post - 1.0) < 1e-4,
            }
            ok = all(checks.values())
            if not ok:
                print("  failed checks:", [k for k, v in checks.items() if not v], "samples", samples, flush=True)
            print("n=%d %s fusion=%s jit=%s world=%d: max err %.3e, norm %.12f, probs/samples/collapse checked, exchanged %.1f MiB/rank -> %s" % (
                n, np.dtype(dtype).name, fusion, jit, world, err, norm_all, exch / 2 ** 20, "OK" if ok else "FAIL"), flush=True)
            failures += 0 if ok else 1
    flag = [failures]
    dist.broadcast_object_list(flag, src=0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if flag[0] else 0)


if __name__ == "__main__":
    main()
