"""N>1 host logic on CPU: two gloo ranks, the state sharded by its top index bit.

* the library's host-side plan (qipb200_plan_exchanges, no GPU needed) must say "no exchange"
  exactly for the ops whose output shard depends on the local input shard only;
* the reference's own distributed hook -- apply_op with input_offset/output_offset and `+=`
  accumulation over input shards (qip-iterators/src/matrix_ops.rs:74-89,98-123) -- must
  reproduce the single-process result when the shards travel over torch.distributed (gloo).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import qip_oracle as qo
    from rustqip_b200 import circuits, gates
    from rustqip_b200.ops import make_matrix_op, make_swap_op
    from rustqip_b200.state import plan_exchanges

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 8
        g = (world - 1).bit_length()
        shard = 1 << (n - g)
        rng = np.random.default_rng(3)
        u2 = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))[0]
        ops = [gates.h(0), gates.h(3), gates.t(0), gates.cz(0, 5), gates.cnot(0, 4), gates.cnot(4, 0),
               gates.cphase(0, 1, 0.7), make_swap_op([0], [6]), make_swap_op([2], [6]), gates.toffoli(0, 2, 5),
               gates.toffoli(2, 5, 0), make_matrix_op([0, 3], u2.reshape(-1)), make_matrix_op([2, 3], u2.reshape(-1)),
               gates.rz(0, 0.3), gates.x(0)] + circuits.random_circuit(n, 3, 11)
        plan = plan_exchanges(ops, n, world)
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        mine = np.ascontiguousarray(psi[rank * shard:(rank + 1) * shard])
        ok = True
        for i, op in enumerate(ops):
            full_in = qo.run_pipeline(n, ops[:i], state=psi)      # reference state before op i
            want = qo.run_pipeline(n, [op], state=full_in)[rank * shard:(rank + 1) * shard]
            # (a) reference-native distributed application: sum over input shards received via gloo
            local_in = torch.from_numpy(np.ascontiguousarray(full_in[rank * shard:(rank + 1) * shard]).view(np.float64))
            gathered = [torch.empty_like(local_in) for _ in range(world)]
            dist.all_gather(gathered, local_in)
            out = np.zeros(shard, dtype=np.complex128)
            for s_rank in range(world):
                inp = gathered[s_rank].numpy().view(np.complex128)
                qo.apply_op(n, op, inp, out, input_offset=s_rank * shard, output_offset=rank * shard)
            ok &= bool(np.max(np.abs(out - want)) < 1e-14)
            # (b) the plan: needs_exchange == 0  <=>  the local input shard alone suffices
            local_only = np.zeros(shard, dtype=np.complex128)
            qo.apply_op(n, op, gathered[rank].numpy().view(np.complex128), local_only,
                        input_offset=rank * shard, output_offset=rank * shard)
            suffices = bool(np.max(np.abs(local_only - want)) < 1e-14)
            flags = [None] * world
            dist.all_gather_object(flags, suffices)
            if plan[i] == 0:
                ok &= all(flags)        # no exchange planned: every rank must be self-sufficient
            else:
                ok &= not all(flags)    # an exchange was planned: some rank really needs remote data
        q.put((rank, ok, [int(x) for x in plan[:15]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_host_logic_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    plans = {tuple(pl) for _, _, pl in res}
    assert len(plans) == 1  # every rank derives the same exchange plan
    plan = list(plans)[0]
    # h(0) needs one exchange, h(3) none, t(0)/cz/cphase none, cnot(0->4) none (control on the rank bit),
    # cnot(4->0) one, swap(0,6) counts two non-diagonal bits of which one is rank-held
    assert plan[0] == 1 and plan[1] == 0 and plan[2] == 0 and plan[3] == 0 and plan[4] == 0 and plan[5] == 1
