"""N > 1 host path on CPU: two processes over gloo exercise exactly what bench.py does between ranks
(ncclUniqueId hand-off, LP-range partitioning of the start-time messages, MAX/SUM reductions).  The data path
itself has no collective except the GVT mailbox round; its multi-partition addressing is covered
by tests/test_emu.py::test_phold_partitioned."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import helpers as H
    from aivika_distributed_b200 import dio as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nid = D.share_nccl_id(dist, rank)
    model = D.PHOLD(1001, population=3, remote_pct=50, seed=5)
    lo, hi = D.partition_range(model, rank, world)
    p = D.make_params(model, D.DIOParams(), D.Specs(0.0, 1.0), rank=rank, world=world, nccl_id=nid)
    buf = (H.ffi.tw_event * (3 * (hi - lo)))(); n = C.c_size_t()
    assert H.ffi.load().tw_model_initial_events_range(C.byref(p), lo, hi, buf, len(buf), C.byref(n)) == 0
    dsts = sorted(set(buf[i].dst for i in range(n.value)))
    ms, e2e, cnt = D.reduce_run(dist, world, [10.0 + rank, 5.0 - rank], [1.0 + rank], [n.value, rank + 1])
    mlo, mhi = D.partition_range(D.MachRep2(7), rank, world)
    q.put((rank, nid, lo, hi, n.value, dsts[0], dsts[-1], ms, e2e, cnt, bytes(p.nccl_id), mlo, mhi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    assert r0[1] == r1[1] and len(r0[1]) == 128 and r0[1] != bytes(128)      # same unique id on both ranks
    assert r0[10] == r0[1] and r1[10] == r1[1]                                # and it reached tw_params.nccl_id
    assert (r0[2], r0[3], r1[2], r1[3]) == (0, 501, 501, 1001)                # contiguous LP blocks
    assert r0[4] + r1[4] == 3003 and (r0[5], r0[6], r1[5], r1[6]) == (0, 500, 501, 1000)
    assert r0[7] == r1[7] == [11.0, 5.0] and r0[8] == [2.0] and r0[9] == [3003, 3]      # MAX of times, SUM of counters
    assert (r0[11], r0[12], r1[11], r1[12]) == (0, 12, 12, 21)                # MachRep2: replicas are not split
