"""Launched under torchrun by tests/test_gpu_multi.py: every rank owns a slice of the LPs, runs the engine with
NVLink peer scatter + the fused GVT mailbox round, and rank 0 checks the gathered per-LP results against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from aivika_distributed_b200 import dio as D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [
        ("phold optimistic", D.PHOLD(30011, remote_pct=50, seed=3), D.DIOParams(dioTimeHorizon=0.7, maxEventsPerStep=8, queueCapacity=256), 6.0),
        ("phold unbounded", D.PHOLD(4096, population=4, remote_pct=100, seed=9), D.DIOParams(dioTimeHorizon=None), 5.0),
        ("phold skewed", D.PHOLD(20000, remote_pct=50, n_hot=256, hot_pct=20, seed=5), D.DIOParams(dioTimeHorizon=0.5, queueCapacity=512, maxEventsPerStep=16), 4.0),
        ("machrep2", D.MachRep2(257), D.DIOParams(queueCapacity=64, stepsPerLaunch=32), 200.0),
    ]
    ok = True
    for name, model, ps, stop in cases:
        nid = D.share_nccl_id(dist, rank, device="cuda")      # an ncclUniqueId serves exactly one communicator
        dio = D.DIO(model, ps, D.Specs(0.0, stop), rank=rank, world=world, device=local, nccl_id=nid)
        assert (dio.lp_base, dio.lp_base + dio.n_local) == D.partition_range(model, rank, world)
        dio.enqueueMessages(dio.initialEvents())
        st = dio.runEventInStopTime()
        n_exec, h = dio.lpStats()
        n_total = dio.params.n_lps
        per = torch.zeros(2, n_total, dtype=torch.int64, device="cuda")
        per[0, dio.lp_base:dio.lp_base + dio.n_local] = torch.from_numpy(n_exec.astype(np.int64)).cuda()
        per[1, dio.lp_base:dio.lp_base + dio.n_local] = torch.from_numpy(h.view(np.int64)).cuda()
        dist.all_reduce(per)
        cnt = torch.tensor([st.sent_remote, st.rolled_back], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        if rank == 0:
            g_exec = per[0].cpu().numpy().astype(np.uint64); g_hash = per[1].cpu().numpy().view(np.uint64)
            if isinstance(model, D.PHOLD):
                o_exec, o_hash, _ = H.oracle_phold(dio.params)
            else:
                res = H.oracle_machrep(dio.params)
                o_exec = np.array([r.n_exec[k] for r in res for k in range(3)], dtype=np.uint64)
                o_hash = np.array([r.hash[k] for r in res for k in range(3)], dtype=np.uint64)
            good = bool((g_exec == o_exec).all() and (g_hash == o_hash).all())
            print(f"[multi-rank x{world}] {name}: {'OK' if good else 'MISMATCH'} committed={int(g_exec.sum())} "
                  f"remote_msgs={int(cnt[0])} rolled_back={int(cnt[1])} steps={st.steps}", flush=True)
            ok &= good and (world == 1 or int(cnt[0]) > 0 or isinstance(model, D.MachRep2))
        if isinstance(model, D.MachRep2):
            mine = dio.machRepResults()
            lo = (dio.lp_base + 2) // 3
            ores = H.oracle_machrep(dio.params, lo, lo + len(mine))
            good = all(mine[i]["up_time_prop"] == ores[i].up_time_prop and mine[i]["immed_prop"] == ores[i].immed_prop for i in range(len(mine)))
            t = torch.tensor([int(good)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"[multi-rank x{world}] machrep2 statistics per replica: {'OK' if int(t) else 'MISMATCH'}", flush=True)
                ok &= bool(int(t))
        dio.close()
    flag = torch.tensor([int(ok)], device="cuda"); dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
