"""Small cases for compute-sanitizer (one tool per gpurun call): PHOLD with rollbacks, PHOLD with hot LPs (whole-region search
and the cooperative bulk selection), MachRep2 through both kernels -- each checked against the CPU oracle.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from aivika_distributed_b200 import dio as D  # noqa: E402


def phold(p, bulk=None):
    if bulk is not None:
        os.environ["TWB200_BULK"] = str(bulk)
    model = D.PHOLD(p.n_lps, p.phold.population, p.phold.remote_pct, p.phold.lookahead, p.phold.mean, p.phold.seed, p.phold.n_hot, p.phold.hot_pct)
    ps = D.DIOParams(dioTimeHorizon=None if p.time_horizon < 0 else p.time_horizon, queueCapacity=p.queue_capacity,
                     maxEventsPerStep=p.max_events_per_step, procLimit=p.proc_limit)
    dio = D.DIO(model, ps, D.Specs(0.0, p.stop_time))
    dio.runEventInStartTime()
    st = dio.runEventInStopTime()
    n_exec, h = dio.lpStats()
    dio.close()
    os.environ.pop("TWB200_BULK", None)
    o_exec, o_hash, _ = H.oracle_phold(p)
    assert (n_exec == o_exec).all() and (h == o_hash).all()
    return st


st = phold(H.phold_params(2048, 3.0, remote_pct=40, horizon=0.6, batch=6, cap=128))
print("phold optimistic:", st.committed, "committed", st.rolled_back, "rolled back", st.steps, "steps")
for bulk in (0, 1):
    st = phold(H.phold_params(256, 2.0, population=8, remote_pct=80, n_hot=2, hot_pct=60, horizon=1.0, cap=2048, batch=16, proc_limit=512), bulk)
    print(f"phold hot LPs (bulk selection {bulk}):", st.committed, "committed", st.rolled_back, "rolled back", st.steps, "steps")
ores = H.oracle_machrep(H.machrep_params(8, 60.0))
for grouped in (True, False):
    dio = D.DIO(D.MachRep2(8), D.DIOParams(queueCapacity=128, stepsPerLaunch=16, closedGroupKernel=grouped), D.Specs(0.0, 60.0))
    dio.enqueueMessages(dio.initialEvents())
    st = dio.runEventInStopTime()
    res = dio.machRepResults()
    dio.close()
    assert all(res[r]["up_time_prop"] == ores[r].up_time_prop for r in range(8))
    print("machrep2", "k_run_groups" if grouped else "k_step", st.committed, "committed", st.rolled_back, "rolled back", st.steps, "steps")
print("sanitize_case: all results equal the oracle")
