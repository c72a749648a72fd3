"""Parity tests proper: the CUDA engine, called through the C ABI (libtwb200.so), against the CPU oracle.

Bit-exact bar: per-LP committed-event counts and order-dependent trace hashes (integer/index results) and the
MachRep2 floating-point statistics (north_star allows 1e-9 relative; we require equality of the bits, which is
stricter).  Full-size configurations are checked through size-independent properties: the committed trace must
not depend on the optimism window, the batch size or the run, and a window equal to the lookahead never rolls back."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def D():
    from aivika_distributed_b200 import dio
    return dio


def run_phold(H, p):
    d = D()
    model = d.PHOLD(p.n_lps, p.phold.population, p.phold.remote_pct, p.phold.lookahead, p.phold.mean, p.phold.seed,
                    p.phold.n_hot, p.phold.hot_pct)
    ps = d.DIOParams(dioTimeHorizon=None if p.time_horizon < 0 else p.time_horizon, queueCapacity=p.queue_capacity,
                     maxEventsPerStep=p.max_events_per_step, stepsPerLaunch=p.steps_per_launch, procLimit=p.proc_limit,
                     fossilThreshold=p.fossil_threshold)
    dio, sim = d.runDIO(model, ps, d.timeServer(p.n_lps), d.Specs(0.0, p.stop_time))
    try:
        return sim()
    finally:
        dio.close()


def check_phold(H, p):
    o_exec, o_hash, o_sum = H.oracle_phold(p)
    st, n_exec, h = run_phold(H, p)
    assert (n_exec == o_exec).all() and (h == o_hash).all()
    assert st.committed == o_sum.committed and st.trace_hash_sum == o_sum.hash_sum
    assert st.executed - st.rolled_back == st.committed
    assert st.gvt > p.stop_time
    return st


@pytest.mark.parametrize("horizon", [0.1, 0.5, 2.0, -1.0])
def test_phold_matches_oracle_for_every_horizon(H, horizon):
    st = check_phold(H, H.phold_params(20000, 10.0, remote_pct=30, horizon=horizon))
    if horizon == 0.1:
        assert st.rolled_back == 0
    else:
        assert st.rolled_back > 0 and 0 <= st.anti_sent - st.annihilated <= 64     # a few antis may still be in flight at the end


@pytest.mark.parametrize("batch", [1, 2, 4, 6, 8, 16, 64])
def test_phold_batch_sizes(H, batch):
    check_phold(H, H.phold_params(5000, 8.0, remote_pct=50, horizon=1.0, batch=batch, cap=128 if batch <= 4 else 512))


@pytest.mark.parametrize("batch", [2, 8, 64])
def test_phold_long_pending_regions(H, batch):
    """Hot LPs with far more than 64 pending events: selection by the warp-cooperative search over the whole region
    (no window mask), with rollbacks, and with every receive time tied (mean 0)."""
    st = check_phold(H, H.phold_params(128, 3.0, population=8, remote_pct=80, n_hot=2, hot_pct=60, horizon=1.0, cap=2048,
                                       batch=batch, proc_limit=512))
    assert st.rolled_back > 0
    check_phold(H, H.phold_params(96, 4.0, population=8, remote_pct=90, n_hot=1, hot_pct=100, lookahead=0.5, mean=0.0,
                                  horizon=-1.0, cap=2048, batch=batch, proc_limit=512))
    check_phold(H, H.phold_params(3000, 4.0, remote_pct=60, n_hot=7, hot_pct=40, horizon=0.5, cap=4096, batch=batch,
                                  proc_limit=1024))


def test_phold_golden_fixture(H):
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "phold_small.json")))
    for case in g["cases"]:
        p = H.phold_params(horizon=0.6, cap=1024, **case["params"])
        st, n_exec, h = run_phold(H, p)
        assert int(st.committed) == case["committed"] and hex(int(st.trace_hash_sum)) == case["hash_sum"]
        assert [int(x) for x in n_exec[:8]] == case["n_exec_head"]
        assert [hex(int(x)) for x in h[:4]] == case["hash_head"]


def test_phold_edge_shapes(H):
    check_phold(H, H.phold_params(1, 30.0, remote_pct=100, population=3))
    check_phold(H, H.phold_params(2, 30.0, remote_pct=100, population=1, horizon=-1.0))
    check_phold(H, H.phold_params(64, 0.0, remote_pct=50))                                   # empty run
    check_phold(H, H.phold_params(129, 5.0, remote_pct=100, lookahead=0.0, horizon=-1.0))    # zero lookahead
    check_phold(H, H.phold_params(4097, 3.0, remote_pct=50, horizon=0.4))                    # ragged last block
    check_phold(H, H.phold_params(512, 5.0, remote_pct=50, n_hot=64, hot_pct=30, horizon=1.0, cap=256))   # skewed fan-out
    check_phold(H, H.phold_params(64, 6.0, remote_pct=50, population=4, horizon=-1.0, cap=64, proc_limit=4, fossil=2))


@pytest.mark.parametrize("grouped", [True, False])
def test_machrep2_reference_configuration(H, grouped):
    """C1: tests/MachRep2DistributedReproducible.hs, 3 LPs, default DIOParams (no time horizon), through the closed-group
    kernel (k_run_groups) and through grid-wide super-steps (k_step)."""
    import json, os
    d = D()
    dio, sim = d.runDIO(d.MachRep2(1), d.DIOParams(queueCapacity=64, stepsPerLaunch=32, closedGroupKernel=grouped), d.timeServer(3),
                        d.Specs(0.0, 1000.0))
    st, n_exec, h, res = sim()
    dio.close()
    o = H.oracle_machrep(H.machrep_params(1, 1000.0))[0]
    assert res[0]["up_time_prop"] == o.up_time_prop and res[0]["immed_prop"] == o.immed_prop
    assert abs(res[0]["up_time_prop"] - 0.6) < 0.03 and abs(res[0]["immed_prop"] - 0.67) < 0.03     # "about 0.6 and 0.67"
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "machrep2_c1.json")))
    assert float(res[0]["up_time_prop"]).hex() == g["up_time_prop"] and [int(x) for x in n_exec] == g["n_exec"]
    assert [hex(int(x)) for x in h] == g["hash"]
    assert st.rolled_back > 0          # unbounded optimism really exercised the rollback path


@pytest.mark.parametrize("grouped", [True, False])
@pytest.mark.parametrize("horizon", [None, 2.0, 0.0])
def test_machrep2_replicas(H, horizon, grouped):
    d = D()
    n = 256
    dio, sim = d.runDIO(d.MachRep2(n), d.DIOParams(dioTimeHorizon=horizon, queueCapacity=64, stepsPerLaunch=32, closedGroupKernel=grouped),
                        d.timeServer(3 * n), d.Specs(0.0, 300.0))
    st, n_exec, h, res = sim()
    dio.close()
    ores = H.oracle_machrep(H.machrep_params(n, 300.0))
    for r in range(n):
        o = ores[r]
        assert res[r]["total_up_time"] == o.total_up_time and res[r]["n_rep"] == o.n_rep and res[r]["n_immed_rep"] == o.n_immed_rep
        rel = abs(res[r]["up_time_prop"] - o.up_time_prop) / o.up_time_prop
        assert rel <= 1e-9 and res[r]["up_time_prop"] == o.up_time_prop          # north_star tolerance, and in fact equal
        assert [int(n_exec[3 * r + k]) for k in range(3)] == [o.n_exec[k] for k in range(3)]
        assert [int(h[3 * r + k]) for k in range(3)] == [o.hash[k] for k in range(3)]


def test_simulation_retry_really_happens(H):
    """a19 (handleEventRetry / SimulationRetry, Internal/Event.hs:247-249, 768-813): unbounded optimism with a large batch
    makes the master see ReleaseRepairPerson before the request it belongs to; the handler returns H_RETRY, the event stays
    pending until the straggler's rollback arrives, and the statistics still equal the oracle's.  Both kernels."""
    d = D()
    n = 512
    ores = H.oracle_machrep(H.machrep_params(n, 200.0))
    for grouped in (True, False):
        dio, sim = d.runDIO(d.MachRep2(n), d.DIOParams(queueCapacity=1024, maxEventsPerStep=16, stepsPerLaunch=32, closedGroupKernel=grouped),
                            d.timeServer(3 * n), d.Specs(0.0, 200.0))
        st, n_exec, h, res = sim()
        dio.close()
        assert st.retries > 0 and st.rolled_back > 0
        assert (res["up_time_prop"] == np.array([r.up_time_prop for r in ores])).all()
        assert (res["immed_prop"] == np.array([r.immed_prop for r in ores])).all()
        assert (h == np.array([r.hash[k] for r in ores for k in range(3)], dtype=np.uint64)).all()


@pytest.mark.parametrize("horizon", [None, 5.0, 0.0])
def test_machrep1_replicas(H, horizon):
    """tests/MachRep1.hs through the handler table (third compiled model): per-replica result, per-LP counts and trace
    hashes equal the sequential oracle's; replica 0 equals the golden fixture."""
    import json, os
    d = D()
    n = 64
    dio, sim = d.runDIO(d.MachRep1(n), d.DIOParams(dioTimeHorizon=horizon, queueCapacity=512, maxEventsPerStep=8, stepsPerLaunch=32),
                        d.timeServer(3 * n), d.Specs(0.0, 400.0))
    st, n_exec, h, res = sim()
    dio.close()
    ores = H.oracle_machrep1(H.machrep1_params(n, 400.0))
    for r in range(n):
        o = ores[r]
        assert res[r]["total_up_time"] == o.total_up_time and res[r]["n_rep"] == o.n_rep
        assert res[r]["up_time_prop"] == o.up_time_prop
        assert [int(n_exec[3 * r + k]) for k in range(3)] == [o.n_exec[k] for k in range(3)]
        assert [int(h[3 * r + k]) for k in range(3)] == [o.hash[k] for k in range(3)]
    if horizon != 0.0:
        assert st.rolled_back > 0
    if horizon is None:
        dio, sim = d.runDIO(d.MachRep1(1), d.DIOParams(queueCapacity=1024, maxEventsPerStep=16, stepsPerLaunch=32), d.timeServer(3),
                            d.Specs(0.0, 10000.0))
        st, n_exec, h, res = sim()
        dio.close()
        g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "machrep1.json")))
        assert float(res[0]["up_time_prop"]).hex() == g["up_time_prop"] and [int(x) for x in n_exec] == g["n_exec"]
        assert abs(res[0]["up_time_prop"] - 0.66) < 0.01          # "Should get value of about 0.66"


def test_two_partitions_under_one_time_server(H):
    """two device partitions and a CPU process registered with one host-side Time Server.  GVT of every
    partition is bounded by the server's GlobalTimeMessage (minimum over all local times), a value above the local time
    is ignored, and the committed traces equal the stand-alone runs (= the oracle)."""
    d = D()

    class CpuProcess:                      # a logical process that only advances its clock
        def __init__(self): self.t = 0.0; self.seen = []
        def localTime(self): return self.t
        def globalTimeMessage(self, t): self.seen.append(t); return True

    pa = H.phold_params(3000, 6.0, remote_pct=40, horizon=0.5, batch=8, seed=11, cap=256)
    pb = H.phold_params(1000, 6.0, remote_pct=20, horizon=1.0, batch=4, seed=12, cap=256)
    mk = lambda p: d.DIO(d.PHOLD(p.n_lps, p.phold.population, p.phold.remote_pct, p.phold.lookahead, p.phold.mean, p.phold.seed),
                         d.DIOParams(dioTimeHorizon=p.time_horizon, queueCapacity=p.queue_capacity,
                                     maxEventsPerStep=p.max_events_per_step, stepsPerLaunch=4), d.Specs(0.0, p.stop_time))
    a, b, cpu = mk(pa), mk(pb), CpuProcess()
    ts = d.timeServer(3)
    for lp in (a, b):
        lp.attachTimeServer()
        lp.enqueueMessages(lp.initialEvents())
        ts.register(lp)
    assert ts.computeGlobalTime() is None          # quorum of 3 not met yet
    ts.register(cpu)
    assert a.globalTimeMessage(1e9) is False        # above the local time: ignored (Internal/Event.hs:533-536)
    rounds, done = 0, False
    while not done:
        sa, sb = a.runEventInStopTime(), b.runEventInStopTime()
        g = ts.globalTime()
        assert g is None or (sa.gvt <= g and sb.gvt <= g) or (sa.gvt > 6.0 and sb.gvt > 6.0)
        cpu.t = min(cpu.t + 0.37, 1e30) if cpu.t < 7.0 else 1e30      # the CPU process lags, then finishes
        g = ts.computeGlobalTime()
        assert g == min(a.localTime(), b.localTime(), cpu.t)
        rounds += 1
        done = sa.gvt > 6.0 and sb.gvt > 6.0
        assert rounds < 5000
    assert rounds > 10 and cpu.seen == sorted(cpu.seen)          # the global time never decreases
    for lp, p in ((a, pa), (b, pb)):
        n_exec, h = lp.lpStats()
        o_exec, o_hash, _ = H.oracle_phold(p)
        assert (n_exec == o_exec).all() and (h == o_hash).all()
        lp.close()


def test_full_size_c3_invariants(H):
    """PHOLD 64K LPs / 10 % remote (BASELINE configs[2]) at a shortened stop time the oracle finishes in seconds,
    then at full horizon sweep: identical committed trace for every window, zero rollbacks when conservative."""
    p = H.phold_params(1 << 16, 5.0, remote_pct=10, horizon=0.4)
    check_phold(H, p)
    ref = None
    for horizon, batch in ((0.1, 4), (0.4, 4), (1.0, 2), (-1.0, 4)):
        st, n_exec, h = run_phold(H, H.phold_params(1 << 16, 20.0, remote_pct=10, horizon=horizon, batch=batch))
        sig = (int(st.committed), int(st.trace_hash_sum), int(n_exec.sum()))
        if horizon == 0.1:
            assert st.rolled_back == 0
        ref = ref or sig
        assert sig == ref
    expect = (1 << 16) * 16 * 20.0 / 1.1
    assert abs(ref[0] - expect) / expect < 0.02


def test_full_size_c4_c5_invariants(H):
    """1M-LP / 50 % remote (configs[3]) and the skewed 256K-LP case (configs[4]) on one GPU, short stop time:
    the committed trace is identical for a conservative and an optimistic window, and run-to-run."""
    for kw, cap, batch in ((dict(n_lps=1 << 20, remote_pct=50), 128, 4), (dict(n_lps=1 << 18, remote_pct=50, n_hot=1024, hot_pct=50), 4096, 64)):
        sigs = []
        for horizon in (0.1, 0.5, 0.5):
            st, n_exec, h = run_phold(H, H.phold_params(stop=2.0, horizon=horizon, cap=cap, batch=batch, **kw))
            sigs.append((int(st.committed), int(st.trace_hash_sum)))
        assert sigs[0] == sigs[1] == sigs[2]


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_benched_configurations_equal_the_oracle_at_full_size(H, name):
    """Every BASELINE configuration exactly as bench.py runs it (full LP count, full stop time, the benched window, batch
    and ring) against the oracle's answer for that configuration (tests/golden/full_configs.json, generated by
    tests/golden/make_golden_full.py from oracle/oracle.c): committed count, trace-hash sum, the LP-bound digest of the
    per-LP (count, hash) pairs and, for MachRep2, the digest of the per-replica results and the documented means."""
    import json, os
    from aivika_distributed_b200 import workloads as W
    d = D()
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "full_configs.json")))[name]
    wl = W.WORKLOADS[name]
    ps = d.DIOParams(dioTimeHorizon=None if wl["horizon"] < 0 else wl["horizon"], queueCapacity=wl["cap"], maxEventsPerStep=wl["batch"])
    if wl["kind"] == "phold":
        model = d.PHOLD(wl["n_lps"], 16, wl["remote_pct"], 0.1, 1.0, W.SEED, wl.get("n_hot", 0), wl.get("hot_pct", 0))
    else:
        model = d.MachRep2(wl["n_replicas"])
    dio, sim = d.runDIO(model, ps, d.timeServer(1), d.Specs(0.0, wl["stop"]))
    try:
        out = sim()
    finally:
        dio.close()
    st, n_exec, h = out[0], out[1], out[2]
    assert int(st.committed) == fx["committed"] and hex(int(st.trace_hash_sum)) == fx["trace_hash_sum"]
    assert hex(W.lp_digest(0, n_exec, h)) == fx["lp_digest"]
    if wl["kind"] == "machrep":
        res = out[3]
        assert hex(W.machrep_digest(0, res)) == fx["replica_digest"]
        assert abs(res["up_time_prop"].mean() - fx["mean_up_time_prop"]) < 1e-12 and abs(res["up_time_prop"].mean() - 0.6) < 0.02
        assert abs(res["immed_prop"].mean() - fx["mean_immed_prop"]) < 1e-12 and abs(res["immed_prop"].mean() - 0.667) < 0.02
    if name == "c5":
        assert st.committed / st.executed < 0.99          # rollback-heavy, as BASELINE configs[4] says


def test_errors_are_reported_through_status_codes(H):
    d = D()
    # ring too small for the population: TW_ERR_QUEUE_OVERFLOW, not corruption
    dio, sim = d.runDIO(d.PHOLD(256, population=16, remote_pct=50), d.DIOParams(queueCapacity=16), None, d.Specs(0.0, 50.0))
    with pytest.raises(d.SimulationAbort) as e:
        sim()
    assert e.value.code == 5
    dio.close()
    # message to a logical process that does not exist / send time after receive time / message in the past of GVT
    dio = d.DIO(d.PHOLD(64), d.DIOParams(), d.Specs(0.0, 5.0))
    with pytest.raises(d.SimulationAbort) as e:
        dio.enqueueMessage(64, 1.0, src=0)
    assert e.value.code == 9
    dio.close()
    dio = d.DIO(d.PHOLD(64), d.DIOParams(), d.Specs(0.0, 5.0))
    with pytest.raises(d.SimulationAbort) as e:
        dio.enqueueMessage(3, 1.0, src=0, send_time=2.0)
    assert e.value.code == 3
    dio.close()
    dio = d.DIO(d.PHOLD(64), d.DIOParams(dioAllowSkippingOutdatedMessage=False), d.Specs(1.0, 5.0))
    with pytest.raises(d.SimulationAbort) as e:
        dio.enqueueMessage(3, 0.5, src=0, send_time=0.5)
    assert e.value.code == 4
    dio.close()
    # the reference's default (dioAllowSkippingOutdatedMessage = True, Internal/DIO.hs:206): logged and skipped
    dio = d.DIO(d.PHOLD(64), d.DIOParams(), d.Specs(1.0, 5.0))
    dio.enqueueMessage(3, 0.5, src=0, send_time=0.5)
    dio.enqueueMessage(3, 1.5, src=0, send_time=1.0)
    assert dio.skippedOutdatedMessages() == 1
    st = dio.runEventInStopTime()
    assert st.committed >= 1
    dio.close()
    # a ring larger than the 16-bit fields of the LP header allow is refused
    with pytest.raises(d.SimulationAbort) as e:
        d.DIO(d.PHOLD(64), d.DIOParams(queueCapacity=65536), d.Specs(0.0, 5.0))
    assert e.value.code == 1


def test_resume_and_time_server_role(H):
    """runEventInStopTime can be called with increasing times; GVT (the Time Server's GlobalTimeMessage) is monotone,
    and the LP state record mirrors LogicalProcessState."""
    d = D()
    ts = d.timeServer(1000)
    dio, _ = d.runDIO(d.PHOLD(1000, remote_pct=40), d.DIOParams(dioTimeHorizon=0.5), ts, d.Specs(0.0, 6.0))
    dio.enqueueMessages(dio.initialEvents())
    g0 = ts.globalTime()
    st1 = dio.runEventInStopTime(2.0); g1 = ts.globalTime()
    st2 = dio.runEventInStopTime(6.0); g2 = ts.globalTime()
    assert g0 <= g1 <= g2 and g1 > 2.0 and g2 > 6.0
    n_exec, h = dio.lpStats()
    o_exec, o_hash, _ = H.oracle_phold(H.phold_params(1000, 6.0, remote_pct=40))
    assert (n_exec == o_exec).all() and (h == o_hash).all()
    s = dio.logicalProcessState(17)
    assert s.lpStateId == 17 and s.lpStateLocalTime <= 6.0 and s.lpStateEventQueueTime > 6.0 and s.lpStateEventQueueSize > 0
    # lpStateRollbackCount is per LP (Optimistic/State.hs:50): the counts of all LPs add up to the engine's total
    per_lp = [dio.logicalProcessState(lp).lpStateRollbackCount for lp in range(1000)]
    assert sum(per_lp) == st1.rollbacks + st2.rollbacks and max(per_lp) > 0 and max(per_lp) < sum(per_lp)
    # reset + rerun gives the same trace
    dio.reset(); dio.enqueueMessages(dio.initialEvents()); dio.runEventInStopTime()
    n2, h2 = dio.lpStats()
    assert (n2 == o_exec).all() and (h2 == o_hash).all()
    dio.close()


def test_kernels_were_launched(H):
    lib = H.ffi.load()
    before = lib.tw_kernel_launch_count()
    run_phold(H, H.phold_params(1000, 2.0))
    assert lib.tw_kernel_launch_count() > before


def test_cpp_host_mirror_runs_machrep2(H):
    """The C++ DIO mirror reproduces the reference program's result through the same C ABI."""
    import os, subprocess
    d = os.path.join(H.PKG_DIR, "host")
    subprocess.run(["make", "-s", "-C", d, "example_machrep2"], check=True)
    r = subprocess.run([os.path.join(d, "example_machrep2")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    o = H.oracle_machrep(H.machrep_params(1, 1000.0))[0]
    assert f"({o.up_time_prop:.17g},{o.immed_prop:.17g})" in r.stdout


def test_monitoring_action(H):
    """dioSimulationMonitoringAction analogue: periodic engine state with monotone GVT and counters."""
    d = D()
    dio = d.DIO(d.PHOLD(4096, remote_pct=30), d.DIOParams(dioTimeHorizon=0.5, stepsPerLaunch=4), d.Specs(0.0, 8.0))
    seen = []
    dio.setMonitoringAction(seen.append)
    dio.enqueueMessages(dio.initialEvents())
    st = dio.runEventInStopTime()
    dio.close()
    assert len(seen) >= 3
    g = [s["gvt"] for s in seen]; ex = [s["executed"] for s in seen]
    assert g == sorted(g) and ex == sorted(ex) and seen[0]["n_lps_local"] == 4096
    assert seen[-1]["executed"] <= st.executed


def test_c2_4096_replicas_match_oracle(H):
    """BASELINE configs[1]: MachRep2 replicated as 4096 independent groups on one GPU (shortened stop time so the
    oracle finishes in seconds): every replica's statistics and per-LP traces equal the sequential oracle's, and the
    population means sit at the documented "about 0.6 and 0.67"."""
    d = D()
    n, stop = 4096, 120.0
    dio, sim = d.runDIO(d.MachRep2(n), d.DIOParams(dioTimeHorizon=2.0, queueCapacity=256, stepsPerLaunch=32),
                        d.timeServer(3 * n), d.Specs(0.0, stop))
    st, n_exec, h, res = sim()
    dio.close()
    ores = H.oracle_machrep(H.machrep_params(n, stop))
    o_exec = np.array([r.n_exec[k] for r in ores for k in range(3)], dtype=np.uint64)
    o_hash = np.array([r.hash[k] for r in ores for k in range(3)], dtype=np.uint64)
    assert (n_exec == o_exec).all() and (h == o_hash).all()
    assert (res["up_time_prop"] == np.array([r.up_time_prop for r in ores])).all()
    assert (res["immed_prop"] == np.array([r.immed_prop for r in ores])).all()
    assert abs(res["up_time_prop"].mean() - 0.6) < 0.02 and abs(res["immed_prop"].mean() - 0.667) < 0.02
    assert st.rolled_back > 0


def test_random_phold_configurations(H):
    """Seeded sweep over shapes the fixed cases do not hit (ragged LP counts, tiny populations, every batch size,
    zero / full remote share, hot sets): the CUDA engine against the oracle, per LP."""
    rng = np.random.default_rng(20261020)
    for _ in range(24):
        n = int(rng.integers(2, 3000)); pop = int(rng.integers(1, 9)); rp = int(rng.choice([0, 5, 50, 100]))
        hz = float(rng.choice([0.05, 0.1, 0.35, 1.5, -1.0])); batch = int(rng.choice([1, 2, 3, 4, 6, 8, 16, 64]))
        n_hot = int(rng.choice([0, 0, 7])); seed = int(rng.integers(0, 2 ** 31))
        p = H.phold_params(n, 4.0, population=pop, remote_pct=rp, horizon=hz, batch=batch, seed=seed,
                           n_hot=min(n_hot, n), hot_pct=5 if n_hot else 0, cap=1024, steps_per_launch=int(rng.choice([1, 4, 16])))
        check_phold(H, p)


def test_phold_equal_timestamps(H):
    """Degenerate time stamps (mean 0: every event of a generation has the same receive time): tie ordering by the
    full key and the more-ties-than-candidates path on the device."""
    for n, pop, rp, hz, batch in ((50, 8, 50, 1.0, 4), (2000, 6, 30, -1.0, 4), (64, 5, 100, 0.5, 2), (5000, 8, 50, 0.5, 3)):
        check_phold(H, H.phold_params(n, 3.0, population=pop, remote_pct=rp, lookahead=0.5, mean=0.0, horizon=hz,
                                      batch=batch, cap=512))


def test_start_time_events_generated_on_the_device(H):
    """tw_populate (runEventInStartTime on every LP, on the device): the same committed trace as the host-built start-time
    messages, for PHOLD (the population) and for both MachRep models (the slaves' `machine` processes), against the oracle."""
    d = D()
    p = H.phold_params(3001, 5.0, remote_pct=40, horizon=0.5, batch=6, cap=256, seed=11)
    o_exec, o_hash, o_sum = H.oracle_phold(p)
    dio = d.DIO(d.PHOLD(p.n_lps, p.phold.population, p.phold.remote_pct, p.phold.lookahead, p.phold.mean, p.phold.seed),
                d.DIOParams(dioTimeHorizon=0.5, queueCapacity=256, maxEventsPerStep=6), d.Specs(0.0, 5.0))
    dio.runEventInStartTime()
    st = dio.runEventInStopTime()
    n_exec, h = dio.lpStats()
    dio.close()
    assert (n_exec == o_exec).all() and (h == o_hash).all() and st.committed == o_sum.committed
    for grouped in (True, False):
        dio = d.DIO(d.MachRep2(64), d.DIOParams(queueCapacity=128, stepsPerLaunch=32, closedGroupKernel=grouped), d.Specs(0.0, 200.0))
        dio.runEventInStartTime()
        dio.runEventInStopTime()
        res = dio.machRepResults()
        dio.close()
        ores = H.oracle_machrep(H.machrep_params(64, 200.0))
        assert all(res[r]["up_time_prop"] == ores[r].up_time_prop and res[r]["immed_prop"] == ores[r].immed_prop for r in range(64))
