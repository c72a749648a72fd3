"""The multi-threaded CPU restatement of the reference's per-LP Time Warp (oracle/tw_ref.cpp: sorted input queue
with straggler / anti-message rules, output queue with aggressive cancellation, undo log, min-reduction GVT, strict
fossil collection) must commit exactly the sequential oracle's schedule for any thread count and horizon."""
import pytest


@pytest.mark.parametrize("threads", [1, 3, 8])
@pytest.mark.parametrize("horizon", [0.1, 1.0, -1.0])
def test_phold_restatement_equals_sequential(H, threads, horizon):
    p = H.phold_params(600, 6.0, remote_pct=40, horizon=horizon)
    n_exec, h, s = H.oracle_phold(p)
    rc, e, hh, st = H.twref_phold(p, threads)
    assert rc == 0
    assert (e == n_exec).all() and (hh == h).all() and st.committed == s.committed
    assert st.executed - st.rolled_back == st.committed
    assert st.gvt > p.stop_time
    if threads > 1 and horizon != 0.1:
        assert st.rolled_back > 0 and st.anti_sent >= st.annihilated > 0


def test_phold_restatement_skewed(H):
    p = H.phold_params(400, 5.0, remote_pct=50, n_hot=16, hot_pct=40, horizon=2.0, seed=4)
    n_exec, h, _ = H.oracle_phold(p)
    rc, e, hh, st = H.twref_phold(p, 4)
    assert rc == 0 and (e == n_exec).all() and (hh == h).all()


@pytest.mark.parametrize("threads,horizon", [(1, -1.0), (3, -1.0), (6, 5.0)])
def test_machrep2_restatement_equals_sequential(H, threads, horizon):
    """Default DIOParams of the reference program = no time horizon; the three LPs of a replica may sit on different threads."""
    p = H.machrep_params(5, 300.0, horizon=horizon)
    res = H.oracle_machrep(p)
    rc, e, hh, out, st = H.twref_machrep(p, threads)
    assert rc == 0
    for r in range(5):
        assert out[r].up_time_prop == res[r].up_time_prop and out[r].immed_prop == res[r].immed_prop
        assert [int(e[3 * r + k]) for k in range(3)] == [res[r].n_exec[k] for k in range(3)]
        assert [int(hh[3 * r + k]) for k in range(3)] == [res[r].hash[k] for k in range(3)]
    if threads > 1:
        assert st.rolled_back > 0


def test_machrep2_c1_three_threads(H):
    """tests/MachRep2DistributedReproducible.hs: one thread per LP, like one node per LP in the reference."""
    p = H.machrep_params(1, 1000.0)
    o = H.oracle_machrep(p)[0]
    rc, e, hh, out, st = H.twref_machrep(p, 3)
    assert rc == 0 and out[0].up_time_prop == o.up_time_prop and out[0].immed_prop == o.immed_prop
    assert abs(out[0].up_time_prop - 0.6) < 0.03 and abs(out[0].immed_prop - 0.67) < 0.03
