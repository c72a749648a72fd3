"""Host-side logic of the engine: the device functions of tw_core.cuh compiled for the CPU (tests/emu) and run
LP by LP in shuffled order must commit exactly the sequential oracle's schedule for every horizon, batch size,
capacity and partitioning.  This is a debugging aid for the rollback / annihilation / GVT logic; the parity
claim for the product is made by the `-m gpu` tests through the C ABI."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def check_phold(H, p, seeds=(0, 1)):
    n_exec, h, s = H.oracle_phold(p)
    out = None
    for sd in seeds:
        rc, e_exec, e_hash, _, stats = H.emu_run(p, sd)
        assert rc == 0
        assert (e_exec == n_exec).all() and (e_hash == h).all()
        assert stats.committed == s.committed
        assert stats.executed - stats.rolled_back == stats.committed
        out = stats
    return out


@pytest.mark.parametrize("horizon", [0.1, 0.5, 2.0, -1.0])
@pytest.mark.parametrize("batch", [1, 4, 6, 16, 64])
def test_phold_all_horizons_commit_the_sequential_schedule(H, horizon, batch):
    st_ = check_phold(H, H.phold_params(300, 6.0, remote_pct=40, horizon=horizon, batch=batch, cap=128 if batch <= 4 else 512))
    if horizon == 0.1:
        assert st_.rolled_back == 0 and st_.anti_sent == 0      # window == lookahead is conservative
    if horizon != 0.1:
        assert st_.rolled_back > 0 and st_.anti_sent == st_.annihilated + st_.kernel_launches   # (+ antis still in flight at the end)


def test_phold_edge_shapes(H):
    check_phold(H, H.phold_params(1, 30.0, remote_pct=100, population=3))          # single LP: never remote
    check_phold(H, H.phold_params(2, 30.0, remote_pct=100, population=1, horizon=-1.0))
    check_phold(H, H.phold_params(64, 4.0, remote_pct=0))                          # no remote messages at all
    check_phold(H, H.phold_params(64, 0.0, remote_pct=50))                         # nothing before the stop time
    check_phold(H, H.phold_params(129, 5.0, remote_pct=100, lookahead=0.0, horizon=-1.0))   # zero lookahead


def test_phold_skewed_fanout_and_small_ring(H):
    st_ = check_phold(H, H.phold_params(512, 5.0, remote_pct=50, n_hot=64, hot_pct=30, horizon=1.0, cap=256))
    assert st_.rolled_back > 0
    # tiny ring: optimism stalls on proc_limit and tombstones are compacted, results must not change
    check_phold(H, H.phold_params(64, 6.0, remote_pct=50, population=4, horizon=-1.0, cap=64, proc_limit=4, fossil=2))


@pytest.mark.parametrize("batch", [2, 8, 64])
def test_phold_long_pending_regions_take_the_whole_region_search(H, batch):
    """A few very hot LPs collect far more than 64 pending events: their selection cannot use the 64-bit window mask and
    goes through the search over the whole region (cooperative on the GPU), including equal-time ties."""
    st_ = check_phold(H, H.phold_params(128, 3.0, population=8, remote_pct=80, n_hot=2, hot_pct=60, horizon=1.0, cap=2048,
                                        batch=batch, proc_limit=512))
    assert st_.rolled_back > 0
    check_phold(H, H.phold_params(96, 4.0, population=8, remote_pct=90, n_hot=1, hot_pct=100, lookahead=0.5, mean=0.0,
                                  horizon=-1.0, cap=2048, batch=batch, proc_limit=512), seeds=(2,))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_phold_partitioned(H, world):
    st_ = check_phold(H, H.phold_params(500, 5.0, remote_pct=50, horizon=0.7, world=world), seeds=(3,))
    assert st_.sent_remote > 0


def test_queue_overflow_is_reported_not_silent(H):
    rc, *_ = H.emu_run(H.phold_params(64, 50.0, remote_pct=50, population=16, horizon=-1.0, cap=16), 1)
    assert rc == 5   # TW_ERR_QUEUE_OVERFLOW


def check_machrep(H, p, seeds=(0, 5)):
    res = H.oracle_machrep(p)
    for sd in seeds:
        rc, n_exec, h, mr, stats = H.emu_run(p, sd)
        assert rc == 0
        for r in range(p.n_lps // 3):
            o, m = res[r], mr[r]
            assert (o.total_up_time, o.n_rep, o.n_immed_rep) == (m.total_up_time, m.n_rep, m.n_immed_rep)
            assert o.up_time_prop == m.up_time_prop and o.immed_prop == m.immed_prop     # bit-exact, stricter than 1e-9
            assert [o.n_exec[k] for k in range(3)] == [int(n_exec[3 * r + k]) for k in range(3)]
            assert [o.hash[k] for k in range(3)] == [int(h[3 * r + k]) for k in range(3)]
    return stats


def test_machrep2_reference_configuration_unbounded_optimism(H):
    st_ = check_machrep(H, H.machrep_params(1, 1000.0))       # dioTimeHorizon = Nothing, as in the reference program
    assert st_.rolled_back > 0


@pytest.mark.parametrize("horizon,batch", [(100.0, 4), (1.0, 1), (0.0, 4)])
def test_machrep2_horizons(H, horizon, batch):
    check_machrep(H, H.machrep_params(6, 150.0, horizon=horizon, batch=batch), seeds=(2,))


def test_machrep2_partitioned_keeps_replicas_whole(H):
    check_machrep(H, H.machrep_params(7, 100.0, horizon=5.0, world=3), seeds=(1,))


def check_machrep1(H, p, seeds=(0, 5)):
    res = H.oracle_machrep1(p)
    for sd in seeds:
        rc, n_exec, h, mr, stats = H.emu_run(p, sd)
        assert rc == 0
        for r in range(p.n_lps // 3):
            o, m = res[r], mr[r]
            assert (o.total_up_time, o.n_rep) == (m.total_up_time, m.n_rep) and o.up_time_prop == m.up_time_prop
            assert [o.n_exec[k] for k in range(3)] == [int(n_exec[3 * r + k]) for k in range(3)]
            assert [o.hash[k] for k in range(3)] == [int(h[3 * r + k]) for k in range(3)]
    return stats


@pytest.mark.parametrize("horizon,batch,world", [(-1.0, 4, 1), (5.0, 8, 1), (0.5, 2, 3), (0.0, 4, 1)])
def test_machrep1_third_model_through_the_handler_table(H, horizon, batch, world):
    """tests/MachRep1.hs: the slaves never receive anything and run ahead, the master gets zero-lookahead messages
    (sendMessage: receive time = send time) from two senders and must commit them in time order."""
    st_ = check_machrep1(H, H.machrep1_params(5, 300.0, horizon=horizon, batch=batch, world=world, cap=512), seeds=(3,))
    if horizon != 0.0:
        assert st_.rolled_back > 0


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 200), pop=st.integers(1, 6), rp=st.integers(0, 100), hor=st.sampled_from([0.05, 0.3, 1.5, -1.0]),
       batch=st.sampled_from([1, 2, 3, 4, 6, 7, 16, 64]), seed=st.integers(0, 2 ** 32), world=st.sampled_from([1, 2, 4]), shuffle=st.integers(0, 99))
def test_phold_random_configurations(H, n, pop, rp, hor, batch, seed, world, shuffle):
    p = H.phold_params(n, 4.0, population=pop, remote_pct=rp, horizon=hor, batch=batch, seed=seed, world=world, cap=128 if batch <= 4 else 512)
    check_phold(H, p, seeds=(shuffle,))


def test_phold_equal_timestamps_exercise_full_key_ordering(H):
    """mean 0 makes every delay equal to the lookahead: whole populations share one receive time, so the engine must
    fall back to the full key (send time, sender, counter) -- including more ties than the candidate list holds."""
    for n, pop, rp, hz, batch in ((50, 8, 50, 1.0, 4), (200, 6, 30, -1.0, 4), (64, 5, 100, 0.5, 2)):
        st_ = check_phold(H, H.phold_params(n, 3.0, population=pop, remote_pct=rp, lookahead=0.5, mean=0.0, horizon=hz,
                                            batch=batch, cap=512), seeds=(3,))
        assert st_.rolled_back > 0


def test_phold_reverse_computation_is_the_exact_inverse(H):
    """PHOLD is rolled back by reverse computation (no saved states): reverse(advance(s, e), e) == s for random words."""
    assert H.emu().emu_phold_reverse_roundtrip(12345, 200000) == 0


@pytest.mark.parametrize("every", [2, 7, 40])
def test_lagging_global_time_of_an_external_time_server(H, every):
    """under an external Time Server the engine only learns the global time now and then
    (GlobalTimeMessage); with any lower bound of the true GVT the committed trace must stay the sequential one."""
    p = H.phold_params(400, 6.0, remote_pct=40, horizon=0.5, batch=8, cap=1024, proc_limit=256)
    n_exec, h, s = H.oracle_phold(p)
    rc, e_exec, e_hash, _, st_ = H.emu_run(p, 3, gvt_every=every)
    assert rc == 0 and (e_exec == n_exec).all() and (e_hash == h).all() and st_.committed == s.committed
    pm = H.machrep_params(4, 120.0, horizon=5.0, batch=4, cap=1024)
    res = H.oracle_machrep(pm)
    rc, n_exec, h, mr, _ = H.emu_run(pm, 2, gvt_every=every)
    assert rc == 0 and all(res[r].up_time_prop == mr[r].up_time_prop and res[r].immed_prop == mr[r].immed_prop for r in range(4))


def test_model_building_blocks(H):
    """tw_dsl.cuh: FcfsResource follows requestResource / releaseResourceWithinEvent (FCFS hand-over, over-release =
    SimulationRetry), Outbox follows enqueueMessage / sendMessage, random_exponential draws from the undoable StdGen."""
    assert H.emu().emu_dsl_selftest() == 0


@pytest.mark.timeout(120)
def test_overflow_inside_a_long_region_is_an_error_code_not_a_hang(H):
    """Regression: a lane searching a long pending region (whole-region search) that hits an inconsistency caused by a
    ring overflow used to keep the warp's round loop alive for ever; it must report TW_ERR_QUEUE_OVERFLOW."""
    p = H.phold_params(1458, 3.0, population=1, remote_pct=50, lookahead=0.0, mean=1.0, seed=177738635, n_hot=16, hot_pct=80,
                       horizon=2.5, cap=64, batch=6)
    for sd in (0, 11, 512):
        rc, *_ = H.emu_run(p, sd, gvt_every=3)
        assert rc == 5


def test_quantum_tie_across_the_candidate_list_boundary(H):
    """Regression (found by tools/stress_emu.py): three entries in one time quantum, two of them a tombstone pair -- the
    last candidate and the first dropped entry looked like a harmless pair, a third, earlier entry behind them was executed
    too late (TW_ERR_TIME_ORDER).  A tie across the list boundary must always take the exact search."""
    p = H.machrep_params(4, 120.0, horizon=-1.0, cap=256, batch=16, world=2, seed1=79928, seed2=149933)
    res = H.oracle_machrep(p)
    for sd in (1, 2, 3, 5, 8):
        rc, n_exec, h, mr, _ = H.emu_run(p, sd, gvt_every=5)
        assert rc == 0 and all(res[r].up_time_prop == mr[r].up_time_prop and res[r].immed_prop == mr[r].immed_prop for r in range(4))
    p = H.machrep_params(11, 120.0, horizon=-1.0, cap=1024, batch=16, seed1=582035, seed2=504955)
    res = H.oracle_machrep(p)
    for sd in (0, 4, 7):
        rc, n_exec, h, mr, _ = H.emu_run(p, sd)
        assert rc == 0 and all(res[r].up_time_prop == mr[r].up_time_prop for r in range(11))


# ---------------------------------------------------------------- the warp-level build (tests/emu/simt.hpp)
def check_phold_simt(H, p, tiles=1, seeds=(0, 3)):
    n_exec, h, s = H.oracle_phold(p)
    for sd in seeds:
        rc, e_exec, e_hash, _, stats = H.emu_simt_run(p, sd, 1, tiles)
        assert rc == 0 and (e_exec == n_exec).all() and (e_hash == h).all() and stats.committed == s.committed
    return stats


@pytest.mark.parametrize("tiles", [1, 2, 3])
def test_warp_level_paths_phold(H, tiles):
    """The code a GPU warp runs -- work queue with deferred delivery, warp-uniform loops, whole-warp search of long
    regions, tie fallbacks, tiles carried through one queue -- executed by 32 host threads as lanes, against the oracle."""
    st_ = check_phold_simt(H, H.phold_params(300, 4.0, remote_pct=50, horizon=0.5, batch=8, cap=256), tiles)
    assert st_.rolled_back > 0 and st_.anti_sent > 0
    check_phold_simt(H, H.phold_params(128, 3.0, population=8, remote_pct=80, n_hot=2, hot_pct=60, horizon=1.0, cap=2048, batch=8,
                                       proc_limit=512), tiles, seeds=(1,))
    check_phold_simt(H, H.phold_params(200, 3.0, population=6, remote_pct=30, lookahead=0.5, mean=0.0, horizon=-1.0, batch=4,
                                       cap=512), tiles, seeds=(2,))
    check_phold_simt(H, H.phold_params(500, 3.0, remote_pct=50, horizon=0.7, world=3, batch=6, cap=256), tiles, seeds=(4,))


@pytest.mark.parametrize("horizon,batch", [(-1.0, 4), (0.5, 4), (2.0, 8), (0.0, 4)])
def test_warp_level_paths_machrep(H, horizon, batch):
    """Regression: the deliveries of one queue flush are unordered, so the copy a rolled-back sender re-sends (same
    counter and times, possibly another payload) can land in front of the anti-message of the old copy; the anti-message
    must annihilate the copy of an EARLIER step, not the newest one (found with this build: wrong MachRep2 statistics)."""
    p = H.machrep_params(3, 40.0, horizon=horizon, batch=batch, cap=256)
    res = H.oracle_machrep(p)
    for sd in (0, 1, 2):
        rc, n_exec, h, mr, _ = H.emu_simt_run(p, sd, 1, 1)
        assert rc == 0
        for r in range(3):
            assert res[r].up_time_prop == mr[r].up_time_prop and res[r].immed_prop == mr[r].immed_prop
            assert [res[r].hash[k] for k in range(3)] == [int(h[3 * r + k]) for k in range(3)]
    p1 = H.machrep1_params(3, 60.0, horizon=horizon, batch=batch, cap=512)
    r1 = H.oracle_machrep1(p1)
    rc, n_exec, h, mr, _ = H.emu_simt_run(p1, 1, 1, 2)
    assert rc == 0 and all(r1[r].up_time_prop == mr[r].up_time_prop for r in range(3))


def test_hot_lp_cooperative_paths(H):
    """One LP receives most of the traffic and holds ~1000 pending events (the shape of C5's hot LPs): the warp-level build
    takes the cooperative bulk selection (coop_select), executes the lined-up records with the whole warp and runs the slow
    path of the inbox drain -- anti-messages, stragglers -- with all 32 lanes (drain_slow_coop, rings of more than 256
    entries).  The committed trace must equal the sequential oracle's for several shuffles and batch sizes."""
    for batch, seed in ((64, 5), (16, 6), (3, 7)):
        p = H.phold_params(256, 2.0, population=8, remote_pct=80, n_hot=1, hot_pct=80, horizon=1.0, cap=4096, batch=batch,
                           proc_limit=1024)
        o_exec, o_hash, o_sum = H.oracle_phold(p)
        rc, e_exec, e_hash, _, stats = H.emu_simt_run(p, seed, 1, 1)
        assert rc == 0 and (e_exec == o_exec).all() and (e_hash == o_hash).all()
        assert stats.rolled_back > 0 and stats.annihilated > 0 and int(o_exec[0]) > 1000      # the hot LP really was hot


def test_bulk_selection_keeps_the_processed_region_sorted_when_every_time_is_tied(H):
    """Regression (found by tools/stress_emu.py on the warp-level build): with mean 0 every delay equals the lookahead, so
    all receive times are tied.  A selection that let ONE live record stay among annihilated records of its own time could
    line it up behind a no-op with a smaller key; the straggler test then stopped at the no-op and missed a rollback
    (wrong committed trace, no error).  Only runs that are annihilated throughout may stay in a selected prefix."""
    p = H.phold_params(214, 2.0, population=7, remote_pct=30, lookahead=0.1, mean=0.0, seed=1915008829, n_hot=16, hot_pct=30,
                       horizon=0.3, cap=512, batch=4, world=2, fossil=2)
    o_exec, o_hash, _ = H.oracle_phold(p)
    rc, e_exec, e_hash, _, _ = H.emu_simt_run(p, 0, 1, 1)         # (shuffles 0, 1, 2, 4 and 7 of 8 failed before the fix)
    assert rc == 0 and (e_exec == o_exec).all() and (e_hash == o_hash).all()


def test_local_time_never_goes_back_under_a_lagging_time_server(H):
    """Regression (same stress run): what a long pending region leaves behind is reported as the lower bound of a time
    quantum of the current window; with GVT refreshed only every 5th step (an external Time Server) that bound could lie
    a hair below the previous step's local time.  The engine publishes the running maximum (block_finish)."""
    p = H.machrep_params(10, 40.0, horizon=50.0, cap=256, batch=4, seed1=211992, seed2=247124)
    res = H.oracle_machrep(p)
    for sd in (0, 1, 3):
        rc, n_exec, h, mr, _ = H.emu_simt_run(p, sd, 5, 1)
        assert rc == 0 and all(res[r].up_time_prop == mr[r].up_time_prop for r in range(10))


def test_bulk_selection_with_equal_time_runs(H):
    """The cooperative bulk selection (coop_select) on a model whose messages come in runs of one receive time: MachRep2 with
    unbounded optimism and a large batch lets the master hold far more than 32 pending events, many of them annihilated
    pairs.  Regression: a run of one time quantum may only stay in the selected prefix when ALL of it is annihilated and
    equal in time -- a pairwise check let an earlier live event of the same quantum be executed after a later no-op
    (TW_ERR_TIME_ORDER on the GPU and here)."""
    p = H.machrep_params(512, 3.0, horizon=-1.0, batch=16, cap=1024)      # (the pairwise check failed at LP 717 in step 17)
    res = H.oracle_machrep(p)
    rc, n_exec, h, mr, stats = H.emu_simt_run(p, 1, 1, 1)
    assert rc == 0 and stats.rolled_back > 0
    assert all(res[r].up_time_prop == mr[r].up_time_prop and res[r].immed_prop == mr[r].immed_prop for r in range(512))


# ---------------------------------------------------------------- closed groups: the warp-local super-step loop (k_run_groups)
@pytest.mark.parametrize("horizon,batch,gpw,budget,every", [(-1.0, 4, 2, 64, 1), (2.0, 4, 10, 8, 1), (0.0, 1, 1, 2, 1), (-1.0, 16, 3, 1000, 3),
                                                            (5.0, 8, 4, 2, 7)])
def test_closed_group_loop_machrep2(H, horizon, batch, gpw, budget, every):
    """MachRep2 replicas never talk to each other: a warp runs the super-steps of its groups back to back, every group
    under its own GVT (group_rounds in tw_core.cuh, the body of k_run_groups), for any number of groups per warp, any
    round budget per launch and a lagging external Time Server bound.  Statistics and traces must equal the oracle's."""
    p = H.machrep_params(7, 120.0, horizon=horizon, batch=batch, cap=1024)
    res = H.oracle_machrep(p)
    rc, n_exec, h, mr, stats = H.emu_groups_run(p, 3, gpw, budget, every)
    assert rc == 0
    for r in range(7):
        assert (res[r].total_up_time, res[r].n_rep, res[r].n_immed_rep) == (mr[r].total_up_time, mr[r].n_rep, mr[r].n_immed_rep)
        assert res[r].up_time_prop == mr[r].up_time_prop and res[r].immed_prop == mr[r].immed_prop
        assert [res[r].n_exec[k] for k in range(3)] == [int(n_exec[3 * r + k]) for k in range(3)]
        assert [res[r].hash[k] for k in range(3)] == [int(h[3 * r + k]) for k in range(3)]
    assert stats.steps % 2 == 0 and stats.executed - stats.rolled_back == stats.committed + 14     # + the 14 START events
    if horizon != 0.0:
        assert stats.rolled_back > 0


def test_closed_group_loop_machrep1_and_reference_configuration(H):
    p = H.machrep1_params(4, 300.0, horizon=5.0, batch=8, cap=1024)
    res = H.oracle_machrep1(p)
    rc, n_exec, h, mr, stats = H.emu_groups_run(p, 3, 2, 64, 1)
    assert rc == 0 and all(res[r].up_time_prop == mr[r].up_time_prop and res[r].n_rep == mr[r].n_rep for r in range(4))
    assert all(res[r].hash[k] == int(h[3 * r + k]) for r in range(4) for k in range(3)) and stats.rolled_back > 0
    p = H.machrep_params(1, 1000.0)                      # C1: tests/MachRep2DistributedReproducible.hs, no time horizon
    o = H.oracle_machrep(p)[0]
    rc, n_exec, h, mr, stats = H.emu_groups_run(p, 1, 1, 4096, 1)
    assert rc == 0 and o.up_time_prop == mr[0].up_time_prop and o.immed_prop == mr[0].immed_prop
    assert [o.hash[k] for k in range(3)] == [int(x) for x in h]


def test_simulation_retry_really_happens(H):
    """a19 (handleEventRetry / SimulationRetry, Internal/Event.hs:247-249, 768-813): with unbounded optimism and a large
    batch the master sees ReleaseRepairPerson before the matching request (resource released above its capacity); the
    event must stay pending until the straggler's rollback arrives -- and the results must still equal the oracle's."""
    p = H.machrep_params(64, 200.0, horizon=-1.0, batch=16, cap=1024)
    res = H.oracle_machrep(p)
    rc, n_exec, h, mr, stats = H.emu_run(p, 1)
    assert rc == 0 and stats.retries > 0
    assert all(res[r].up_time_prop == mr[r].up_time_prop and res[r].immed_prop == mr[r].immed_prop for r in range(64))
    rc, n_exec, h, mr, stats = H.emu_groups_run(H.machrep_params(8, 200.0, horizon=-1.0, batch=16, cap=1024), 2, 3, 256, 1)
    assert rc == 0 and stats.retries > 0
    assert all(res[r].up_time_prop == mr[r].up_time_prop and res[r].immed_prop == mr[r].immed_prop for r in range(8))
