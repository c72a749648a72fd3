"""Host mirror of the reference interface (dio.py): parameter mapping, names, partition arithmetic."""
import math

import numpy as np


def test_dioparams_defaults_follow_the_reference(H):
    from aivika_distributed_b200 import dio as D
    ps = D.defaultDIOParams
    # defaultDIOParams, Internal/DIO.hs:196-217 -- every field the mirror carries
    assert ps.dioLoggingPriority == "WARNING" and ps.dioName == "LP"
    assert ps.dioTimeHorizon is None and ps.dioUndoableLogSizeThreshold == 10_000_000
    assert ps.dioOutputMessageQueueSizeThreshold == 10_000 and ps.dioTransientMessageQueueSizeThreshold == 5_000
    assert ps.dioSyncTimeout == 60_000_000
    assert ps.dioAllowPrematureIO is False and ps.dioAllowSkippingOutdatedMessage is True       # :205-206
    assert D.defaultTimeServerParams.tsTimeSyncDelay == 100_000      # Internal/TimeServer.hs:140
    assert D.defaultTimeServerParams.tsLoggingPriority == "WARNING" and D.defaultTimeServerParams.tsName == "Time Server"


def test_dioparams_fields_reach_the_engine_parameters(H):
    """No decorative fields: what the mirror accepts is what tw_params carries."""
    from aivika_distributed_b200 import dio as D
    f = H.ffi
    m = D.PHOLD(1000)
    p = D.make_params(m, D.DIOParams(), D.Specs(0.0, 1.0))
    assert p.flags & f.TW_FLAG_SKIP_OUTDATED and not (p.flags & f.TW_FLAG_NO_GROUP_KERNEL)
    assert p.proc_limit == 0 and p.sync_timeout_ms == 60_000 and p.outbox_capacity == 0       # engine defaults, 60 s
    p = D.make_params(m, D.DIOParams(dioAllowSkippingOutdatedMessage=False, closedGroupKernel=False, dioSyncTimeout=5_000_000,
                                     outboxCapacity=4096), D.Specs(0.0, 1.0))
    assert p.flags == f.TW_FLAG_NO_GROUP_KERNEL and p.sync_timeout_ms == 5000 and p.outbox_capacity == 4096
    # isEventOverflow thresholds (Internal/Event.hs:289-306) bound the processed-but-uncommitted entries of an LP
    assert D.make_params(m, D.DIOParams(queueCapacity=1024, dioUndoableLogSizeThreshold=100), D.Specs()).proc_limit == 100
    assert D.make_params(m, D.DIOParams(queueCapacity=1024, dioOutputMessageQueueSizeThreshold=50), D.Specs()).proc_limit == 50
    assert D.make_params(D.MachRep2(2), D.DIOParams(queueCapacity=1024, dioOutputMessageQueueSizeThreshold=60), D.Specs()).proc_limit == 20
    assert D.make_params(m, D.DIOParams(queueCapacity=1024, procLimit=64, dioUndoableLogSizeThreshold=100), D.Specs()).proc_limit == 64
    assert D.make_params(m, D.DIOParams(queueCapacity=1024, procLimit=500, dioUndoableLogSizeThreshold=100), D.Specs()).proc_limit == 100


def test_benched_workloads_have_oracle_fixtures(H):
    """Every configuration bench.py can run is pinned to the oracle (tests/golden/full_configs.json), and the LP digest is
    additive over partitions (what lets N ranks digest their own slices)."""
    import json, os
    import numpy as np
    from aivika_distributed_b200 import workloads as W
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "full_configs.json")))
    for name, wl in W.WORKLOADS.items():
        assert name in fx and fx[name]["workload"] == wl["desc"], name
        assert fx[name]["committed"] > 0 and fx[name]["trace_hash_sum"].startswith("0x") and fx[name]["lp_digest"].startswith("0x")
    assert fx["c4"]["committed"] == 305102840 and abs(fx["c2"]["mean_up_time_prop"] - 0.6) < 0.01
    n_exec, h, s = H.oracle_phold(H.phold_params(1000, 3.0, remote_pct=50, seed=3))
    whole = W.lp_digest(0, n_exec, h)
    parts = sum(W.lp_digest(lo, n_exec[lo:hi], h[lo:hi]) for lo, hi in ((0, 334), (334, 668), (668, 1000))) & ((1 << 64) - 1)
    assert whole == parts and whole != W.lp_digest(0, n_exec[::-1], h[::-1])
    # the small golden fixture and the full-size one agree with a fresh oracle run where that is cheap (c1)
    r = H.oracle_machrep(H.machrep_params(1, 1000.0))[0]
    assert float(r.up_time_prop).hex() == fx["c1"]["replica0"]["up_time_prop"] and sum(r.n_exec) == fx["c1"]["committed"]


def test_make_params_maps_horizon_and_specs(H):
    from aivika_distributed_b200 import dio as D
    p = D.make_params(D.PHOLD(1000, remote_pct=33), D.DIOParams(dioTimeHorizon=None), D.Specs(0.0, 7.5))
    assert p.time_horizon < 0 and p.stop_time == 7.5 and p.n_lps == 1000 and p.phold.remote_pct == 33
    p = D.make_params(D.MachRep2(5), D.DIOParams(dioTimeHorizon=100.0), D.Specs(0.0, 1000.0), rank=1, world=2)
    assert p.time_horizon == 100.0 and p.n_lps == 15 and p.rank == 1 and p.world == 2
    assert (p.machrep.seed1, p.machrep.seed2, p.machrep.delta) == (456, 789, 1e-6)   # tests/MachRep2DistributedReproducible.hs:71-74


def test_event_record_dtype(H):
    from aivika_distributed_b200 import dio as D
    assert D.EVENT_DTYPE.itemsize == 48 and D.EVENT_DTYPE.fields["value"][1] == 40


def test_initial_events_ranges_tile_the_population(H):
    import ctypes as C
    lib = H.ffi.load()
    p = H.phold_params(100, 1.0, population=3)
    full = (H.ffi.tw_event * 300)(); n = C.c_size_t()
    lib.tw_model_initial_events(C.byref(p), full, 300, C.byref(n)); assert n.value == 300
    parts = b""
    for lo, hi in ((0, 34), (34, 68), (68, 100)):
        buf = (H.ffi.tw_event * 300)()
        lib.tw_model_initial_events_range(C.byref(p), lo, hi, buf, 300, C.byref(n))
        parts += bytes(buf)[: n.value * 48]
    assert parts == bytes(full)
    m = H.machrep_params(4)
    buf = (H.ffi.tw_event * 16)()
    lib.tw_model_initial_events(C.byref(m), buf, 16, C.byref(n))
    assert n.value == 8 and [buf[i].dst for i in range(8)] == [1, 2, 4, 5, 7, 8, 10, 11]


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(H):
    """aivika-distributed_b200/host/dio.hpp (C++ twin of the Haskell surface) builds against the C ABI; without a
    GPU the example must abort with the engine's error, not fall back to anything."""
    import os, subprocess, torch
    d = os.path.join(H.PKG_DIR, "host")
    subprocess.run(["make", "-s", "-C", d, "example_machrep2"], check=True)
    if torch.cuda.is_available():
        return
    r = subprocess.run([os.path.join(d, "example_machrep2")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
