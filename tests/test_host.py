"""Host mirror of the reference interface (dio.py): parameter mapping, names, partition arithmetic."""
import math

import numpy as np


def test_dioparams_defaults_follow_the_reference(H):
    from aivika_distributed_b200 import dio as D
    ps = D.defaultDIOParams
    # Internal/DIO.hs:196-217
    assert ps.dioTimeHorizon is None and ps.dioUndoableLogSizeThreshold == 10_000_000
    assert ps.dioOutputMessageQueueSizeThreshold == 10_000 and ps.dioTransientMessageQueueSizeThreshold == 5_000
    assert ps.dioSyncTimeout == 60_000_000
    assert D.defaultTimeServerParams.tsTimeSyncDelay == 100_000      # Internal/TimeServer.hs:140


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
