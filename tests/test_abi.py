"""The C-ABI library loads, exports every symbol include/twb200.h declares, and its struct layouts are the ones
the Python binding assumes.  No compute calls: this runs without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(H):
    hdr = open(os.path.join(ROOT, "include", "twb200.h")).read()
    declared = set(re.findall(r"\b(tw_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(H.ffi.EXPORTS), declared ^ set(H.ffi.EXPORTS)
    lib = H.ffi.load()
    for s in declared:
        assert getattr(lib, s) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", H.ffi.LIB_PATH], capture_output=True, text=True).stdout
    for s in declared:
        assert re.search(rf"\bT {s}\b", out), s


def test_struct_layouts_match_the_header(H, tmp_path):
    src = tmp_path / "sz.c"
    names = ["tw_event", "tw_phold_params", "tw_machrep_params", "tw_params", "tw_run_stats", "tw_lp_state", "tw_machrep_result"]
    src.write_text('#include <stdio.h>\n#include "twb200.h"\nint main(){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) +
                   'printf("%zu %zu\\n", offsetof(tw_params, phold), offsetof(tw_params, nccl_id));return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    for n, sz in zip(names, lines):
        assert C.sizeof(getattr(H.ffi, n)) == int(sz), n
    a, b = lines[len(names)].split()
    assert H.ffi.tw_params.phold.offset == int(a) and H.ffi.tw_params.nccl_id.offset == int(b)


def test_version_and_error_strings(H):
    lib = H.ffi.load()
    assert b"sm_100a" in lib.tw_version()
    assert b"Send time" in lib.tw_strerror(3) and b"outdated" in lib.tw_strerror(4)
    assert lib.tw_kernel_launch_count() >= 0


def test_library_is_sm100a_only(H):
    out = subprocess.run(["cuobjdump", "-lelf", H.ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_no_gpu_means_loud_failure_not_fallback(H):
    """Without a usable sm_100 device tw_create must fail with TW_ERR_CUDA (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = H.ffi.load()
    p = H.phold_params(64, 1.0)
    h = C.c_void_p()
    rc = lib.tw_create(C.byref(p), C.byref(h))
    assert rc == 2 and not h.value
    assert b"no CPU fallback" in lib.tw_last_error(None) or b"CUDA" in lib.tw_last_error(None)
    from aivika_distributed_b200 import dio as D
    with pytest.raises(D.SimulationAbort):
        D.DIO(D.PHOLD(64))


def test_invalid_arguments(H):
    lib = H.ffi.load()
    h = C.c_void_p()
    assert lib.tw_create(None, C.byref(h)) == 1
    p = H.phold_params(64, 1.0); p.model = 99
    assert lib.tw_create(C.byref(p), C.byref(h)) == 1
    p = H.machrep_params(1); p.n_lps = 4
    assert lib.tw_create(C.byref(p), C.byref(h)) == 1
    assert lib.tw_run(None, 1.0, None) == 1


def test_haskell_binding_constants_match_the_abi(H):
    """aivika-distributed_b200/hs/...B200.hs cannot be compiled here (no GHC); at least its struct sizes, the field
    offsets it pokes and the imported symbols are checked against the real ABI."""
    src = open(os.path.join(H.PKG_DIR, "hs", "Simulation", "Aivika", "Distributed", "Optimistic", "B200.hs")).read()
    m = re.search(r"sizeofParams = (\d+); sizeofEvent = (\d+); sizeofRunStats = (\d+); sizeofMachRep = (\d+)", src)
    f = H.ffi
    assert [int(x) for x in m.groups()] == [C.sizeof(f.tw_params), C.sizeof(f.tw_event), C.sizeof(f.tw_run_stats), C.sizeof(f.tw_machrep_result)]
    for sym in re.findall(r'foreign import ccall \w+ "(tw_\w+)"', src):
        assert sym in f.EXPORTS
    P = f.tw_params
    expect = {"queueCapacity": P.queue_capacity.offset, "maxEventsPerStep": P.max_events_per_step.offset,
              "stepsPerLaunch": P.steps_per_launch.offset, "spcStartTime": P.start_time.offset, "spcStopTime": P.stop_time.offset,
              "pholdPopulation": P.phold.offset + f.tw_phold_params.population.offset,
              "pholdSeed": P.phold.offset + f.tw_phold_params.seed.offset,
              "mrMeanUpTime": P.machrep.offset + f.tw_machrep_params.mean_up_time.offset,
              "mrSeedStride": P.machrep.offset + f.tw_machrep_params.seed_stride.offset}
    for name, off in expect.items():
        assert re.search(rf"pokeByteOff p {off}\s+{name}\b", src), (name, off)
