"""Shared test helpers: load the package (its directory name has a hyphen), the CPU oracle and the emulator."""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "aivika-distributed_b200")


def load_pkg():
    """import aivika-distributed_b200/ as module `aivika_distributed_b200`."""
    name = "aivika_distributed_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load_pkg()
ffi = pkg._ffi


def ensure_built():
    """The CUDA library is a build artefact (git-ignored): compile it when a fresh checkout runs the CPU suite."""
    if not os.path.exists(ffi.LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(PKG_DIR, "csrc"), "libtwb200.so"], check=True, stdout=subprocess.DEVNULL)


ensure_built()


# ------------------------------------------------------------------ oracle (plain C, test infrastructure)
class or_event(C.Structure):
    _fields_ = ffi.tw_event._fields_


class or_phold_params(C.Structure):
    _fields_ = [("n_lps", C.c_uint32), ("population", C.c_uint32), ("lookahead", C.c_double), ("mean", C.c_double),
                ("stop_time", C.c_double), ("remote_pct", C.c_uint32), ("n_hot", C.c_uint32), ("hot_pct", C.c_uint32),
                ("pad", C.c_uint32), ("seed", C.c_uint64)]


class or_summary(C.Structure):
    _fields_ = [("committed", C.c_uint64), ("hash_sum", C.c_uint64), ("remote", C.c_uint64), ("last_time", C.c_double)]


class or_trace_rec(C.Structure):
    _fields_ = [("dst", C.c_uint32), ("src", C.c_uint32), ("ctr", C.c_uint64), ("recv_bits", C.c_uint64)]


class or_machrep_params(C.Structure):
    _fields_ = [("n_replicas", C.c_uint32), ("pad", C.c_uint32), ("mean_up_time", C.c_double),
                ("mean_repair_time", C.c_double), ("delta", C.c_double), ("stop_time", C.c_double),
                ("seed1", C.c_int64), ("seed2", C.c_int64), ("seed_stride", C.c_int64)]


class or_machrep_result(C.Structure):
    _fields_ = [("total_up_time", C.c_double), ("n_rep", C.c_int64), ("n_immed_rep", C.c_int64),
                ("up_time_prop", C.c_double), ("immed_prop", C.c_double), ("n_exec", C.c_uint64 * 3),
                ("hash", C.c_uint64 * 3)]


def _make(target_dir, target):
    subprocess.run(["make", "-s", "-C", target_dir, target], check=True, stdout=subprocess.DEVNULL)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        d = os.path.join(ROOT, "oracle")
        so = os.path.join(d, "liboracle.so")
        if not os.path.exists(so) or any(os.path.getmtime(os.path.join(d, f)) > os.path.getmtime(so) for f in ("oracle.c", "oracle_math.h")):
            _make(d, "liboracle.so")
        L = C.CDLL(os.path.join(d, "liboracle.so"))
        L.oracle_log.restype = C.c_double; L.oracle_log.argtypes = [C.c_double]
        L.oracle_u01.restype = C.c_double; L.oracle_u01.argtypes = [C.c_uint64]
        L.oracle_draw.restype = C.c_uint64; L.oracle_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.oracle_trace.restype = C.c_uint64; L.oracle_trace.argtypes = [C.c_uint64, C.c_double, C.c_uint32, C.c_uint64]
        L.oracle_stdgen_doubles.argtypes = [C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_stdgen_next.argtypes = [C.c_int64, C.c_uint32, C.c_void_p]
        L.oracle_phold_run.argtypes = [C.POINTER(or_phold_params), C.c_void_p, C.c_void_p, C.POINTER(or_summary),
                                       C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_phold_initial_events.argtypes = [C.POINTER(or_phold_params), C.c_void_p]
        L.oracle_machrep_run.argtypes = [C.POINTER(or_machrep_params), C.c_uint32, C.c_uint32, C.c_void_p,
                                         C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_machrep1_run.argtypes = [C.POINTER(or_machrep_params), C.c_uint32, C.c_uint32, C.c_void_p]
        _oracle = L
    return _oracle


_emu = None


def emu():
    """CPU build of the engine's device functions (tests/emu/emu.cpp): debugging aid, not a product path."""
    global _emu
    if _emu is None:
        d = os.path.join(ROOT, "tests", "emu")
        so = os.path.join(d, "libtwemu.so")
        srcs = [os.path.join(d, "emu.cpp")] + [os.path.join(PKG_DIR, "csrc", f) for f in ("tw_core.cuh", "tw_models.cuh", "tw_math.cuh", "tw_dsl.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                            os.path.join(d, "emu.cpp"), "-o", so], check=True)
        L = C.CDLL(so)
        L.emu_run.argtypes = [C.POINTER(ffi.tw_params), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.POINTER(ffi.tw_run_stats)]
        L.emu_run.restype = C.c_int
        L.emu_run_lagged.argtypes = [C.POINTER(ffi.tw_params), C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(ffi.tw_run_stats)]
        L.emu_run_lagged.restype = C.c_int
        _emu = L
    return _emu


# ------------------------------------------------------------------ parameter builders
def phold_params(n_lps, stop, population=16, remote_pct=10, lookahead=0.1, mean=1.0, seed=1, n_hot=0, hot_pct=0,
                 horizon=-1.0, cap=128, batch=4, world=1, rank=0, steps_per_launch=16, proc_limit=0, fossil=0):
    p = ffi.tw_params()
    p.model = ffi.TW_MODEL_PHOLD; p.n_lps = n_lps; p.rank = rank; p.world = world; p.device = -1
    p.queue_capacity = cap; p.max_events_per_step = batch; p.steps_per_launch = steps_per_launch
    p.start_time = 0.0; p.stop_time = stop; p.time_horizon = horizon; p.proc_limit = proc_limit; p.fossil_threshold = fossil
    p.phold.population = population; p.phold.remote_pct = remote_pct; p.phold.n_hot = n_hot; p.phold.hot_pct = hot_pct
    p.phold.lookahead = lookahead; p.phold.mean = mean; p.phold.seed = seed
    return p


def machrep_params(n_replicas, stop=1000.0, horizon=-1.0, cap=64, batch=4, world=1, rank=0, steps_per_launch=32,
                   seed1=456, seed2=789, stride=1000, proc_limit=0, fossil=0):
    p = ffi.tw_params()
    p.model = ffi.TW_MODEL_MACHREP2; p.n_lps = 3 * n_replicas; p.rank = rank; p.world = world; p.device = -1
    p.queue_capacity = cap; p.max_events_per_step = batch; p.steps_per_launch = steps_per_launch
    p.start_time = 0.0; p.stop_time = stop; p.time_horizon = horizon; p.proc_limit = proc_limit; p.fossil_threshold = fossil
    p.machrep.mean_up_time = 1.0; p.machrep.mean_repair_time = 0.5; p.machrep.delta = 1e-6
    p.machrep.seed1 = seed1; p.machrep.seed2 = seed2; p.machrep.seed_stride = stride
    return p


def machrep1_params(n_replicas, stop=1000.0, horizon=-1.0, cap=64, batch=4, world=1, rank=0, steps_per_launch=32,
                    seed1=456, seed2=789, stride=1000):
    """tests/MachRep1.hs replicated (constants :39-40); the StdGen seeds follow the reproducible MachRep2 variant."""
    p = machrep_params(n_replicas, stop, horizon, cap, batch, world, rank, steps_per_launch, seed1, seed2, stride)
    p.model = ffi.TW_MODEL_MACHREP1
    p.machrep.delta = 0.0
    return p


def oracle_machrep1(p, r0=0, r1=None):
    n_rep = p.n_lps // 3
    r1 = n_rep if r1 is None else r1
    op = or_machrep_params(n_rep, 0, p.machrep.mean_up_time, p.machrep.mean_repair_time, p.machrep.delta,
                           p.stop_time, p.machrep.seed1, p.machrep.seed2, p.machrep.seed_stride)
    res = (or_machrep_result * (r1 - r0))()
    rc = oracle().oracle_machrep1_run(C.byref(op), r0, r1, res)
    assert rc == 0, rc
    return res


def oracle_phold(p):
    """Sequential ground truth for a tw_params PHOLD configuration -> (n_exec[], hash[], summary)."""
    import numpy as np
    op = or_phold_params(p.n_lps, p.phold.population, p.phold.lookahead, p.phold.mean, p.stop_time,
                         p.phold.remote_pct, p.phold.n_hot, p.phold.hot_pct, 0, p.phold.seed)
    n_exec = np.zeros(p.n_lps, dtype=np.uint64); h = np.zeros(p.n_lps, dtype=np.uint64)
    s = or_summary()
    rc = oracle().oracle_phold_run(C.byref(op), n_exec.ctypes.data, h.ctypes.data, C.byref(s), None, 0, None)
    assert rc == 0
    return n_exec, h, s


def oracle_machrep(p, r0=0, r1=None):
    n_rep = p.n_lps // 3
    r1 = n_rep if r1 is None else r1
    op = or_machrep_params(n_rep, 0, p.machrep.mean_up_time, p.machrep.mean_repair_time, p.machrep.delta,
                           p.stop_time, p.machrep.seed1, p.machrep.seed2, p.machrep.seed_stride)
    res = (or_machrep_result * (r1 - r0))()
    rc = oracle().oracle_machrep_run(C.byref(op), r0, r1, res, None, 0, None)
    assert rc == 0, rc
    return res


_emu_simt = None


def emu_simt():
    """The warp-level build of the same file (-DTW_SIMT_EMU): 32 host threads play the lanes of a warp (tests/emu/simt.hpp),
    so the device code paths -- work queue, deferred delivery, whole-warp search -- run on the CPU."""
    global _emu_simt
    if _emu_simt is None:
        d = os.path.join(ROOT, "tests", "emu")
        so = os.path.join(d, "libtwemu_simt.so")
        srcs = [os.path.join(d, "emu.cpp"), os.path.join(d, "simt.hpp")] + \
               [os.path.join(PKG_DIR, "csrc", f) for f in ("tw_core.cuh", "tw_models.cuh", "tw_math.cuh", "tw_dsl.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-DTW_SIMT_EMU",
                            "-x", "c++", os.path.join(d, "emu.cpp"), "-o", so], check=True)
        L = C.CDLL(so)
        L.emu_run_tiles.argtypes = [C.POINTER(ffi.tw_params), C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(ffi.tw_run_stats)]
        L.emu_run_tiles.restype = C.c_int
        L.emu_run_groups.argtypes = [C.POINTER(ffi.tw_params), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.POINTER(ffi.tw_run_stats)]
        L.emu_run_groups.restype = C.c_int
        assert L.emu_is_simt() == 1
        _emu_simt = L
    return _emu_simt


def emu_simt_run(p, shuffle_seed=1, gvt_every=1, tiles=1):
    import numpy as np
    n_exec = np.zeros(p.n_lps, dtype=np.uint64); h = np.zeros(p.n_lps, dtype=np.uint64)
    mr = (ffi.tw_machrep_result * max(1, p.n_lps // 3))()
    st = ffi.tw_run_stats()
    rc = emu_simt().emu_run_tiles(C.byref(p), shuffle_seed, gvt_every, tiles, n_exec.ctypes.data, h.ctypes.data, mr, C.byref(st))
    return rc, n_exec, h, mr, st


def emu_groups_run(p, shuffle_seed=1, gpw=2, budget=64, gvt_every=1):
    """Closed-group models through the warp-local super-step loop (k_run_groups) on the warp-level emulation."""
    import numpy as np
    n_exec = np.zeros(p.n_lps, dtype=np.uint64); h = np.zeros(p.n_lps, dtype=np.uint64)
    mr = (ffi.tw_machrep_result * max(1, p.n_lps // 3))()
    st = ffi.tw_run_stats()
    rc = emu_simt().emu_run_groups(C.byref(p), shuffle_seed, gpw, budget, gvt_every, n_exec.ctypes.data, h.ctypes.data, mr, C.byref(st))
    return rc, n_exec, h, mr, st


def emu_run(p, shuffle_seed=1, gvt_every=1):
    import numpy as np
    n_exec = np.zeros(p.n_lps, dtype=np.uint64); h = np.zeros(p.n_lps, dtype=np.uint64)
    mr = (ffi.tw_machrep_result * max(1, p.n_lps // 3))()
    st = ffi.tw_run_stats()
    if gvt_every == 1:
        rc = emu().emu_run(C.byref(p), shuffle_seed, n_exec.ctypes.data, h.ctypes.data, mr, C.byref(st))
    else:
        rc = emu().emu_run_lagged(C.byref(p), shuffle_seed, gvt_every, n_exec.ctypes.data, h.ctypes.data, mr, C.byref(st))
    return rc, n_exec, h, mr, st


# ------------------------------------------------------------------ CPU Time Warp restatement (oracle/tw_ref.cpp)
class twref_stats(C.Structure):
    _fields_ = [("committed", C.c_uint64), ("executed", C.c_uint64), ("rolled_back", C.c_uint64), ("rollbacks", C.c_uint64),
                ("anti_sent", C.c_uint64), ("annihilated", C.c_uint64), ("stragglers", C.c_uint64), ("gvt_rounds", C.c_uint64),
                ("gvt", C.c_double)]


_twref = None


def twref():
    global _twref
    if _twref is None:
        d = os.path.join(ROOT, "oracle")
        so = os.path.join(d, "libtwref.so")
        if not os.path.exists(so) or os.path.getmtime(os.path.join(d, "tw_ref.cpp")) > os.path.getmtime(so):
            _make(d, "libtwref.so")
        L = C.CDLL(so)
        L.twref_phold_run.argtypes = [C.POINTER(or_phold_params), C.c_int, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p,
                                      C.POINTER(twref_stats)]
        L.twref_machrep_run.argtypes = [C.POINTER(or_machrep_params), C.c_int, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.POINTER(twref_stats)]
        L.twref_hardware_threads.restype = C.c_int
        _twref = L
    return _twref


def twref_phold(p, threads, horizon=None, batch=4):
    import numpy as np
    op = or_phold_params(p.n_lps, p.phold.population, p.phold.lookahead, p.phold.mean, p.stop_time,
                         p.phold.remote_pct, p.phold.n_hot, p.phold.hot_pct, 0, p.phold.seed)
    n_exec = np.zeros(p.n_lps, dtype=np.uint64); h = np.zeros(p.n_lps, dtype=np.uint64)
    st = twref_stats()
    hz = p.time_horizon if horizon is None else horizon
    rc = twref().twref_phold_run(C.byref(op), threads, hz, batch, n_exec.ctypes.data, h.ctypes.data, C.byref(st))
    return rc, n_exec, h, st


def twref_machrep(p, threads, horizon=None, batch=4):
    import numpy as np
    n_rep = p.n_lps // 3
    op = or_machrep_params(n_rep, 0, p.machrep.mean_up_time, p.machrep.mean_repair_time, p.machrep.delta,
                           p.stop_time, p.machrep.seed1, p.machrep.seed2, p.machrep.seed_stride)
    n_exec = np.zeros(p.n_lps, dtype=np.uint64); h = np.zeros(p.n_lps, dtype=np.uint64)
    res = (ffi.tw_machrep_result * n_rep)()
    st = twref_stats()
    hz = p.time_horizon if horizon is None else horizon
    rc = twref().twref_machrep_run(C.byref(op), threads, hz, batch, n_exec.ctypes.data, h.ctypes.data, res, C.byref(st))
    return rc, n_exec, h, res, st
