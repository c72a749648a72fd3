"""Host-side mirror of the reference's DIO surface for the batched Time Warp path.

Names, argument meaning and error behaviour follow the reference (paths relative to /root/reference):
  DIOParams / defaultDIOParams   Simulation/Aivika/Distributed/Optimistic/Internal/DIO.hs:70-110, 196-217
  Specs                          aivika-transformers (spcStartTime / spcStopTime), used at tests/MachRep2DistributedReproducible.hs:64-68
  runDIO                         Internal/DIO.hs:327-328
  enqueueMessage / sendMessage   Optimistic/Message.hs:41-69
  timeServer / TimeServerParams  Internal/TimeServer.hs:47-72, 485
  LogicalProcessState            Optimistic/State.hs:27-54
The reference runs one Haskell closure per event; the device runs a *compiled* handler for a whole batch of
homogeneous LPs, so a run is described by a model descriptor (PHOLD / MachRep2 / MachRep1) instead of a closure.
Everything executes on the GPU through libtwb200.so (include/twb200.h); there is no CPU path here.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _ffi as ffi


class SimulationAbort(RuntimeError):
    """Raised when the engine reports a non-zero status (the reference raises `error`/SimulationAbort)."""

    def __init__(self, code, message):
        super().__init__(f"{ffi.ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


# ---------------------------------------------------------------- parameters (reference field names)
@dataclass
class DIOParams:
    """defaultDIOParams (Internal/DIO.hs:196-217); only the fields that touch the event-processing path are carried.

    How the reference's thresholds act here (isEventOverflow, Internal/Event.hs:289-306, throttles optimism when the
    undo log or the output queue of an LP grows past its threshold): an LP's undo log and output queue are its
    processed-but-uncommitted entries, bounded by `procLimit` (default: a quarter of the ring); a threshold below that
    lowers the bound (`make_params`).  There is no transient-message queue on the device (messages are delivered between
    two super-steps, none is in flight at a GVT cut), so dioTransientMessageQueueSizeThreshold has nothing to bound.
    dioSyncTimeout bounds how long a GPU waits for its peers' GVT mailbox entries (reference: the synchronisation
    timeout of blocking operations, in microseconds)."""
    dioLoggingPriority: str = "WARNING"
    dioName: str = "LP"
    dioTimeHorizon: Optional[float] = None                  # Nothing = unbounded optimism
    dioUndoableLogSizeThreshold: int = 10_000_000
    dioOutputMessageQueueSizeThreshold: int = 10_000
    dioTransientMessageQueueSizeThreshold: int = 5_000
    dioSyncTimeout: int = 60_000_000
    dioAllowPrematureIO: bool = False                       # results are only read at GVT == stop: never premature
    dioAllowSkippingOutdatedMessage: bool = True            # Internal/DIO.hs:206; False: TW_ERR_OUTDATED_MESSAGE
    # engine-specific knobs (no reference counterpart)
    queueCapacity: int = 128
    maxEventsPerStep: int = 4
    stepsPerLaunch: int = 16
    procLimit: int = 0
    fossilThreshold: int = 0
    outboxCapacity: int = 0                                 # multi-GPU staging ring records (0 = engine default)
    closedGroupKernel: bool = True                          # MachRep1/2: warp-local super-step loop (k_run_groups)


defaultDIOParams = DIOParams()


@dataclass
class Specs:
    spcStartTime: float = 0.0
    spcStopTime: float = 1000.0
    spcDT: float = 1.0


@dataclass
class TimeServerParams:
    tsLoggingPriority: str = "WARNING"
    tsName: str = "Time Server"
    tsTimeSyncDelay: int = 100_000      # the device engine reduces GVT every super-step instead


defaultTimeServerParams = TimeServerParams()


@dataclass
class TimeServer:
    """The Time Server role (Internal/TimeServer.hs:485): quorum of LPs + the global virtual time.

    With a single device partition the role collapses into the in-kernel reduction (`globalTime`).  With several
    registered logical processes -- device partitions (DIO objects after `attachTimeServer`) or CPU processes, anything
    with `localTime()` and `globalTimeMessage(t)` -- `computeGlobalTime` is one round of the reference's protocol:
    ComputeLocalTimeMessage to everybody, LocalTimeMessage back, minimum, GlobalTimeMessage to everybody
    (computeTimeServerGlobalTime / timeServerGlobalTime / provideTimeServerGlobalTime, Internal/TimeServer.hs:369-421)."""
    quorum: int
    params: TimeServerParams = field(default_factory=TimeServerParams)
    _dio: Optional["DIO"] = None
    _processes: list = field(default_factory=list)
    _global_time: Optional[float] = None

    def globalTime(self):
        if self._processes:
            return self._global_time
        if self._dio is None:
            return None
        return self._dio.globalTime()

    def register(self, lp):
        """RegisterLogicalProcessMessage (Internal/TimeServer.hs:196-222)."""
        self._processes.append(lp)

    def computeGlobalTime(self):
        """One GVT round; returns the global time, or None while the quorum is not met (:311-326)."""
        if len(self._processes) < self.quorum:
            return None
        local = [lp.localTime() for lp in self._processes]
        t0 = min(local)
        self._global_time = t0
        for lp in self._processes:
            lp.globalTimeMessage(t0)
        return t0


def timeServer(n, params=defaultTimeServerParams):
    return TimeServer(n, params)


# ---------------------------------------------------------------- model descriptors (compiled handlers)
@dataclass
class PHOLD:
    n_lps: int
    population: int = 16
    remote_pct: int = 10
    lookahead: float = 0.1
    mean: float = 1.0
    seed: int = 1
    n_hot: int = 0
    hot_pct: int = 0

    def fill(self, p):
        p.model = ffi.TW_MODEL_PHOLD; p.n_lps = self.n_lps
        p.phold.population = self.population; p.phold.remote_pct = self.remote_pct
        p.phold.n_hot = self.n_hot; p.phold.hot_pct = self.hot_pct
        p.phold.lookahead = self.lookahead; p.phold.mean = self.mean; p.phold.seed = self.seed


@dataclass
class MachRep2:
    """tests/MachRep2DistributedReproducible.hs replicated `n_replicas` times (LP 3r master, 3r+1/3r+2 slaves)."""
    n_replicas: int = 1
    meanUpTime: float = 1.0
    meanRepairTime: float = 0.5
    delta: float = 1e-6
    seeds: tuple = (456, 789)
    seed_stride: int = 1000

    def fill(self, p):
        p.model = ffi.TW_MODEL_MACHREP2; p.n_lps = 3 * self.n_replicas
        p.machrep.mean_up_time = self.meanUpTime; p.machrep.mean_repair_time = self.meanRepairTime
        p.machrep.delta = self.delta; p.machrep.seed1 = self.seeds[0]; p.machrep.seed2 = self.seeds[1]
        p.machrep.seed_stride = self.seed_stride


@dataclass
class MachRep1(MachRep2):
    """tests/MachRep1.hs replicated `n_replicas` times: two machines, two repair persons (LP 3r master, 3r+1/3r+2 slaves).

    The second model of the reference's tests compiled through the handler table (csrc/engine.cu kModels); its slaves
    report the up time with `sendMessage` (receive time = send time, no lookahead).  `delta` is unused."""
    delta: float = 0.0

    def fill(self, p):
        super().fill(p)
        p.model = ffi.TW_MODEL_MACHREP1


@dataclass
class LogicalProcessState:
    lpStateId: int
    lpStateLocalTime: float
    lpStateEventQueueTime: float
    lpStateEventQueueSize: int
    lpStateLogSize: int
    lpStateInputMessageCount: int
    lpStateOutputMessageCount: int
    lpStateRollbackCount: int


def make_params(model, ps: DIOParams, specs: Specs, rank=0, world=1, device=-1, nccl_id=None):
    p = ffi.tw_params()
    model.fill(p)
    p.rank = rank; p.world = world; p.device = device
    p.queue_capacity = ps.queueCapacity; p.max_events_per_step = ps.maxEventsPerStep
    p.steps_per_launch = ps.stepsPerLaunch; p.fossil_threshold = ps.fossilThreshold
    # isEventOverflow: the undo log / output queue of an LP = its processed-but-uncommitted entries (<= 3 messages each)
    sends = 1 if isinstance(model, PHOLD) else 3
    limit = ps.procLimit if ps.procLimit else max(1, ps.queueCapacity // 4)
    limit = min(limit, max(1, ps.dioUndoableLogSizeThreshold), max(1, ps.dioOutputMessageQueueSizeThreshold // sends))
    p.proc_limit = 0 if (not ps.procLimit and limit == max(1, ps.queueCapacity // 4)) else limit
    p.outbox_capacity = ps.outboxCapacity
    p.sync_timeout_ms = max(1, min(ps.dioSyncTimeout // 1000, 0xFFFFFFFF))        # microseconds -> milliseconds
    p.flags = (ffi.TW_FLAG_SKIP_OUTDATED if ps.dioAllowSkippingOutdatedMessage else 0) | \
              (0 if ps.closedGroupKernel else ffi.TW_FLAG_NO_GROUP_KERNEL)
    p.start_time = specs.spcStartTime; p.stop_time = specs.spcStopTime
    p.time_horizon = -1.0 if ps.dioTimeHorizon is None else float(ps.dioTimeHorizon)
    if nccl_id is not None:
        C.memmove(p.nccl_id, bytes(nccl_id), 128)
    return p


class DIO:
    """A batch of logical processes living on one GPU partition (the device-side `EventQueue DIO`s)."""

    def __init__(self, model, ps: DIOParams = defaultDIOParams, specs: Specs = Specs(), rank=0, world=1, device=-1,
                 nccl_id=None):
        self.lib = ffi.load()
        self.model, self.ps, self.specs = model, ps, specs
        self.params = make_params(model, ps, specs, rank, world, device, nccl_id)
        h = C.c_void_p()
        rc = self.lib.tw_create(C.byref(self.params), C.byref(h))
        if rc != 0:
            raise SimulationAbort(rc, self.lib.tw_last_error(None).decode())
        self.h = h
        base, n = C.c_uint32(), C.c_uint32()
        self.lib.tw_partition(self.h, C.byref(base), C.byref(n))
        self.lp_base, self.n_local = base.value, n.value

    # -- lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.tw_destroy(self.h); self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SimulationAbort(rc, self.lib.tw_last_error(self.h).decode())

    def reset(self):
        self._check(self.lib.tw_reset(self.h))

    # -- messages
    def initialEvents(self, out=None):
        """Synthetic start-time messages addressed to this partition's LPs, as a host record array.

        `out` may be a preallocated (e.g. pinned) EVENT_DTYPE array; otherwise a numpy array is allocated."""
        per_lp = self.model.population if isinstance(self.model, PHOLD) else 1
        cap = max(1, self.n_local * per_lp)
        buf = np.zeros(cap, dtype=EVENT_DTYPE) if out is None else out
        assert buf.dtype == EVENT_DTYPE and len(buf) >= cap
        n = C.c_size_t()
        rc = self.lib.tw_model_initial_events_range(C.byref(self.params), self.lp_base, self.lp_base + self.n_local,
                                                    buf.ctypes.data, len(buf), C.byref(n))
        self._check(rc)
        return buf[: n.value]

    def runEventInStartTime(self):
        """Every LP enqueues its model's own start-time events on the device (tw_populate): what the reference's programs do
        with `runEventInStartTime` / `runProcessInStartTime` on each LP.  Equivalent to enqueueMessages(initialEvents())."""
        self._check(self.lib.tw_populate(self.h))

    def enqueueMessages(self, events: np.ndarray):
        """QueueMessageBulk: `events` is a host array of EVENT_DTYPE records."""
        assert events.dtype == EVENT_DTYPE and events.flags["C_CONTIGUOUS"]
        self._check(self.lib.tw_enqueue_messages(self.h, events.ctypes.data, len(events)))

    def enqueueMessage(self, pid, t, src=None, send_time=None, ctr=0, msg_type=0, value=0.0):
        """enqueueMessage pid t a (Optimistic/Message.hs:48-69), issued from the start time by the host."""
        ev = ffi.tw_event()
        ev.recv_ts = t; ev.send_ts = self.specs.spcStartTime if send_time is None else send_time
        ev.ctr = ctr; ev.src = pid if src is None else src; ev.dst = pid; ev.type = msg_type; ev.value = value
        self._check(self.lib.tw_enqueue_message(self.h, C.byref(ev)))

    def sendMessage(self, pid, **kw):
        """sendMessage pid a = enqueueMessage pid <now> a (Optimistic/Message.hs:41-45)."""
        self.enqueueMessage(pid, self.globalTime(), **kw)

    # -- running
    def runEventInStopTime(self, until=None):
        st = ffi.tw_run_stats()
        t = self.specs.spcStopTime if until is None else until
        self._check(self.lib.tw_run(self.h, t, C.byref(st)))
        return st

    def globalTime(self):
        g = C.c_double()
        self._check(self.lib.tw_gvt(self.h, C.byref(g)))
        return g.value

    # -- the LP side of the Time Server protocol (Internal/TimeServer.hs; tw_attach_time_server & co.)
    def attachTimeServer(self, on=True):
        """registerDIO: GVT is bounded by the Time Server's GlobalTimeMessage from now on (resets the run)."""
        self._check(self.lib.tw_attach_time_server(self.h, 1 if on else 0))

    def localTime(self):
        """The answer to ComputeLocalTimeMessage: the partition's local minimum time (LocalTimeMessage)."""
        t = C.c_double()
        self._check(self.lib.tw_local_time(self.h, C.byref(t)))
        return t.value

    def skippedOutdatedMessages(self):
        """Messages dropped because their receive time was below the global time (dioAllowSkippingOutdatedMessage)."""
        n = C.c_uint64()
        self._check(self.lib.tw_skipped_outdated(self.h, C.byref(n)))
        return n.value

    def maxRingOccupancy(self):
        """Largest number of ring slots an LP with a long pending region had in use (0: no such LP); see tw_max_ring_occupancy."""
        n = C.c_uint32()
        self._check(self.lib.tw_max_ring_occupancy(self.h, C.byref(n)))
        return n.value

    def globalTimeMessage(self, t):
        """GlobalTimeMessage t -> updateGlobalTime; returns False when the value was ignored (above the local time)."""
        ok = C.c_int()
        self._check(self.lib.tw_global_time_message(self.h, t, C.byref(ok)))
        return bool(ok.value)

    def logicalProcessState(self, lp):
        s = ffi.tw_lp_state()
        self._check(self.lib.tw_lp_state_get(self.h, lp, C.byref(s)))
        return LogicalProcessState(s.lp, s.local_time, s.event_queue_time, s.event_queue_size, s.log_size,
                                   s.input_message_count, s.output_message_count, s.rollback_count)

    def setMonitoringAction(self, action):
        """dioSimulationMonitoringAction (Internal/DIO.hs:453-476): `action(state_dict)` is called every time the
        engine inspects the device control block during runEventInStopTime (TimeServerState + LP counters)."""
        def _cb(ptr, _user):
            st = ptr.contents
            action({k: getattr(st, k) for k, _ in ffi.tw_monitor_state._fields_})
        self._monitor_cb = ffi.MONITOR_FN(_cb)           # keep the trampoline alive
        self._check(self.lib.tw_set_monitor(self.h, self._monitor_cb, None))

    # -- results
    def lpStats(self, out=None):
        """Per-LP committed-event counts and trace hashes.  `out` = (n_exec, hash) preallocated uint64 arrays of at least
        n_local entries (e.g. views of pinned memory: the device -> host copy then needs no staging)."""
        if out is None:
            n_exec = np.zeros(self.n_local, dtype=np.uint64); h = np.zeros(self.n_local, dtype=np.uint64)
        else:
            n_exec, h = out
            assert n_exec.dtype == np.uint64 and h.dtype == np.uint64 and len(n_exec) >= self.n_local and len(h) >= self.n_local
        self._check(self.lib.tw_read_lp_stats(self.h, n_exec.ctypes.data, h.ctypes.data, self.n_local))
        return n_exec[:self.n_local], h[:self.n_local]

    def machRepResults(self):
        first_master = ((self.lp_base + 2) // 3) * 3
        end = self.lp_base + self.n_local
        n = max(0, (end - first_master + 2) // 3) if first_master < end else 0
        out = np.zeros(n, dtype=MACHREP_DTYPE)
        if n:
            self._check(self.lib.tw_read_machrep(self.h, out.ctypes.data, n))
        return out


EVENT_DTYPE = np.dtype([("recv_ts", "<f8"), ("send_ts", "<f8"), ("ctr", "<u8"), ("src", "<u4"), ("dst", "<u4"),
                        ("type", "<u4"), ("pad", "<u4"), ("value", "<f8")])
MACHREP_DTYPE = np.dtype([("total_up_time", "<f8"), ("n_rep", "<i8"), ("n_immed_rep", "<i8"),
                          ("up_time_prop", "<f8"), ("immed_prop", "<f8")])
assert EVENT_DTYPE.itemsize == C.sizeof(ffi.tw_event) == 48
assert MACHREP_DTYPE.itemsize == C.sizeof(ffi.tw_machrep_result) == 40


def runDIO(model, ps: DIOParams = defaultDIOParams, timeServerId: Optional[TimeServer] = None, specs: Specs = Specs(),
           **partition):
    """runDIO m ps timeServerId (Internal/DIO.hs:327-328): returns (dio, simulation).

    `simulation()` registers with the time server, loads the model's start-time messages, runs to the stop time
    (runEventInStopTime) and returns (run statistics, per-LP committed counts, per-LP trace hashes[, MachRep2 results]).
    """
    dio = DIO(model, ps, specs, **partition)
    if timeServerId is not None:
        timeServerId._dio = dio

    def simulation():
        dio.enqueueMessages(dio.initialEvents())
        st = dio.runEventInStopTime()
        n_exec, h = dio.lpStats()
        if isinstance(model, MachRep2):
            return st, n_exec, h, dio.machRepResults()
        return st, n_exec, h

    return dio, simulation


# ---------------------------------------------------------------- multi-rank host logic (one process per GPU)
def partition_range(model, rank, world):
    """LP range [lo, hi) owned by `rank`: contiguous blocks of ceil(n/world) LPs, MachRep2 replicas kept whole.
    Must agree with Engine::init in csrc/engine.cu (checked by tests on both sides)."""
    n = model.n_lps if isinstance(model, PHOLD) else 3 * model.n_replicas
    n_per = (n + world - 1) // world
    if isinstance(model, MachRep2) and world > 1:
        n_per = ((n_per + 2) // 3) * 3
    lo = min(n, n_per * rank)
    return lo, min(n, lo + n_per)


def share_nccl_id(dist, rank, device=None):
    """Rank 0 creates the ncclUniqueId (tw_nccl_unique_id) and broadcasts its 128 bytes over torch.distributed."""
    import torch
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        rc = ffi.load().tw_nccl_unique_id(C.byref(buf))
        if rc != 0:
            raise SimulationAbort(rc, "ncclGetUniqueId failed")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=device or "cpu")
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def reduce_run(dist, world, device_ms, e2e_s, counters, device=None):
    """bench.py's cross-rank reduction: MAX over ranks of the per-step times, SUM of the event counters."""
    import torch
    tms = torch.tensor(device_ms, dtype=torch.float64, device=device or "cpu")
    te2e = torch.tensor(e2e_s, dtype=torch.float64, device=device or "cpu")
    tcm = torch.tensor(counters, dtype=torch.int64, device=device or "cpu")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(te2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(tcm, op=dist.ReduceOp.SUM)
    return tms.tolist(), te2e.tolist(), [int(x) for x in tcm.tolist()]
