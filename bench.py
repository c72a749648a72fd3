#!/usr/bin/env python
"""bench.py -- committed events/sec of the Time Warp hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--workload c4|c3|c5|c2] [--impl reference]

A "step" is one complete run of the synthetic workload (load the start-time messages, advance until GVT passes
the stop time, read the per-LP results).  `value` is device-timed (CUDA events inside tw_run, inputs resident in
HBM); `e2e` is the same metric through the public API with HOST buffers (pinned host -> device copy of the
start-time messages and device -> host read of the per-LP results inside the timed region).
For N > 1 launch with torchrun: one rank per GPU, the LP range is sharded, messages to remote LPs are NVLink peer
stores from inside the step kernel into staging rings in the receiver's memory (no atomic crosses NVLink; the receiver
routes them with k_deliver) and GVT is a peer-memory mailbox round run by the step kernel's last block.

--impl reference times the CPU restatement of the reference algorithm (oracle/, kind "port"; the Haskell engine
itself cannot be built in this image: no GHC) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))



def load_pkg():
    import importlib.util
    name = "aivika_distributed_b200"
    if name in sys.modules:
        return sys.modules[name]
    d = os.path.join(ROOT, "aivika-distributed_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load_pkg()
from aivika_distributed_b200 import workloads as W  # noqa: E402

WORKLOADS = W.WORKLOADS
ALGO_BYTES_PHOLD, ALGO_BYTES_MACHREP = W.ALGO_BYTES_PHOLD, W.ALGO_BYTES_MACHREP


def oracle_fixture(args, wl):
    """The oracle's answer for this exact workload (tests/golden/full_configs.json, written by
    tests/golden/make_golden_full.py), or None when the run departs from the benched configuration."""
    if args.seed != W.SEED or args.stop is not None or args.n_lps is not None or args.remote_pct is not None:
        return None
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "full_configs.json"))).get(args.workload)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_sample(wl, args):
    """CPU baseline on a bounded sample of the same workload: the multi-threaded Time Warp restatement of the
    reference algorithm (oracle/tw_ref.cpp) on all host threads.  Returns (events/s, threads, sample description)."""
    import helpers as H
    threads = H.twref().twref_hardware_threads()
    hz = wl["horizon"]
    if wl["kind"] == "phold":
        kw = dict(remote_pct=wl["remote_pct"], n_hot=wl.get("n_hot", 0), hot_pct=wl.get("hot_pct", 0), seed=args.seed, horizon=hz)
        t = time.perf_counter(); H.twref_phold(H.phold_params(wl["n_lps"], -1.0, **kw), threads); t_init = time.perf_counter() - t
        t = time.perf_counter(); rc, _, _, st = H.twref_phold(H.phold_params(wl["n_lps"], wl["cpu_stop"], **kw), threads)
        dt = time.perf_counter() - t
        assert rc == 0
        n = st.committed
        sample = (f"same workload truncated to stop time {wl['cpu_stop']} ({n} committed, {st.rolled_back} rolled back), "
                  f"LP/population set-up ({t_init:.1f} s) subtracted")
    else:
        nrep = min(wl["n_replicas"], 512)
        t_init = 0.0
        t = time.perf_counter(); rc, _, _, _, st = H.twref_machrep(H.machrep_params(nrep, wl["cpu_stop"], horizon=hz), threads)
        dt = time.perf_counter() - t
        assert rc == 0
        n = st.committed
        sample = f"first {nrep} replicas of the same workload ({n} committed, {st.rolled_back} rolled back)"
    cpu_sample.last_ms = (dt - t_init) * 1e3
    desc = (sample + f"; C++ Time Warp restatement of the reference algorithm (oracle/tw_ref.cpp), {threads} threads; "
            "the Haskell engine cannot be built here (no GHC)")
    return n / max(dt - t_init, 1e-9), threads, desc


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    vals = []
    sample, threads = "", 1
    for i in range(args.warmup + args.steps):
        v, threads, sample = cpu_sample(wl, args)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "committed events/sec", "value": v, "unit": "events/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu_sample.last_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "time_horizon": wl["horizon"],
                   "max_events_per_lp_per_superstep": args.batch, "parallelism": f"{threads} host threads"},
        "cpu_baseline": {"value": v, "unit": "events/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--horizon", type=float, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--cap", type=int, default=None)
    ap.add_argument("--stop", type=float, default=None)
    ap.add_argument("--n-lps", type=int, default=None, help="diagnostic: override the LP count of a PHOLD workload")
    ap.add_argument("--remote-pct", type=int, default=None, help="diagnostic: override the remote share of a PHOLD workload")
    ap.add_argument("--steps-per-launch", type=int, default=16)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-group-kernel", action="store_true",
                    help="diagnostic: run MachRep2 through grid-wide super-steps (k_step) instead of the closed-group kernel")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.stop is not None:
        wl["stop"] = args.stop
    if args.horizon is not None:
        wl["horizon"] = args.horizon
    if args.cap is not None:
        wl["cap"] = args.cap
    if args.batch is None:
        args.batch = wl["batch"]
    if args.n_lps is not None or args.remote_pct is not None:
        if args.n_lps is not None:
            wl["n_lps"] = args.n_lps
        if args.remote_pct is not None:
            wl["remote_pct"] = args.remote_pct
        wl["desc"] += f" [diagnostic override: n_lps={wl['n_lps']}, remote_pct={wl['remote_pct']}]"

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = load_pkg()
    from aivika_distributed_b200 import dio as D
    lib = pkg._ffi.load()
    nccl_id = D.share_nccl_id(dist, rank, device="cuda") if world > 1 else None

    if wl["kind"] == "phold":
        model = D.PHOLD(n_lps=wl["n_lps"], population=16, remote_pct=wl["remote_pct"], lookahead=0.1, mean=1.0,
                        seed=args.seed, n_hot=wl.get("n_hot", 0), hot_pct=wl.get("hot_pct", 0))
        algo_bytes = ALGO_BYTES_PHOLD
    else:
        model = D.MachRep2(n_replicas=wl["n_replicas"])
        algo_bytes = ALGO_BYTES_MACHREP
    ps = D.DIOParams(dioTimeHorizon=None if wl["horizon"] < 0 else wl["horizon"], queueCapacity=wl["cap"],
                     maxEventsPerStep=args.batch, stepsPerLaunch=args.steps_per_launch,
                     closedGroupKernel=not args.no_group_kernel)
    specs = D.Specs(0.0, wl["stop"])
    dio = D.DIO(model, ps, specs, rank=rank, world=world, device=local_rank, nccl_id=nccl_id)

    # synthetic start-time messages in PINNED host memory (torch is only the allocator here)
    per_lp = 16 if wl["kind"] == "phold" else 1
    n_init_cap = max(1, dio.n_local * per_lp)
    pinned = torch.empty(n_init_cap * 48, dtype=torch.uint8).pin_memory()
    ev_view = pinned.numpy().view(D.EVENT_DTYPE)
    events = dio.initialEvents(out=ev_view)
    h2d = int(events.nbytes); d2h = int(dio.n_local * 16 + C.sizeof(pkg._ffi.tw_run_stats))
    # the per-LP results land in pinned host memory as well (no pageable staging inside the timed region)
    res_pinned = torch.empty(2 * max(1, dio.n_local), dtype=torch.int64).pin_memory()
    res_view = res_pinned.numpy().view(np.uint64)
    res_out = (res_view[:max(1, dio.n_local)], res_view[max(1, dio.n_local):])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    dev_ms, e2e_s, committed, hashes, stats, digests = [], [], [], [], [], []
    launches0 = None
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            sampler.start()
            launches0 = lib.tw_kernel_launch_count()
        dio.reset()
        barrier()
        t0 = time.perf_counter()
        dio.enqueueMessages(events)            # pinned host -> device + k_route
        st = dio.runEventInStopTime()          # the timed device region (CUDA events inside tw_run)
        n_exec, h = dio.lpStats(out=res_out)   # device -> host (pinned)
        t1 = time.perf_counter()
        barrier()
        if i >= args.warmup:
            dev_ms.append(st.device_ms); e2e_s.append(t1 - t0); committed.append(int(st.committed))
            hashes.append(int(st.trace_hash_sum)); stats.append(st.as_dict())
            assert int(n_exec.sum()) == int(st.committed)
            digests.append(W.lp_digest(dio.lp_base, n_exec, h))
    launches = lib.tw_kernel_launch_count() - launches0
    max_occ = dio.maxRingOccupancy()
    # the same end-to-end step with the start-time events created on the device (tw_populate: every LP's own
    # runEventInStartTime), i.e. without shipping the synthetic population over PCIe -- reported beside `e2e`, not instead
    e2e_dev_s = []
    for i in range(args.steps):
        dio.reset()
        barrier()
        t0 = time.perf_counter()
        dio.runEventInStartTime()
        st2 = dio.runEventInStopTime()
        n2, h2 = dio.lpStats(out=res_out)
        t1 = time.perf_counter()
        barrier()
        assert int(st2.committed) == committed[0] and int(st2.trace_hash_sum) == hashes[0], "device-side population must give the same trace"
        e2e_dev_s.append(t1 - t0)
    sampler.stop_flag = True
    if sampler.is_alive():
        sampler.join(timeout=2)
    assert len(set(committed)) == 1 and len(set(hashes)) == 1 and len(set(digests)) == 1, "the committed trace must not depend on the run"
    rep_digest = 0
    if wl["kind"] == "machrep":
        res = dio.machRepResults()
        rep_digest = W.machrep_digest((dio.lp_base + 2) // 3, res)

    # max over ranks of the per-step device time, sum over ranks of the committed events
    _, e2e_dev_l, _ = D.reduce_run(dist, world, dev_ms, e2e_dev_s, [0], device="cuda")
    occ_l, _, _ = D.reduce_run(dist, world, [float(max_occ)], [0.0], [0], device="cuda")     # MAX over ranks
    ms_l, e2e_l, cnt = D.reduce_run(dist, world, dev_ms, e2e_s,
                                    [committed[0], stats[-1]["executed"], stats[-1]["rolled_back"], stats[-1]["anti_sent"],
                                     stats[-1]["sent_remote"], stats[-1]["sent_local"], launches,
                                     hashes[0] & 0xFFFFFFFF, hashes[0] >> 32, digests[0] & 0xFFFFFFFF, digests[0] >> 32,
                                     rep_digest & 0xFFFFFFFF, rep_digest >> 32], device="cuda")
    total_ms = float(sum(ms_l)); total_e2e = float(sum(e2e_l))
    tot_committed, tot_exec, tot_rb, tot_anti, tot_remote, tot_local, tot_launch, h_lo, h_hi, d_lo, d_hi, r_lo, r_hi = cnt
    M64 = (1 << 64) - 1
    trace_hash_total = ((h_hi << 32) + h_lo) & M64                   # sum over all LPs of the box, mod 2^64
    lp_digest_total = ((d_hi << 32) + d_lo) & M64
    rep_digest_total = ((r_hi << 32) + r_lo) & M64
    K = args.steps
    value = tot_committed * K / (total_ms / 1e3)
    e2e_value = tot_committed * K / total_e2e

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM bytes per launch of the dominant kernel: only from an ncu --set full capture of THIS workload at THIS
        # partition size (profiles/traffic.json, written by tools/profile_round.sh + tools/ncu_traffic.py); null otherwise
        traffic, traffic_src = None, None
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:lps{dio.n_local}:world{world}")
            if t and args.stop is None and args.n_lps is None and args.remote_pct is None:
                traffic, traffic_src = t["dram_bytes_per_launch"], t.get("source")
        except Exception:
            pass
        steps_per_run = stats[-1]["steps"]
        grouped = wl["kind"] == "machrep" and not args.no_group_kernel
        kern = "k_run_groups<MachRepModel>" if grouped else ("k_step<PholdModel>" if wl["kind"] == "phold" else "k_step<MachRepModel>")
        # launches of the dominant kernel in one run: one per super-step, or (closed-group kernel) a handful per run
        n_launch = max(1, stats[-1]["kernel_launches"] - 2) if grouped else max(1, steps_per_run)
        # parity with the CPU oracle at the benched size: committed count, trace-hash sum, LP-bound digest (+ MachRep2 results)
        fx = oracle_fixture(args, wl)
        if fx is None:
            parity = {"checked": False, "why": "diagnostic override: not one of the pinned configurations"}
        else:
            parity = {"checked": True, "fixture": "tests/golden/full_configs.json (oracle/oracle.c, sequential ground truth)",
                      "committed": tot_committed == fx["committed"], "trace_hash_sum": hex(trace_hash_total) == fx["trace_hash_sum"],
                      "lp_digest": hex(lp_digest_total) == fx["lp_digest"]}
            if wl["kind"] == "machrep":
                parity["replica_digest"] = hex(rep_digest_total) == fx["replica_digest"]
            bad = [k for k, v in parity.items() if v is False]
            assert not bad, f"PARITY FAILURE against the oracle fixture for {args.workload}: {bad} " \
                            f"(committed {tot_committed} vs {fx['committed']}, hash {hex(trace_hash_total)} vs {fx['trace_hash_sum']})"
        achieved = (tot_committed / world) * algo_bytes * K / (total_ms / 1e3) / 1e9
        line = {
            "metric": "committed events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "time_horizon": wl["horizon"],
                       "max_events_per_lp_per_superstep": args.batch, "queue_capacity": wl["cap"],
                       "parallelism": f"lp-sharding x{world}",
                       "l2": "per-LP event rings (n_lps x capacity x 32 B) exceed the 126 MB L2 for c4/c5; c3's live set is L2-resident"},
            "committed_events_per_step": tot_committed, "executed_events_per_step": tot_exec,
            "rollback_efficiency": tot_committed / max(1, tot_exec),
            "rolled_back_per_step": tot_rb, "anti_messages_per_step": tot_anti,
            "anti_messages_per_sec": tot_anti * K / (total_ms / 1e3),
            "max_ring_occupancy": {"slots": int(occ_l[0]), "queue_capacity": wl["cap"],
                                   "note": "largest ring fill of an LP with more than 256 pending events (0: none had)"},
            "remote_messages_per_step": tot_remote, "local_messages_per_step": tot_local,
            "supersteps_per_step": steps_per_run, "trace_hash_sum": hex(trace_hash_total), "lp_digest": hex(lp_digest_total),
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world},
            "e2e_device_population": {"value": tot_committed * K / float(sum(e2e_dev_l)), "unit": "events/s", "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": d2h * world,
                                      "note": "same step with every LP creating its own start-time events on the device (tw_populate = "
                                              "runEventInStartTime); `e2e` ships them from pinned host memory as the contract asks"},
            "gpu_launches": tot_launch,
            "nvlink": None if world == 1 else {
                "achieved_gbs_per_gpu": tot_remote * (32 if wl["kind"] == "phold" else 48) * K / (total_ms / 1e3) / 1e9 / world,
                "peak_gbs_per_direction": 770.0,
                "note": "32 B (PHOLD) / 48 B (MachRep2) per message whose destination LP lives on another GPU, written as peer "
                        "stores from k_step into the receiver's staging ring (no atomic crosses NVLink); algorithmic bytes / "
                        "time, not a counter; peak = measured peer copy (B200_PROFILING.md)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kern,
                         "note": f"{kern}: {algo_bytes} algorithmic B per committed event (SURVEY.md 8d) x "
                                 f"{tot_committed / world / max(1, n_launch):.0f} events per launch, "
                                 f"{total_ms / K / max(1, n_launch) * 1e3:.2f} us per launch; peak = measured copy bandwidth"
                                 + ("" if peaks else " (fallback)")},
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            v, threads, sample = cpu_sample(wl, args)
            line["cpu_baseline"] = {"value": v, "unit": "events/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    dio.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
