"""The BASELINE.json configurations as host-side descriptions (one place: bench.py, the parity tests and the script that
generates the oracle fixtures all read this table).

  c1  tests/MachRep2DistributedReproducible.hs itself: 3 LPs, seeds 456/789, stop 1000 (configs[0])
  c2  the same model replicated as 4096 independent 3-LP groups on one GPU (configs[1])
  c3  PHOLD 64K LPs, 10 % remote, lookahead 0.1 (configs[2])
  c4  PHOLD 1M LPs, 50 % remote (configs[3]) -- the configuration the headline metric is quoted on
  c5  rollback-heavy PHOLD 256K LPs: SURVEY.md 8(d) fan-out (half of the remote messages go to 1024 hot LPs), time
      horizon 0.5 = five lookaheads (configs[4])
`horizon`, `batch` and `cap` are engine operating points (they never change a committed trace: that is the invariant the
tests check); `cpu_stop` bounds the CPU-baseline sample.
"""

ALGO_BYTES_PHOLD = 208      # algorithmic bytes per committed PHOLD event
ALGO_BYTES_MACHREP = 288

WORKLOADS = {
    "c4": dict(desc="PHOLD 1M LPs, 50% remote, P=16, lookahead 0.1, exp(1) delays, T=20 (BASELINE configs[3])",
               kind="phold", n_lps=1 << 20, remote_pct=50, stop=20.0, horizon=0.3, batch=6, cap=128, cpu_stop=0.5),
    "c3": dict(desc="PHOLD 64K LPs, 10% remote, P=16, lookahead 0.1, exp(1) delays, T=100 (BASELINE configs[2])",
               kind="phold", n_lps=1 << 16, remote_pct=10, stop=100.0, horizon=0.2, batch=6, cap=128, cpu_stop=10.0),
    "c5": dict(desc="rollback-heavy PHOLD 256K LPs, 50% remote, half of the remote messages to 1024 hot LPs (each holds ~1000 "
                    "pending events), time horizon 0.5 = 5 lookaheads, T=20 (BASELINE configs[4], SURVEY.md 8d fan-out)",
               kind="phold", n_lps=1 << 18, remote_pct=50, n_hot=1024, hot_pct=50, stop=20.0, horizon=0.5, batch=64, cap=4096,
               cpu_stop=1.0),
    "c2": dict(desc="MachRep2DistributedReproducible replicated as 4096 independent 3-LP groups, stop 1000 (BASELINE configs[1])",
               kind="machrep", n_replicas=4096, stop=1000.0, horizon=1e-4, batch=4, cap=256, cpu_stop=1000.0),
    "c1": dict(desc="tests/MachRep2DistributedReproducible.hs: master + 2 slaves, seeds 456/789, stop 1000 (BASELINE configs[0])",
               kind="machrep", n_replicas=1, stop=1000.0, horizon=-1.0, batch=4, cap=64, cpu_stop=1000.0),
}

SEED = 1                    # PHOLD seed of every benched configuration

_M64 = (1 << 64) - 1


def lp_digest(lp_base, n_exec, trace_hash):
    """Order-independent digest that binds every LP's committed count and trace hash to its id: sum over LPs (mod 2^64)
    of fin(hash ^ fin(n_exec + C * (lp + 1))).  Partial sums of partitions add up, so N ranks can each digest their slice."""
    import numpy as np

    def fin(z):                                   # splitmix64 finaliser on uint64 arrays (wraps mod 2^64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        lp = np.arange(lp_base + 1, lp_base + 1 + len(n_exec), dtype=np.uint64)
        a = fin(np.asarray(n_exec, dtype=np.uint64) + lp * np.uint64(0x9E3779B97F4A7C15))
        d = fin(np.asarray(trace_hash, dtype=np.uint64) ^ a)
        return int(d.sum(dtype=np.uint64)) & _M64


def machrep_digest(first_replica, results):
    """The same for per-replica MachRep results: digest over the bit patterns of (total_up_time, n_rep, n_immed_rep)."""
    import numpy as np
    with np.errstate(over="ignore"):
        r = np.arange(first_replica + 1, first_replica + 1 + len(results), dtype=np.uint64)
        t = np.ascontiguousarray(results["total_up_time"]).view(np.uint64)
        a = t * np.uint64(0x9E3779B97F4A7C15) + np.asarray(results["n_rep"]).astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F) \
            + np.asarray(results["n_immed_rep"]).astype(np.uint64) * np.uint64(0x165667B19E3779F9) + r
        a = (a ^ (a >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
        a = a ^ (a >> np.uint64(32))
        return int(a.sum(dtype=np.uint64)) & _M64
