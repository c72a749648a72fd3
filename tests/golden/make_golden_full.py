"""Pins every BENCHED configuration (BASELINE.json configs c1..c5 at FULL size) to the CPU oracle.

    python tests/golden/make_golden_full.py [c4 c3 ...]     # several minutes (305 M sequential events for c4)

Writes tests/golden/full_configs.json: per workload the committed-event count, the sum of the per-LP trace hashes, an
LP-bound digest of the per-LP (count, hash) pairs and, for MachRep2, a digest of the per-replica results plus the
population means.  bench.py asserts its run against these values for every workload and GPU count, and
tests/test_gpu_parity.py does the same at full size -- so the number that is benched is a number that was checked.

These fixtures pin the ORACLE (sequential ground truth, oracle/oracle.c), not the Haskell engine: the reference ships no
golden vectors and cannot be built in this image (no GHC) -- parity stays "unpinned against the reference" (DESIGN.md 6).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402
from aivika_distributed_b200 import workloads as W  # noqa: E402

OUT = os.path.join(HERE, "full_configs.json")


def run(name):
    wl = W.WORKLOADS[name]
    t0 = time.time()
    if wl["kind"] == "phold":
        p = H.phold_params(wl["n_lps"], wl["stop"], remote_pct=wl["remote_pct"], n_hot=wl.get("n_hot", 0),
                           hot_pct=wl.get("hot_pct", 0), seed=W.SEED)
        n_exec, h, s = H.oracle_phold(p)
        rec = {"committed": int(s.committed), "trace_hash_sum": hex(int(s.hash_sum)), "lp_digest": hex(W.lp_digest(0, n_exec, h)),
               "remote_messages": int(s.remote), "max_committed_per_lp": int(n_exec.max())}
    else:
        p = H.machrep_params(wl["n_replicas"], wl["stop"])
        res = H.oracle_machrep(p)
        n = wl["n_replicas"]
        n_exec = np.array([r.n_exec[k] for r in res for k in range(3)], dtype=np.uint64)
        h = np.array([r.hash[k] for r in res for k in range(3)], dtype=np.uint64)
        arr = np.zeros(n, dtype=[("total_up_time", "<f8"), ("n_rep", "<i8"), ("n_immed_rep", "<i8")])
        arr["total_up_time"] = [r.total_up_time for r in res]; arr["n_rep"] = [r.n_rep for r in res]
        arr["n_immed_rep"] = [r.n_immed_rep for r in res]
        with np.errstate(over="ignore"):
            hsum = int(h.sum(dtype=np.uint64))
        rec = {"committed": int(n_exec.sum()), "trace_hash_sum": hex(hsum), "lp_digest": hex(W.lp_digest(0, n_exec, h)),
               "replica_digest": hex(W.machrep_digest(0, arr)),
               "mean_up_time_prop": float(np.mean([r.up_time_prop for r in res])),
               "mean_immed_prop": float(np.mean([r.immed_prop for r in res])),
               "replica0": {"up_time_prop": float(res[0].up_time_prop).hex(), "immed_prop": float(res[0].immed_prop).hex()}}
    rec["workload"] = wl["desc"]
    rec["oracle_seconds"] = round(time.time() - t0, 1)
    return rec


def main():
    names = sys.argv[1:] or ["c1", "c2", "c3", "c5", "c4"]
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        data[name] = run(name)
        print(name, data[name], flush=True)
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
