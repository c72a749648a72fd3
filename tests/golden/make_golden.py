"""Regenerates the golden fixtures from the CPU oracle (python tests/golden/make_golden.py).

The reference (aivika-distributed) has no golden vectors, no assertions and cannot be built in this image
(no GHC), so these fixtures pin the ORACLE, not the reference: they detect accidental changes of the model
definitions (RNG streams, software log, tie-breaking, handler order)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as H  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))

r = H.oracle_machrep(H.machrep_params(1, 1000.0))[0]
json.dump({"config": "tests/MachRep2DistributedReproducible.hs: seeds 456/789, stop 1000, delta 1e-6, means 1.0/0.5",
           "up_time_prop": r.up_time_prop.hex(), "immed_prop": r.immed_prop.hex(),
           "up_time_prop_decimal": r.up_time_prop, "immed_prop_decimal": r.immed_prop,
           "total_up_time": r.total_up_time.hex(), "n_rep": r.n_rep, "n_immed_rep": r.n_immed_rep,
           "n_exec": list(r.n_exec), "hash": [hex(x) for x in r.hash]},
          open(os.path.join(here, "machrep2_c1.json"), "w"), indent=1)

r = H.oracle_machrep1(H.machrep1_params(1, 10000.0))[0]
json.dump({"config": "tests/MachRep1.hs: two machines, two repair persons, means 1.0/0.5, stop 10000; StdGen seeds 456/789 "
                     "(the reference draws from the run's default generator, which is not reproducible)",
           "up_time_prop": r.up_time_prop.hex(), "up_time_prop_decimal": r.up_time_prop,
           "total_up_time": r.total_up_time.hex(), "n_messages": r.n_rep,
           "n_exec": list(r.n_exec), "hash": [hex(x) for x in r.hash]},
          open(os.path.join(here, "machrep1.json"), "w"), indent=1)

cases = []
for params in (dict(n_lps=64, stop=5.0, remote_pct=10), dict(n_lps=1000, stop=3.0, remote_pct=50, seed=3),
               dict(n_lps=512, stop=4.0, remote_pct=50, n_hot=8, hot_pct=50, seed=11),
               dict(n_lps=3, stop=20.0, remote_pct=100, population=4)):
    n_exec, h, s = H.oracle_phold(H.phold_params(**params))
    cases.append({"params": params, "committed": int(s.committed), "hash_sum": hex(int(s.hash_sum)),
                  "n_exec_head": [int(x) for x in n_exec[:8]], "hash_head": [hex(int(x)) for x in h[:4]]})
json.dump({"cases": cases}, open(os.path.join(here, "phold_small.json"), "w"), indent=1)
print("golden fixtures written")
