"""Writes tests/golden/wire_vectors.json: byte-level vectors of the reference's `Binary` wire format, produced by the Python
restatement of the documented `Data.Binary` generic encoding in tests/test_wire.py (NOT by GHC: there is none in this image;
parity with the Haskell packages is unpinned and says so wherever it is reported).  python tests/golden/make_wire_vectors.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_wire as T  # noqa: E402


class Ev:
    def __init__(self, recv, send, ctr, src, dst, typ, value, anti=False):
        self.recv_ts, self.send_ts, self.src, self.dst, self.type, self.value = recv, send, src, dst, typ, value
        self.ctr = ctr | ((1 << 63) if anti else 0)


msgs = [Ev(3.5, 2.25, 41, 1, 0, 1, 0.7310585786300049), Ev(1.000001, 1.0, 0, 2, 0, 2, 1.0), Ev(1000.0, 999.999999, 2 ** 40 + 5, 0, 1, 7, 0.0),
        Ev(12.0, 12.0, 9, 5, 4, 10, 1.0, True), Ev(0.5, 0.0, 0, 3, 3, 0, 0.0), Ev(7.25, 7.249999, 123456789, 1, 0, 6, 0.0)]
out = {"note": "Message / LogicalProcessMessage / TimeServerMessage as their GHC.Generics Binary instances lay them out "
               "(Internal/Message.hs:37-54, 113-156); ProcessIds and fingerprints are the synthetic ones of tests/test_wire.py",
       "messages": [], "lp_messages": [], "ts_messages": []}
for e in msgs:
    out["messages"].append({"recv_ts": e.recv_ts.hex(), "send_ts": e.send_ts.hex(), "ctr": e.ctr & ((1 << 61) - 1), "anti": bool(e.ctr >> 63),
                            "src": e.src, "dst": e.dst, "type": e.type, "value": float(e.value).hex(), "bytes": T.py_message(e).hex()})
out["lp_messages"] = [
    {"kind": 0, "n": 1, "bytes": (b"\x00" + T.py_message(msgs[0])).hex()},
    {"kind": 1, "n": 3, "bytes": (b"\x01" + T.b_int(3) + b"".join(T.py_message(m) for m in msgs[:3])).hex()},
    {"kind": 1, "n": 0, "bytes": (b"\x01" + T.b_int(0)).hex()},
    {"kind": 4, "n": 0, "bytes": "04"}, {"kind": 9, "n": 0, "bytes": "09"},
    {"kind": 5, "n": 0, "time": (17.25).hex(), "bytes": (b"\x05" + T.b_double(17.25)).hex()}]
for k, lp, t in ((5, 2, 3.75), (0, 0, 0.0), (3, 5, 0.0), (1, 4, 0.0), (5, 0, 999.999999)):
    out["ts_messages"].append({"kind": k, "lp": lp, "time": float(t).hex(),
                               "bytes": (bytes([k]) + T.b_pid(T.PIDS[lp]) + (T.b_double(t) if k == 5 else b"")).hex()})
json.dump(out, open(os.path.join(HERE, "wire_vectors.json"), "w"), indent=1)
print("wire vectors written")
