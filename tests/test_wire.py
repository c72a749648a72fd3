"""the reference's `Binary` wire format of Message / LogicalProcessMessage / TimeServerMessage
(Internal/Message.hs:37-54, 113-156) as produced by csrc/tw_wire.cpp, against an independent Python restatement of the
`Data.Binary` generic encoding written here, against hand-derived known answers, and against the committed byte fixtures
(tests/golden/wire_vectors.json).  PARITY UNPINNED against GHC: `binary` and `distributed-process` are not vendored
and there is no Haskell toolchain in this image; the fixtures pin the documented layout, not a run of the reference."""
import ctypes as C
import json
import math
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ independent restatement of Data.Binary (class instances)
def b_int(n):                      # instance Binary Int: Int64 big endian
    return struct.pack(">q", n)


def b_integer(n):                  # instance Binary Integer
    if -2 ** 31 <= n <= 2 ** 31 - 1:
        return b"\x00" + struct.pack(">i", n)
    mag, out = abs(n), []
    while mag:
        out.append(mag & 0xFF); mag >>= 8
    return b"\x01" + (b"\x01" if n > 0 else b"\xff") + b_int(len(out)) + bytes(out)


def b_double(d):                   # instance Binary Double: put (decodeFloat d)
    if d == 0.0:
        return b_integer(0) + b_int(0)
    m, e = math.frexp(d)           # d = m * 2^e, 0.5 <= |m| < 1  ->  mantissa in [2^52, 2^53)
    return b_integer(int(m * 2 ** 53)) + b_int(e - 53)


def b_bytestring(b):
    return b_int(len(b)) + b


def b_pid(pid):                    # ProcessId = NodeId (EndPointAddress) + LocalProcessId (two Int32)
    addr, unique, counter = pid
    return b_bytestring(addr) + struct.pack(">ii", unique, counter)


def b_message(seq, send_t, recv_t, sender, receiver, anti, fingerprint, payload):
    chunks = [fingerprint] + ([payload] if payload else [])
    data = b_int(len(chunks)) + b"".join(b_bytestring(c) for c in chunks)
    return struct.pack(">Q", seq) + b_double(send_t) + b_double(recv_t) + b_pid(sender) + b_pid(receiver) + bytes([1 if anti else 0]) + data


PIDS = [(b"127.0.0.1:8080:0", 10 + i, 3 * i) for i in range(5)] + [(b"10.1.2.3:9000:0", -7, 2 ** 31 - 1)]
FP = [bytes((37 * k + j) & 0xFF for j in range(16)) for k in range(13)]
# MachRep2 kinds (include/twb200.h TW_MR_*) -> payload shapes of tests/MachRep2DistributedReproducible.hs:88-104
SHAPES = [0, 2, 3, 3, 1, 1, 1, 1, 1, 1, 3, 1, 1]


@pytest.fixture(scope="module")
def wire(H):
    f = H.ffi; lib = f.load()
    pids = (f.tw_wire_pid * len(PIDS))(*[f.tw_wire_pid(a, len(a), u, c) for a, u, c in PIDS])
    types = (f.tw_wire_type * 13)()
    for k in range(13):
        types[k].fingerprint[:] = list(FP[k]); types[k].payload = SHAPES[k]
    ctx = C.c_void_p()
    assert lib.tw_wire_create(pids, len(PIDS), types, 13, C.byref(ctx)) == 0
    yield lib, ctx, f
    lib.tw_wire_destroy(ctx)


def mk_event(f, recv, send, ctr, src, dst, typ, value, anti=False):
    e = f.tw_event(); e.recv_ts = recv; e.send_ts = send; e.ctr = ctr | ((1 << 63) if anti else 0)
    e.src = src; e.dst = dst; e.type = typ; e.value = value
    return e


def py_payload(ev):
    shape = SHAPES[ev.type]
    if shape == 1: return b_pid(PIDS[ev.src])
    if shape == 2: return b_pid(PIDS[ev.src]) + b_double(ev.value)
    if shape == 3: return b_pid(PIDS[ev.src]) + b_int(int(ev.value))
    return b""


def py_message(ev):
    return b_message(ev.ctr & ((1 << 61) - 1), ev.send_ts, ev.recv_ts, PIDS[ev.src], PIDS[ev.dst], bool(ev.ctr >> 63), FP[ev.type], py_payload(ev))


def c_encode(lib, ctx, ev):
    n = C.c_size_t()
    assert lib.tw_wire_encode_message(ctx, C.byref(ev), None, 0, C.byref(n)) == 1 and n.value > 0     # size query
    buf = (C.c_uint8 * n.value)()
    assert lib.tw_wire_encode_message(ctx, C.byref(ev), buf, n.value, C.byref(n)) == 0
    return bytes(buf)


def test_known_answers_of_the_binary_package():
    """Hand-derived from the class instances of `binary`: encode (1.0 :: Double) is the 25-byte string every Haskell
    programmer has seen in GHCi (tag 1, sign 1, 7 mantissa bytes little endian, exponent -52 as Int64)."""
    assert b_double(1.0).hex() == "0101" + "0000000000000007" + "00000000000010" + "ffffffffffffffcc"
    assert b_double(0.0).hex() == "0000000000" + "0000000000000000"
    assert b_double(-0.5).hex() == "01ff" + "0000000000000007" + "00000000000010" + "ffffffffffffffcb"
    assert b_double(1e-6).hex()[:4] == "0101" and len(b_double(1e-6)) == 25
    assert b_integer(5).hex() == "0000000005" and b_integer(-1).hex() == "00ffffffff" and b_integer(2 ** 31).hex() == "0101" + "0000000000000004" + "00000080"


def test_message_bytes_equal_the_restatement(wire):
    lib, ctx, f = wire
    rng = np.random.default_rng(7)
    cases = [mk_event(f, 3.5, 2.25, 41, 1, 0, 1, 0.7310585786300049), mk_event(f, 1e-6, 0.0, 0, 2, 0, 2, 1.0),
             mk_event(f, 1000.0, 999.999999, 2 ** 40 + 5, 0, 1, 7, 0.0), mk_event(f, 12.0, 12.0, 9, 5, 4, 10, 1.0, anti=True),
             mk_event(f, 0.0, 0.0, 0, 3, 3, 0, 0.0), mk_event(f, 5e-324, 0.0, 1, 0, 2, 1, 2.2250738585072014e-308)]
    for _ in range(200):
        cases.append(mk_event(f, float(rng.exponential(50.0)), float(rng.exponential(1.0)), int(rng.integers(0, 2 ** 60)), int(rng.integers(0, 6)),
                              int(rng.integers(0, 6)), int(rng.integers(0, 13)), float(rng.integers(-5, 5)) if rng.random() < 0.5 else float(rng.normal()),
                              anti=bool(rng.integers(0, 2))))
    for ev in cases:
        wire_bytes = c_encode(lib, ctx, ev)
        if ev.recv_ts > 1e-300:                                   # (math.frexp restatement: normal numbers)
            assert wire_bytes == py_message(ev)
        out = f.tw_event(); used = C.c_size_t()
        assert lib.tw_wire_decode_message(ctx, wire_bytes, len(wire_bytes), C.byref(out), C.byref(used)) == 0 and used.value == len(wire_bytes)
        assert (out.recv_ts, out.send_ts, out.ctr, out.src, out.dst, out.type) == (ev.recv_ts, ev.send_ts, ev.ctr, ev.src, ev.dst, ev.type)
        if SHAPES[ev.type] == 2: assert out.value == ev.value
        if SHAPES[ev.type] == 3: assert out.value == float(int(ev.value))


def test_committed_byte_fixtures(wire):
    """tests/golden/wire_vectors.json (written by tests/golden/make_wire_vectors.py from the Python restatement above):
    the C encoder reproduces every vector and the decoder maps it back."""
    lib, ctx, f = wire
    g = json.load(open(os.path.join(HERE, "golden", "wire_vectors.json")))
    for v in g["messages"]:
        ev = mk_event(f, float.fromhex(v["recv_ts"]), float.fromhex(v["send_ts"]), v["ctr"], v["src"], v["dst"], v["type"], float.fromhex(v["value"]), v["anti"])
        assert c_encode(lib, ctx, ev).hex() == v["bytes"]
    for v in g["lp_messages"]:
        kind = C.c_uint32(); n = C.c_size_t(); t = C.c_double(); evs = (f.tw_event * 8)()
        raw = bytes.fromhex(v["bytes"])
        assert lib.tw_wire_decode_lp_message(ctx, raw, len(raw), C.byref(kind), evs, 8, C.byref(n), C.byref(t)) == 0
        assert kind.value == v["kind"] and n.value == v["n"]
        if v["kind"] == 5: assert t.value == float.fromhex(v["time"])
    for v in g["ts_messages"]:
        n = C.c_size_t(); buf = (C.c_uint8 * 256)()
        assert lib.tw_wire_encode_ts_message(ctx, v["kind"], v["lp"], float.fromhex(v["time"]), buf, 256, C.byref(n)) == 0
        assert bytes(buf[: n.value]).hex() == v["bytes"]


def test_logical_process_and_time_server_messages(wire):
    lib, ctx, f = wire
    evs = (f.tw_event * 3)(mk_event(f, 2.0, 1.0, 5, 1, 0, 1, 0.25), mk_event(f, 2.0, 1.0, 6, 1, 0, 2, 1.0), mk_event(f, 2.0, 1.0, 7, 1, 0, 4, 0.0))
    buf = (C.c_uint8 * 4096)(); n = C.c_size_t()
    # QueueMessageBulk [m1, m2, m3] = constructor 1, list length, the messages
    assert lib.tw_wire_encode_lp_message(ctx, 1, evs, 3, 0.0, buf, 4096, C.byref(n)) == 0
    raw = bytes(buf[: n.value])
    assert raw == b"\x01" + b_int(3) + b"".join(py_message(e) for e in evs)
    kind = C.c_uint32(); cnt = C.c_size_t(); t = C.c_double(); out = (f.tw_event * 3)()
    assert lib.tw_wire_decode_lp_message(ctx, raw, len(raw), C.byref(kind), out, 2, C.byref(cnt), C.byref(t)) == 1 and cnt.value == 3   # too small
    assert lib.tw_wire_decode_lp_message(ctx, raw, len(raw), C.byref(kind), out, 3, C.byref(cnt), C.byref(t)) == 0
    assert kind.value == 1 and [out[i].ctr for i in range(3)] == [5, 6, 7] and out[0].value == 0.25
    # QueueMessage m = constructor 0; ComputeLocalTimeMessage = 4; GlobalTimeMessage t = 5; AbortSimulationMessage = 9
    assert lib.tw_wire_encode_lp_message(ctx, 0, evs, 1, 0.0, buf, 4096, C.byref(n)) == 0 and bytes(buf[: n.value]) == b"\x00" + py_message(evs[0])
    assert lib.tw_wire_encode_lp_message(ctx, 4, None, 0, 0.0, buf, 4096, C.byref(n)) == 0 and bytes(buf[: n.value]) == b"\x04"
    assert lib.tw_wire_encode_lp_message(ctx, 5, None, 0, 17.25, buf, 4096, C.byref(n)) == 0 and bytes(buf[: n.value]) == b"\x05" + b_double(17.25)
    assert lib.tw_wire_encode_lp_message(ctx, 9, None, 0, 0.0, buf, 4096, C.byref(n)) == 0 and bytes(buf[: n.value]) == b"\x09"
    assert lib.tw_wire_encode_lp_message(ctx, 2, None, 0, 0.0, buf, 4096, C.byref(n)) == 12       # acknowledgement traffic: not on this path
    # TimeServerMessage: LocalTimeMessage pid t = constructor 5; RegisterLogicalProcessMessage pid = 0; ...Acknowledgement = 3
    for k, lp, tm in ((5, 2, 3.75), (0, 0, 0.0), (3, 5, 0.0), (1, 4, 0.0)):
        assert lib.tw_wire_encode_ts_message(ctx, k, lp, tm, buf, 4096, C.byref(n)) == 0
        raw = bytes(buf[: n.value])
        assert raw == bytes([k]) + b_pid(PIDS[lp]) + (b_double(tm) if k == 5 else b"")
        kk = C.c_uint32(); ll = C.c_uint32(); tt = C.c_double()
        assert lib.tw_wire_decode_ts_message(ctx, raw, len(raw), C.byref(kk), C.byref(ll), C.byref(tt)) == 0
        assert (kk.value, ll.value) == (k, lp) and (k != 5 or tt.value == tm)


def test_malformed_input_is_rejected(wire):
    lib, ctx, f = wire
    good = c_encode(lib, ctx, mk_event(f, 3.5, 2.25, 41, 1, 0, 1, 0.5))
    out = f.tw_event()
    for bad in (good[:-1], good[:20], b"", good[:8] + b"\x07" + good[9:], good.replace(PIDS[1][0], b"127.0.0.9:8080:0")):
        assert lib.tw_wire_decode_message(ctx, bad, len(bad), C.byref(out), None) == 1
    ev = mk_event(f, 1.0, 0.0, 1, 9, 0, 1, 0.0)                       # an LP without a ProcessId
    n = C.c_size_t()
    assert lib.tw_wire_encode_message(ctx, C.byref(ev), None, 0, C.byref(n)) == 9


@pytest.mark.gpu
def test_device_partition_under_a_stock_time_server_protocol(H, wire):
    """The byte-level Time Server round (Internal/TimeServer.hs:369-403; LP side Internal/Event.hs:425-449): the server's
    ComputeLocalTimeMessage is answered with ComputeLocalTimeAcknowledgementMessage + LocalTimeMessage pid t, its
    GlobalTimeMessage bounds GVT, start-time messages arrive as one QueueMessageBulk -- and the committed trace of the
    partition equals the oracle's."""
    lib, ctx, f = wire
    from aivika_distributed_b200 import dio as D
    n_rep = 2
    dio = D.DIO(D.MachRep2(n_rep), D.DIOParams(queueCapacity=256, stepsPerLaunch=4), D.Specs(0.0, 60.0))
    dio.attachTimeServer()
    init = dio.initialEvents()                                        # START events, kind 0 (no payload)
    evs = (f.tw_event * len(init)).from_buffer_copy(init.tobytes())
    buf = (C.c_uint8 * 65536)(); n = C.c_size_t(); rl = C.c_size_t(); nr = C.c_uint32(); rep = (C.c_uint8 * 1024)()
    assert lib.tw_wire_encode_lp_message(ctx, 1, evs, len(init), 0.0, buf, 65536, C.byref(n)) == 0
    assert lib.tw_wire_handle_lp_message(dio.h, ctx, 0, buf, n.value, rep, 1024, C.byref(rl), C.byref(nr)) == 0 and nr.value == 0
    rounds, last = 0, 0.0
    while True:
        st = dio.runEventInStopTime()
        if st.gvt > 60.0:
            break
        assert lib.tw_wire_handle_lp_message(dio.h, ctx, 0, b"\x04", 1, rep, 1024, C.byref(rl), C.byref(nr)) == 0 and nr.value == 2
        raw = bytes(rep[: rl.value])
        l0 = int.from_bytes(raw[:4], "little"); ack = raw[4:4 + l0]; l1 = int.from_bytes(raw[4 + l0:8 + l0], "little"); lt = raw[8 + l0:8 + l0 + l1]
        assert ack == b"\x03" + b_pid(PIDS[0]) and lt[:1] == b"\x05" and lt[1:1 + len(b_pid(PIDS[0]))] == b_pid(PIDS[0])
        kk = C.c_uint32(); ll = C.c_uint32(); tt = C.c_double()
        assert lib.tw_wire_decode_ts_message(ctx, lt, len(lt), C.byref(kk), C.byref(ll), C.byref(tt)) == 0 and tt.value >= last
        last = tt.value
        g = b"\x05" + b_double(tt.value)                              # the server's GlobalTimeMessage (one LP: its own local time)
        assert lib.tw_wire_handle_lp_message(dio.h, ctx, 0, g, len(g), rep, 1024, C.byref(rl), C.byref(nr)) == 0
        rounds += 1
        assert rounds < 100000
    assert rounds > 5
    res = dio.machRepResults(); ores = H.oracle_machrep(H.machrep_params(n_rep, 60.0))
    assert all(res[r]["up_time_prop"] == ores[r].up_time_prop and res[r]["immed_prop"] == ores[r].immed_prop for r in range(n_rep))
    assert lib.tw_wire_handle_lp_message(dio.h, ctx, 0, b"\x09", 1, rep, 1024, C.byref(rl), C.byref(nr)) == 13      # AbortSimulationMessage
    dio.close()
