// tw_wire.cpp -- the reference's wire format at the host boundary: host-side C++ only, no
// device code; linked into libtwb200.so next to engine.cu.
//
// What is encoded (all `instance Binary X` via GHC.Generics, so the layout is the one `Data.Binary` documents for
// generic instances: products = the fields in order, sums = a Word8 constructor index in declaration order):
//   Message                 Simulation/Aivika/Distributed/Optimistic/Internal/Message.hs:37-54
//   LogicalProcessMessage   :113-137  (QueueMessage = 0, QueueMessageBulk = 1, ComputeLocalTimeMessage = 4,
//                                      GlobalTimeMessage = 5, AbortSimulationMessage = 9)
//   TimeServerMessage       :140-156  (RegisterLogicalProcessMessage = 0, UnregisterLogicalProcessMessage = 1,
//                                      ComputeLocalTimeAcknowledgementMessage = 3, LocalTimeMessage = 5)
// Primitive encodings (package `binary`, class Binary): Word64 / Int = 8 bytes big endian, Int32 = 4 bytes big endian,
// Bool = one byte, [a] = Int length + elements, strict ByteString = Int length + bytes, Double = (Integer, Int) from
// `decodeFloat`, Integer = tag 0 + Int32 when it fits, else tag 1 + sign byte (1 or 0xFF) + the magnitude as a
// little-endian byte list.  Third-party records (package `distributed-process`, Control.Distributed.Process.Internal.Types):
//   ProcessId      = NodeId (EndPointAddress = strict ByteString) + LocalProcessId (lpidUnique, lpidCounter :: Int32)
//   DP.Message     = `put (messageToPayload msg)` = [ByteString]: the 16-byte type fingerprint, then the chunks of the
//                    payload's own Binary encoding
// The 16 fingerprint bytes of a payload type come from GHC's Typeable and cannot be computed outside a Haskell program:
// the caller registers them per message type (tw_wire_type), together with the shape of the payload.
// PARITY UNPINNED: neither `binary` nor `distributed-process` is vendored under /root/reference and no GHC exists in this
// image; the layouts above are restated from the packages' published sources and the reference's own type
// declarations, and pinned only by the byte-level fixtures of tests/test_wire.py (an independent Python restatement).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/twb200.h"

namespace {

constexpr uint64_t kAnti = 1ull << 63, kCtrBits = ~(7ull << 61);

struct Writer {
    uint8_t *buf; size_t cap, len = 0;
    Writer(uint8_t *b, size_t c) : buf(b), cap(c) {}
    void u8(uint8_t v) { if (len < cap) buf[len] = v; ++len; }
    void be(uint64_t v, int n) { for (int i = n - 1; i >= 0; --i) u8((uint8_t)(v >> (8 * i))); }
    void bytes(const uint8_t *p, size_t n) { for (size_t i = 0; i < n; ++i) u8(p[i]); }
    void i64(int64_t v) { be((uint64_t)v, 8); }
    void integer(bool neg, uint64_t mag)                 // instance Binary Integer
    {
        if (mag <= 0x7FFFFFFFull || (neg && mag == 0x80000000ull)) { u8(0); be((uint32_t)(neg ? 0 - (int64_t)mag : (int64_t)mag), 4); return; }
        u8(1); u8(neg ? 0xFF : 1);
        int n = 0; for (uint64_t m = mag; m; m >>= 8) ++n;
        i64(n);
        for (int i = 0; i < n; ++i) u8((uint8_t)(mag >> (8 * i)));
    }
    void f64(double d)                                    // instance Binary Double: put (decodeFloat d)
    {
        uint64_t b; memcpy(&b, &d, 8);
        const bool neg = (b >> 63) != 0;
        int64_t e = (int64_t)((b >> 52) & 0x7FF);
        uint64_t m = b & ((1ull << 52) - 1);
        if (e == 0 && m == 0) { integer(false, 0); i64(0); return; }      // decodeFloat 0 = (0, 0) (and -0.0 as well)
        if (e == 0) { e = 1; while (!(m >> 52)) { m <<= 1; --e; } }       // denormal: normalised mantissa
        else m |= 1ull << 52;                                             // (inf / nan decode like finite numbers in GHC)
        integer(neg, m); i64(e - 1075);
    }
};

struct Reader {
    const uint8_t *buf; size_t len, pos = 0; bool ok = true;
    Reader(const uint8_t *b, size_t n) : buf(b), len(n) {}
    uint8_t u8() { if (pos >= len) { ok = false; return 0; } return buf[pos++]; }
    uint64_t be(int n) { uint64_t v = 0; for (int i = 0; i < n; ++i) v = (v << 8) | u8(); return v; }
    int64_t i64() { return (int64_t)be(8); }
    bool integer(bool &neg, uint64_t &mag)
    {
        const uint8_t tag = u8();
        if (tag == 0) { const int32_t v = (int32_t)(uint32_t)be(4); neg = v < 0; mag = neg ? (uint64_t)(-(int64_t)v) : (uint64_t)v; return ok; }
        if (tag != 1) return ok = false;
        const uint8_t s = u8(); neg = s == 0xFF;
        const int64_t n = i64();
        if (!ok || n < 0 || n > 8) return ok = false;
        mag = 0; for (int64_t i = 0; i < n; ++i) mag |= (uint64_t)u8() << (8 * i);
        return ok;
    }
    double f64()                                           // encodeFloat mantissa exponent
    {
        bool neg; uint64_t m; if (!integer(neg, m)) return 0.0;
        const int64_t e = i64();
        if (!ok) return 0.0;
        if (e < -2000 || e > 2000) { ok = false; return 0.0; }
        const double v = ldexp((double)m, (int)e);        // m < 2^53: exact
        return neg ? -v : v;
    }
};

struct Pid { std::string addr; int32_t unique, counter; };

}  // namespace

struct tw_wire_ctx {
    std::vector<Pid> pids;                                 // LP id -> ProcessId
    std::map<std::string, uint32_t> lp_of;                 // encoded ProcessId -> LP id
    std::vector<tw_wire_type> types;                       // message kind -> fingerprint + payload shape
    std::string err;
    std::string key(const Pid &p) const
    {
        std::string k = p.addr; k.push_back('\0');
        k.append((const char *)&p.unique, 4); k.append((const char *)&p.counter, 4);
        return k;
    }
    void put_pid(Writer &w, uint32_t lp) const
    {
        const Pid &p = pids[lp];
        w.i64((int64_t)p.addr.size()); w.bytes((const uint8_t *)p.addr.data(), p.addr.size());
        w.be((uint32_t)p.unique, 4); w.be((uint32_t)p.counter, 4);
    }
    bool get_pid(Reader &r, uint32_t &lp) const
    {
        const int64_t n = r.i64();
        if (!r.ok || n < 0 || (size_t)n > r.len - r.pos) return r.ok = false;
        Pid p; p.addr.assign((const char *)r.buf + r.pos, (size_t)n); r.pos += (size_t)n;
        p.unique = (int32_t)(uint32_t)r.be(4); p.counter = (int32_t)(uint32_t)r.be(4);
        if (!r.ok) return false;
        auto it = lp_of.find(key(p));
        if (it == lp_of.end()) return r.ok = false;
        lp = it->second;
        return true;
    }
    // the payload's own Binary encoding (what ends up in the chunks of DP.Message)
    void put_payload(Writer &w, const tw_event &ev, const tw_wire_type &t) const
    {
        switch (t.payload) {
        case TW_WIRE_PAYLOAD_SRC_PID: put_pid(w, ev.src); break;
        case TW_WIRE_PAYLOAD_SRC_PID_DOUBLE: put_pid(w, ev.src); w.f64(ev.value); break;
        case TW_WIRE_PAYLOAD_SRC_PID_INT: put_pid(w, ev.src); w.i64((int64_t)ev.value); break;
        case TW_WIRE_PAYLOAD_DOUBLE: w.f64(ev.value); break;
        case TW_WIRE_PAYLOAD_INT: w.i64((int64_t)ev.value); break;
        case TW_WIRE_PAYLOAD_SRC_INT_CTR: w.i64((int64_t)ev.src); w.be(ev.ctr & kCtrBits, 8); break;
        default: break;
        }
    }
    bool get_payload(Reader &r, tw_event &ev, const tw_wire_type &t) const
    {
        uint32_t lp = 0;
        switch (t.payload) {
        case TW_WIRE_PAYLOAD_SRC_PID: return get_pid(r, lp) && lp == ev.src;
        case TW_WIRE_PAYLOAD_SRC_PID_DOUBLE: if (!get_pid(r, lp) || lp != ev.src) return false; ev.value = r.f64(); return r.ok;
        case TW_WIRE_PAYLOAD_SRC_PID_INT: if (!get_pid(r, lp) || lp != ev.src) return false; ev.value = (double)r.i64(); return r.ok;
        case TW_WIRE_PAYLOAD_DOUBLE: ev.value = r.f64(); return r.ok;
        case TW_WIRE_PAYLOAD_INT: ev.value = (double)r.i64(); return r.ok;
        case TW_WIRE_PAYLOAD_SRC_INT_CTR: { const int64_t s = r.i64(); const uint64_t c = r.be(8); return r.ok && (uint32_t)s == ev.src && c == (ev.ctr & kCtrBits); }
        default: return true;
        }
    }
    void put_message(Writer &w, const tw_event &ev) const     // Message, Internal/Message.hs:37-54
    {
        w.be(ev.ctr & kCtrBits, 8);                            // messageSequenceNo :: Word64
        w.f64(ev.send_ts); w.f64(ev.recv_ts);                  // messageSendTime, messageReceiveTime
        put_pid(w, ev.src); put_pid(w, ev.dst);                // messageSenderId, messageReceiverId
        w.u8((ev.ctr & kAnti) ? 1 : 0);                        // messageAntiToggle
        const tw_wire_type &t = types[ev.type];
        Writer probe(nullptr, 0); put_payload(probe, ev, t);   // messageData = [fingerprint, chunk]
        w.i64(probe.len ? 2 : 1);
        w.i64(16); w.bytes(t.fingerprint, 16);
        if (probe.len) { w.i64((int64_t)probe.len); put_payload(w, ev, t); }
    }
    bool get_message(Reader &r, tw_event &ev) const
    {
        memset(&ev, 0, sizeof ev);
        ev.ctr = r.be(8);
        ev.send_ts = r.f64(); ev.recv_ts = r.f64();
        if (!r.ok || !get_pid(r, ev.src) || !get_pid(r, ev.dst)) return false;
        const uint8_t anti = r.u8();
        if (!r.ok || anti > 1 || (ev.ctr & ~kCtrBits)) return r.ok = false;
        if (anti) ev.ctr |= kAnti;
        const int64_t n_chunks = r.i64();
        if (!r.ok || n_chunks < 1 || n_chunks > 1024) return r.ok = false;
        if (r.i64() != 16 || !r.ok || r.len - r.pos < 16) return r.ok = false;
        uint32_t type = 0xFFFFFFFFu;
        for (uint32_t k = 0; k < types.size(); ++k) if (!memcmp(types[k].fingerprint, r.buf + r.pos, 16)) { type = k; break; }
        if (type == 0xFFFFFFFFu) return r.ok = false;
        r.pos += 16; ev.type = type;
        std::string payload;                                   // BSL.fromChunks: any chunking of the same bytes is the same message
        for (int64_t c = 1; c < n_chunks; ++c) {
            const int64_t n = r.i64();
            if (!r.ok || n < 0 || (size_t)n > r.len - r.pos) return r.ok = false;
            payload.append((const char *)r.buf + r.pos, (size_t)n); r.pos += (size_t)n;
        }
        Reader pr((const uint8_t *)payload.data(), payload.size());
        if (!get_payload(pr, ev, types[type]) || !pr.ok || pr.pos != pr.len) return r.ok = false;
        return true;
    }
};

static thread_local std::string g_wire_err;

extern "C" {

int tw_wire_create(const tw_wire_pid *pids, uint32_t n_lps, const tw_wire_type *types, uint32_t n_types, tw_wire_ctx **out)
{
    if (!pids || !types || !out || n_lps == 0 || n_types == 0) return TW_ERR_INVALID_ARG;
    tw_wire_ctx *c = new tw_wire_ctx();
    for (uint32_t i = 0; i < n_lps; ++i) {
        if (!pids[i].node_address && pids[i].address_len) { delete c; return TW_ERR_INVALID_ARG; }
        Pid p; p.addr.assign((const char *)pids[i].node_address, pids[i].address_len); p.unique = pids[i].lpid_unique; p.counter = pids[i].lpid_counter;
        if (!c->lp_of.emplace(c->key(p), i).second) { delete c; g_wire_err = "two logical processes share one ProcessId"; return TW_ERR_INVALID_ARG; }
        c->pids.push_back(p);
    }
    c->types.assign(types, types + n_types);
    *out = c;
    return TW_OK;
}
int tw_wire_destroy(tw_wire_ctx *c) { if (!c) return TW_ERR_INVALID_ARG; delete c; return TW_OK; }

static int check_event(const tw_wire_ctx *c, const tw_event &ev)
{
    if (ev.src >= c->pids.size() || ev.dst >= c->pids.size()) return TW_ERR_BAD_DESTINATION;
    if (ev.type >= c->types.size()) return TW_ERR_INVALID_ARG;
    return TW_OK;
}

int tw_wire_encode_message(const tw_wire_ctx *c, const tw_event *ev, uint8_t *buf, size_t cap, size_t *len)
{
    if (!c || !ev || !len || (!buf && cap)) return TW_ERR_INVALID_ARG;
    int rc = check_event(c, *ev); if (rc) return rc;
    Writer w(buf, cap); c->put_message(w, *ev);
    *len = w.len;
    return w.len <= cap ? TW_OK : TW_ERR_INVALID_ARG;          // *len = bytes needed
}
int tw_wire_decode_message(const tw_wire_ctx *c, const uint8_t *buf, size_t len, tw_event *ev, size_t *consumed)
{
    if (!c || !buf || !ev) return TW_ERR_INVALID_ARG;
    Reader r(buf, len);
    if (!c->get_message(r, *ev)) return TW_ERR_INVALID_ARG;
    if (consumed) *consumed = r.pos;
    return TW_OK;
}

int tw_wire_encode_lp_message(const tw_wire_ctx *c, uint32_t kind, const tw_event *evs, size_t n, double time,
                              uint8_t *buf, size_t cap, size_t *len)
{
    if (!c || !len || (!buf && cap)) return TW_ERR_INVALID_ARG;
    Writer w(buf, cap);
    w.u8((uint8_t)kind);
    switch (kind) {
    case TW_WIRE_QUEUE_MESSAGE:
        if (!evs || n != 1) return TW_ERR_INVALID_ARG;
        if (int rc = check_event(c, evs[0])) return rc;
        c->put_message(w, evs[0]); break;
    case TW_WIRE_QUEUE_MESSAGE_BULK:
        if (!evs && n) return TW_ERR_INVALID_ARG;
        w.i64((int64_t)n);
        for (size_t i = 0; i < n; ++i) { if (int rc = check_event(c, evs[i])) return rc; c->put_message(w, evs[i]); }
        break;
    case TW_WIRE_COMPUTE_LOCAL_TIME: case TW_WIRE_ABORT_SIMULATION: break;
    case TW_WIRE_GLOBAL_TIME: w.f64(time); break;
    default: return TW_ERR_NOT_SUPPORTED;
    }
    *len = w.len;
    return w.len <= cap ? TW_OK : TW_ERR_INVALID_ARG;
}
int tw_wire_decode_lp_message(const tw_wire_ctx *c, const uint8_t *buf, size_t len, uint32_t *kind, tw_event *evs, size_t evs_cap,
                              size_t *n, double *time)
{
    if (!c || !buf || !kind || !n) return TW_ERR_INVALID_ARG;
    Reader r(buf, len);
    *kind = r.u8(); *n = 0;
    if (!r.ok) return TW_ERR_INVALID_ARG;
    switch (*kind) {
    case TW_WIRE_QUEUE_MESSAGE:
        if (!evs || evs_cap < 1 || !c->get_message(r, evs[0])) return TW_ERR_INVALID_ARG;
        *n = 1; break;
    case TW_WIRE_QUEUE_MESSAGE_BULK: {
        const int64_t cnt = r.i64();
        if (!r.ok || cnt < 0) return TW_ERR_INVALID_ARG;
        if ((size_t)cnt > evs_cap || (!evs && cnt)) { *n = (size_t)cnt; return TW_ERR_INVALID_ARG; }   // *n = events needed
        for (int64_t i = 0; i < cnt; ++i) if (!c->get_message(r, evs[i])) return TW_ERR_INVALID_ARG;
        *n = (size_t)cnt; break; }
    case TW_WIRE_COMPUTE_LOCAL_TIME: case TW_WIRE_ABORT_SIMULATION: break;
    case TW_WIRE_GLOBAL_TIME: { const double t = r.f64(); if (!r.ok) return TW_ERR_INVALID_ARG; if (time) *time = t; break; }
    default: return TW_ERR_NOT_SUPPORTED;                      // acknowledgement / monitoring / reconnect traffic: not this path
    }
    return r.pos == len ? TW_OK : TW_ERR_INVALID_ARG;
}

int tw_wire_encode_ts_message(const tw_wire_ctx *c, uint32_t kind, uint32_t lp, double time, uint8_t *buf, size_t cap, size_t *len)
{
    if (!c || !len || (!buf && cap) || lp >= c->pids.size()) return TW_ERR_INVALID_ARG;
    if (kind > TW_WIRE_TS_PROVIDE_STATE) return TW_ERR_NOT_SUPPORTED;
    Writer w(buf, cap);
    w.u8((uint8_t)kind); c->put_pid(w, lp);
    if (kind == TW_WIRE_TS_LOCAL_TIME) w.f64(time);
    *len = w.len;
    return w.len <= cap ? TW_OK : TW_ERR_INVALID_ARG;
}
int tw_wire_decode_ts_message(const tw_wire_ctx *c, const uint8_t *buf, size_t len, uint32_t *kind, uint32_t *lp, double *time)
{
    if (!c || !buf || !kind || !lp) return TW_ERR_INVALID_ARG;
    Reader r(buf, len);
    *kind = r.u8();
    if (!r.ok || *kind > TW_WIRE_TS_PROVIDE_STATE || !c->get_pid(r, *lp)) return TW_ERR_INVALID_ARG;
    if (*kind == TW_WIRE_TS_LOCAL_TIME) { const double t = r.f64(); if (time) *time = t; }
    return (r.ok && r.pos == len) ? TW_OK : TW_ERR_INVALID_ARG;
}

// The LP side of the reference's protocol for a whole device partition that registered with a stock `timeServer` (and
// with CPU logical processes) under the ProcessId of LP `self_lp`: one serialized LogicalProcessMessage in, the serialized
// TimeServerMessages it answers with out (each preceded by its length as a little-endian uint32: framing is the
// transport's business in Cloud Haskell).  processChannelMessage, Internal/Event.hs:358-449; the acknowledgement that
// precedes the local time: processTimeServerMessage, Internal/DIO.hs:560-566.
int tw_wire_handle_lp_message(tw_engine *e, const tw_wire_ctx *c, uint32_t self_lp, const uint8_t *buf, size_t len,
                              uint8_t *replies, size_t cap, size_t *replies_len, uint32_t *n_replies)
{
    if (!e || !c || !buf || !replies_len || !n_replies || (!replies && cap)) return TW_ERR_INVALID_ARG;
    *replies_len = 0; *n_replies = 0;
    uint32_t kind = 0; size_t n = 0; double t = 0.0;
    std::vector<tw_event> evs(1);
    int rc = tw_wire_decode_lp_message(c, buf, len, &kind, evs.data(), evs.size(), &n, &t);
    if (rc == TW_ERR_INVALID_ARG && kind == TW_WIRE_QUEUE_MESSAGE_BULK && n > evs.size()) {
        evs.resize(n);
        rc = tw_wire_decode_lp_message(c, buf, len, &kind, evs.data(), evs.size(), &n, &t);
    }
    if (rc) return rc;
    switch (kind) {
    case TW_WIRE_QUEUE_MESSAGE: case TW_WIRE_QUEUE_MESSAGE_BULK:
        return n ? tw_enqueue_messages(e, evs.data(), n) : TW_OK;
    case TW_WIRE_GLOBAL_TIME: { int accepted = 0; return tw_global_time_message(e, t, &accepted); }
    case TW_WIRE_COMPUTE_LOCAL_TIME: {
        double lt = 0.0; rc = tw_local_time(e, &lt); if (rc) return rc;
        const uint32_t kinds[2] = {TW_WIRE_TS_COMPUTE_LOCAL_TIME_ACK, TW_WIRE_TS_LOCAL_TIME};
        size_t pos = 0;
        for (int k = 0; k < 2; ++k) {
            size_t m = 0;
            uint8_t *dst = (pos + 4 <= cap) ? replies + pos + 4 : nullptr;
            rc = tw_wire_encode_ts_message(c, kinds[k], self_lp, lt, dst, dst ? cap - pos - 4 : 0, &m);
            if (dst && rc == TW_OK) { replies[pos] = (uint8_t)m; replies[pos + 1] = (uint8_t)(m >> 8); replies[pos + 2] = (uint8_t)(m >> 16); replies[pos + 3] = (uint8_t)(m >> 24); }
            else if (rc != TW_OK && rc != TW_ERR_INVALID_ARG) return rc;
            pos += 4 + m;
        }
        *replies_len = pos; *n_replies = 2;
        return pos <= cap ? TW_OK : TW_ERR_INVALID_ARG;       // *replies_len = bytes needed
    }
    case TW_WIRE_ABORT_SIMULATION: return TW_ERR_ABORTED;       // SimulationAbort, Internal/Event.hs:476-484
    default: return TW_ERR_NOT_SUPPORTED;
    }
}

}  // extern "C"
