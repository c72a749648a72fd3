// tw_models.cuh -- compiled model handlers (device + host).
//
// A handler is the device form of `handleSignal messageReceived (\msg -> ... enqueueMessage ...)`
// (Simulation/Aivika/Distributed/Optimistic/Message.hs:48-86): a pure function
//   (params, lp, state-before, received event) -> (state-after, sent events)
// Purity is what lets the engine regenerate anti-messages by replaying the handler on the saved
// pre-state instead of keeping an OutputMessageQueue (Internal/OutputMessageQueue.hs:81-104).
//
// The engine runs a handler in two parts so that the expensive one can be spread over all lanes of a warp
// (tw_core.cuh, WorkQ):
//   advance(params, lp, state, event) -> status   the sequential state transition of the LP (cheap, one LP = one lane)
//   Work                                           what is needed to produce the messages of that execution
//   emit(params, work, out[], n_out)               the messages themselves; pure, runs on whichever lane is free
// A model with kReversible = true undoes an execution by reverse computation (reverse(params, lp, state, event) restores
// the state before `event`), so the engine keeps no saved states for it: the UndoableLog (Internal/UndoableLog.hs)
// shrinks to the processed event records themselves.  Otherwise the pre-state of every processed event is saved.
// A model with kSplit = true derives Work from (pre-state, event) and does the heavy arithmetic in emit (PHOLD:
// the draws, the logarithm and the destination); a model with kSplit = false computes its messages inside
// handle() and a Work item is simply one finished message.
//
// Event.ctr: bits 0..60 = per-sender send counter (part of the saved state, so it is rollback
// invariant, unlike generateMessageSequenceNo, Internal/OutputMessageQueue.hs:119-122),
// bit 63 = anti-message toggle (Internal/Message.hs:48), bit 62 = annihilated, bit 61 = processed as no-op.
#pragma once
#include "tw_dsl.cuh"
#include "../../include/twb200.h"

namespace twb {

constexpr uint64_t kAntiBit = 1ull << 63;
constexpr uint64_t kDeadBit = 1ull << 62;
constexpr uint64_t kNoopBit = 1ull << 61;
constexpr uint64_t kCtrMask = ~(kAntiBit | kDeadBit | kNoopBit);

// ------------------------------------------------------------------ PHOLD
struct PholdModel {
    static constexpr int kMaxSends = 1;
    static constexpr uint32_t kGroup = 1;
    struct alignas(32) Event {             // 32 B: one DRAM sector
        double   recv_ts, send_ts;
        uint64_t ctr;
        uint32_t src, dst;
    };
    struct alignas(32) State {             // 32 B
        uint64_t rng_ctr, send_ctr, n_exec, hash;
    };
    struct Params {
        uint32_t n_lps, population, remote_pct, n_hot, hot_pct;
        double   lookahead, mean;
        uint64_t seed;
    };
    static Params make_params(const tw_params &p)
    {
        Params m;
        m.n_lps = p.n_lps; m.population = p.phold.population; m.remote_pct = p.phold.remote_pct;
        m.n_hot = p.phold.n_hot; m.hot_pct = p.phold.hot_pct;
        m.lookahead = p.phold.lookahead; m.mean = p.phold.mean; m.seed = p.phold.seed;
        return m;
    }
    TW_HD static State initial_state(const Params &p, uint32_t lp)
    {
        State s;
        s.rng_ctr = p.population; s.send_ctr = p.population; s.n_exec = 0;
        s.hash = mix_fin(lp_stream_key(p.seed, lp) + 1ull);
        return s;
    }
    TW_HD static Event from_wire(const tw_event &w)
    {
        Event e; e.recv_ts = w.recv_ts; e.send_ts = w.send_ts; e.ctr = w.ctr; e.src = w.src; e.dst = w.dst;
        return e;
    }
    // One event execution: draw the delay and the destination, send one message.
    static constexpr bool kSplit = true;
    struct alignas(16) Work {              // 32 B: the part of (pre-state, event) the sent message depends on
        double   now;
        uint64_t rng_ctr, send_ctr;
        uint32_t lp, anti;
    };
    TW_HD static Work make_work(uint32_t lp, const State &pre, const Event &e, bool anti)
    {
        Work w; w.now = e.recv_ts; w.rng_ctr = pre.rng_ctr; w.send_ctr = pre.send_ctr; w.lp = lp; w.anti = anti ? 1u : 0u;
        return w;
    }
    TW_HD static uint32_t work_lp(const Work &w) { return w.lp; }
    TW_HD static int advance(const Params &, uint32_t, State &s, const Event &e)
    {
        s.hash = trace_fold(s.hash, e.recv_ts, e.src, e.ctr & kCtrMask);
        s.rng_ctr += 4; s.send_ctr += 1; s.n_exec += 1;
        return H_OK;
    }
    static constexpr bool kReversible = true;
    TW_HD static void reverse(const Params &, uint32_t, State &s, const Event &e)      // exact inverse of advance
    {
        s.hash = trace_unfold(s.hash, e.recv_ts, e.src, e.ctr & kCtrMask);
        s.rng_ctr -= 4; s.send_ctr -= 1; s.n_exec -= 1;
    }
    TW_HD static void emit(const Params &p, const Work &w, Event *out, int &n_out)
    {
        const uint32_t lp = w.lp;
        const uint64_t key = lp_stream_key(p.seed, lp);
        const uint64_t d1 = stream_draw(key, w.rng_ctr);
        const uint64_t d2 = stream_draw(key, w.rng_ctr + 1);
        const double now = w.now;
        const double delay = fadd(p.lookahead, expo_variate(uniform01(d1), p.mean));
        uint32_t dst = lp;
        const uint64_t N = p.n_lps;
        if (N > 1 && (((d2 >> 32) * 100ull) >> 32) < p.remote_pct) {
            const uint64_t d3 = stream_draw(key, w.rng_ctr + 2);
            const uint64_t d4 = stream_draw(key, w.rng_ctr + 3);
            if (p.n_hot > 0 && (((d4 >> 32) * 100ull) >> 32) < p.hot_pct) {
                const uint64_t k = mulhi64(d3, p.n_hot);
                dst = (uint32_t)((k * N) / p.n_hot);
            } else {
                const uint64_t k = mulhi64(d3, N - 1);
                dst = (uint32_t)(k + (k >= lp ? 1u : 0u));
            }
        }
        Event &o = out[0];
        o.recv_ts = fadd(now, delay); o.send_ts = now; o.ctr = w.send_ctr | (w.anti ? kAntiBit : 0ull); o.src = lp; o.dst = dst;
        n_out = 1;
    }
    // the undivided handler (advance after emit): what one execution does as a whole
    TW_HD static int handle(const Params &p, uint32_t lp, State &s, const Event &e, Event *out, int &n_out)
    {
        emit(p, make_work(lp, s, e, false), out, n_out);
        return advance(p, lp, s, e);
    }
    // The j-th start-time event of LP lp (what the LP's model enqueues from runEventInStartTime); false: it has fewer.
    TW_HD static bool start_event(const Params &p, double /*start_time*/, uint32_t lp, uint32_t j, tw_event &e)
    {
        if (j >= p.population) return false;
        e.recv_ts = expo_variate(uniform01(stream_draw(lp_stream_key(p.seed, lp), j)), p.mean);
        e.send_ts = 0.0; e.ctr = j; e.src = lp; e.dst = lp; e.type = 0; e.pad = 0; e.value = 0.0;
        return true;
    }
    static size_t initial_events(const tw_params &tp, tw_event *out, size_t cap, uint32_t lp0 = 0, uint32_t lp1 = 0xFFFFFFFFu)
    {
        Params p = make_params(tp);
        size_t n = 0;
        for (uint32_t i = lp0; i < p.n_lps && i < lp1; ++i)
            for (uint32_t j = 0; ; ++j) {
                tw_event e;
                if (!start_event(p, tp.start_time, i, j, e)) break;
                if (n >= cap) return n;
                out[n++] = e;
            }
        return n;
    }
};

// ------------------------------------------------------------------ MachRep2
// LP 3r = masterModel, LP 3r+1 / 3r+2 = slaveModel 1 / 2 of replica r
// (tests/MachRep2DistributedReproducible.hs:125-229).
struct MachRepModel {
    static constexpr int kMaxSends = 3;
    static constexpr bool kSplit = false;      // messages are computed inside handle(); a Work item is one finished message
    static constexpr bool kReversible = false; // FCFS wait list, accumulators: the pre-state of every execution is saved
    struct alignas(16) Event {             // 48 B
        double   recv_ts, send_ts;
        uint64_t ctr;
        uint32_t src, dst;
        uint32_t type, pad;
        double   value;
    };
    using Work = Event;
    struct alignas(16) State {             // 64 B; master and slave views share the record
        double   total_up_time;            // master :167
        int64_t  n_rep, n_immed;           // master :168-169
        dsl::FcfsResource<3> repair_person;    // master: newFCFSResource 1 (:171-172), tw_dsl.cuh
        StdGen   gen;                      // slave: Ref DIO StdGen (:77-78), rolled back with the state
        uint64_t send_ctr, n_exec, hash;
    };
    struct Params {
        uint32_t n_lps;
        double mean_up, mean_rep, delta;
        int64_t seed1, seed2, stride;
    };
    static Params make_params(const tw_params &p)
    {
        Params m; m.n_lps = p.n_lps; m.mean_up = p.machrep.mean_up_time; m.mean_rep = p.machrep.mean_repair_time;
        m.delta = p.machrep.delta; m.seed1 = p.machrep.seed1; m.seed2 = p.machrep.seed2; m.stride = p.machrep.seed_stride;
        return m;
    }
    TW_HD static State initial_state(const Params &p, uint32_t lp)
    {
        State s;
        s.total_up_time = 0.0; s.n_rep = 0; s.n_immed = 0; s.repair_person.init(1);
        const uint32_t role = lp % 3u, r = lp / 3u;
        s.gen.s1 = 0; s.gen.s2 = 0;
        if (role == 1) s.gen = mk_std_gen(p.seed1 + (int64_t)r * p.stride);
        if (role == 2) s.gen = mk_std_gen(p.seed2 + (int64_t)r * p.stride);
        s.send_ctr = role ? 1ull : 0ull;   // slaves: ctr 0 is the START self-event
        s.n_exec = 0;
        s.hash = mix_fin(lp_stream_key(0, lp) + 1ull);
        return s;
    }
    TW_HD static Event from_wire(const tw_event &w)
    {
        Event e; e.recv_ts = w.recv_ts; e.send_ts = w.send_ts; e.ctr = w.ctr; e.src = w.src; e.dst = w.dst;
        e.type = w.type; e.pad = 0; e.value = w.value;
        return e;
    }
    TW_HD static uint32_t work_lp(const Event &w) { return w.src; }
    TW_HD static void emit(const Params &, const Event &w, Event *out, int &n_out) { out[0] = w; n_out = 1; }
    TW_HD static int handle(const Params &p, uint32_t lp, State &s, const Event &e, Event *out, int &n_out)
    {
        const uint32_t role = lp % 3u, master = lp - role;
        const double t = e.recv_ts, td = fadd(t, p.delta);
        n_out = 0;
        dsl::Outbox<Event> ob(out, n_out, s.send_ctr, lp, t);
        if (role != 0) {
            switch (e.type) {
            case TW_MR_START:
            case TW_MR_RELEASE_REPAIR_PERSON_RESP: {          // machine, :130-135 / :150-154
                const double up = dsl::random_exponential(s.gen, p.mean_up);
                const double at = fadd(td, up);
                ob.enqueueMessage(master, at, TW_MR_TOTAL_UP_TIME_CHANGE, up);
                ob.enqueueMessage(master, at, TW_MR_NREP_CHANGE, 1.0);
                ob.enqueueMessage(master, at, TW_MR_REPAIR_PERSON_COUNT, 0.0);
                break; }
            case TW_MR_REPAIR_PERSON_COUNT_RESP:              // :137-142
                if ((int64_t)e.value == 1) ob.enqueueMessage(master, td, TW_MR_NIMMED_REP_CHANGE, 1.0);
                ob.enqueueMessage(master, td, TW_MR_REQUEST_REPAIR_PERSON, 0.0);
                break;
            case TW_MR_REQUEST_REPAIR_PERSON_RESP: {          // :144-148
                const double rt = dsl::random_exponential(s.gen, p.mean_rep);
                ob.enqueueMessage(master, fadd(td, rt), TW_MR_RELEASE_REPAIR_PERSON, 0.0);
                break; }
            default: break;                                    // responses nobody subscribed to
            }
        } else {
            const uint32_t from = e.src;
            switch (e.type) {
            case TW_MR_TOTAL_UP_TIME_CHANGE:                  // :176-180
                s.total_up_time = fadd(s.total_up_time, e.value);
                ob.enqueueMessage(from, td, TW_MR_TOTAL_UP_TIME_CHANGE_RESP, 0.0); break;
            case TW_MR_NREP_CHANGE:                           // :182-186
                s.n_rep += (int64_t)e.value;
                ob.enqueueMessage(from, td, TW_MR_NREP_CHANGE_RESP, 0.0); break;
            case TW_MR_NIMMED_REP_CHANGE:                     // :188-192
                s.n_immed += (int64_t)e.value;
                ob.enqueueMessage(from, td, TW_MR_NIMMED_REP_CHANGE_RESP, 0.0); break;
            case TW_MR_REPAIR_PERSON_COUNT:                   // :194-198
                ob.enqueueMessage(from, td, TW_MR_REPAIR_PERSON_COUNT_RESP, (double)s.repair_person.count); break;
            case TW_MR_REQUEST_REPAIR_PERSON: {               // :200-206 requestResource
                bool granted;
                if (s.repair_person.request((uint8_t)(from - master), granted) != H_OK) return H_RETRY;
                if (granted) ob.enqueueMessage(from, td, TW_MR_REQUEST_REPAIR_PERSON_RESP, 0.0);
                break; }
            case TW_MR_RELEASE_REPAIR_PERSON: {               // :208-212 releaseResourceWithinEvent
                bool resumed; uint8_t w;
                if (s.repair_person.release(1, resumed, w) != H_OK) return H_RETRY;   // SimulationRetry (Internal/Event.hs:247-249, 768-813)
                ob.enqueueMessage(from, td, TW_MR_RELEASE_REPAIR_PERSON_RESP, 0.0);
                if (resumed) ob.enqueueMessage(master + w, td, TW_MR_REQUEST_REPAIR_PERSON_RESP, 0.0);
                break; }
            default: break;
            }
        }
        s.hash = trace_fold(s.hash, e.recv_ts, e.src, e.ctr & kCtrMask);
        if (e.type != TW_MR_START) s.n_exec += 1;
        return H_OK;
    }
    static constexpr uint32_t kGroup = 3;      // LPs of one closed group (kept on one GPU)
    TW_HD static tw_machrep_result result(const State &s, double t_stop)
    {
        tw_machrep_result o;
        o.total_up_time = s.total_up_time; o.n_rep = s.n_rep; o.n_immed_rep = s.n_immed;
        o.up_time_prop = s.total_up_time / (2.0 * t_stop);             // tests/MachRep2DistributedReproducible.hs:214-217
        o.immed_prop = (double)s.n_immed / (double)s.n_rep;            // :219-224
        return o;
    }
    TW_HD static bool start_event(const Params &, double start_time, uint32_t lp, uint32_t j, tw_event &e)
    {
        if (lp % 3u == 0 || j > 0) return false;               // runEventInStartTime machine, :154 (slaves only)
        e.recv_ts = start_time; e.send_ts = start_time; e.ctr = 0; e.src = lp; e.dst = lp;
        e.type = TW_MR_START; e.pad = 0; e.value = 0.0;
        return true;
    }
    static size_t initial_events(const tw_params &tp, tw_event *out, size_t cap, uint32_t lp0 = 0, uint32_t lp1 = 0xFFFFFFFFu)
    {
        const Params p = make_params(tp);
        size_t n = 0;
        for (uint32_t lp = lp0; lp < tp.n_lps && lp < lp1; ++lp) {
            tw_event e;
            if (!start_event(p, tp.start_time, lp, 0, e)) continue;
            if (n >= cap) return n;
            out[n++] = e;
        }
        return n;
    }
};

// ------------------------------------------------------------------ MachRep1
// tests/MachRep1.hs: two machines and two repair persons (no contention).  LP 3r = masterModel (:75-93), LP 3r+1 / 3r+2 =
// slaveModel (:52-72) of replica r.  The `machine` process becomes three self-scheduled events (holdProcess =
// enqueueEvent at now + dt): START / REPAIR_END draw the up time, UP_END reports it to the master with sendMessage
// (receive time = now: no lookahead at all) and draws the repair time.  The reference draws from the run's default
// generator (not reproducible); here each slave owns a StdGen like tests/MachRep2Reproducible.hs:58-67.
struct MachRep1Model {
    static constexpr int kMaxSends = 2;
    static constexpr bool kSplit = false;
    static constexpr bool kReversible = false;
    static constexpr uint32_t kGroup = 3;
    using Event = MachRepModel::Event;
    using Work = Event;
    using Params = MachRepModel::Params;
    struct alignas(16) State {             // 48 B
        double   total_up_time;            // master: Ref totalUpTime (:78)
        int64_t  n_rep;                    // master: messages received (not in the reference; a check value)
        StdGen   gen;                      // slave
        uint64_t send_ctr, n_exec, hash;
    };
    static Params make_params(const tw_params &p) { return MachRepModel::make_params(p); }
    TW_HD static State initial_state(const Params &p, uint32_t lp)
    {
        State s;
        s.total_up_time = 0.0; s.n_rep = 0;
        const uint32_t role = lp % 3u, r = lp / 3u;
        s.gen.s1 = 0; s.gen.s2 = 0;
        if (role == 1) s.gen = mk_std_gen(p.seed1 + (int64_t)r * p.stride);
        if (role == 2) s.gen = mk_std_gen(p.seed2 + (int64_t)r * p.stride);
        s.send_ctr = role ? 1ull : 0ull;   // slaves: ctr 0 is the START self-event
        s.n_exec = 0;
        s.hash = mix_fin(lp_stream_key(0, lp) + 1ull);
        return s;
    }
    TW_HD static Event from_wire(const tw_event &w) { return MachRepModel::from_wire(w); }
    TW_HD static uint32_t work_lp(const Event &w) { return w.src; }
    TW_HD static void emit(const Params &, const Event &w, Event *out, int &n_out) { out[0] = w; n_out = 1; }
    TW_HD static int handle(const Params &p, uint32_t lp, State &s, const Event &e, Event *out, int &n_out)
    {
        const uint32_t role = lp % 3u, master = lp - role;
        const double t = e.recv_ts;
        n_out = 0;
        dsl::Outbox<Event> ob(out, n_out, s.send_ctr, lp, t);
        if (role != 0) {
            switch (e.type) {
            case TW_M1_START:
            case TW_M1_REPAIR_END: {                           // machine :55-58: upTime <- randomExponential; holdProcess upTime
                const double up = dsl::random_exponential(s.gen, p.mean_up);
                ob.enqueueMessage(lp, fadd(t, up), TW_M1_UP_END, up);
                break; }
            case TW_M1_UP_END: {                               // :60-66: sendMessage masterId (TotalUpTimeChange upTime); repair
                ob.sendMessage(master, TW_M1_TOTAL_UP_TIME_CHANGE, e.value);
                const double rt = dsl::random_exponential(s.gen, p.mean_rep);
                ob.enqueueMessage(lp, fadd(t, rt), TW_M1_REPAIR_END, 0.0);
                break; }
            default: break;
            }
        } else if (e.type == TW_M1_TOTAL_UP_TIME_CHANGE) {     // :84-86: modifyRef totalUpTime (+ x)
            s.total_up_time = fadd(s.total_up_time, e.value);
            s.n_rep += 1;
        }
        s.hash = trace_fold(s.hash, e.recv_ts, e.src, e.ctr & kCtrMask);
        if (e.type != TW_M1_START) s.n_exec += 1;
        return H_OK;
    }
    TW_HD static tw_machrep_result result(const State &s, double t_stop)
    {
        tw_machrep_result o;
        o.total_up_time = s.total_up_time; o.n_rep = s.n_rep; o.n_immed_rep = 0;
        o.up_time_prop = s.total_up_time / (2.0 * t_stop);             // tests/MachRep1.hs:88-91, n = 2
        o.immed_prop = 0.0;
        return o;
    }
    TW_HD static bool start_event(const Params &, double start_time, uint32_t lp, uint32_t j, tw_event &e)
    {
        if (lp % 3u == 0 || j > 0) return false;               // runProcessInStartTime machine, :68 (slaves only)
        e.recv_ts = start_time; e.send_ts = start_time; e.ctr = 0; e.src = lp; e.dst = lp;
        e.type = TW_M1_START; e.pad = 0; e.value = 0.0;
        return true;
    }
    static size_t initial_events(const tw_params &tp, tw_event *out, size_t cap, uint32_t lp0 = 0, uint32_t lp1 = 0xFFFFFFFFu)
    {
        const Params p = make_params(tp);
        size_t n = 0;
        for (uint32_t lp = lp0; lp < tp.n_lps && lp < lp1; ++lp) {
            tw_event e;
            if (!start_event(p, tp.start_time, lp, 0, e)) continue;
            if (n >= cap) return n;
            out[n++] = e;
        }
        return n;
    }
};

}  // namespace twb
