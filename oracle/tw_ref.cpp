/*
 * tw_ref.cpp -- TEST INFRASTRUCTURE / CPU BASELINE.  Not part of the product path.
 *
 * A multi-threaded CPU restatement of the reference's per-LP Time Warp algorithm.  It exists (a) to check the
 * reference's rollback / anti-message / fossil-collection rules independently of the GPU engine -- its committed
 * trace must equal the sequential oracle's (oracle.c) for any thread count and any delivery interleaving -- and
 * (b) as the multi-core CPU baseline of bench.py (`--impl reference`, kind "port"): the Haskell engine itself cannot
 * be built in this image (no GHC).
 *
 * Structure, file for file (paths relative to /root/reference/Simulation/Aivika/Distributed/Optimistic/Internal/):
 *   InputQueue   = InputMessageQueue.hs: receive-time-sorted vector of messages with `processed` / `annihilated`
 *                  flags (:48-75); enqueue = duplicate / anti-message / straggler / future case analysis (:120-162);
 *                  binary search for the insertion point (:329-348); annihilation in place (:269-279);
 *                  reduceInputMessages strict `< GVT` (:385-398)
 *   OutputQueue  = OutputMessageQueue.hs: sent messages in send-time order (:40-48); sendMessage validation
 *                  (:65-78); rollbackOutputMessages pops the tail while sendTime > t (>= t if including) and turns
 *                  them into anti-messages (:81-104); reduceOutputMessages (:107-116)
 *   UndoLog      = UndoableLog.hs: time-stamped undo records, LIFO rollback to time t incl./excl. (:93-118),
 *                  reduceLog (:121-141); here one record per event = the state before it (copy state saving)
 *   LP::step     = Event.hs processPendingEventsCore (:196-250): drain the channel, then pop the earliest pending
 *                  event inside the window and run the handler; isTimeHorizonExceeded (:309-324)
 *   GVT          = Event.hs getLocalTime (:496-514) + TimeServer.hs timeServerGlobalTime (:406-421): min over LPs
 *                  and in-flight messages.  Samadi's acknowledgement protocol is replaced by a quiescent barrier
 *                  between worker threads (no message is in flight when the minimum is taken), which yields the
 *                  same bound; fossil collection is strict (`< GVT`, Event.hs:639-649).
 * Deliberate deviation: equal receive times are ordered by (send time, sender, send counter)
 * instead of arrival order, and the send counter is part of the rolled-back state (A.3), so that the committed
 * trace is deterministic.
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -pthread -fPIC -shared tw_ref.cpp -o libtwref.so
 */
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

extern "C" {
#include "oracle_math.h"
}

namespace {

struct Msg {
    double recv, send;
    uint64_t ctr;
    uint32_t src, dst;
    uint32_t type;
    bool anti;
    double value;
};

inline bool key_less(const Msg &a, const Msg &b)
{
    if (a.recv != b.recv) return a.recv < b.recv;
    if (a.send != b.send) return a.send < b.send;
    if (a.src != b.src) return a.src < b.src;
    return a.ctr < b.ctr;
}
inline bool same_message(const Msg &a, const Msg &b)      // Eq Message, Message.hs:70-78 (payload ignored)
{
    return a.ctr == b.ctr && a.src == b.src && a.recv == b.recv && a.send == b.send;
}

struct InItem { Msg m; bool annihilated, processed; };

// ---- model states and handlers (pure: state x message -> state', sends) ----
struct PholdParams { uint32_t n_lps, population; double lookahead, mean, stop; uint32_t remote_pct, n_hot, hot_pct; uint64_t seed; };
struct MachParams { double mean_up, mean_rep, delta, stop; int64_t seed1, seed2, stride; };

struct State {
    // PHOLD
    uint64_t rng_ctr, send_ctr, n_exec, hash;
    // MachRep2 master / slave
    double total_up; int64_t n_rep, n_immed; int32_t count; uint32_t wait[4]; uint32_t wait_len; or_stdgen gen;
};

enum { MR_START = 0, MR_TUT = 1, MR_NREP = 2, MR_NIMM = 3, MR_COUNT = 4, MR_REQ = 5, MR_REL = 6,
       MR_TUT_R = 7, MR_NREP_R = 8, MR_NIMM_R = 9, MR_COUNT_R = 10, MR_REQ_R = 11, MR_REL_R = 12 };
enum { H_OK = 0, H_RETRY = 1 };

struct Model {
    int kind;             // 1 PHOLD, 2 MachRep2
    PholdParams ph; MachParams mr;
    double stop() const { return kind == 1 ? ph.stop : mr.stop; }

    void init_state(uint32_t lp, State &s) const
    {
        memset(&s, 0, sizeof s);
        if (kind == 1) { s.rng_ctr = ph.population; s.send_ctr = ph.population; s.hash = or_fin(or_lpkey(ph.seed, lp) + 1ULL); }
        else {
            s.count = 1;
            uint32_t role = lp % 3, r = lp / 3;
            if (role == 1) s.gen = or_mkstdgen(mr.seed1 + (int64_t)r * mr.stride);
            if (role == 2) s.gen = or_mkstdgen(mr.seed2 + (int64_t)r * mr.stride);
            s.send_ctr = role ? 1 : 0;
            s.hash = or_fin(or_lpkey(0, lp) + 1ULL);
        }
    }
    static void send(State &s, uint32_t from, uint32_t to, double now, double at, uint32_t type, double value, std::vector<Msg> &out)
    {
        Msg m; m.recv = at; m.send = now; m.ctr = s.send_ctr++; m.src = from; m.dst = to; m.type = type; m.anti = false; m.value = value;
        out.push_back(m);
    }
    int handle(uint32_t lp, State &s, const Msg &e, std::vector<Msg> &out) const
    {
        if (kind == 1) {
            uint64_t key = or_lpkey(ph.seed, lp);
            uint64_t d1 = or_draw(key, s.rng_ctr), d2 = or_draw(key, s.rng_ctr + 1), d3 = or_draw(key, s.rng_ctr + 2), d4 = or_draw(key, s.rng_ctr + 3);
            double delay = ph.lookahead + or_expo(or_u01(d1), ph.mean);
            uint32_t dst = lp; uint64_t N = ph.n_lps;
            if (N > 1 && (((d2 >> 32) * 100ULL) >> 32) < ph.remote_pct) {
                if (ph.n_hot > 0 && (((d4 >> 32) * 100ULL) >> 32) < ph.hot_pct) { uint64_t k = or_mulhi(d3, ph.n_hot); dst = (uint32_t)((k * N) / ph.n_hot); }
                else { uint64_t k = or_mulhi(d3, N - 1); dst = (uint32_t)(k + (k >= lp ? 1 : 0)); }
            }
            send(s, lp, dst, e.recv, e.recv + delay, 0, 0.0, out);
            s.hash = or_trace(s.hash, e.recv, e.src, e.ctr);
            s.rng_ctr += 4; s.n_exec += 1;
            return H_OK;
        }
        const uint32_t role = lp % 3, master = lp - role;
        const double t = e.recv, td = t + mr.delta;
        if (role != 0) {
            switch (e.type) {
            case MR_START: case MR_REL_R: {
                double up = or_expo(or_stdrandom_double(&s.gen), mr.mean_up); double at = td + up;
                send(s, lp, master, t, at, MR_TUT, up, out); send(s, lp, master, t, at, MR_NREP, 1.0, out); send(s, lp, master, t, at, MR_COUNT, 0.0, out);
                break; }
            case MR_COUNT_R:
                if ((int64_t)e.value == 1) send(s, lp, master, t, td, MR_NIMM, 1.0, out);
                send(s, lp, master, t, td, MR_REQ, 0.0, out); break;
            case MR_REQ_R: { double rt = or_expo(or_stdrandom_double(&s.gen), mr.mean_rep); send(s, lp, master, t, td + rt, MR_REL, 0.0, out); break; }
            default: break;
            }
        } else {
            const uint32_t from = e.src;
            switch (e.type) {
            case MR_TUT: s.total_up += e.value; send(s, lp, from, t, td, MR_TUT_R, 0.0, out); break;
            case MR_NREP: s.n_rep += (int64_t)e.value; send(s, lp, from, t, td, MR_NREP_R, 0.0, out); break;
            case MR_NIMM: s.n_immed += (int64_t)e.value; send(s, lp, from, t, td, MR_NIMM_R, 0.0, out); break;
            case MR_COUNT: send(s, lp, from, t, td, MR_COUNT_R, (double)s.count, out); break;
            case MR_REQ:
                if (s.count == 0) { if (s.wait_len >= 3) return H_RETRY; s.wait[s.wait_len++] = from; }
                else { s.count -= 1; send(s, lp, from, t, td, MR_REQ_R, 0.0, out); }
                break;
            case MR_REL: {
                if (s.count + 1 > 1) return H_RETRY;               // SimulationRetry, Event.hs:247-249
                uint32_t waiter = 0; bool resumed = false;
                if (s.wait_len == 0) s.count += 1;
                else { waiter = s.wait[0]; for (uint32_t k = 1; k < s.wait_len; ++k) s.wait[k - 1] = s.wait[k]; s.wait_len--; resumed = true; }
                send(s, lp, from, t, td, MR_REL_R, 0.0, out);
                if (resumed) send(s, lp, waiter, t, td, MR_REQ_R, 0.0, out);
                break; }
            default: break;
            }
        }
        s.hash = or_trace(s.hash, e.recv, e.src, e.ctr);
        if (e.type != MR_START) s.n_exec += 1;
        return H_OK;
    }
};

struct Stats { std::atomic<uint64_t> executed{0}, rolled_back{0}, rollbacks{0}, anti{0}, annihilated{0}, stragglers{0}, gvt_rounds{0}; };

struct LP {
    uint32_t id;
    State st;
    std::vector<InItem> input;                      // InputMessageQueue: sorted by key; [0, nproc) processed
    size_t nproc = 0;
    struct LogRec { double t; State before; uint64_t out_idx; };
    std::vector<LogRec> log;                        // UndoableLog: (event time, state before it, #messages sent before it), parallel to processed items
    std::deque<Msg> output;                         // OutputMessageQueue, send-time order
    uint64_t out_base = 0;                          // messages already fossil-collected from the front of `output`
    double lvt = -__builtin_inf();                  // queueTime
};

struct Engine;
struct Worker {
    std::mutex mu; std::vector<Msg> mailbox;        // Channel.hs: multi-producer mailbox
    std::vector<Msg> inbox_local;
};

struct Engine {
    Model model; int T; double horizon; uint32_t batch;
    std::vector<LP> lps; std::vector<Worker> workers; std::vector<uint32_t> lo, hi;
    Stats stats;
    std::atomic<uint64_t> sent_in_round{0};
    // barrier
    std::mutex bmu; std::condition_variable bcv; int bcount = 0; uint64_t bgen = 0;
    std::vector<double> local_min; double gvt = 0.0; std::atomic<int> error{0}; uint64_t stalled = 0;

    void barrier()
    {
        std::unique_lock<std::mutex> lk(bmu);
        uint64_t g = bgen;
        if (++bcount == T) { bcount = 0; bgen++; bcv.notify_all(); }
        else bcv.wait(lk, [&] { return bgen != g; });
    }
    int owner(uint32_t lp) const { return (int)(std::upper_bound(hi.begin(), hi.end(), lp) - hi.begin()); }

    void deliver(int self, const Msg &m)                // sendMessageDIO: into the owner's mailbox
    {
        int w = owner(m.dst);
        if (w == self) { workers[w].inbox_local.push_back(m); }
        else { std::lock_guard<std::mutex> g(workers[w].mu); workers[w].mailbox.push_back(m); }
        sent_in_round.fetch_add(1, std::memory_order_relaxed);
    }

    // rollbackInputMessages / rollbackLog / rollbackOutputMessages: un-process items [pos, nproc)
    void rollback(int self, LP &lp, size_t pos, double t, bool including)
    {
        if (pos >= lp.nproc) return;
        stats.rollbacks++;
        uint64_t undone = 0;
        for (size_t i = pos; i < lp.nproc; ++i) { lp.input[i].processed = false; if (!lp.input[i].m.anti && !lp.input[i].annihilated) undone++; }
        stats.rolled_back += undone;
        lp.st = lp.log[pos].before;                       // rollbackLog: undo to before the first rolled-back event
        const uint64_t keep = lp.log[pos].out_idx;
        lp.log.resize(pos);
        lp.nproc = pos;
        lp.lvt = pos ? lp.input[pos - 1].m.recv : -__builtin_inf();
        // rollbackOutputMessages: the reference pops the tail while sendTime > t (>= t when including).  With the
        // deterministic tie-break an event at exactly time t may be rolled back as well, so the cut is taken at the
        // first rolled-back event instead (same set whenever receive times are distinct).
        (void)t; (void)including;
        std::vector<Msg> antis;
        while (lp.out_base + lp.output.size() > keep) {
            Msg a = lp.output.back(); a.anti = true; antis.push_back(a); lp.output.pop_back();
        }
        for (auto it = antis.rbegin(); it != antis.rend(); ++it) { deliver(self, *it); stats.anti++; }
    }

    // InputMessageQueue.enqueueMessage
    void enqueue(int self, LP &lp, const Msg &m)
    {
        if (m.recv < gvt) { error = 4; return; }           // "Received the outdated message"
        if (m.anti) {
            size_t i = 0; bool found = false;
            for (; i < lp.input.size(); ++i) if (!lp.input[i].m.anti && !lp.input[i].annihilated && same_message(lp.input[i].m, m)) { found = true; break; }
            if (!found) { error = 6; return; }
            if (i < lp.nproc) rollback(self, lp, i, m.recv, true);      // twin processed: rollback including its time
            lp.input.erase(lp.input.begin() + (long)i);                  // annihilateMessage: vectorDeleteAt
            stats.annihilated++;
            return;
        }
        // lookupRightMessageIndex: insertion point by key (binary search)
        InItem it{m, false, false};
        auto pos = std::upper_bound(lp.input.begin(), lp.input.end(), it, [](const InItem &a, const InItem &b) { return key_less(a.m, b.m); });
        size_t p = (size_t)(pos - lp.input.begin());
        if (p < lp.nproc) { stats.stragglers++; rollback(self, lp, p, m.recv, false); }   // straggler: rollback not including
        lp.input.insert(lp.input.begin() + (long)p, it);
    }

    // processPendingEventsCore: up to `batch` events with time <= limit
    void advance(int self, LP &lp, double limit, std::vector<Msg> &out)
    {
        for (uint32_t b = 0; b < batch && lp.nproc < lp.input.size(); ++b) {
            InItem &item = lp.input[lp.nproc];
            if (!(item.m.recv <= limit)) break;
            State before = lp.st;
            out.clear();
            int rc = model.handle(lp.id, lp.st, item.m, out);
            if (rc != H_OK) { lp.st = before; break; }                   // retry later: wait for the rollback
            lp.log.push_back(LP::LogRec{item.m.recv, before, lp.out_base + lp.output.size()});
            item.processed = true; lp.nproc++; lp.lvt = item.m.recv;
            stats.executed++;
            for (const Msg &m : out) {
                if (m.send > m.recv) { error = 3; return; }
                lp.output.push_back(m);
                deliver(self, m);
            }
        }
    }

    void fossil(LP &lp)                                    // reduceEvents, strict < gvt
    {
        size_t k = 0;
        while (k < lp.nproc && lp.input[k].m.recv < gvt) ++k;
        if (k) { lp.input.erase(lp.input.begin(), lp.input.begin() + (long)k); lp.log.erase(lp.log.begin(), lp.log.begin() + (long)k); lp.nproc -= k; }
        while (!lp.output.empty() && lp.output.front().send < gvt) { lp.output.pop_front(); lp.out_base++; }
    }

    bool drain(int self)
    {
        Worker &w = workers[self];
        std::vector<Msg> got;
        { std::lock_guard<std::mutex> g(w.mu); got.swap(w.mailbox); }
        bool any = !got.empty() || !w.inbox_local.empty();
        while (!w.inbox_local.empty() || !got.empty()) {
            std::vector<Msg> loc; loc.swap(w.inbox_local);
            for (const Msg &m : got) enqueue(self, lps[m.dst], m);
            for (const Msg &m : loc) enqueue(self, lps[m.dst], m);
            got.clear();
        }
        return any;
    }

    void run_worker(int self)
    {
        const double stop = model.stop();
        std::vector<Msg> out;
        for (;;) {
            // ---- optimistic phase: a few sweeps over the local LPs
            for (int sweep = 0; sweep < 4; ++sweep) {
                drain(self);
                double limit = stop;
                if (horizon >= 0.0 && gvt + horizon < limit) limit = gvt + horizon;
                for (uint32_t i = lo[self]; i < hi[self]; ++i) advance(self, lps[i], limit, out);
            }
            // ---- GVT round: quiesce (no message in flight), then min over LPs
            for (;;) {
                barrier();
                if (self == 0) sent_in_round = 0;
                barrier();
                drain(self);                                 // may roll back and send anti-messages
                barrier();
                if (sent_in_round.load() == 0) break;
            }
            double m = __builtin_inf();
            for (uint32_t i = lo[self]; i < hi[self]; ++i) { LP &lp = lps[i]; if (lp.nproc < lp.input.size()) m = std::min(m, lp.input[lp.nproc].m.recv); }
            local_min[self] = m;
            barrier();
            if (self == 0) {
                double g = __builtin_inf(); for (double v : local_min) g = std::min(g, v);
                if (g == gvt) { if (++stalled > 200000) error = 7; } else stalled = 0;       // "deadlock when retrying"
                gvt = g; stats.gvt_rounds++;
            }
            barrier();
            for (uint32_t i = lo[self]; i < hi[self]; ++i) fossil(lps[i]);
            if (!(gvt <= stop) || error.load()) break;
        }
    }
};

struct twref_stats { uint64_t committed, executed, rolled_back, rollbacks, anti_sent, annihilated, stragglers, gvt_rounds; double gvt; };

int run(Engine &E, const std::vector<Msg> &initial, uint64_t *n_exec, uint64_t *hash, twref_stats *st)
{
    const uint32_t N = (uint32_t)E.lps.size();
    if (E.T < 1) E.T = 1;
    if ((uint32_t)E.T > N) E.T = (int)N;
    E.workers = std::vector<Worker>((size_t)E.T);
    E.local_min.assign((size_t)E.T, 0.0);
    for (int t = 0; t < E.T; ++t) { E.lo.push_back((uint32_t)((uint64_t)N * t / E.T)); E.hi.push_back((uint32_t)((uint64_t)N * (t + 1) / E.T)); }
    for (uint32_t i = 0; i < N; ++i) { E.lps[i].id = i; E.model.init_state(i, E.lps[i].st); }
    for (const Msg &m : initial) E.workers[(size_t)E.owner(m.dst)].mailbox.push_back(m);
    std::vector<std::thread> th;
    for (int t = 0; t < E.T; ++t) th.emplace_back([&E, t] { E.run_worker(t); });
    for (auto &x : th) x.join();
    uint64_t committed = 0;
    for (uint32_t i = 0; i < N; ++i) { if (n_exec) n_exec[i] = E.lps[i].st.n_exec; if (hash) hash[i] = E.lps[i].st.hash; committed += E.lps[i].st.n_exec; }
    if (st) {
        st->committed = committed; st->executed = E.stats.executed; st->rolled_back = E.stats.rolled_back; st->rollbacks = E.stats.rollbacks;
        st->anti_sent = E.stats.anti; st->annihilated = E.stats.annihilated; st->stragglers = E.stats.stragglers; st->gvt_rounds = E.stats.gvt_rounds; st->gvt = E.gvt;
    }
    return E.error.load();
}

}  // namespace

extern "C" {

struct twref_phold_params { uint32_t n_lps, population; double lookahead, mean, stop_time; uint32_t remote_pct, n_hot, hot_pct, pad; uint64_t seed; };
struct twref_machrep_params { uint32_t n_replicas, pad; double mean_up_time, mean_repair_time, delta, stop_time; int64_t seed1, seed2, seed_stride; };
struct twref_machrep_result { double total_up_time; int64_t n_rep, n_immed_rep; double up_time_prop, immed_prop; };

int twref_phold_run(const twref_phold_params *p, int threads, double horizon, uint32_t batch, uint64_t *n_exec, uint64_t *hash, twref_stats *st)
{
    Engine E; E.model.kind = 1;
    E.model.ph = PholdParams{p->n_lps, p->population, p->lookahead, p->mean, p->stop_time, p->remote_pct, p->n_hot, p->hot_pct, p->seed};
    E.T = threads; E.horizon = horizon; E.batch = batch ? batch : 4;
    E.lps.resize(p->n_lps);
    std::vector<Msg> init; init.reserve((size_t)p->n_lps * p->population);
    for (uint32_t i = 0; i < p->n_lps; ++i) {
        uint64_t key = or_lpkey(p->seed, i);
        for (uint32_t j = 0; j < p->population; ++j) {
            Msg m; m.recv = or_expo(or_u01(or_draw(key, j)), p->mean); m.send = 0.0; m.ctr = j; m.src = i; m.dst = i; m.type = 0; m.anti = false; m.value = 0.0;
            init.push_back(m);
        }
    }
    return run(E, init, n_exec, hash, st);
}

int twref_machrep_run(const twref_machrep_params *p, int threads, double horizon, uint32_t batch, uint64_t *n_exec, uint64_t *hash,
                      twref_machrep_result *res, twref_stats *st)
{
    Engine E; E.model.kind = 2;
    E.model.mr = MachParams{p->mean_up_time, p->mean_repair_time, p->delta, p->stop_time, p->seed1, p->seed2, p->seed_stride};
    E.T = threads; E.horizon = horizon; E.batch = batch ? batch : 4;
    E.lps.resize((size_t)3 * p->n_replicas);
    std::vector<Msg> init;
    for (uint32_t lp = 0; lp < 3 * p->n_replicas; ++lp) {
        if (lp % 3 == 0) continue;
        Msg m; m.recv = 0.0; m.send = 0.0; m.ctr = 0; m.src = lp; m.dst = lp; m.type = MR_START; m.anti = false; m.value = 0.0;
        init.push_back(m);
    }
    int rc = run(E, init, n_exec, hash, st);
    if (res)
        for (uint32_t r = 0; r < p->n_replicas; ++r) {
            const State &s = E.lps[(size_t)3 * r].st;
            res[r].total_up_time = s.total_up; res[r].n_rep = s.n_rep; res[r].n_immed_rep = s.n_immed;
            res[r].up_time_prop = s.total_up / (2.0 * p->stop_time); res[r].immed_prop = (double)s.n_immed / (double)s.n_rep;
        }
    return rc;
}

int twref_hardware_threads(void) { unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }

}  // extern "C"
