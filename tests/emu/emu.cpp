// emu.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libtwb200.so, never used by bench.py).
//
// Compiles the engine's device functions (tw_core.cuh: lp_step / emit_event / route_event and the model
// handlers) with g++ and drives them sequentially on the CPU: one "kernel launch" = a loop over all LPs of
// all emulated partitions in a shuffled order.  Purpose: debug the rollback / annihilation / GVT logic of
// the exact source the GPU runs, in a container that has no GPU, and exercise the multi-partition (world > 1)
// addressing without NVLink.  It is a development aid and a `-m "not gpu"` test of the host-side logic; parity
// claims are made only by the `-m gpu` tests that call the CUDA build through the C ABI.
//
// Two builds (tests/helpers.py): libtwemu.so runs the scalar paths LP by LP; libtwemu_simt.so (-DTW_SIMT_EMU) runs the
// warp-level device paths -- work queue with deferred delivery, whole-warp region search, reconvergence points -- with 32
// host threads as the lanes of a warp (simt.hpp), one warp at a time.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef TW_SIMT_EMU
#include "simt.hpp"
#endif
#include "../../aivika-distributed_b200/csrc/tw_core.cuh"

using namespace twb;

template <class M>
struct Part {
    Dev<M> d;
    std::vector<typename M::Event> pages;
    std::vector<typename M::State> saved, state;
    std::vector<LPHdr> hdr;
    std::vector<LPIn> in;
    std::vector<typename M::Event> stage;      // incoming staging rings [sender][parity][stage_cap]
    std::vector<uint32_t> stage_cnt;           // this partition's sender-side counters
    std::vector<uint32_t> recv_cnt;            // what the "mailbox" told this partition: messages staged by every sender
    std::vector<unsigned long long> grp_gvt;
};

static uint64_t rng_next(uint64_t &s) { s += kGolden; return mix_fin(s); }

#ifdef TW_SIMT_EMU
// 32 lane threads that run one warp-visit (d.tiles tiles of 32 LPs, then the queue drain) per job, like k_step does.
template <class M>
struct WarpPool {
    simt::Warp warp;
    simt::Barrier start{33}, done{33};
    std::vector<std::thread> lanes;
    std::vector<char> smem;
    Acc accs[32];
    const Dev<M> *d = nullptr; uint32_t tile0 = 0; double gvt = 0, wend = 0; uint32_t step = 0; bool quit = false;
    // group mode (k_run_groups): the warp runs the super-steps of groups [g0, g0 + gpw) back to back
    bool group_mode = false; uint32_t g0 = 0, gpw = 1, max_rounds = 0; double cap = 0; double graw[32]; uint32_t rounds[32]; bool mine[32];
    WarpPool() : smem(WorkQ<M>::kWarpBytes + 16)
    {
        for (unsigned lane = 0; lane < 32; ++lane) lanes.emplace_back([this, lane] {
            simt::g_warp = &warp; simt::g_lane = lane; threadIdx.x = lane;
            for (;;) {
                start.wait();
                if (quit) return;
                Acc a; a.init();
                WorkQ<M> q; q.init((void *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15));
                if (group_mode) {
                    if constexpr (M::kGroup > 1)
                        mine[lane] = group_rounds<M>(*d, a, q, g0, gpw, max_rounds, step, cap, graw[lane], rounds[lane]);
                } else {
                    for (uint32_t t = 0; t < d->tiles; ++t) lp_step<M>(*d, a, q, (tile0 + t) * 32u + lane, gvt, wend, step);
                    q.drain(*d, a, step & 1u);
                }
                accs[lane] = a;
                done.wait();
            }
        });
    }
    void run(const Dev<M> &dev, uint32_t first_tile, double g, double w, uint32_t s, Acc &acc)
    {
        d = &dev; tile0 = first_tile; gvt = g; wend = w; step = s;
        start.wait(); done.wait();
        for (int i = 0; i < 32; ++i) acc.merge(accs[i]);
    }
    // one warp-visit of k_run_groups: returns the rounds run; the groups' GVTs are written back like the kernel does
    uint32_t run_groups(Dev<M> &dev, uint32_t first_group, uint32_t groups_per_warp, uint32_t budget, uint32_t step0, double bound, Acc &acc)
    {
        d = &dev; group_mode = true; g0 = first_group; gpw = groups_per_warp; max_rounds = budget; step = step0; cap = bound;
        start.wait(); done.wait();
        group_mode = false;
        for (int i = 0; i < 32; ++i) {
            Acc a = accs[i]; a.pend_min = pos_inf(); a.emit_min = pos_inf(); acc.merge(a);
            if (mine[i] && i % (int)M::kGroup == 0) dev.grp_gvt[first_group + i / M::kGroup] = ordered_bits(graw[i]);
        }
        return rounds[0];
    }
    ~WarpPool() { quit = true; start.wait(); for (auto &t : lanes) t.join(); }
};
#endif

template <class M>
static int emu_run_t(const tw_params *p, uint64_t shuffle_seed, uint64_t *n_exec, uint64_t *hash,
                     tw_machrep_result *mr, tw_run_stats *st, uint32_t gvt_every = 1, uint32_t tiles = 1)
{
    const uint32_t W = p->world ? p->world : 1;
    if (W > (uint32_t)kMaxWorld) return TW_ERR_INVALID_ARG;
    const uint32_t cap = p->queue_capacity ? p->queue_capacity : 128;
    std::vector<Part<M>> parts(W);
    uint32_t n_per = (p->n_lps + W - 1) / W;
    if (M::kGroup > 1 && W > 1) n_per = ((n_per + M::kGroup - 1) / M::kGroup) * M::kGroup;
    for (uint32_t r = 0; r < W; ++r) {
        Part<M> &P = parts[r]; Dev<M> &d = P.d;
        memset(&d, 0, sizeof d);
        d.n_global = p->n_lps; d.world = W; d.rank = r; d.n_per = n_per; d.lp_base = n_per * r;
        d.n_local = d.lp_base >= p->n_lps ? 0 : std::min(n_per, p->n_lps - d.lp_base);
        d.cap = cap; d.mask = cap - 1; d.tiles = tiles ? tiles : 1;
        d.max_batch = p->max_events_per_step ? std::min<uint32_t>(p->max_events_per_step, kMaxBatch) : 8;
        d.proc_limit = p->proc_limit ? p->proc_limit : cap / 4;
        d.fossil_thresh = p->fossil_threshold ? p->fossil_threshold : cap / 8;
        if (d.fossil_thresh > d.proc_limit / 2) d.fossil_thresh = d.proc_limit / 2 ? d.proc_limit / 2 : 1;      // as Engine::init
        d.stop = p->stop_time; d.horizon = (p->time_horizon >= 0.0) ? p->time_horizon : -1.0;
        d.mp = M::make_params(*p);
        const size_t nl = std::max<size_t>(d.n_local, 1);
        P.pages.resize(nl * cap); P.saved.resize(nl * cap); P.state.resize(nl); P.hdr.resize(nl); P.in.resize(nl);
        d.pages = P.pages.data(); d.saved = P.saved.data(); d.state = P.state.data(); d.hdr = P.hdr.data(); d.in = P.in.data();
        for (uint32_t l = 0; l < d.n_local; ++l) {
            LPHdr h; h.head = h.np_hp = h.consumed = h.rollbacks = 0; h.lvt = neg_inf(); h.min_pend = pos_inf();
            P.hdr[l] = h; LPIn in; in.word = 0; P.in[l] = in;
            P.state[l] = M::initial_state(d.mp, d.lp_base + l);
        }
    }
    // cross-partition exchange as on the GPUs: staging rings in the receiver's memory, sender-local counters
    const uint32_t stage_cap = p->outbox_capacity ? p->outbox_capacity : std::max<uint32_t>(8192u, 2u * n_per);
    for (uint32_t r = 0; r < W; ++r) {
        parts[r].stage.resize((size_t)W * 2u * stage_cap); parts[r].stage_cnt.assign(kMaxWorld * kCntStride, 0u);
        parts[r].recv_cnt.assign(kMaxWorld, 0u);
        parts[r].d.stage_in = parts[r].stage.data(); parts[r].d.stage_cnt = parts[r].stage_cnt.data(); parts[r].d.stage_cap = stage_cap;
        parts[r].d.flags = p->flags;
    }
    for (uint32_t r = 0; r < W; ++r)
        for (uint32_t q = 0; q < W; ++q) parts[r].d.stage_out[q] = parts[q].stage.data() + (size_t)r * 2u * stage_cap;

    // initial events through the same routing function the k_route kernel uses
    std::vector<tw_event> init((size_t)p->n_lps * (p->model == TW_MODEL_PHOLD ? p->phold.population : 1) + 8);
    size_t n_init = 0;
    n_init = M::initial_events(*p, init.data(), init.size());
    uint32_t step = 0; double gvt = p->start_time;
    Acc acc; acc.init();
    for (uint32_t r = 0; r < W; ++r)
        for (size_t i = 0; i < n_init; ++i) route_event<M>(parts[r].d, acc, (step & 1u) ^ 1u, init[i]);
    if (acc.err) return (int)acc.err;
    gvt = std::min(gvt, acc.emit_min);

    std::vector<std::pair<uint32_t, uint32_t>> order;
#ifdef TW_SIMT_EMU
    WarpPool<M> pool;                       // order = (partition, first tile of a warp)
    for (uint32_t r = 0; r < W; ++r)
        for (uint32_t t = 0; t * 32u < std::max<uint32_t>(parts[r].d.n_local, 1u); t += parts[r].d.tiles) order.push_back({r, t});
#else
    for (uint32_t r = 0; r < W; ++r) for (uint32_t l = 0; l < parts[r].d.n_local; ++l) order.push_back({r, l});
#endif
    uint64_t rs = shuffle_seed;
    tw_run_stats S; memset(&S, 0, sizeof S);
    uint64_t stall = 0;
    double local_min = gvt;                 // what the partition would report as its local time; `gvt` may lag behind it
    while (gvt <= p->stop_time) {
        double wend = p->stop_time;
        if (p->time_horizon >= 0.0) wend = std::min(wend, gvt + p->time_horizon);
        if (shuffle_seed) for (size_t i = order.size(); i > 1; --i) std::swap(order[i - 1], order[rng_next(rs) % i]);
        acc.init();
        // k_deliver: every partition routes what the others staged for it during the previous super-step
        for (uint32_t r = 0; r < W; ++r)
            for (uint32_t q = 0; q < W; ++q)
                for (uint32_t j = 0; j < parts[r].recv_cnt[q]; ++j) deliver_staged<M>(parts[r].d, acc, (step & 1u) ^ 1u, q, j);
#ifdef TW_SIMT_EMU
        for (auto &o : order) pool.run(parts[o.first].d, o.second, gvt, wend, step, acc);
#else
        for (auto &o : order) { WorkQ<M> q; q.init(nullptr); lp_step<M>(parts[o.first].d, acc, q, o.second, gvt, wend, step); }
#endif
        // gvt_exchange_warp: the counts travel with the mailbox entries, the sender's counters start again at zero
        for (uint32_t r = 0; r < W; ++r)
            for (uint32_t q = 0; q < W; ++q) {
                parts[q].recv_cnt[r] = std::min(parts[r].stage_cnt[q * kCntStride], stage_cap);
                parts[r].stage_cnt[q * kCntStride] = 0u;
            }
        if (acc.err == TW_ERR_ANTI_UNMATCHED || acc.err == TW_ERR_TIME_ORDER || acc.err == TW_ERR_OUTDATED_MESSAGE) {
            // engine.cu k_diagnose_overflow: an overflowed ring elsewhere is the cause of such an inconsistency
            for (uint32_t r = 0; r < W; ++r)
                for (uint32_t l = 0; l < parts[r].d.n_local; ++l)
                    if (in_total(parts[r].in[l].word) - hdr_head_prev(parts[r].hdr[l]) > cap) { acc.err = TW_ERR_QUEUE_OVERFLOW; acc.err_lp = parts[r].d.lp_base + l; }
        }
        if (acc.err) { fprintf(stderr, "emu: error %u at LP %u step %u\n", acc.err, acc.err_lp, step); return (int)acc.err; }
        if (getenv("EMU_TRACE")) fprintf(stderr, "step %u gvt %.17g exec %u rb %u rollbacks %u strag %u anti %u annih %u retry %u sent %u pend %.17g emit %.17g\n", step, gvt, acc.exec, acc.rb_events, acc.rollbacks, acc.strag, acc.anti, acc.annih, acc.retry, acc.sent_local + acc.sent_remote, acc.pend_min, acc.emit_min);
        double g = std::min(acc.pend_min, acc.emit_min);
        // (a quantum lower bound taken over a lagging window base may lie a hair below the previous step's value; the engine
        // publishes the running maximum -- engine.cu, block_finish --, and so does the emulation, within a relative 1e-6)
        if (g < local_min && local_min - g <= 1e-6 * std::max(1.0, std::fabs(local_min))) g = local_min;
        if (g < local_min) { fprintf(stderr, "emu: GVT went backwards %.17g -> %.17g at step %u\n", local_min, g, step); return 100; }
        if (g == local_min) { if (++stall > 100000) return TW_ERR_DEADLOCK; } else stall = 0;
        local_min = g; step++;
        // an external Time Server hands out the global time only now and then (tw_global_time_message): until then the
        // engine keeps working with the old, lower bound
        if (step % gvt_every == 0 || !(g <= p->stop_time)) gvt = g;
        S.steps++; S.executed += acc.exec; S.rolled_back += acc.rb_events; S.rollbacks += acc.rollbacks; S.stragglers += acc.strag;
        S.anti_sent += acc.anti; S.annihilated += acc.annih; S.retries += acc.retry; S.sent_local += acc.sent_local; S.sent_remote += acc.sent_remote;
    }
    S.gvt = gvt;
    // as Engine::run: an overflow of the last super-steps has no next drain that would notice it
    for (uint32_t r = 0; r < W; ++r)
        for (uint32_t l = 0; l < parts[r].d.n_local; ++l)
            if (in_total(parts[r].in[l].word) - hdr_head_prev(parts[r].hdr[l]) > cap) {
                fprintf(stderr, "emu: error %u at LP %u after the last step\n", (unsigned)TW_ERR_QUEUE_OVERFLOW, parts[r].d.lp_base + l);
                return TW_ERR_QUEUE_OVERFLOW;
            }
    // anti-messages sent in the final steps can still sit unconsumed in inboxes (their twins lie beyond the stop time);
    // reported through the otherwise unused kernel_launches field so tests can close the balance anti_sent == annihilated + left
    for (uint32_t r = 0; r < W; ++r)
        for (uint32_t l = 0; l < parts[r].d.n_local; ++l)
            for (uint32_t a = parts[r].hdr[l].consumed; a != in_total(parts[r].in[l].word); ++a)
                if (parts[r].pages[(size_t)l * cap + (a & (cap - 1))].ctr & kAntiBit) S.kernel_launches++;
    for (uint32_t r = 0; r < W; ++r)
        for (uint32_t l = 0; l < parts[r].d.n_local; ++l) {
            const uint32_t lp = parts[r].d.lp_base + l;
            if (n_exec) n_exec[lp] = parts[r].state[l].n_exec;
            if (hash) hash[lp] = parts[r].state[l].hash;
            S.committed += parts[r].state[l].n_exec; S.trace_hash_sum += parts[r].state[l].hash;
        }
    if (st) *st = S;
    if (mr) {
        if constexpr (M::kGroup == 3) {
            for (uint32_t lp = 0; lp < p->n_lps; lp += 3) {
                const uint32_t r = lp / n_per;
                mr[lp / 3] = M::result(parts[r].state[lp - parts[r].d.lp_base], p->stop_time);
            }
        }
    }
    return TW_OK;
}

#ifdef TW_SIMT_EMU
// Closed-group models through the warp-local super-step loop (k_run_groups): launches of `budget` rounds until every group
// has passed the stop time.  gvt_every > 1 plays an external Time Server whose bound lags behind.
template <class M>
static int emu_run_groups_t(const tw_params *p, uint64_t shuffle_seed, uint32_t gpw, uint32_t budget, uint64_t *n_exec, uint64_t *hash,
                            tw_machrep_result *mr, tw_run_stats *st, uint32_t gvt_every)
{
    if (p->world > 1 || M::kGroup < 2) return TW_ERR_NOT_SUPPORTED;
    const uint32_t cap = p->queue_capacity ? p->queue_capacity : 128;
    Part<M> P; Dev<M> &d = P.d;
    memset(&d, 0, sizeof d);
    d.n_global = p->n_lps; d.world = 1; d.rank = 0; d.n_per = p->n_lps; d.lp_base = 0; d.n_local = p->n_lps;
    d.cap = cap; d.mask = cap - 1; d.tiles = 1; d.closed = 1; d.flags = p->flags;
    d.max_batch = p->max_events_per_step ? std::min<uint32_t>(p->max_events_per_step, kMaxBatch) : 8;
    d.proc_limit = p->proc_limit ? p->proc_limit : cap / 4;
    d.fossil_thresh = p->fossil_threshold ? p->fossil_threshold : cap / 8;
    if (d.fossil_thresh > d.proc_limit / 2) d.fossil_thresh = d.proc_limit / 2 ? d.proc_limit / 2 : 1;
    d.stop = p->stop_time; d.horizon = (p->time_horizon >= 0.0) ? p->time_horizon : -1.0;
    d.mp = M::make_params(*p);
    const size_t nl = std::max<size_t>(d.n_local, 1);
    P.pages.resize(nl * cap); P.saved.resize(nl * cap); P.state.resize(nl); P.hdr.resize(nl); P.in.resize(nl);
    P.stage_cnt.assign(kMaxWorld * kCntStride, 0u);
    const uint32_t n_groups = d.n_local / M::kGroup;
    P.grp_gvt.assign(std::max<uint32_t>(n_groups, 1u), ordered_bits(p->start_time));
    d.pages = P.pages.data(); d.saved = P.saved.data(); d.state = P.state.data(); d.hdr = P.hdr.data(); d.in = P.in.data();
    d.stage_cnt = P.stage_cnt.data(); d.grp_gvt = P.grp_gvt.data();
    for (uint32_t l = 0; l < d.n_local; ++l) {
        LPHdr h; h.head = h.np_hp = h.consumed = h.rollbacks = 0; h.lvt = neg_inf(); h.min_pend = pos_inf();
        P.hdr[l] = h; LPIn in; in.word = 0; P.in[l] = in;
        P.state[l] = M::initial_state(d.mp, l);
    }
    std::vector<tw_event> init((size_t)p->n_lps + 8);
    const size_t n_init = M::initial_events(*p, init.data(), init.size());
    const uint32_t step0 = 2;
    Acc acc; acc.init();
    for (size_t i = 0; i < n_init; ++i) route_event<M>(d, acc, (step0 & 1u) ^ 1u, init[i]);
    if (acc.err) return (int)acc.err;
    WarpPool<M> pool;
    std::vector<uint32_t> order;
    for (uint32_t g = 0; g < n_groups; g += gpw) order.push_back(g);
    uint64_t rs = shuffle_seed;
    tw_run_stats S; memset(&S, 0, sizeof S);
    double bound = gvt_every > 1 ? p->start_time : pos_inf();
    for (uint64_t launch = 0;; ++launch) {
        if (launch > 2000000) return TW_ERR_DEADLOCK;
        if (shuffle_seed) for (size_t i = order.size(); i > 1; --i) std::swap(order[i - 1], order[rng_next(rs) % i]);
        acc.init();
        uint32_t max_rounds = 0;
        for (uint32_t g : order) max_rounds = std::max(max_rounds, pool.run_groups(d, g, gpw, budget, step0, bound, acc));
        if (acc.err) { fprintf(stderr, "emu groups: error %u at LP %u\n", acc.err, acc.err_lp); return (int)acc.err; }
        S.steps += max_rounds; S.executed += acc.exec; S.rolled_back += acc.rb_events; S.rollbacks += acc.rollbacks; S.stragglers += acc.strag;
        S.anti_sent += acc.anti; S.annihilated += acc.annih; S.retries += acc.retry; S.sent_local += acc.sent_local;
        double local = pos_inf();
        for (uint32_t g = 0; g < n_groups; ++g) local = std::min(local, from_ordered(P.grp_gvt[g]));
        S.gvt = std::min(local, bound);
        if (!(S.gvt <= p->stop_time)) break;
        if (gvt_every > 1 && launch % gvt_every == gvt_every - 1) bound = local;      // the next GlobalTimeMessage
    }
    for (uint32_t l = 0; l < d.n_local; ++l)
        if (in_total(P.in[l].word) - hdr_head_prev(P.hdr[l]) > cap) return TW_ERR_QUEUE_OVERFLOW;
    for (uint32_t l = 0; l < d.n_local; ++l) {
        if (n_exec) n_exec[l] = P.state[l].n_exec;
        if (hash) hash[l] = P.state[l].hash;
        S.committed += P.state[l].n_exec; S.trace_hash_sum += P.state[l].hash;
    }
    if (st) *st = S;
    if (mr) { if constexpr (M::kGroup == 3) for (uint32_t lp = 0; lp < p->n_lps; lp += 3) mr[lp / 3] = M::result(P.state[lp], p->stop_time); }
    return TW_OK;
}
extern "C" int emu_run_groups(const tw_params *p, uint64_t shuffle_seed, uint32_t gpw, uint32_t budget, uint32_t gvt_every,
                              uint64_t *n_exec, uint64_t *hash, tw_machrep_result *mr, tw_run_stats *st)
{
    if (gpw < 1) gpw = 1;
    if (budget < 2) budget = 2;
    if (p->model == TW_MODEL_MACHREP2) return emu_run_groups_t<MachRepModel>(p, shuffle_seed, gpw, budget & ~1u, n_exec, hash, mr, st, gvt_every);
    if (p->model == TW_MODEL_MACHREP1) return emu_run_groups_t<MachRep1Model>(p, shuffle_seed, gpw, budget & ~1u, n_exec, hash, mr, st, gvt_every);
    return TW_ERR_NOT_SUPPORTED;
}
#endif

// building blocks of tw_dsl.cuh against the semantics of aivika-transformers' FCFS resource:
// returns 0 when every scripted step behaves as documented, otherwise the number of the failing check.
extern "C" int emu_dsl_selftest()
{
    dsl::FcfsResource<3> r; r.init(1);
    bool granted = false, resumed = false; uint8_t w = 0;
    if (r.request(1, granted) != H_OK || !granted || r.count != 0) return 1;          // count 1 -> granted, count 0
    if (r.request(2, granted) != H_OK || granted || r.wait_len != 1) return 2;         // busy -> queued
    if (r.request(3, granted) != H_OK || granted || r.wait_len != 2) return 3;
    if (r.release(1, resumed, w) != H_OK || !resumed || w != 2 || r.count != 0 || r.wait_len != 1) return 4;   // FCFS: first waiter, count unchanged
    if (r.release(1, resumed, w) != H_OK || !resumed || w != 3 || r.wait_len != 0) return 5;
    if (r.release(1, resumed, w) != H_OK || resumed || r.count != 1) return 6;         // nobody waits: count' = 1
    if (r.release(1, resumed, w) != H_RETRY || r.count != 1) return 7;                 // over-release: SimulationRetry, state untouched
    r.init(0);
    for (uint8_t k = 1; k <= 3; ++k) if (r.request(k, granted) != H_OK || granted) return 8;
    if (r.request(4, granted) != H_RETRY || r.wait_len != 3) return 9;                 // wait list full: only a stale optimistic state
    MachRepModel::Event out[4]; int n = 0; uint64_t ctr = 41;
    dsl::Outbox<MachRepModel::Event> ob(out, n, ctr, 7, 2.5);
    ob.enqueueMessage(9, 3.5, 5, 1.25); ob.sendMessage(3, 6, 0.0);
    if (n != 2 || ctr != 43 || out[0].recv_ts != 3.5 || out[0].send_ts != 2.5 || out[0].ctr != 41 || out[0].src != 7 || out[0].dst != 9 ||
        out[0].type != 5 || out[0].value != 1.25) return 10;
    if (out[1].recv_ts != 2.5 || out[1].send_ts != 2.5 || out[1].ctr != 42 || out[1].dst != 3) return 11;    // sendMessage: receive time = now
    StdGen g1 = mk_std_gen(456), g2 = mk_std_gen(456);
    if (dsl::random_exponential(g1, 0.5) != expo_variate(std_random_double(g2), 0.5) || g1.s1 != g2.s1 || g1.s2 != g2.s2) return 12;
    return 0;
}

// reverse computation: PholdModel::reverse must be the exact inverse of PholdModel::advance (returns mismatches)
extern "C" int emu_phold_reverse_roundtrip(uint64_t seed, int n)
{
    int bad = 0; uint64_t s = seed;
    PholdModel::Params p; memset(&p, 0, sizeof p);
    for (int i = 0; i < n; ++i) {
        PholdModel::State a; a.rng_ctr = rng_next(s); a.send_ctr = rng_next(s); a.n_exec = rng_next(s); a.hash = rng_next(s);
        PholdModel::Event e; e.recv_ts = bits2d(rng_next(s) >> 2); e.send_ts = 0.0; e.ctr = rng_next(s); e.src = (uint32_t)rng_next(s); e.dst = 0;
        PholdModel::State b = a;
        PholdModel::advance(p, 7, b, e);
        PholdModel::reverse(p, 7, b, e);
        if (b.rng_ctr != a.rng_ctr || b.send_ctr != a.send_ctr || b.n_exec != a.n_exec || b.hash != a.hash) ++bad;
        if (mix_fin_inv(mix_fin(a.hash)) != a.hash) ++bad;
    }
    return bad;
}

// the same run with GVT refreshed only every `gvt_every` super-steps (an external Time Server's bound)
extern "C" int emu_run_lagged(const tw_params *p, uint64_t shuffle_seed, uint32_t gvt_every, uint64_t *n_exec, uint64_t *hash,
                              tw_machrep_result *mr, tw_run_stats *st)
{
    if (gvt_every == 0) gvt_every = 1;
    if (p->model == TW_MODEL_PHOLD) return emu_run_t<PholdModel>(p, shuffle_seed, n_exec, hash, mr, st, gvt_every);
    if (p->model == TW_MODEL_MACHREP2) return emu_run_t<MachRepModel>(p, shuffle_seed, n_exec, hash, mr, st, gvt_every);
    if (p->model == TW_MODEL_MACHREP1) return emu_run_t<MachRep1Model>(p, shuffle_seed, n_exec, hash, mr, st, gvt_every);
    return TW_ERR_INVALID_ARG;
}

// warp-level build: LP tiles per warp as in Engine::init (the work queue is carried from tile to tile)
extern "C" int emu_run_tiles(const tw_params *p, uint64_t shuffle_seed, uint32_t gvt_every, uint32_t tiles, uint64_t *n_exec,
                             uint64_t *hash, tw_machrep_result *mr, tw_run_stats *st)
{
    if (gvt_every == 0) gvt_every = 1;
    if (p->model == TW_MODEL_PHOLD) return emu_run_t<PholdModel>(p, shuffle_seed, n_exec, hash, mr, st, gvt_every, tiles);
    if (p->model == TW_MODEL_MACHREP2) return emu_run_t<MachRepModel>(p, shuffle_seed, n_exec, hash, mr, st, gvt_every, tiles);
    if (p->model == TW_MODEL_MACHREP1) return emu_run_t<MachRep1Model>(p, shuffle_seed, n_exec, hash, mr, st, gvt_every, tiles);
    return TW_ERR_INVALID_ARG;
}
extern "C" int emu_is_simt(void)
{
#ifdef TW_SIMT_EMU
    return 1;
#else
    return 0;
#endif
}

extern "C" int emu_run(const tw_params *p, uint64_t shuffle_seed, uint64_t *n_exec, uint64_t *hash,
                       tw_machrep_result *mr, tw_run_stats *st)
{
    if (p->model == TW_MODEL_PHOLD) return emu_run_t<PholdModel>(p, shuffle_seed, n_exec, hash, mr, st);
    if (p->model == TW_MODEL_MACHREP2) return emu_run_t<MachRepModel>(p, shuffle_seed, n_exec, hash, mr, st);
    if (p->model == TW_MODEL_MACHREP1) return emu_run_t<MachRep1Model>(p, shuffle_seed, n_exec, hash, mr, st);
    return TW_ERR_INVALID_ARG;
}
