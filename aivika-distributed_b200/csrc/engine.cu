// engine.cu -- sm_100a kernels, host-side engine and the C ABI of libtwb200.so.
//
// Kernels (all hand-written, one thread per logical process, HBM-bound integer/f64 work):
//   k_step<M>        one super-step of every LP of this partition (tw_core.cuh::lp_step), followed by the
//                    in-kernel GVT reduction: warp shuffle min -> one atomicMin per block -> the last block
//                    publishes GVT (Samadi's round trip through the Time Server,
//                    Simulation/Aivika/Distributed/Optimistic/Internal/TimeServer.hs:369-421, collapses into this)
//   k_run_groups<M>  models whose LPs form closed groups (MachRep1/2 replicas: kGroup = 3): a warp owns a few groups and
//                    runs their super-steps back to back inside ONE launch, every group with its own GVT -- the same
//                    lp_step, no grid-wide step per causal hop
//   k_deliver<M>     multi-GPU: routes the staging rings the other GPUs filled during the previous super-step into the
//                    inboxes of this partition's LPs (the inbox relay of Internal/DIO.hs:355-370)
//   k_route<M>       host-injected messages (enqueueMessage from runEventInStartTime / QueueMessageBulk) -> inboxes
//   k_init<M>        newEventQueue for every LP (Internal/Event.hs:92-116)
//   k_collect<M>     committed-event count and trace-hash sum
// Multi-GPU: LP ranges are sharded (LP i lives on rank i / n_per).  A message for another GPU is a peer store over
// NVLink into a staging ring in the receiver's memory; the slot comes from a sender-local counter (warp-aggregated, no
// atomic crosses NVLink) and the count travels with the GVT mailbox entry.  GVT is a peer-memory mailbox exchange run by
// the last block of k_step (gvt_exchange_warp); NCCL is used for set-up and the collective host-side barriers only.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "tw_core.cuh"

namespace twb {

// Launch shape of k_step, chosen by the sweep in profiles/r1_tuning.md
#ifndef TW_BLOCK
#define TW_BLOCK 128
#endif
#ifndef TW_MINBLOCKS
#define TW_MINBLOCKS 4
#endif
constexpr int kBlock = TW_BLOCK;
static std::atomic<uint64_t> g_launches{0};

// ------------------------------------------------------------------ device helpers

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

enum { C_EXEC = 0, C_RB, C_ROLLBACKS, C_STRAG, C_ANTI, C_ANNIH, C_RETRY, C_SLOCAL, C_SREMOTE, C_FOSSIL, C_LONG, C_COUNT };

// Cross-GPU GVT, run by one warp: every rank publishes its local minimum for step `tag` -- and the number of messages it
// staged for that peer during the step -- into the mailbox of every peer (NVLink peer stores) and waits until all `world`
// entries for that step have arrived, then takes the minimum: the LocalTimeMessage / GlobalTimeMessage round of the Time
// Server (Internal/TimeServer.hs:253-270, 384-421) without a server and without NCCL on the per-step path (an 8-byte
// ncclAllReduce costs ~20-40 us, this ~5 us).  It is also the barrier that orders the peer stores of super-step s
// before the loads of super-step s+1.  The wait is bounded by wall-clock time (globaltimer; tw_params / TWB200_PEER_TIMEOUT_MS)
// so that a lost peer becomes an error code, never a hang.  One rank per GPU is required (and enforced by the host).
__device__ __forceinline__ void gvt_exchange_warp(Ctl *ctl, GvtSlot *const *peer_slots, uint32_t *stage_cnt, uint32_t rank,
                                                  uint32_t world, uint32_t tag, double local_min, unsigned long long timeout_ns)
{
    const uint32_t r = threadIdx.x & 31u;
    const uint32_t par = tag & 1u;
    const bool aborted = *(volatile uint32_t *)&ctl->error != 0;
    if (r < world) {
        GvtSlot *dst = peer_slots[r] + par * kMaxWorld + rank;           // my entry in rank r's mailbox
        if (aborted) {
            *(volatile uint32_t *)&(peer_slots[r] + 0 * kMaxWorld + rank)->tag = kSlotAbort;
            *(volatile uint32_t *)&(peer_slots[r] + 1 * kMaxWorld + rank)->tag = kSlotAbort;
        } else {
            *(volatile double *)&dst->value = local_min;
            *(volatile uint32_t *)&dst->count = *(volatile uint32_t *)&stage_cnt[r * kCntStride];   // staged for rank r
            __threadfence_system();
            *(volatile uint32_t *)&dst->tag = tag;
        }
        *(volatile uint32_t *)&stage_cnt[r * kCntStride] = 0u;           // the next step starts a fresh ring
    }
    if (aborted) return;
    double v = pos_inf(); uint32_t bad = 0;
    if (r < world) {
        const GvtSlot *src = peer_slots[rank] + par * kMaxWorld + r;      // rank r's entry in my mailbox
        uint32_t t = *(volatile const uint32_t *)&src->tag, spins = 0;
        unsigned long long t0 = 0;
        while (t != tag) {
            if (t == kSlotAbort) { bad = TW_ERR_CUDA; break; }
            if ((++spins & 1023u) == 0u) {                                 // wall clock, not iterations: clocks and sleeps vary
                const unsigned long long now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > timeout_ns) { bad = TW_ERR_DEADLOCK; break; }
            }
            __nanosleep(64);
            t = *(volatile const uint32_t *)&src->tag;
        }
        __threadfence_system();
        v = *(volatile const double *)&src->value;
    }
    bad = __reduce_max_sync(0xffffffffu, bad);
    const unsigned long long m = warp_min_u64(ordered_bits(v));
    if (r == 0) {
        if (bad) { if (atomicCAS(&ctl->error, 0u, bad) == 0u) ctl->err_lp = 0xFFFFFFFFu; }
        else {
            double g = from_ordered(m);
            const double cap = *(volatile double *)&ctl->gvt_cap, prev = *(volatile double *)&ctl->gvt;
            g = g < cap ? g : cap;
            ctl->gvt = g < prev ? prev : g;          // the global time never goes back (see block_finish)
        }
    }
}

// Block-level reduction of the per-thread accumulators, then the "last block" publishes GVT (and, in a multi-GPU run
// with fused exchange, does the cross-GPU round itself: no separate kernel between two super-steps).
template <class M>
__device__ __forceinline__ void block_finish(const Dev<M> &d, const Acc &acc, uint32_t step)
{
    Ctl *ctl = d.ctl;
    __shared__ unsigned long long s_pend[kBlock / 32], s_emit[kBlock / 32];
    __shared__ unsigned int s_cnt[C_COUNT];
    __shared__ unsigned int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < C_COUNT) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long p = warp_min_u64(ordered_bits(acc.pend_min));
    unsigned long long e = warp_min_u64(ordered_bits(acc.emit_min));
    const uint32_t vals[C_COUNT] = {acc.exec, acc.rb_events, acc.rollbacks, acc.strag, acc.anti, acc.annih,
                                    acc.retry, acc.sent_local, acc.sent_remote, acc.fossil, acc.long_lanes};
#pragma unroll
    for (int i = 0; i < C_COUNT; ++i) {
        const unsigned int s = __reduce_add_sync(0xffffffffu, vals[i]);
        if (lane == 0 && s) atomicAdd(&s_cnt[i], s);
    }
    if (acc.err) { if (atomicCAS(&ctl->error, 0u, acc.err) == 0u) ctl->err_lp = acc.err_lp; }
    if (acc.max_occ) atomicMax(&ctl->max_occ, acc.max_occ);
    if (lane == 0) { s_pend[warp] = p; s_emit[warp] = e; }
    __syncthreads();
    if (threadIdx.x < C_COUNT) {
        const unsigned int s = s_cnt[threadIdx.x];
        if (s) atomicAdd(&ctl->c_exec + threadIdx.x, (unsigned long long)s);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kBlock / 32; ++w) { p = s_pend[w] < p ? s_pend[w] : p; e = s_emit[w] < e ? s_emit[w] : e; }
        if (p != ordered_bits(pos_inf())) atomicMin(&ctl->pend_min, p);
        if (e != ordered_bits(pos_inf())) atomicMin(&ctl->emit_min, e);
        __threadfence();
        const unsigned int ticket = atomicAdd(&ctl->done_blocks, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
        double g = 0.0;
        if (lane == 0) {
            __threadfence();
            const unsigned long long pm = *(volatile unsigned long long *)&ctl->pend_min;
            const unsigned long long em = *(volatile unsigned long long *)&ctl->emit_min;
            g = from_ordered(pm < em ? pm : em);
            // What an LP leaves pending is sometimes reported as the lower bound of a time quantum of the current window
            // (tw_core.cuh: `drop`, coop_select's `rest`).  Under an external Time Server the window's base lags behind, and
            // such a bound can lie a fraction of a quantum below what the partition reported a step earlier: the local
            // time handed out never goes back (the previous value is itself a lower bound of everything still unprocessed).
            // (One partition: its local time IS the global one.  Several: only the global minimum is monotone -- a peer may
            // send this partition a message below its last local minimum --, so the clamp sits in gvt_exchange_warp.)
            if (d.world == 1) { const double prev = *(volatile double *)&ctl->gvt_local; if (g < prev) g = prev; }
            ctl->gvt_local = g;
            if (d.world == 1) { const double cap = *(volatile double *)&ctl->gvt_cap; ctl->gvt = g < cap ? g : cap; }   // otherwise the cross-GPU minimum follows
            ctl->pend_min = ordered_bits(pos_inf());
            ctl->emit_min = ordered_bits(pos_inf());
            ctl->step = step + 1; ctl->c_steps += 1;
            ctl->done_blocks = 0; ctl->tile_next = 0; ctl->hot_cnt[step & 1u] = 0;    // (this step's list is used up)
            __threadfence();
        }
        __syncwarp();
        if (d.world > 1 && d.fused_gvt) {
            g = __shfl_sync(0xffffffffu, g, 0);
            gvt_exchange_warp(ctl, d.peer_slots, d.stage_cnt, d.rank, d.world, step + 1u, g, d.peer_timeout_ns);
        }
    }
}

template <class M, bool kBulk>
__global__ void __launch_bounds__(kBlock, TW_MINBLOCKS) k_step(const __grid_constant__ Dev<M> d)
{
    Ctl *ctl = d.ctl;
    const uint4 head = __ldcg(reinterpret_cast<const uint4 *>(ctl));          // gvt | step | error in one load (L2: written by the last launch)
    const double gvt = __hiloint2double((int)head.y, (int)head.x);
    if (!(gvt <= d.stop)) return;                 // GVT passed the stop time: the run is complete (same answer in every block)
    const uint32_t step = head.z;
    Acc acc; acc.init();
    // An error recorded by an EARLIER launch: nothing to do.  The error word can also be set by a block of THIS launch
    // while later blocks are still being scheduled; those must still take their ticket in block_finish, or no block would
    // be the last one and neither the step counter nor (multi-GPU) the abort tag would be published.
    if (head.w == 0u) {
        double wend = d.stop;
        if (d.horizon >= 0.0) { const double w = gvt + d.horizon; wend = w < wend ? w : wend; }
        extern __shared__ uint4 tw_queue_smem[];             // per warp: work queue + held messages (tw_core.cuh, WorkQ)
        WorkQ<M> q;
        q.init(reinterpret_cast<char *>(tw_queue_smem) + (threadIdx.x >> 5) * WorkQ<M>::kWarpBytes);
        // Persistent warps: the grid is sized to the resident warps of the GPU, not to the partition.  A warp starts with
        // the d.tiles consecutive tiles of 32 LPs its position in the grid gives it and then takes further ones from the
        // partition's tile counter until none is left -- no warp slot idles at a block's end while the slowest warp of that
        // block finishes.  The next assignment is requested before the current tiles are run, so its round trip is hidden.
        // The work queue is carried from tile to tile.
        const uint32_t n_tiles = (d.n_local + 31u) >> 5;
        // (the build with the bulk selection also runs the tiles that held a long pending region in the previous step FIRST:
        // a hot LP's tile takes ten to twenty times longer than an ordinary one, and met late in the tile order it alone
        // would be the step's tail.  Such tiles register themselves for the next step in hot_list; the ordinary pass skips them.)
        const bool hot_first = kBulk && d.tiles == 1u && !(d.flags & kFlagNoHotFirst);
        const uint32_t hc = hot_first ? __ldcg(&ctl->hot_cnt[step & 1u]) : 0u;
        const uint32_t n_virtual = n_tiles + hc;
        const uint32_t first_dyn = gridDim.x * (kBlock / 32) * d.tiles;        // tiles beyond the statically assigned ones
        uint32_t tile0 = (blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5)) * d.tiles;
        while (tile0 < n_virtual) {
            uint32_t next = 0xFFFFFFFFu;
            if (first_dyn < n_virtual && (threadIdx.x & 31u) == 0u) next = first_dyn + atomicAdd(&ctl->tile_next, d.tiles);
            uint32_t tile = tile0; bool skip = false;
            if constexpr (kBulk) {
                if (hc) {
                    if (tile0 < hc) tile = __ldcg(&d.hot_list[(size_t)(step & 1u) * n_tiles + tile0]);
                    else { tile = tile0 - hc; skip = __ldcg(&d.hot_mark[(size_t)(step & 1u) * n_tiles + tile]) == step; }
                }
            }
            if (!skip) {
                const uint32_t long0 = acc.long_lanes;
                for (uint32_t t = 0; t < d.tiles; ++t)
                    lp_step<M, kBulk>(d, acc, q, (tile + t) * 32u + (threadIdx.x & 31u), gvt, wend, step);   // all lanes enter: phases reconverge inside
                if constexpr (kBulk) {
                    const bool lng = __any_sync(0xffffffffu, acc.long_lanes != long0);
                    if (lng && hot_first && (threadIdx.x & 31u) == 0u) {
                        const uint32_t idx = atomicAdd(&ctl->hot_cnt[(step + 1u) & 1u], 1u);
                        d.hot_list[(size_t)((step + 1u) & 1u) * n_tiles + idx] = tile;
                        d.hot_mark[(size_t)((step + 1u) & 1u) * n_tiles + tile] = step + 1u;    // (per parity: this step's marks are still being read)
                    }
                }
            }
            tile0 = __shfl_sync(0xffffffffu, next, 0);
        }
        q.drain(d, acc, step & 1u);
        // peer stores must be performed before the cross-GPU reduction publishes this step's tag
        if (d.world > 1) {
            if (d.flags & kFlagFenceGpu) __threadfence();
            else if (!(d.flags & kFlagFenceNone)) __threadfence_system();
        }
    }
    block_finish<M>(d, acc, step);
}

// Multi-GPU, between two super-steps: route what the other GPUs staged for this partition during the previous step.
// The counts were published with the GVT mailbox entries of that step (tag == current step number).
template <class M>
__global__ void __launch_bounds__(kBlock) k_deliver(const __grid_constant__ Dev<M> d)
{
    Ctl *ctl = d.ctl;
    const uint4 head = __ldcg(reinterpret_cast<const uint4 *>(ctl));          // gvt | step | error
    if (head.w) return;
    const uint32_t step = head.z;
    const uint32_t par_prev = (step & 1u) ^ 1u;
    // the counts: ONE thread per sender and block reads the mailbox (every warp polling the same few words made this
    // kernel cost 45 us on eight GPUs -- one L2 slice served 150K identical requests), the block shares them
    const GvtSlot *mine = d.peer_slots[d.rank] + (step & 1u) * kMaxWorld;
    __shared__ uint32_t s_cnt[kMaxWorld];
    if (threadIdx.x < (uint32_t)kMaxWorld) {
        const uint32_t r = threadIdx.x;
        uint32_t c = 0;
        if (r < d.world && r != d.rank && *(volatile const uint32_t *)&mine[r].tag == step) c = *(volatile const uint32_t *)&mine[r].count;
        if (c > d.stage_cap) c = d.stage_cap;         // the sender has flagged the overflow; never read past the ring
        s_cnt[r] = c;
    }
    __syncthreads();
    uint32_t cnt[kMaxWorld]; uint32_t total = 0;
#pragma unroll
    for (uint32_t r = 0; r < (uint32_t)kMaxWorld; ++r) { cnt[r] = s_cnt[r]; total += cnt[r]; }
    Acc acc; acc.init();
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < total; i += gridDim.x * kBlock) {
        uint32_t j = i, src = 0;
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)kMaxWorld; ++r) if (src == r && j >= cnt[r]) { j -= cnt[r]; src = r + 1; }
        deliver_staged<M>(d, acc, par_prev, src, j);
    }
    if (acc.err) { if (atomicCAS(&ctl->error, 0u, acc.err) == 0u) ctl->err_lp = acc.err_lp; }
}

// Closed groups (M::kGroup LPs that only ever talk to each other: the replicas of MachRep1/2): the causal chain of such
// a model is one message hop per super-step (delta = 1e-6 look-ahead), which as grid-wide launches costs a launch and a
// full GPU's latency chain per hop.  Here a warp owns `gpw` whole groups (one lane = one LP) and iterates their
// super-steps inside the kernel -- same lp_step, same rollback and annihilation machinery, messages through the same
// inbox rings --, synchronised by __syncwarp only.  Every group advances under its OWN GVT (the minimum over its LPs:
// nobody else can ever send to it), bounded by the external Time Server's value when one is attached.  A launch runs an
// even number of rounds per warp, so the arrival parity of every inbox stays aligned with the engine's step counter.
template <class M>
__global__ void __launch_bounds__(32) k_run_groups(const __grid_constant__ Dev<M> d, uint32_t gpw, uint32_t max_rounds)
{
    Ctl *ctl = d.ctl;
    if (*(volatile uint32_t *)&ctl->error) return;
    constexpr uint32_t G = M::kGroup;
    const uint32_t lane = threadIdx.x;
    const double cap = *(volatile double *)&ctl->gvt_cap;
    const uint32_t step0 = *(volatile uint32_t *)&ctl->step;
    extern __shared__ uint4 tw_queue_smem[];
    WorkQ<M> q; q.init(tw_queue_smem);
    Acc acc; acc.init();
    double graw; uint32_t rounds;
    const bool mine = group_rounds<M>(d, acc, q, blockIdx.x * gpw, gpw, max_rounds, step0, cap, graw, rounds);
    const uint32_t base_lane = lane - lane % G;
    const uint32_t grp = blockIdx.x * gpw + lane / G;
    if (mine && lane == base_lane) d.grp_gvt[grp] = ordered_bits(graw);
    // totals: counters, the partition's local time = minimum over its groups, the round count of the slowest warp
    const uint32_t vals[C_COUNT] = {acc.exec, acc.rb_events, acc.rollbacks, acc.strag, acc.anti, acc.annih,
                                    acc.retry, acc.sent_local, acc.sent_remote, acc.fossil, acc.long_lanes};
#pragma unroll
    for (int i = 0; i < C_COUNT; ++i) {
        const unsigned int s = __reduce_add_sync(0xffffffffu, vals[i]);
        if (lane == 0 && s) atomicAdd(&ctl->c_exec + i, (unsigned long long)s);
    }
    if (acc.err) { if (atomicCAS(&ctl->error, 0u, acc.err) == 0u) ctl->err_lp = acc.err_lp; }
    const unsigned long long wm = warp_min_u64(ordered_bits(graw));
    if (lane == 0) {
        atomicMin(&ctl->pend_min, wm);
        atomicMax(&ctl->c_rounds, (unsigned long long)rounds);
        __threadfence();
        const unsigned int ticket = atomicAdd(&ctl->done_blocks, 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            const double g = from_ordered(*(volatile unsigned long long *)&ctl->pend_min);
            ctl->gvt_local = g; ctl->gvt = g < cap ? g : cap;
            ctl->pend_min = ordered_bits(pos_inf());
            ctl->c_steps += *(volatile unsigned long long *)&ctl->c_rounds; ctl->c_rounds = 0;
            ctl->done_blocks = 0;
            __threadfence();
        }
    }
}

template <class M>
__global__ void __launch_bounds__(kBlock) k_route(const Dev<M> d, const tw_event *evs, size_t n)
{
    Ctl *ctl = d.ctl;
    const uint32_t step = ctl->step;
    const double gvt = ctl->gvt;
    Acc acc; acc.init();
    uint32_t skipped = 0;
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        const tw_event w = evs[i];
        if (w.recv_ts < gvt) {
            // "Received the outdated message" (Internal/Event.hs:370-374): skipped when dioAllowSkippingOutdatedMessage
            const uint32_t owner = d.world > 1 ? w.dst / d.n_per : 0u;
            if (d.flags & kFlagSkipOutdated) { if (owner == d.rank) skipped++; }
            else acc.fail(TW_ERR_OUTDATED_MESSAGE, w.dst);
            continue;
        }
        route_event<M>(d, acc, (step & 1u) ^ 1u, w);
        if (d.grp_gvt && w.dst >= d.lp_base && w.dst - d.lp_base < d.n_local)          // closed groups keep their own GVT
            atomicMin(&d.grp_gvt[(w.dst - d.lp_base) / M::kGroup], ordered_bits(w.recv_ts));
    }
    if (skipped) atomicAdd(&ctl->c_skipped, (unsigned long long)skipped);
    // fold the injected minimum into GVT: gvt = min(gvt, min recv_ts)
    __shared__ unsigned long long s_emit[kBlock / 32];
    unsigned long long e = warp_min_u64(ordered_bits(acc.emit_min));
    if ((threadIdx.x & 31) == 0) s_emit[threadIdx.x >> 5] = e;
    if (acc.err) { if (atomicCAS(&ctl->error, 0u, acc.err) == 0u) ctl->err_lp = acc.err_lp; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kBlock / 32; ++w) e = s_emit[w] < e ? s_emit[w] : e;
        if (e != ordered_bits(pos_inf())) atomicMin(&ctl->emit_min, e);
    }
}

// runEventInStartTime on the device: every LP of this partition creates its own start-time events (M::start_event: the
// PHOLD population, the `machine` processes of the MachRep slaves) and delivers them like host-injected messages -- the
// synthetic population never crosses PCIe.  One thread per LP.
template <class M>
__global__ void __launch_bounds__(kBlock) k_populate(const Dev<M> d, double start_time)
{
    Ctl *ctl = d.ctl;
    const uint32_t step = ctl->step;
    Acc acc; acc.init();
    const uint32_t l = blockIdx.x * kBlock + threadIdx.x;
    if (l < d.n_local) {
        const uint32_t lp = d.lp_base + l;
        for (uint32_t j = 0; ; ++j) {
            tw_event w;
            if (!M::start_event(d.mp, start_time, lp, j, w)) break;
            route_event<M>(d, acc, (step & 1u) ^ 1u, w);
            if (d.grp_gvt) atomicMin(&d.grp_gvt[l / M::kGroup], ordered_bits(w.recv_ts));
        }
    }
    __shared__ unsigned long long s_emit[kBlock / 32];
    unsigned long long e = warp_min_u64(ordered_bits(acc.emit_min));
    if ((threadIdx.x & 31) == 0) s_emit[threadIdx.x >> 5] = e;
    if (acc.err) { if (atomicCAS(&ctl->error, 0u, acc.err) == 0u) ctl->err_lp = acc.err_lp; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kBlock / 32; ++w) e = s_emit[w] < e ? s_emit[w] : e;
        if (e != ordered_bits(pos_inf())) atomicMin(&ctl->emit_min, e);
    }
}

// The same exchange as a kernel of its own (TWB200_GVT_KERNEL=1: comparison with the fused form; also the mode to use
// under ncu or compute-sanitizer, which serialise kernels and would make the fused wait time out).
__global__ void k_gvt_exchange(Ctl *ctl, GvtSlot *const *peer_slots_dev, uint32_t *stage_cnt, uint32_t rank, uint32_t world,
                               unsigned long long timeout_ns)
{
    const uint32_t tag = *(volatile uint32_t *)&ctl->step;
    gvt_exchange_warp(ctl, peer_slots_dev, stage_cnt, rank, world, tag, *(volatile double *)&ctl->gvt_local, timeout_ns);
}

// GlobalTimeMessage from an external Time Server (single thread): the bound moves, GVT follows up to the local minimum
__global__ void k_set_gvt_cap(Ctl *ctl, double cap, uint32_t world)
{
    ctl->gvt_cap = cap;
    if (world == 1) ctl->gvt = ctl->gvt_local < cap ? ctl->gvt_local : cap;
    else if (cap < ctl->gvt) ctl->gvt = cap;         // multi-GPU: the next mailbox round raises it again
}

// after k_route: gvt = min(gvt, injected minimum)   (single thread)
__global__ void k_fold_inject(Ctl *ctl)
{
    const double m = from_ordered(ctl->emit_min);
    if (m < ctl->gvt) ctl->gvt = m;
    if (m < ctl->gvt_local) ctl->gvt_local = m;
    ctl->emit_min = ordered_bits(pos_inf());
}

template <class M>
__global__ void __launch_bounds__(kBlock) k_init(const Dev<M> d, double start_time, double cap0)
{
    const uint32_t l = blockIdx.x * kBlock + threadIdx.x;
    if (l == 0) {
        // the step counter is never rewound: it tags the cross-GPU GVT mailboxes, and a tag must not repeat across runs
        // (+4: a tile registered as hot for the step after the last one of the previous run must not match a step of this run)
        const uint32_t next_step = (d.ctl->step + 4u) & ~1u;
        Ctl c; memset(&c, 0, sizeof c);
        c.step = next_step;
        c.gvt = start_time; c.gvt_local = start_time;
        c.gvt_cap = cap0;
        c.pend_min = ordered_bits(pos_inf()); c.emit_min = ordered_bits(pos_inf());
        *d.ctl = c;
        for (uint32_t r = 0; r < (uint32_t)kMaxWorld; ++r) d.stage_cnt[r * kCntStride] = 0u;
    }
    if (l >= d.n_local) return;
    LPHdr h; h.head = 0; h.np_hp = 0; h.consumed = 0; h.rollbacks = 0; h.lvt = neg_inf(); h.min_pend = pos_inf();
    d.hdr[l] = h;
    LPIn in; in.word = 0;
    d.in[l] = in;
    d.state[l] = M::initial_state(d.mp, d.lp_base + l);
    // a group's own GVT starts at "nothing known yet"; k_route / the first round lower it to the earliest message
    if (d.grp_gvt && l % M::kGroup == 0) d.grp_gvt[l / M::kGroup] = ordered_bits(start_time);
}

// After a consistency error (unmatched anti-message, time order, outdated message): was a ring overflow the cause?  An
// LP whose ring overflowed during a super-step notices it only at its next drain, but by then the garbage it read may
// already have produced bogus (anti-)messages that make *other* LPs fail first.  Report the cause.
template <class M>
__global__ void __launch_bounds__(kBlock) k_diagnose_overflow(const Dev<M> d)
{
    const uint32_t l = blockIdx.x * kBlock + threadIdx.x;
    if (l >= d.n_local) return;
    // the same conservative rule as the LP's own check: everything reserved so far against the head at the start of the
    // LP's last completed step (senders of that step may have raced with its fossil collection)
    if (in_total(d.in[l].word) - hdr_head_prev(d.hdr[l]) > d.cap) { d.ctl->error = TW_ERR_QUEUE_OVERFLOW; d.ctl->err_lp = d.lp_base + l; }
}

template <class M>
__global__ void __launch_bounds__(kBlock) k_collect(const Dev<M> d, unsigned long long *out /* [2] */)
{
    unsigned long long n = 0, hsum = 0;
    for (uint32_t l = blockIdx.x * kBlock + threadIdx.x; l < d.n_local; l += gridDim.x * kBlock) {
        n += d.state[l].n_exec; hsum += d.state[l].hash;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); hsum += __shfl_xor_sync(0xffffffffu, hsum, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], n); atomicAdd(&out[1], hsum); }
}

template <class M>
__global__ void __launch_bounds__(kBlock) k_gather_stats(const Dev<M> d, unsigned long long *n_exec, unsigned long long *hash)
{
    const uint32_t l = blockIdx.x * kBlock + threadIdx.x;
    if (l >= d.n_local) return;
    n_exec[l] = d.state[l].n_exec; hash[l] = d.state[l].hash;
}

template <class M>
__global__ void k_gather_replicas(const Dev<M> d, tw_machrep_result *out, uint32_t n_rep, double t_stop)
{
    const uint32_t r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= n_rep) return;
    // masters (first LP of each group) owned by this partition, in order
    const uint32_t first_master = ((d.lp_base + M::kGroup - 1u) / M::kGroup) * M::kGroup;
    const uint32_t lp = first_master + M::kGroup * r;
    if (lp < d.lp_base || lp >= d.lp_base + d.n_local) return;
    out[r] = M::result(d.state[lp - d.lp_base], t_stop);
}

// ------------------------------------------------------------------ NCCL through dlopen (torch-bundled or system libnccl.so.2)

struct NcclId128 { char b[128]; };      // ncclUniqueId is passed by value
struct NcclApi {
    typedef NcclId128 Id128;
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId128, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string &err)
    {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define TW_SYM(field, name) *(void **)(&field) = dlsym(lib, name); if (!field) { err = std::string("missing NCCL symbol ") + name; return false; }
        TW_SYM(GetUniqueId, "ncclGetUniqueId");
        TW_SYM(CommInitRank, "ncclCommInitRank");
        TW_SYM(CommDestroy, "ncclCommDestroy");
        TW_SYM(AllReduce, "ncclAllReduce");
        TW_SYM(AllGather, "ncclAllGather");
        TW_SYM(GetErrorString, "ncclGetErrorString");
#undef TW_SYM
        return true;
    }
};
static NcclApi g_nccl;
enum { kNcclInt8 = 0, kNcclFloat64 = 8, kNcclMin = 3 };   // ncclDataType_t / ncclRedOp_t values (nccl.h)

// ------------------------------------------------------------------ host engine

struct EngineBase {
    std::string err;
    virtual ~EngineBase() {}
    virtual int reset() = 0;
    virtual int enqueue(const tw_event *evs, size_t n) = 0;
    virtual int populate() = 0;
    virtual int run(double until, tw_run_stats *st) = 0;
    virtual int gvt(double *g) = 0;
    virtual int lp_state(uint32_t lp, tw_lp_state *out) = 0;
    virtual int read_lp_stats(uint64_t *n_exec, uint64_t *hash, size_t n) = 0;
    virtual int read_machrep(tw_machrep_result *out, size_t n) = 0;
    virtual void partition(uint32_t *base, uint32_t *n) const = 0;
    virtual int local_time(double *t) = 0;
    virtual int global_time_message(double t, int *accepted) = 0;
    virtual int attach_time_server(int on) = 0;
    virtual int skipped_outdated(uint64_t *n) = 0;
    virtual int max_ring_occupancy(uint32_t *n) = 0;
    // dioSimulationMonitoringAction: the callback runs on a worker thread of its own, never on the thread that launches
    // the super-steps (a slow callback must not delay the launches: other GPUs wait for this one in their mailbox round)
    tw_monitor_fn monitor = nullptr; void *monitor_user = nullptr;
    std::thread mon_thread; std::mutex mon_mu; std::condition_variable mon_cv;
    std::deque<tw_monitor_state> mon_q; bool mon_quit = false, mon_busy = false;
    void monitor_post(const tw_monitor_state &ms)
    {
        if (!monitor) return;
        std::lock_guard<std::mutex> g(mon_mu);
        if (!mon_thread.joinable()) mon_thread = std::thread([this] { monitor_loop(); });
        mon_q.push_back(ms); mon_cv.notify_all();
    }
    void monitor_loop()
    {
        std::unique_lock<std::mutex> lk(mon_mu);
        for (;;) {
            mon_cv.wait(lk, [this] { return mon_quit || !mon_q.empty(); });
            if (mon_q.empty()) { if (mon_quit) return; continue; }
            const tw_monitor_state ms = mon_q.front(); mon_q.pop_front();
            tw_monitor_fn fn = monitor; void *user = monitor_user;
            mon_busy = true; lk.unlock();
            if (fn) fn(&ms, user);
            lk.lock(); mon_busy = false; mon_cv.notify_all();
        }
    }
    void monitor_flush()          // tw_run returns only after every posted state has been delivered
    {
        std::unique_lock<std::mutex> lk(mon_mu);
        mon_cv.wait(lk, [this] { return mon_q.empty() && !mon_busy; });
    }
    void monitor_stop()
    {
        { std::lock_guard<std::mutex> g(mon_mu); mon_quit = true; mon_cv.notify_all(); }
        if (mon_thread.joinable()) mon_thread.join();
    }
};

#define TW_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            char _b[512]; snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            err = _b; return TW_ERR_CUDA;                                                               \
        }                                                                                               \
    } while (0)
#define TW_NCCL(call)                                                                                   \
    do {                                                                                                \
        int _e = (call);                                                                                \
        if (_e != 0) {                                                                                  \
            char _b[512]; snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(_e), __FILE__, __LINE__); \
            err = _b; return TW_ERR_CUDA;                                                               \
        }                                                                                               \
    } while (0)

static const char *code_name(int c)
{
    switch (c) {
    case TW_OK: return "ok";
    case TW_ERR_INVALID_ARG: return "invalid argument";
    case TW_ERR_CUDA: return "CUDA/NCCL error or no usable sm_100 device";
    case TW_ERR_SEND_AFTER_RECV: return "The Send time cannot be greater than the Receive message time";
    case TW_ERR_OUTDATED_MESSAGE: return "Received the outdated message (receive time < GVT)";
    case TW_ERR_QUEUE_OVERFLOW: return "per-LP queue capacity exceeded (event overflow)";
    case TW_ERR_ANTI_UNMATCHED: return "Cannot annihilate another message (anti-message without twin)";
    case TW_ERR_DEADLOCK: return "Detected a deadlock when retrying the computations (GVT stalled)";
    case TW_ERR_TIME_ORDER: return "The time value is too small (pending event older than the LP clock)";
    case TW_ERR_BAD_DESTINATION: return "receiver id outside [0, n_lps)";
    case TW_ERR_MODEL: return "a model declared as closed groups sent a message outside the sender's group";
    case TW_ERR_OUTBOX_OVERFLOW: return "cross-GPU staging ring full (raise tw_params.outbox_capacity)";
    case TW_ERR_NOT_SUPPORTED: return "not supported";
    case TW_ERR_ABORTED: return "Aborted the simulation (AbortSimulationMessage)";
    default: return "unknown error";
    }
}

template <class M>
struct Engine : EngineBase {
    tw_params P;
    Dev<M> d;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_chk[2] = {nullptr, nullptr};
    Ctl *h_ctl = nullptr;                 // pinned, 2 slots
    tw_event *d_stage = nullptr; size_t stage_cap = 0;
    unsigned long long *d_tmp = nullptr;  // [2]
    unsigned long long *d_stats = nullptr; // per-LP result staging (tw_read_lp_stats), allocated on first use
    cudaGraphExec_t graph = nullptr, graph_bulk = nullptr;      // super-steps with / without the bulk selection compiled in
    uint32_t graph_steps = 0; double graph_stop = 0.0;
    int bulk_mode = -1;                                           // TWB200_BULK=0/1 forces a build; -1: follow the workload
    void *comm = nullptr;                 // ncclComm_t
    std::vector<void *> peer_ptrs;        // opened IPC mappings
    uint64_t launches = 0;
    int grid = 0, step_grid = 0;
    bool external_time = false;           // GVT bounded by an external Time Server (tw_attach_time_server)
    void *meta = nullptr; size_t meta_bytes = 0, l2_window_bytes = 0, l2_persist_bytes = 0;
    GvtSlot *d_slots = nullptr; GvtSlot **d_peer_slots = nullptr;      // my GVT mailbox; device copy of the peer table
    bool group_kernel = false; uint32_t gpw = 1, group_grid = 0;      // closed-group models: k_run_groups instead of k_step
    int deliver_grid = 0;

    int init(const tw_params &p)
    {
        P = p;
        if (p.n_lps == 0 || p.world == 0 || p.world > (uint32_t)kMaxWorld || p.rank >= p.world) { err = "bad n_lps/rank/world"; return TW_ERR_INVALID_ARG; }
        uint32_t cap = p.queue_capacity ? p.queue_capacity : 128;
        if (cap & (cap - 1)) { err = "queue_capacity must be a power of two"; return TW_ERR_INVALID_ARG; }
        if (cap < 8) { err = "queue_capacity must be >= 8"; return TW_ERR_INVALID_ARG; }
        // nproc, the head delta and the per-parity arrival counts are 16-bit fields (LPHdr, LPIn)
        if (cap > kMaxCap) { err = "queue_capacity must be <= 32768"; return TW_ERR_INVALID_ARG; }
        uint32_t mb = p.max_events_per_step ? p.max_events_per_step : 8;
        if (mb > (uint32_t)kMaxBatch) mb = kMaxBatch;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { err = "no CUDA device: the engine has no CPU fallback"; return TW_ERR_CUDA; }
        if (p.device >= 0) TW_CUDA(cudaSetDevice(p.device));
        int dev = 0; TW_CUDA(cudaGetDevice(&dev));
        cudaDeviceProp prop; TW_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10) { char b[128]; snprintf(b, sizeof b, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor); err = b; return TW_ERR_CUDA; }
        memset(&d, 0, sizeof d);
        d.n_global = p.n_lps; d.world = p.world; d.rank = p.rank;
        d.n_per = (p.n_lps + p.world - 1) / p.world;
        if (M::kGroup > 1 && p.world > 1) d.n_per = ((d.n_per + M::kGroup - 1) / M::kGroup) * M::kGroup;   // keep replicas whole
        d.lp_base = d.n_per * p.rank;
        d.n_local = d.lp_base >= p.n_lps ? 0 : (p.n_lps - d.lp_base < d.n_per ? p.n_lps - d.lp_base : d.n_per);
        d.cap = cap; d.mask = cap - 1; d.max_batch = mb;
        d.proc_limit = p.proc_limit ? p.proc_limit : cap / 4;
        d.fossil_thresh = p.fossil_threshold ? p.fossil_threshold : cap / 8;
        // fossil collection must set in before the throttle does, or a throttled LP only ever advances one event per GVT
        if (d.fossil_thresh > d.proc_limit / 2) d.fossil_thresh = d.proc_limit / 2 ? d.proc_limit / 2 : 1;
        d.stop = p.stop_time;
        d.horizon = (p.time_horizon >= 0.0) ? p.time_horizon : -1.0;     // NaN compares false -> unbounded
        d.mp = M::make_params(p);
        d.flags = p.flags & 0xFFFFu;
        if (const char *e = getenv("TWB200_HOT_FIRST")) { if (!atoi(e)) d.flags |= kFlagNoHotFirst; }
        if (const char *e = getenv("TWB200_FENCE")) { if (!strcmp(e, "gpu")) d.flags |= kFlagFenceGpu; else if (!strcmp(e, "none")) d.flags |= kFlagFenceNone; }
        {
            double ms = p.sync_timeout_ms ? (double)p.sync_timeout_ms : 20000.0;   // a peer that stays away this long is gone
            if (const char *e = getenv("TWB200_PEER_TIMEOUT_MS")) { const double v = atof(e); if (v > 0.0) ms = v; }
            d.peer_timeout_ns = (unsigned long long)(ms * 1e6);
        }
        const size_t nl = d.n_local ? d.n_local : 1;
        TW_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        TW_CUDA(cudaEventCreate(&ev0)); TW_CUDA(cudaEventCreate(&ev1));
        TW_CUDA(cudaEventCreateWithFlags(&ev_chk[0], cudaEventDisableTiming));
        TW_CUDA(cudaEventCreateWithFlags(&ev_chk[1], cudaEventDisableTiming));
        TW_CUDA(cudaMalloc(&d.pages, nl * cap * sizeof(typename M::Event)));
        // saved pre-states (the UndoableLog) only for models that are not undone by reverse computation
        TW_CUDA(cudaMalloc(&d.saved, (M::kReversible ? (size_t)1 : nl * cap) * sizeof(typename M::State)));
        // per-LP metadata in ONE allocation, hottest first: inbox counters | headers | model state.  An L2 persisting
        // window is laid over it below (the slot-reservation atomics and header reads then hit L2).
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t off_hdr = up(nl * sizeof(LPIn)), off_state = off_hdr + up(nl * sizeof(LPHdr));
        meta_bytes = off_state + up(nl * sizeof(typename M::State));
        TW_CUDA(cudaMalloc(&meta, meta_bytes));
        d.in = (LPIn *)meta; d.hdr = (LPHdr *)((char *)meta + off_hdr); d.state = (typename M::State *)((char *)meta + off_state);
        TW_CUDA(cudaMalloc(&d.ctl, sizeof(Ctl)));
        TW_CUDA(cudaMemset(d.ctl, 0, sizeof(Ctl)));
        {
            const size_t nt = (nl + 31) / 32;
            TW_CUDA(cudaMalloc(&d.hot_list, 2 * nt * sizeof(uint32_t)));
            TW_CUDA(cudaMalloc(&d.hot_mark, 2 * nt * sizeof(uint32_t)));
            TW_CUDA(cudaMemset(d.hot_mark, 0, 2 * nt * sizeof(uint32_t)));   // step numbers start at 2: nothing is marked
        }
        TW_CUDA(cudaMalloc(&d.stage_cnt, kMaxWorld * kCntStride * sizeof(uint32_t)));
        TW_CUDA(cudaMemset(d.stage_cnt, 0, kMaxWorld * kCntStride * sizeof(uint32_t)));
        // Closed groups (replicas): a warp iterates the super-steps of its groups inside one launch (k_run_groups).  A round
        // costs the same instructions however many lanes are busy, so lanes are packed -- but not to the point where an SM
        // has fewer warps than it needs to hide their latency: about six warps per SM (measured on c2: 2 groups per warp
        // 0.26, 5 groups 0.42, 10 groups 0.37 G events/s; profiles/r2_tuning.md).
        if constexpr (M::kGroup > 1) if (!(p.flags & kFlagNoGroupKernel) && !getenv("TWB200_NO_GROUP_KERNEL")) {
            group_kernel = true;
            const uint32_t n_groups = d.n_local / M::kGroup;
            const uint32_t resident = (uint32_t)prop.multiProcessorCount * 6u;
            gpw = (n_groups + resident - 1) / resident; if (gpw < 1) gpw = 1;
            const uint32_t max_gpw = 32u / M::kGroup; if (gpw > max_gpw) gpw = max_gpw;
            if (const char *e = getenv("TWB200_GROUPS_PER_WARP")) { const long v = atol(e); if (v >= 1 && v <= (long)max_gpw) gpw = (uint32_t)v; }
            group_grid = (n_groups + gpw - 1) / gpw; if (group_grid < 1) group_grid = 1;
            TW_CUDA(cudaMalloc(&d.grp_gvt, (size_t)(n_groups ? n_groups : 1) * sizeof(unsigned long long)));
            TW_CUDA(cudaFuncSetAttribute(k_run_groups<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WorkQ<M>::kWarpBytes));
        }
        TW_CUDA(cudaMalloc(&d_tmp, 2 * sizeof(unsigned long long)));
        TW_CUDA(cudaHostAlloc(&h_ctl, 2 * sizeof(Ctl), cudaHostAllocDefault));
        grid = (int)((nl + kBlock - 1) / kBlock);
        // tiles per warp: large partitions run several tiles per warp (fewer partial queue flushes and block epilogues),
        // small ones keep one so that every SM gets work.  TWB200_TILES overrides.
        // (measured on B200: c4 1M LPs 5.18 -> 5.53 G ev/s and c5 256K LPs 2.53 -> 2.77 with two tiles, 64K LPs slower)
        // (with persistent warps -- see k_step -- the queue is carried over all tiles of a warp anyway: one tile per grab,
        // c4 6.12 against 5.99 G ev/s with two, profiles/r2_tuning.md)
        d.tiles = 1u;
        if (const char *e = getenv("TWB200_TILES")) { const long v = atol(e); if (v >= 1 && v <= 64) d.tiles = (uint32_t)v; }
        step_grid = (int)((nl + (size_t)kBlock * d.tiles - 1) / ((size_t)kBlock * d.tiles));
        {   // persistent warps: at most the resident blocks of the device (TWB200_STATIC_TILES=1: one block per tile group)
            int resident = prop.multiProcessorCount * TW_MINBLOCKS;
            if (const char *e = getenv("TWB200_GRID_BLOCKS_PER_SM")) { const long v = atol(e); if (v >= 1 && v <= 32) resident = prop.multiProcessorCount * (int)v; }
            if (!getenv("TWB200_STATIC_TILES") && step_grid > resident) step_grid = resident;
        }
        TW_CUDA(cudaFuncSetAttribute(k_step<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes()));
        TW_CUDA(cudaFuncSetAttribute(k_step<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes()));
        if (const char *e = getenv("TWB200_BULK")) bulk_mode = atoi(e) ? 1 : 0;
        // Keep the per-LP metadata resident in L2 with a persisting access-policy window (kernel nodes captured from
        // this stream inherit it).  TWB200_L2_WINDOW_MB overrides the size (0 = off); default = the inbox counters.
        {
            size_t want = nl * sizeof(LPIn);               // measured best: pin the inbox counters only (profiles/r1_tuning.md)
            if (const char *e = getenv("TWB200_L2_WINDOW_MB")) want = (size_t)atol(e) << 20;
            int max_win = 0, max_persist = 0;
            cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            if (want > meta_bytes) want = meta_bytes;
            if (want > (size_t)max_win) want = (size_t)max_win;
            if (want > 0 && max_persist > 0) {
                const size_t persist = want < (size_t)max_persist ? want : (size_t)max_persist;
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
                cudaStreamAttrValue av; memset(&av, 0, sizeof av);
                av.accessPolicyWindow.base_ptr = meta; av.accessPolicyWindow.num_bytes = want;
                av.accessPolicyWindow.hitRatio = (float)((double)persist / (double)want);
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
                cudaGetLastError();
                l2_window_bytes = want; l2_persist_bytes = persist;
            }
        }
        // multi-GPU GVT: fused into k_step's last block unless the separate-kernel mode is requested (ncu, compute-sanitizer)
        d.fused_gvt = (p.world > 1 && !getenv("TWB200_GVT_KERNEL")) ? 1u : 0u;
        if (p.world > 1) { int rc = init_peers(); if (rc) return rc; }
        return reset();
    }

    // exchange cudaIpc handles of pages / inbox counters through NCCL and map every peer (NVLink P2P)
    static size_t stage_bytes() { return (size_t)(kBlock / 32) * WorkQ<M>::kWarpBytes; }   // work queues + held messages

    // Map every peer's staging rings and GVT mailbox (cudaIpc handles exchanged through NCCL; NVLink P2P).
    int init_peers()
    {
        if (!g_nccl.load(err)) return TW_ERR_CUDA;
        NcclApi::Id128 id; memcpy(id.b, P.nccl_id, 128);
        TW_NCCL(g_nccl.CommInitRank(&comm, (int)P.world, id, (int)P.rank));
        // staging rings for what the other GPUs send here: [sender][parity][stage_cap] records.  Default capacity: two
        // records per local LP and sender (a first super-step that executes a whole window at once fits).
        d.stage_cap = P.outbox_capacity ? P.outbox_capacity : (d.n_per * 2u > 8192u ? d.n_per * 2u : 8192u);
        const size_t ring_bytes = (size_t)P.world * 2u * d.stage_cap * sizeof(typename M::Event);
        TW_CUDA(cudaMalloc(&d.stage_in, ring_bytes));
        struct Handles { cudaIpcMemHandle_t stage, slots; };
        Handles mine; Handles *d_all = nullptr;
        TW_CUDA(cudaIpcGetMemHandle(&mine.stage, d.stage_in));
        TW_CUDA(cudaMalloc(&d_slots, 2 * kMaxWorld * sizeof(GvtSlot)));
        TW_CUDA(cudaMemset(d_slots, 0xFE, 2 * kMaxWorld * sizeof(GvtSlot)));
        TW_CUDA(cudaIpcGetMemHandle(&mine.slots, d_slots));
        d.peer_slots[d.rank] = d_slots;
        TW_CUDA(cudaMalloc(&d_all, sizeof(Handles) * P.world));
        TW_CUDA(cudaMemcpyAsync(d_all + P.rank, &mine, sizeof mine, cudaMemcpyHostToDevice, stream));
        TW_NCCL(g_nccl.AllGather(d_all + P.rank, d_all, sizeof(Handles), kNcclInt8, comm, stream));
        std::vector<Handles> all(P.world);
        TW_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * P.world, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        TW_CUDA(cudaFree(d_all));
        for (uint32_t r = 0; r < P.world; ++r) {
            if (r == P.rank) continue;
            void *pst = nullptr, *ps = nullptr;
            TW_CUDA(cudaIpcOpenMemHandle(&pst, all[r].stage, cudaIpcMemLazyEnablePeerAccess));
            TW_CUDA(cudaIpcOpenMemHandle(&ps, all[r].slots, cudaIpcMemLazyEnablePeerAccess));
            peer_ptrs.push_back(pst); peer_ptrs.push_back(ps);
            d.stage_out[r] = (typename M::Event *)pst + (size_t)P.rank * 2u * d.stage_cap;    // rank r's rings for sender `rank`
            d.peer_slots[r] = (GvtSlot *)ps;
        }
        TW_CUDA(cudaMalloc(&d_peer_slots, kMaxWorld * sizeof(GvtSlot *)));
        TW_CUDA(cudaMemcpy(d_peer_slots, d.peer_slots, kMaxWorld * sizeof(GvtSlot *), cudaMemcpyHostToDevice));
        // k_deliver: one staged message per thread for the usual step (a 1M-LP c4 run stages ~250K messages per GPU and
        // step at 2 GPUs, ~110K at 8); the grid-stride loop takes whatever is beyond.  TWB200_DELIVER_BLOCKS_PER_SM overrides.
        {
            cudaDeviceProp prop; int dev = 0; cudaGetDevice(&dev); cudaGetDeviceProperties(&prop, dev);
            int per_sm = 16;
            if (const char *e = getenv("TWB200_DELIVER_BLOCKS_PER_SM")) { const long v = atol(e); if (v >= 1 && v <= 64) per_sm = (int)v; }
            deliver_grid = prop.multiProcessorCount * per_sm;
        }
        // one rank per GPU: the mailbox exchange spins on peers, two ranks sharing a GPU could starve each other
        {
            int dev = 0; cudaGetDevice(&dev);
            char mine_id[64] = {0}; cudaDeviceGetPCIBusId(mine_id, sizeof mine_id, dev);
            char *d_ids = nullptr; std::vector<char> ids(64 * P.world);
            TW_CUDA(cudaMalloc(&d_ids, 64 * P.world));
            TW_CUDA(cudaMemcpyAsync(d_ids + 64 * P.rank, mine_id, 64, cudaMemcpyHostToDevice, stream));
            TW_NCCL(g_nccl.AllGather(d_ids + 64 * P.rank, d_ids, 64, kNcclInt8, comm, stream));
            TW_CUDA(cudaMemcpyAsync(ids.data(), d_ids, 64 * P.world, cudaMemcpyDeviceToHost, stream));
            TW_CUDA(cudaStreamSynchronize(stream));
            TW_CUDA(cudaFree(d_ids));
            for (uint32_t r = 0; r < P.world; ++r)
                if (r != P.rank && !strncmp(ids.data() + 64 * r, mine_id, 64)) { err = "two ranks share one GPU: multi-rank runs need one GPU per rank"; return TW_ERR_INVALID_ARG; }
        }
        return TW_OK;
    }

    int barrier_all()
    {
        if (P.world <= 1) return TW_OK;
        // one-double all-reduce doubles as the cross-GPU barrier
        TW_NCCL(g_nccl.AllReduce(&d.ctl->gvt_local, &d.ctl->gvt, 1, kNcclFloat64, kNcclMin, comm, stream));
        return TW_OK;
    }

    int rendezvous()                     // cross-GPU barrier that leaves the control block alone
    {
        if (P.world <= 1) return TW_OK;
        double *scratch = reinterpret_cast<double *>(d_tmp);
        TW_NCCL(g_nccl.AllReduce(scratch, scratch + 1, 1, kNcclFloat64, kNcclMin, comm, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        return TW_OK;
    }

    int reset() override
    {
        // an engine attached to an external Time Server starts with GVT pinned to the start time
        k_init<M><<<grid, kBlock, 0, stream>>>(d, P.start_time, external_time ? P.start_time : (double)INFINITY); count(1);
        TW_CUDA(cudaGetLastError());
        if (d_slots) TW_CUDA(cudaMemsetAsync(d_slots, 0xFE, 2 * kMaxWorld * sizeof(GvtSlot), stream));   // tags = kSlotEmpty
        TW_CUDA(cudaStreamSynchronize(stream));
        if (P.world > 1) { int rc = barrier_all(); if (rc) return rc; TW_CUDA(cudaStreamSynchronize(stream)); }
        return TW_OK;
    }

    void count(uint64_t n) { launches += n; g_launches += n; }

    int enqueue(const tw_event *evs, size_t n) override
    {
        if (n == 0) return TW_OK;
        if (!evs) { err = "null event buffer"; return TW_ERR_INVALID_ARG; }
        if (n > stage_cap) {
            if (d_stage) TW_CUDA(cudaFree(d_stage));
            d_stage = nullptr; stage_cap = 0;
            TW_CUDA(cudaMalloc(&d_stage, n * sizeof(tw_event)));
            stage_cap = n;
        }
        TW_CUDA(cudaMemcpyAsync(d_stage, evs, n * sizeof(tw_event), cudaMemcpyHostToDevice, stream));
        size_t blocks = (n + kBlock - 1) / kBlock; if (blocks > 148 * 16) blocks = 148 * 16;
        k_route<M><<<(int)blocks, kBlock, 0, stream>>>(d, d_stage, n);
        k_fold_inject<<<1, 1, 0, stream>>>(d.ctl); count(2);
        TW_CUDA(cudaGetLastError());
        if (P.world > 1) { int rc = barrier_all(); if (rc) return rc; }
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        return check_device_error(h_ctl[0]);
    }

    int populate() override
    {
        k_populate<M><<<grid, kBlock, 0, stream>>>(d, P.start_time);
        k_fold_inject<<<1, 1, 0, stream>>>(d.ctl); count(2);
        TW_CUDA(cudaGetLastError());
        if (P.world > 1) { int rc = barrier_all(); if (rc) return rc; }
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        return check_device_error(h_ctl[0]);
    }

    int check_device_error(const Ctl &c_in)
    {
        Ctl c = c_in;
        if (c.error == TW_ERR_ANTI_UNMATCHED || c.error == TW_ERR_TIME_ORDER || c.error == TW_ERR_OUTDATED_MESSAGE) {
            k_diagnose_overflow<M><<<grid, kBlock, 0, stream>>>(d); count(1);
            Ctl *tmp = &h_ctl[0];
            if (cudaMemcpyAsync(tmp, d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
                cudaStreamSynchronize(stream) == cudaSuccess && tmp->error == TW_ERR_QUEUE_OVERFLOW) { c.error = tmp->error; c.err_lp = tmp->err_lp; }
        }
        if (c.error) {
            char b[256]; snprintf(b, sizeof b, "%s (LP %u, step %u)", code_name((int)c.error), c.err_lp, c.step);
            err = b; return (int)c.error;
        }
        return TW_OK;
    }

    int build_graph(uint32_t steps)
    {
        // the kernel parameters (d, including the stop time) are baked into the captured nodes: reuse the graph as long
        // as neither changes (a bench or a Time Server loop calls tw_run many times with the same stop time)
        if (graph && graph_steps == steps && graph_stop == d.stop) return TW_OK;
        if (graph) { cudaGraphExecDestroy(graph); graph = nullptr; }
        if (graph_bulk) { cudaGraphExecDestroy(graph_bulk); graph_bulk = nullptr; }
        for (int bulk = 0; bulk < 2; ++bulk) {
            cudaGraph_t g = nullptr;
            TW_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            for (uint32_t i = 0; i < steps; ++i) {
                if (P.world > 1) k_deliver<M><<<deliver_grid, kBlock, 0, stream>>>(d);        // what the peers staged in the last step
                if (bulk) k_step<M, true><<<step_grid, kBlock, stage_bytes(), stream>>>(d);
                else k_step<M, false><<<step_grid, kBlock, stage_bytes(), stream>>>(d);
                if (P.world > 1 && !d.fused_gvt) k_gvt_exchange<<<1, 32, 0, stream>>>(d.ctl, d_peer_slots, d.stage_cnt, d.rank, d.world, d.peer_timeout_ns);
            }
            TW_CUDA(cudaStreamEndCapture(stream, &g));
            TW_CUDA(cudaGraphInstantiate(bulk ? &graph_bulk : &graph, g, 0));
            TW_CUDA(cudaGraphDestroy(g));
        }
        graph_steps = steps; graph_stop = d.stop;
        return TW_OK;
    }
    uint64_t launches_per_graph(uint32_t B) const { return P.world > 1 ? (d.fused_gvt ? 2ull * B : 3ull * B) : B; }

    void post_monitor(const Ctl &c)
    {
        if (!monitor) return;
        tw_monitor_state ms; memset(&ms, 0, sizeof ms);
        ms.gvt = c.gvt; ms.steps = c.c_steps; ms.executed = c.c_exec; ms.rolled_back = c.c_rb_events;
        ms.rollbacks = c.c_rollbacks; ms.anti_sent = c.c_anti; ms.annihilated = c.c_annih;
        ms.sent_local = c.c_sent_local; ms.sent_remote = c.c_sent_remote; ms.n_lps_local = d.n_local; ms.rank = d.rank;
        monitor_post(ms);
    }

    // closed-group models: every launch of k_run_groups advances all groups by up to `rounds` super-steps of their own
    int run_groups(double until, const Ctl &c0, Ctl &last, bool &finished)
    {
        Dev<M> dg = d; dg.closed = 1u;
        const uint32_t rounds = (P.steps_per_launch ? P.steps_per_launch : 16) * 64u;     // even
        double last_gvt = c0.gvt; uint64_t stall = 0;
        while (!finished) {
            k_run_groups<M><<<group_grid, 32, WorkQ<M>::kWarpBytes, stream>>>(dg, gpw, rounds); count(1);
            TW_CUDA(cudaGetLastError());
            TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
            TW_CUDA(cudaStreamSynchronize(stream));
            last = h_ctl[0];
            post_monitor(last);
            int rc = check_device_error(last); if (rc) return rc;
            if (!(last.gvt <= until)) { finished = true; break; }
            if (external_time && !(last.gvt < last.gvt_cap)) break;           // waiting for the next GlobalTimeMessage
            if (last.gvt == last_gvt) { if (++stall > 64) { err = code_name(TW_ERR_DEADLOCK); return TW_ERR_DEADLOCK; } }
            else { stall = 0; last_gvt = last.gvt; }
        }
        return TW_OK;
    }

    // Diagnostic (TWB200_TIME_KERNELS=1): the super-steps launched one by one with CUDA events around every kernel, no graph,
    // no overlap between consecutive launches.  Prints the average duration of k_deliver and k_step to stderr.
    int run_timed(double until, Ctl &last, bool &finished)
    {
        cudaEvent_t e[3]; for (auto &x : e) TW_CUDA(cudaEventCreate(&x));
        double t_deliver = 0, t_step = 0; uint64_t n = 0;
        while (!finished) {
            TW_CUDA(cudaEventRecord(e[0], stream));
            if (P.world > 1) { k_deliver<M><<<deliver_grid, kBlock, 0, stream>>>(d); count(1); }
            TW_CUDA(cudaEventRecord(e[1], stream));
            k_step<M, false><<<step_grid, kBlock, stage_bytes(), stream>>>(d); count(1);
            if (P.world > 1 && !d.fused_gvt) { k_gvt_exchange<<<1, 32, 0, stream>>>(d.ctl, d_peer_slots, d.stage_cnt, d.rank, d.world, d.peer_timeout_ns); count(1); }
            TW_CUDA(cudaEventRecord(e[2], stream));
            TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
            TW_CUDA(cudaStreamSynchronize(stream));
            float a = 0, b = 0; cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[1], e[2]);
            t_deliver += a; t_step += b; ++n;
            last = h_ctl[0];
            int rc = check_device_error(last); if (rc) return rc;
            if (!(last.gvt <= until)) finished = true;
            if (n > 1000000) break;
        }
        fprintf(stderr, "[twb200 rank %u] %llu super-steps: k_deliver %.2f us, k_step%s %.2f us on average\n", d.rank, (unsigned long long)n,
                1e3 * t_deliver / (double)(n ? n : 1), d.fused_gvt ? " (+ fused GVT round)" : " + k_gvt_exchange", 1e3 * t_step / (double)(n ? n : 1));
        for (auto &x : e) cudaEventDestroy(x);
        return TW_OK;
    }

    int run(double until, tw_run_stats *st) override
    {
        if (!(until <= P.stop_time)) until = P.stop_time;
        d.stop = until;
        const uint32_t B = P.steps_per_launch ? P.steps_per_launch : 16;
        if (!group_kernel) { int rc = build_graph(B); if (rc) return rc; }
        // the ranks enter their first super-step together: the in-kernel mailbox wait of a rank that started early must
        // not run into its timeout because a peer is still instantiating its graph
        if (P.world > 1) { int rc = rendezvous(); if (rc) return rc; }
        Ctl c0;
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        c0 = h_ctl[0];
        const uint64_t launches0 = launches;
        TW_CUDA(cudaEventRecord(ev0, stream));
        // pipelined: batch k+1 is enqueued before batch k's control block is inspected
        int slot = 0; bool pending[2] = {false, false}; bool long_seen = false;
        double last_gvt = c0.gvt; uint64_t stall = 0; int result = TW_OK; bool finished = !(c0.gvt <= until);
        Ctl last = c0;
        if (group_kernel) { if constexpr (M::kGroup > 1) result = run_groups(until, c0, last, finished); }
        else if (getenv("TWB200_TIME_KERNELS")) result = run_timed(until, last, finished);
        else while (!finished) {
            // long pending regions (hot LPs of a skewed workload) in the last inspected batch: the build with the
            // cooperative bulk selection; otherwise the leaner one (either is correct for any state)
            const bool use_bulk = bulk_mode >= 0 ? bulk_mode == 1 : long_seen;
            TW_CUDA(cudaGraphLaunch(use_bulk ? graph_bulk : graph, stream)); count(launches_per_graph(B));
            TW_CUDA(cudaMemcpyAsync(&h_ctl[slot], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
            TW_CUDA(cudaEventRecord(ev_chk[slot], stream));
            pending[slot] = true;
            const int other = slot ^ 1;
            if (pending[other]) {
                TW_CUDA(cudaEventSynchronize(ev_chk[other]));
                pending[other] = false;
                long_seen = h_ctl[other].c_long - last.c_long >= (unsigned long long)B;      // at least one such LP per step
                last = h_ctl[other];
                post_monitor(last);
                if ((result = check_device_error(last)) != TW_OK) break;
                if (!(last.gvt <= until)) { finished = true; break; }
                if (external_time && !(last.gvt < last.gvt_cap)) break;       // waiting for the next GlobalTimeMessage
                if (last.gvt == last_gvt) { if (++stall > 4096) { err = code_name(TW_ERR_DEADLOCK); result = TW_ERR_DEADLOCK; break; } }
                else { stall = 0; last_gvt = last.gvt; }
            }
            slot = other;
        }
        TW_CUDA(cudaEventRecord(ev1, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        for (int s = 0; s < 2; ++s) if (pending[s]) { last = h_ctl[s].c_steps >= last.c_steps ? h_ctl[s] : last; }
        if (result == TW_OK) result = check_device_error(last);
        float ms = 0.f; TW_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        // An overflow of the last super-steps has no next drain that would notice it: check every ring once more before
        // the results are handed out (the kernel turns an overflow into the engine's error state).
        if (result == TW_OK) { k_diagnose_overflow<M><<<grid, kBlock, 0, stream>>>(d); count(1); }
        // final control block + committed totals
        TW_CUDA(cudaMemsetAsync(d_tmp, 0, 2 * sizeof(unsigned long long), stream));
        k_collect<M><<<grid < 1184 ? grid : 1184, kBlock, 0, stream>>>(d, d_tmp); count(1);
        unsigned long long tot[2];
        TW_CUDA(cudaMemcpyAsync(tot, d_tmp, sizeof tot, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        const Ctl &c = h_ctl[0];
        if (result == TW_OK) result = check_device_error(c);
        if (monitor) monitor_flush();
#ifdef TW_PHASE_TIMERS
        fprintf(stderr, "[twb200 phase timers, cycles of lanes with > 256 ring entries] fossil %llu drain %llu drain_slow %llu rollback %llu compact %llu select+execute %llu | selections %llu selected %llu | long lane-steps %llu\n",
                c.t_phase[0], c.t_phase[1], c.t_phase[2], c.t_phase[3], c.t_phase[4], c.t_phase[5], c.t_phase[6], c.t_phase[7], c.c_long);
#endif
        if (st) {
            memset(st, 0, sizeof *st);
            st->gvt = c.gvt; st->device_ms = ms; st->steps = c.c_steps - c0.c_steps;
            st->committed = tot[0]; st->trace_hash_sum = tot[1];
            st->executed = c.c_exec - c0.c_exec; st->rolled_back = c.c_rb_events - c0.c_rb_events;
            st->rollbacks = c.c_rollbacks - c0.c_rollbacks; st->stragglers = c.c_strag - c0.c_strag;
            st->anti_sent = c.c_anti - c0.c_anti; st->annihilated = c.c_annih - c0.c_annih; st->retries = c.c_retry - c0.c_retry;
            st->sent_local = c.c_sent_local - c0.c_sent_local; st->sent_remote = c.c_sent_remote - c0.c_sent_remote;
            st->kernel_launches = launches - launches0;
        }
        return result;
    }

    int gvt(double *g) override
    {
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        *g = h_ctl[0].gvt;
        return TW_OK;
    }

    int lp_state(uint32_t lp, tw_lp_state *out) override
    {
        if (lp < d.lp_base || lp >= d.lp_base + d.n_local) { err = "LP not owned by this partition"; return TW_ERR_INVALID_ARG; }
        const uint32_t l = lp - d.lp_base;
        LPHdr h; LPIn in; typename M::State s;
        TW_CUDA(cudaMemcpyAsync(&h, d.hdr + l, sizeof h, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaMemcpyAsync(&in, d.in + l, sizeof in, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaMemcpyAsync(&s, d.state + l, sizeof s, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        memset(out, 0, sizeof *out);
        out->lp = lp; out->local_time = h.lvt; out->event_queue_time = h.min_pend;
        const uint32_t nproc = hdr_nproc(h);
        out->event_queue_size = in_total(in.word) - h.head - nproc;
        out->log_size = nproc; out->input_message_count = in_total(in.word) - h.head; out->output_message_count = nproc;
        out->rollback_count = h.rollbacks; out->executed = s.n_exec; out->trace_hash = s.hash;
        return TW_OK;
    }

    int read_lp_stats(uint64_t *n_exec, uint64_t *hash, size_t n) override
    {
        if (n < d.n_local) { err = "output arrays shorter than the local LP count"; return TW_ERR_INVALID_ARG; }
        if (d.n_local == 0) return TW_OK;
        if (!d_stats) TW_CUDA(cudaMalloc(&d_stats, 2 * sizeof(unsigned long long) * d.n_local));   // kept for the engine's life
        unsigned long long *dn = d_stats, *dh = d_stats + d.n_local;
        k_gather_stats<M><<<grid, kBlock, 0, stream>>>(d, dn, dh); count(1);
        TW_CUDA(cudaGetLastError());
        TW_CUDA(cudaMemcpyAsync(n_exec, dn, sizeof(uint64_t) * d.n_local, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaMemcpyAsync(hash, dh, sizeof(uint64_t) * d.n_local, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        return TW_OK;
    }

    int read_machrep(tw_machrep_result *out, size_t n) override;

    void partition(uint32_t *base, uint32_t *n) const override { if (base) *base = d.lp_base; if (n) *n = d.n_local; }

    // ---- the Time Server role across partitions
    // getLocalTime / LocalTimeMessage (Internal/Event.hs:425-436, 496-514): this partition's minimum over unprocessed and
    // in-flight messages, whatever the bound says.
    int local_time(double *t) override
    {
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        *t = h_ctl[0].gvt_local;
        return TW_OK;
    }
    // GlobalTimeMessage / updateGlobalTime (Internal/Event.hs:437-449, 528-540): a global time above the local time is
    // ignored, otherwise it becomes the bound of GVT (fossil collection, commitment, termination).
    int global_time_message(double t, int *accepted) override
    {
        if (!external_time) { err = "no external Time Server attached (tw_attach_time_server)"; return TW_ERR_INVALID_ARG; }
        double lt = 0.0; int rc = local_time(&lt); if (rc) return rc;
        const bool ok = !(t > lt) && !(t < h_ctl[0].gvt_cap);          // never backwards either
        if (accepted) *accepted = ok ? 1 : 0;
        if (!ok) return TW_OK;
        k_set_gvt_cap<<<1, 1, 0, stream>>>(d.ctl, t, d.world); count(1);
        TW_CUDA(cudaGetLastError());
        TW_CUDA(cudaStreamSynchronize(stream));
        return TW_OK;
    }
    int attach_time_server(int on) override { external_time = on != 0; return reset(); }
    int max_ring_occupancy(uint32_t *n) override
    {
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        *n = h_ctl[0].max_occ;
        return TW_OK;
    }
    int skipped_outdated(uint64_t *n) override
    {
        TW_CUDA(cudaMemcpyAsync(&h_ctl[0], d.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        *n = h_ctl[0].c_skipped;
        return TW_OK;
    }

    ~Engine() override
    {
        monitor_stop();
        if (stream) cudaStreamSynchronize(stream);
        if (graph) cudaGraphExecDestroy(graph);
        if (graph_bulk) cudaGraphExecDestroy(graph_bulk);
        for (void *p : peer_ptrs) cudaIpcCloseMemHandle(p);
        if (comm) g_nccl.CommDestroy(comm);
        cudaFree(d.pages); cudaFree(d.saved); cudaFree(meta); cudaFree(d.ctl);
        cudaFree(d_stage); cudaFree(d_tmp); cudaFree(d_stats); cudaFree(d_slots); cudaFree(d_peer_slots);
        cudaFree(d.stage_cnt); cudaFree(d.stage_in); cudaFree(d.grp_gvt); cudaFree(d.hot_list); cudaFree(d.hot_mark);
        if (h_ctl) cudaFreeHost(h_ctl);
        if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1);
        if (ev_chk[0]) cudaEventDestroy(ev_chk[0]); if (ev_chk[1]) cudaEventDestroy(ev_chk[1]);
        if (stream) cudaStreamDestroy(stream);
    }
};

template <class M>
int Engine<M>::read_machrep(tw_machrep_result *out, size_t n)
{
    if constexpr (M::kGroup == 1) { (void)out; (void)n; err = "this model has no per-replica result"; return TW_ERR_NOT_SUPPORTED; }
    else {
        const uint32_t first_master = ((d.lp_base + M::kGroup - 1u) / M::kGroup) * M::kGroup;
        const uint32_t end = d.lp_base + d.n_local;
        const uint32_t n_rep = first_master < end ? (end - first_master + M::kGroup - 1u) / M::kGroup : 0u;
        if (n < n_rep) { err = "output array shorter than the local replica count"; return TW_ERR_INVALID_ARG; }
        if (!n_rep) return TW_OK;
        tw_machrep_result *dr = nullptr;
        TW_CUDA(cudaMalloc(&dr, sizeof(tw_machrep_result) * n_rep));
        k_gather_replicas<M><<<(n_rep + kBlock - 1) / kBlock, kBlock, 0, stream>>>(d, dr, n_rep, P.stop_time); count(1);
        TW_CUDA(cudaMemcpyAsync(out, dr, sizeof(tw_machrep_result) * n_rep, cudaMemcpyDeviceToHost, stream));
        TW_CUDA(cudaStreamSynchronize(stream));
        TW_CUDA(cudaFree(dr));
        return TW_OK;
    }
}

// ------------------------------------------------------------------ the handler table
// One row per compiled model: a model is a struct of tw_models.cuh (State, Event payload, handle or advance/emit, built
// from the primitives of tw_dsl.cuh); registering it here is all the engine needs.
struct ModelEntry {
    uint32_t id; const char *name; uint32_t group;
    EngineBase *(*make)(const tw_params &, int *rc);
    size_t (*initial_events)(const tw_params &, tw_event *, size_t, uint32_t, uint32_t);
};
template <class M> static EngineBase *make_engine(const tw_params &p, int *rc) { auto *e = new Engine<M>(); *rc = e->init(p); return e; }
static const ModelEntry kModels[] = {
    {TW_MODEL_PHOLD,    "PHOLD",    PholdModel::kGroup,    make_engine<PholdModel>,    PholdModel::initial_events},
    {TW_MODEL_MACHREP2, "MachRep2", MachRepModel::kGroup,  make_engine<MachRepModel>,  MachRepModel::initial_events},
    {TW_MODEL_MACHREP1, "MachRep1", MachRep1Model::kGroup, make_engine<MachRep1Model>, MachRep1Model::initial_events},
};
static const ModelEntry *find_model(uint32_t id)
{
    for (const ModelEntry &m : kModels) if (m.id == id) return &m;
    return nullptr;
}

}  // namespace twb

// ------------------------------------------------------------------ C ABI

using namespace twb;

struct tw_engine { EngineBase *impl; std::string create_err; };
static thread_local std::string g_create_err;

extern "C" {

int tw_create(const tw_params *p, tw_engine **out)
{
    if (!p || !out) return TW_ERR_INVALID_ARG;
    *out = nullptr;
    const ModelEntry *m = find_model(p->model);
    if (!m) { g_create_err = "unknown model"; return TW_ERR_INVALID_ARG; }
    if (p->n_lps % m->group) { g_create_err = std::string(m->name) + " needs n_lps = " + std::to_string(m->group) + " * replicas"; return TW_ERR_INVALID_ARG; }
    int rc = TW_OK;
    EngineBase *impl = m->make(*p, &rc);
    if (rc != TW_OK) { g_create_err = impl->err; delete impl; return rc; }
    tw_engine *h = new tw_engine(); h->impl = impl; *out = h;
    return TW_OK;
}
int tw_destroy(tw_engine *e) { if (!e) return TW_ERR_INVALID_ARG; delete e->impl; delete e; return TW_OK; }
int tw_reset(tw_engine *e) { return e ? e->impl->reset() : TW_ERR_INVALID_ARG; }
int tw_enqueue_message(tw_engine *e, const tw_event *ev) { return e ? e->impl->enqueue(ev, 1) : TW_ERR_INVALID_ARG; }
int tw_enqueue_messages(tw_engine *e, const tw_event *evs, size_t n) { return e ? e->impl->enqueue(evs, n) : TW_ERR_INVALID_ARG; }
int tw_populate(tw_engine *e) { return e ? e->impl->populate() : TW_ERR_INVALID_ARG; }
int tw_run(tw_engine *e, double until, tw_run_stats *st) { return e ? e->impl->run(until, st) : TW_ERR_INVALID_ARG; }
int tw_gvt(tw_engine *e, double *g) { return (e && g) ? e->impl->gvt(g) : TW_ERR_INVALID_ARG; }
int tw_lp_state_get(tw_engine *e, uint32_t lp, tw_lp_state *out) { return (e && out) ? e->impl->lp_state(lp, out) : TW_ERR_INVALID_ARG; }
int tw_read_lp_stats(tw_engine *e, uint64_t *n_exec, uint64_t *hash, size_t n) { return (e && n_exec && hash) ? e->impl->read_lp_stats(n_exec, hash, n) : TW_ERR_INVALID_ARG; }
int tw_read_machrep(tw_engine *e, tw_machrep_result *out, size_t n) { return (e && out) ? e->impl->read_machrep(out, n) : TW_ERR_INVALID_ARG; }
int tw_set_monitor(tw_engine *e, tw_monitor_fn fn, void *user)
{
    if (!e) return TW_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->impl->mon_mu);
    e->impl->monitor = fn; e->impl->monitor_user = user;
    return TW_OK;
}
int tw_skipped_outdated(tw_engine *e, uint64_t *n) { return (e && n) ? e->impl->skipped_outdated(n) : TW_ERR_INVALID_ARG; }
int tw_max_ring_occupancy(tw_engine *e, uint32_t *n) { return (e && n) ? e->impl->max_ring_occupancy(n) : TW_ERR_INVALID_ARG; }
int tw_attach_time_server(tw_engine *e, int on) { return e ? e->impl->attach_time_server(on) : TW_ERR_INVALID_ARG; }
int tw_local_time(tw_engine *e, double *t) { return (e && t) ? e->impl->local_time(t) : TW_ERR_INVALID_ARG; }
int tw_global_time_message(tw_engine *e, double t, int *accepted) { return e ? e->impl->global_time_message(t, accepted) : TW_ERR_INVALID_ARG; }
int tw_partition(const tw_engine *e, uint32_t *base, uint32_t *n) { if (!e) return TW_ERR_INVALID_ARG; e->impl->partition(base, n); return TW_OK; }

int tw_model_initial_events(const tw_params *p, tw_event *out, size_t cap, size_t *n)
{
    if (!p || !out || !n) return TW_ERR_INVALID_ARG;
    const ModelEntry *m = find_model(p->model);
    if (!m) return TW_ERR_INVALID_ARG;
    *n = m->initial_events(*p, out, cap, 0u, 0xFFFFFFFFu);
    return TW_OK;
}

int tw_model_initial_events_range(const tw_params *p, uint32_t lp0, uint32_t lp1, tw_event *out, size_t cap, size_t *n)
{
    if (!p || !out || !n) return TW_ERR_INVALID_ARG;
    const ModelEntry *m = find_model(p->model);
    if (!m) return TW_ERR_INVALID_ARG;
    *n = m->initial_events(*p, out, cap, lp0, lp1);
    return TW_OK;
}

int tw_nccl_unique_id(uint8_t out[128])
{
    std::string err;
    if (!g_nccl.load(err)) { g_create_err = err; return TW_ERR_CUDA; }
    return g_nccl.GetUniqueId(out) == 0 ? TW_OK : TW_ERR_CUDA;
}

const char *tw_last_error(const tw_engine *e) { return e ? e->impl->err.c_str() : g_create_err.c_str(); }
const char *tw_strerror(int code) { return code_name(code); }
const char *tw_version(void) { return "twb200 0.2 sm_100a"; }
uint64_t tw_kernel_launch_count(void) { return g_launches.load(); }

}  // extern "C"
