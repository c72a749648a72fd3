// tw_core.cuh -- the per-LP super-step of the batched Time Warp engine (device code).
//
// One lane owns one logical process (LP) for one super-step; a warp owns one or two tiles of 32 LPs.  The step is the
// device form of one pass of `processPendingEventsCore` / `call` (Simulation/Aivika/Distributed/Optimistic/Internal/
// Event.hs:196-250) over a whole batch of homogeneous LPs:
//
//   1. fossil collection, strict `< GVT`           reduceEvents            Internal/Event.hs:639-649
//   2. drain the inbox                             processChannelMessages  Internal/Event.hs:338-405
//        anti-message  -> find twin, annihilate    findAntiMessage / annihilateMessage
//                                                                          Internal/InputMessageQueue.hs:245-279
//        straggler     -> mark rollback position   enqueueMessage          Internal/InputMessageQueue.hs:146-160
//   3. rollback: undo the executions (rollbackLog, Internal/UndoableLog.hs:93-118) -- by reverse computation for a
//      reversible model, from the saved pre-state otherwise --, un-process the rolled-back inputs
//      (rollbackInputMessages :190-201) and send anti-messages (rollbackOutputMessages,
//      Internal/OutputMessageQueue.hs:81-104), regenerated from the pre-state of each undone event: no output queue
//   4. select the pending events inside the optimistic window [GVT, GVT + horizon] (isTimeHorizonExceeded,
//      Internal/Event.hs:309-324) in one pass (packed-key candidate list) and execute them in key order, at most
//      `max_batch` per step: the owner lane does the LP's state transition, the production and delivery of the
//      messages goes through the warp's work queue (WorkQ) and runs with all 32 lanes
//   5. contribute min(pending, sent) to the GVT reduction (getLocalTime, Internal/Event.hs:496-514)
//
// Per-LP storage (all in HBM, capacity `cap` = power of two slots, ring addressed by monotone indices):
//   pages[lp][slot]  event records.  [head, head+nproc) = processed, sorted by key (the reference's
//                    itemProcessed=True prefix of the InputMessageQueue vector); [head+nproc, consumed) =
//                    pending, unsorted; [consumed, total) = inbox written by senders.
//   saved[lp][slot]  state *before* the processed event in that slot (copy state saving; the UndoableLog) -- only
//                    for models that are not reversible.
//   hdr[lp]          owner-private header; in[lp] = counters other LPs touch with atomics.
// Nothing is ever moved except one swap per pop; annihilated pairs stay as tombstones that are "processed"
// as no-ops when their time comes (what the reference does for an anti-message that arrives first,
// Internal/InputMessageQueue.hs:293-301).
//
// Inbox protocol: a sender reserves a slot and counts the arrival with ONE 64-bit atomicAdd on the LP's inbox word
// (total | cnt[parity], see LPIn) and stores the record at slot idx & mask.  The owner, one step later, consumes exactly
// cnt[previous parity] entries:
// all of them were written by a kernel that has completed, so no flag or fence is needed, and an
// anti-message can never overtake its positive twin (same counter, program order).
//
// The code is written against tiny atomic shims so that tests/emu can compile the very same functions with
// g++ and run them sequentially on the CPU as a debugging aid; the product only ever runs the CUDA build.
#pragma once
#include "tw_models.cuh"

// TW_SIMT: the warp-level code paths (work queue, whole-warp search, reconvergence).  They are what the GPU runs; a host
// build can select them too (-DTW_SIMT_EMU, tests/emu/simt.hpp: 32 host threads play the lanes of a warp and the warp
// intrinsics are rendezvous points), so that the very code the device executes is tested against the oracle on the CPU.
#if defined(__CUDA_ARCH__) || defined(TW_SIMT_EMU)
#define TW_SIMT 1
#else
#define TW_SIMT 0
#endif

namespace twb {

// Inbox words are touched by the owning GPU only (messages from other GPUs arrive through the staging rings and are
// routed by the owner, see StageOut / k_deliver): device scope is the right scope for them.
#if defined(__CUDA_ARCH__)
TW_HD unsigned long long atomic_add_u64(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
TW_HD uint32_t atomic_add_u32(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
TW_HD void atomic_min_u64(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
#elif defined(TW_SIMT_EMU)
TW_HD void atomic_min_u64(unsigned long long *p, unsigned long long v)
{
    unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}
TW_HD unsigned long long atomic_add_u64(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
TW_HD uint32_t atomic_add_u32(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
#else
TW_HD unsigned long long atomic_add_u64(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
TW_HD uint32_t atomic_add_u32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
#endif

constexpr int kMaxBatch = 64;     // upper bound of events one LP executes per super-step (tw_params.max_events_per_step)
constexpr int kMaxWorld = 8;      // GPUs of one box
#ifndef TW_SCAN_BATCH
#define TW_SCAN_BATCH 8
#endif
constexpr int kScanBatch = TW_SCAN_BATCH;   // pending-record loads issued back to back in the selection scan
constexpr uint32_t kMaskBits = 32;          // pending regions up to this long are selected through the packed-key
                                            // candidate list (5-bit entry index); longer ones (hot LPs of a skewed
                                            // workload) are searched by the whole warp
#ifndef TW_CAND
#define TW_CAND 6
#endif
constexpr int kCand = TW_CAND;              // candidates kept per selection scan (5 bits each in the 32-bit displacement map)
static_assert(kCand >= 2 && kCand <= 6, "the displaced-entry map holds 6 x 5 bits");
constexpr uint32_t kQuantBits = 27;         // receive times are quantised to this many bits over [gvt, window end]
constexpr uint32_t kNoCand = 0xFFFFFFFFu;
constexpr int kQueueSlots = 64;             // per-warp work queue (shared memory), a power of two >= 2 * 32

TW_HD double dmin(double a, double b) { return a < b ? a : b; }
TW_HD double pos_inf() { return bits2d(0x7FF0000000000000ull); }
TW_HD double neg_inf() { return bits2d(0xFFF0000000000000ull); }
// order-preserving map double -> uint64 (for atomicMin)
TW_HD uint64_t ordered_bits(double x) { uint64_t b = d2bits(x); return (b >> 63) ? ~b : (b | 0x8000000000000000ull); }
TW_HD double   from_ordered(uint64_t o) { return bits2d((o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o); }

struct alignas(32) LPHdr {       // 32 B, owner-private
    uint32_t head;               // monotone index of the oldest retained entry
    uint32_t np_hp;              // bits 0..15: processed entries [head, head+nproc); bits 16..31: head - head_prev, where
                                 // head_prev = head at the start of the previous step (conservative overflow check)
    uint32_t consumed;           // entries [head, consumed) are known to the owner
    uint32_t rollbacks;          // rollback episodes of this LP (lpStateRollbackCount, Optimistic/State.hs:50)
    double   lvt;                // receive time of the last processed entry (queueTime), -inf if none
    double   min_pend;           // exact min receive time over pending entries, +inf if none
};
constexpr uint32_t kMaxCap = 32768;          // ring capacity bound: nproc and the head delta fit 16 bits, arrival counts too
TW_HD uint32_t hdr_nproc(const LPHdr &h) { return h.np_hp & 0xFFFFu; }
TW_HD uint32_t hdr_head_prev(const LPHdr &h) { return h.head - (h.np_hp >> 16); }
TW_HD uint32_t hdr_pack(uint32_t nproc, uint32_t head, uint32_t head_prev) { return nproc | ((head - head_prev) << 16); }
// Inbox word of an LP, touched by senders with ONE 64-bit atomic per message:
//   bits 32..63  total : monotone slot-reservation counter (wraps harmlessly)
//   bits  0..15  cnt0  : arrivals during even super-steps      bits 16..31  cnt1 : during odd super-steps
// A sender adds (1 << 32) | (1 << 16*parity) and takes its slot index from the old value; the owner subtracts the
// count it consumed (atomically, other fields keep changing).  Per LP and step at most 65535 arrivals (> any ring).
struct alignas(8) LPIn { unsigned long long word; };
TW_HD uint32_t in_total(unsigned long long w) { return (uint32_t)(w >> 32); }
TW_HD uint32_t in_count(unsigned long long w, uint32_t parity) { return (uint32_t)(w >> (16u * parity)) & 0xFFFFu; }
TW_HD unsigned long long in_add(uint32_t parity) { return (1ull << 32) | (1ull << (16u * parity)); }
TW_HD unsigned long long load_inbox_word(const LPIn *p)
{
#if defined(__CUDA_ARCH__)
    return __ldcg(&p->word);                 // L2 is where the atomics happen
#else
    return *(const volatile unsigned long long *)&p->word;
#endif
}

// Cross-GPU GVT mailbox entry (peer-written over NVLink): value first, system fence, then the tag (= step number).
// `count` = messages the sender staged for this receiver during the step (StageOut): the receiver routes exactly that many.
struct alignas(16) GvtSlot { double value; uint32_t tag; uint32_t count; };
constexpr uint32_t kSlotEmpty = 0xFFFFFFFEu, kSlotAbort = 0xFFFFFFFFu;

struct alignas(16) Ctl {         // device-resident control block
    double   gvt;                // GVT the next step uses (queueGlobalTime)
    uint32_t step, error;        // (gvt, step, error: the 16 bytes every warp reads at the start of a super-step -- one load)
    double   gvt_local;          // this partition's min before the cross-GPU reduction
    unsigned long long pend_min; // ordered bits, atomicMin target
    unsigned long long emit_min;
    uint32_t finished, err_lp;
    uint32_t done_blocks;
    uint32_t tile_next;          // k_step: next tile of 32 LPs to hand out (warps fetch tiles until none is left)
    unsigned long long c_exec, c_rb_events, c_rollbacks, c_strag, c_anti, c_annih, c_retry,
                       c_sent_local, c_sent_remote, c_fossil, c_long, c_steps;   // order == enum in engine.cu
    double   gvt_cap;            // GlobalTimeMessage of an external Time Server: gvt never exceeds it (+inf = none)
    unsigned long long c_skipped; // host-injected messages older than GVT that were skipped (dioAllowSkippingOutdatedMessage)
    uint32_t max_occ, pad1;       // most ring slots any LP of this partition had in use at the start of a super-step
    uint32_t hot_cnt[2];          // tiles with a long pending region, registered for the step of that parity (run first)
#ifdef TW_PHASE_TIMERS
    unsigned long long t_phase[8];
#endif
    unsigned long long c_rounds;  // closed-group kernel: rounds of the slowest warp, summed over launches
};

// per-thread accumulators, reduced per block by the kernel wrapper
struct Acc {
    double pend_min, emit_min;
    uint32_t exec, rb_events, rollbacks, strag, anti, annih, retry, sent_local, sent_remote, fossil;
    uint32_t long_lanes;             // LPs that started the step with more than 8 * kMaskBits pending entries
    uint32_t max_occ;                // most ring slots in use by an LP of this thread (retained entries + arrivals)
    uint32_t err, err_lp;
    TW_HD void init()
    {
        pend_min = pos_inf(); emit_min = pos_inf();
        exec = rb_events = rollbacks = strag = anti = annih = retry = sent_local = sent_remote = fossil = long_lanes = max_occ = 0;
        err = 0; err_lp = 0;
    }
    TW_HD void fail(uint32_t code, uint32_t lp) { if (!err) { err = code; err_lp = lp; } }
    TW_HD void merge(const Acc &o)
    {
        pend_min = dmin(pend_min, o.pend_min); emit_min = dmin(emit_min, o.emit_min);
        exec += o.exec; rb_events += o.rb_events; rollbacks += o.rollbacks; strag += o.strag; anti += o.anti; annih += o.annih;
        retry += o.retry; sent_local += o.sent_local; sent_remote += o.sent_remote; fossil += o.fossil; long_lanes += o.long_lanes;
        max_occ = o.max_occ > max_occ ? o.max_occ : max_occ;
        if (!err && o.err) { err = o.err; err_lp = o.err_lp; }
    }
};

constexpr uint32_t kCntStride = 32;          // staging counters sit 128 bytes apart (one L2 line each: no false contention)
constexpr uint32_t kFlagSkipOutdated = 1u;   // tw_params.flags bit 0: dioAllowSkippingOutdatedMessage (Internal/DIO.hs:96-98)
constexpr uint32_t kFlagNoGroupKernel = 2u;  // tw_params.flags bit 1: run closed-group models through k_step as well
constexpr uint32_t kFlagFenceGpu = 1u << 16; // engine-internal (TWB200_FENCE=gpu|none): scope of the per-thread fence after a
constexpr uint32_t kFlagFenceNone = 1u << 17; // multi-GPU super-step's peer stores (default: system)
constexpr uint32_t kFlagNoHotFirst = 1u << 18; // engine-internal (TWB200_HOT_FIRST=0): tiles in index order in every build

template <class M>
struct Dev {                     // passed to kernels by value
    typename M::Event *pages;
    typename M::State *saved;
    typename M::State *state;
    LPHdr *hdr;
    LPIn  *in;
    Ctl   *ctl;
    uint32_t n_local, lp_base, n_global, n_per;
    uint32_t cap, mask, max_batch, proc_limit, fossil_thresh;
    uint32_t rank, world;
    uint32_t fused_gvt;          // multi-GPU: the last block of k_step runs the GVT mailbox exchange itself
    uint32_t tiles;              // tiles of 32 LPs one warp runs per super-step (large partitions: fewer, fuller queue flushes)
    uint32_t flags;              // kFlag*
    uint32_t closed;             // closed-group kernel: a message must stay inside the sender's group of M::kGroup LPs
    double   stop, horizon;      // horizon < 0: unbounded
    typename M::Params mp;
    // Cross-GPU exchange (StageOut below): messages for LPs of rank r are appended to a staging ring that lives in r's
    // memory (one ring per sender and step parity), slots taken from a sender-local counter.
    typename M::Event *stage_in;                // my incoming rings: [sender rank][parity][stage_cap]
    typename M::Event *stage_out[kMaxWorld];    // peer-mapped (NVLink P2P): rank r's rings for messages from me, [parity][stage_cap]
    uint32_t          *stage_cnt;               // sender-local slot counters, [rank r * kCntStride], reset every step
    uint32_t           stage_cap;
    GvtSlot           *peer_slots[kMaxWorld];   // every rank's GVT mailbox: slots[parity][sender rank]
    unsigned long long *grp_gvt;                // closed-group kernel: ordered bits of every group's own GVT
    uint32_t *hot_list;                         // [2][tiles]: tiles to run first in the step of that parity (longest work first)
    uint32_t *hot_mark;                         // [2][tiles]: the step a tile is registered for, by that step's parity
    unsigned long long peer_timeout_ns;         // wall-clock bound of the cross-GPU mailbox wait (a lost peer becomes an error)
};

template <class E>
TW_HD bool key_less(const E &a, const E &b)
{
    if (a.recv_ts != b.recv_ts) return a.recv_ts < b.recv_ts;
    if (a.send_ts != b.send_ts) return a.send_ts < b.send_ts;
    if (a.src != b.src) return a.src < b.src;
    return (a.ctr & kCtrMask) < (b.ctr & kCtrMask);
}
template <class E>
TW_HD bool same_message(const E &a, const E &b)     // Eq Message ignores the payload, Internal/Message.hs:70-78
{
    return a.src == b.src && ((a.ctr ^ b.ctr) & kCtrMask) == 0 && a.recv_ts == b.recv_ts && a.send_ts == b.send_ts;
}

// sendMessage (Internal/OutputMessageQueue.hs:65-78) + deliverMessage (Internal/DIO.hs:605-614): the scatter.
//   receiver on this GPU      one 64-bit atomic on the receiver's inbox word reserves the ring slot and counts the arrival
//   receiver on another GPU   the record goes into the staging ring the receiver keeps for this sender (peer store over
//                             NVLink); the slot comes from a counter in the SENDER's memory -- no atomic crosses NVLink,
//                             and the warp's work queue reserves once per destination GPU and flush (warp-aggregated).
//                             The receiver routes the ring into its inboxes before its next super-step (deliver_staged);
//                             the count travels with the GVT mailbox entry (GvtSlot::count).
// emit_classify checks the destination and returns the owning rank and the local index there.
template <class M>
TW_HD bool emit_classify(const Dev<M> &d, Acc &acc, uint32_t lp, const typename M::Event &ev, uint32_t &owner, uint32_t &l)
{
    const uint32_t dst = ev.dst;
    if (dst >= d.n_global) { acc.fail(TW_ERR_BAD_DESTINATION, lp); return false; }
    if (ev.recv_ts < ev.send_ts) { acc.fail(TW_ERR_SEND_AFTER_RECV, lp); return false; }
    if (d.closed && dst / M::kGroup != lp / M::kGroup) { acc.fail(TW_ERR_MODEL, lp); return false; }   // not a closed group
    acc.emit_min = dmin(acc.emit_min, ev.recv_ts);
    if (ev.ctr & kAntiBit) acc.anti++;
    owner = d.rank; l = dst - d.lp_base;
    if (d.world > 1) { owner = dst / d.n_per; l = dst - owner * d.n_per; }
    if (owner != d.rank) acc.sent_remote++; else acc.sent_local++;
    return true;
}
// scalar form (CPU emulation without warps, k_route-like callers): reserve and store at once
template <class M>
TW_HD void emit_event(const Dev<M> &d, Acc &acc, uint32_t parity, uint32_t lp, const typename M::Event &ev)
{
    uint32_t owner, l;
    if (!emit_classify<M>(d, acc, lp, ev, owner, l)) return;
    if (owner == d.rank) {
        const uint32_t idx = in_total(atomic_add_u64(&d.in[l].word, in_add(parity)));
        d.pages[(size_t)l * d.cap + (idx & d.mask)] = ev;
    } else {
        const uint32_t idx = atomic_add_u32(&d.stage_cnt[owner * kCntStride], 1u);
        if (idx >= d.stage_cap) { acc.fail(TW_ERR_OUTBOX_OVERFLOW, lp); return; }
        d.stage_out[owner][(size_t)parity * d.stage_cap + idx] = ev;
    }
}
// The receiving side of the exchange: entry j of the ring that rank `src` filled during the previous super-step
// (parity par_prev) goes into the destination LP's inbox exactly as if a local sender had delivered it in that step.
// It is already part of GVT (the sender counted it in its emit minimum).
template <class M>
TW_HD void deliver_staged(const Dev<M> &d, Acc &acc, uint32_t par_prev, uint32_t src, uint32_t j)
{
    const typename M::Event ev = d.stage_in[((size_t)src * 2u + par_prev) * d.stage_cap + j];
    const uint32_t l = ev.dst - d.lp_base;
    if (ev.dst < d.lp_base || l >= d.n_local) { acc.fail(TW_ERR_BAD_DESTINATION, ev.dst); return; }
    const uint32_t idx = in_total(atomic_add_u64(&d.in[l].word, in_add(par_prev)));
    d.pages[(size_t)l * d.cap + (idx & d.mask)] = ev;
}

// An inconsistency seen by an LP (time order, unmatched anti-message, outdated message) can be the consequence of a
// ring overflow that happened earlier in the same super-step (a sender overwrote a live slot; the owner's own check
// runs at its next drain).  Report the cause: the overflow.
template <class M>
TW_HD void fail_checked(const Dev<M> &d, Acc &acc, uint32_t code, uint32_t l, uint32_t head_prev)
{
    const uint32_t total = in_total(*(volatile unsigned long long *)&d.in[l].word);
    acc.fail(total - head_prev > d.cap ? (uint32_t)TW_ERR_QUEUE_OVERFLOW : code, d.lp_base + l);
}

// Warp reconvergence point.  lp_step is entered by all 32 lanes of a warp with uniform control flow between
// phases; data-dependent loops inside a phase diverge, and this barrier brings the lanes back together so that the
// next phase runs with full SIMT width.  No-op in the CPU emulation build.
#if defined(__CUDA_ARCH__)
#define TW_RECONVERGE() __syncwarp()
#define TW_NOINLINE __device__ __noinline__
TW_HD bool warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
TW_HD uint32_t bitlen64(uint64_t m) { return 64u - (uint32_t)__clzll((long long)m); }
#elif defined(TW_SIMT_EMU)
#define TW_RECONVERGE() __syncwarp()
#define TW_NOINLINE
TW_HD bool warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
TW_HD uint32_t bitlen64(uint64_t m) { return 64u - (uint32_t)__clzll((long long)m); }
#elif defined(__CUDACC__)
#define TW_RECONVERGE() ((void)0)
#define TW_NOINLINE __device__ __noinline__
TW_HD bool warp_any(bool p) { return p; }
TW_HD uint32_t bitlen64(uint64_t m) { return m ? 64u - (uint32_t)__builtin_clzll(m) : 0u; }
#else
#define TW_RECONVERGE() ((void)0)
#define TW_NOINLINE
TW_HD bool warp_any(bool p) { return p; }
TW_HD uint32_t bitlen64(uint64_t m) { return m ? 64u - (uint32_t)__builtin_clzll(m) : 0u; }
#endif

// Diagnostic build (-DTW_PHASE_TIMERS): SM clock cycles a lane with a long pending region spends in each phase of its
// super-step, summed into Ctl::t_phase (printed by tw_run); compiled out otherwise.
#if defined(TW_PHASE_TIMERS) && defined(__CUDA_ARCH__)
#define TW_T0() long long tw_t_ = clock64(); const bool tw_hot_ = live && (hdr_consumed_guess > 256u)
#define TW_TICK(i) do { if (tw_hot_) { const long long n_ = clock64(); atomicAdd(&d.ctl->t_phase[i], (unsigned long long)(n_ - tw_t_)); tw_t_ = n_; } } while (0)
#else
#define TW_T0() ((void)0)
#define TW_TICK(i) ((void)0)
#endif

// ------------------------------------------------------------------ the work queue of a warp
// An executed event is split into the LP's sequential state transition (M::advance, run by the lane that owns the
// LP) and the production + delivery of its messages (M::emit + emit_event).  The second, expensive part does not
// have to run on the owner's lane: every lane that executed (or rolled back) an event pushes a Work item into the
// warp's queue in shared memory, and as soon as 32 items are queued the warp runs them with all 32 lanes -- one
// item per lane -- instead of looping max-over-lanes times with a third of the lanes active.  Anti-messages of a
// rollback (rollbackOutputMessages, Internal/OutputMessageQueue.hs:81-104) go through the same queue.
// The CPU emulation runs every item at once.
template <class M>
TW_HD void run_work(const Dev<M> &d, Acc &acc, uint32_t parity, const typename M::Work &w)
{
    typename M::Event out[M::kMaxSends]; int n = 0;
    M::emit(d.mp, w, out, n);
    const uint32_t lp = M::work_lp(w);
#pragma unroll
    for (int k = 0; k < M::kMaxSends; ++k)
        if (k < n) emit_event<M>(d, acc, parity, lp, out[k]);
}

template <class M>
struct WorkQ {
    using Work = typename M::Work;
    using Event = typename M::Event;
    static constexpr int G = (int)(sizeof(Event) / 16);      // 16-byte granules of one event record
    static_assert(sizeof(Event) % 16 == 0, "event records are held in 16-byte granules");
    // shared memory of one warp: the queue, then the held events (lane-interleaved granules), then their ring pointers
    // + 32 x 8 bytes: the per-group minimum of the messages sent in a round (closed-group kernel)
    static constexpr size_t kWarpBytes = kQueueSlots * sizeof(Work) + (size_t)M::kMaxSends * 32 * (sizeof(Event) + 8) + 32 * 8;
#if TW_SIMT
    Work *buf;                       // this warp's ring of kQueueSlots items in shared memory
    uint32_t head, tail;             // warp-uniform monotone indices
    // Deferred delivery: the messages a lane produces in one flush are held in shared memory while their slot
    // reservations are in flight; the record stores are issued at the next flush (or at drain), when the indices have
    // long arrived.  The warp therefore never waits for a reservation round trip.
    uint4 *hold;                     // this lane's granule 0 of held event 0
    unsigned long long *hring;       // this lane's ring pointer of held event 0 (bit 0 set: a staging ring of another GPU)
    uint32_t hmask;                  // bit k: held event k is valid
    uint32_t hidx[M::kMaxSends];     // local receiver: the reserved ring index; staged: the base index of the lane's group
                                     // (valid on the group's leader lane) -- results of atomics, possibly still in flight
    uint32_t hgrp[M::kMaxSends];     // staged: leader lane of the group | rank of this lane's message in the group << 5
    // Closed-group kernel: a work item runs on whichever lane is free, but the group's GVT needs the minimum over the
    // messages sent by ITS LPs: gmin[g] (shared memory) collects it per group of the warp, l0 = the warp's first local LP.
    unsigned long long *gmin; uint32_t l0;
    __device__ __forceinline__ void init(void *warp_smem)
    {
        buf = reinterpret_cast<Work *>(warp_smem); head = tail = 0; hmask = 0;
        const uint32_t lane = threadIdx.x & 31u;
        hold = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(warp_smem) + kQueueSlots * sizeof(Work)) + lane;
        hring = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(warp_smem) + kQueueSlots * sizeof(Work)
                                                       + (size_t)M::kMaxSends * 32 * sizeof(Event)) + lane;
#pragma unroll
        for (int k = 0; k < M::kMaxSends; ++k) { hidx[k] = 0; hgrp[k] = lane; }
        gmin = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(warp_smem) + kQueueSlots * sizeof(Work)
                                                      + (size_t)M::kMaxSends * 32 * (sizeof(Event) + 8));
        l0 = 0;
    }
    // issue the record stores of the previous flush (all lanes of the warp call it)
    __device__ __forceinline__ void complete(const Dev<M> &d, Acc &acc)
    {
#pragma unroll
        for (int k = 0; k < M::kMaxSends; ++k) {
            uint32_t idx = hidx[k];
            const bool held = (hmask & (1u << k)) != 0u;
            unsigned long long rp = held ? hring[k * 32] : 0ull;
            if (d.world > 1) {                               // kernel-uniform: staged messages take their slot from the group's base
                const uint32_t base = __shfl_sync(0xffffffffu, hidx[k], (int)(hgrp[k] & 31u));
                if (rp & 1ull) idx = base + (hgrp[k] >> 5);
            }
            if (held) {
                Event e; uint4 *w = reinterpret_cast<uint4 *>(&e);
#pragma unroll
                for (int g = 0; g < G; ++g) w[g] = hold[(k * G + g) * 32];
                if (rp & 1ull) {
                    if (idx >= d.stage_cap) acc.fail(TW_ERR_OUTBOX_OVERFLOW, e.src);
                    else reinterpret_cast<Event *>(rp & ~1ull)[idx] = e;
                } else reinterpret_cast<Event *>(rp)[idx & d.mask] = e;
            }
        }
        hmask = 0;
    }
    // run the first n (<= 32) queued items, one per lane
    __device__ __forceinline__ void run(const Dev<M> &d, Acc &acc, uint32_t parity, uint32_t n)
    {
        const uint32_t lane = threadIdx.x & 31u;
        complete(d, acc);                                    // the previous flush's messages
        Event out[M::kMaxSends]; int nn = 0; uint32_t lp = 0;
        if (lane < n) {
            const Work w = buf[(head + lane) & (kQueueSlots - 1)];
            M::emit(d.mp, w, out, nn);
            lp = M::work_lp(w);
        }
#pragma unroll
        for (int k = 0; k < M::kMaxSends; ++k) {
            uint32_t owner = d.rank, l = 0;
            bool valid = k < nn;
            if (valid) valid = emit_classify<M>(d, acc, lp, out[k], owner, l);
            if (valid && d.closed) atomic_min_u64(&gmin[(lp - d.lp_base - l0) / M::kGroup], ordered_bits(out[k].recv_ts));
            const bool staged = valid && owner != d.rank;
            unsigned long long rp = 0ull;
            if (d.world > 1) {
                // warp-aggregated reservation: the messages of this flush that go to the same GPU take consecutive slots of
                // its staging ring from ONE atomic on the sender-local counter (issued by the group's first lane)
                const unsigned grp = __match_any_sync(0xffffffffu, staged ? owner : 0xFFFFFFFFu);
                if (staged) {
                    const uint32_t leader = (uint32_t)__ffs((int)grp) - 1u;
                    const uint32_t rk = (uint32_t)__popc(grp & ((1u << lane) - 1u));
                    hgrp[k] = leader | (rk << 5);
                    if (rk == 0u) hidx[k] = atomic_add_u32(&d.stage_cnt[owner * kCntStride], (uint32_t)__popc(grp));
                    rp = (unsigned long long)(d.stage_out[owner] + (size_t)parity * d.stage_cap) | 1ull;
                }
            }
            if (valid && !staged) {
                hidx[k] = in_total(atomic_add_u64(&d.in[l].word, in_add(parity)));
                hgrp[k] = lane;
                rp = (unsigned long long)(d.pages + (size_t)l * d.cap);
            }
            if (valid) {
                const uint4 *src = reinterpret_cast<const uint4 *>(&out[k]);
#pragma unroll
                for (int g = 0; g < G; ++g) hold[(k * G + g) * 32] = src[g];
                hring[k * 32] = rp;
                hmask |= 1u << k;
            }
        }
        head += n;
        __syncwarp();
    }
    // every lane of the warp calls push; lanes with has == true contribute one item
    __device__ __forceinline__ void push(const Dev<M> &d, Acc &acc, uint32_t parity, bool has, const Work &w)
    {
        const unsigned b = __ballot_sync(0xffffffffu, has);
        if (b == 0u) return;
        if (has) buf[(tail + __popc(b & ((1u << (threadIdx.x & 31u)) - 1u))) & (kQueueSlots - 1)] = w;
        tail += __popc(b);
        __syncwarp();
        if (tail - head >= 32u) run(d, acc, parity, 32u);
    }
    __device__ __forceinline__ void drain(const Dev<M> &d, Acc &acc, uint32_t parity)
    {
        if (tail != head) run(d, acc, parity, tail - head);
        complete(d, acc);
    }
#else
    void init(void *) {}
    void push(const Dev<M> &d, Acc &acc, uint32_t parity, bool has, const Work &w) { if (has) run_work<M>(d, acc, parity, w); }
    void drain(const Dev<M> &, Acc &, uint32_t) {}
#endif
};

// Mutable per-LP step context handed to the out-of-line slow path.
template <class M>
struct StepCtx {
    double lvt;
    uint32_t head, nproc, consumed0, consumed, head_prev;
    uint32_t rb;                                      // out: processed position to roll back to (nproc = none)
    bool live;
};

// Slow path of the inbox drain (taken only when an arrival is an anti-message or may be a straggler): exactly the
// reference's enqueueMessage case analysis (Internal/InputMessageQueue.hs:120-162).  It annihilates pairs and
// determines the rollback position; the rollback itself is done by the caller (its anti-messages go through the
// warp's work queue).
template <class M>
TW_NOINLINE void drain_slow(const Dev<M> &d, Acc &acc, StepCtx<M> &c, uint32_t l)
{
    using Event = typename M::Event;
    const uint32_t mask = d.mask;
    Event *pg = d.pages + (size_t)l * d.cap;
    const uint32_t head = c.head, nproc = c.nproc, consumed = c.consumed;
    uint32_t rb = nproc;
    for (uint32_t a = c.consumed0; a != consumed; ++a) {
        const Event e = pg[a & mask];
        if (e.ctr & kAntiBit) {
            // find the positive twin, newest first, among what arrived in EARLIER super-steps: an anti-message is sent at
            // least one step after its twin, and there is at most one live copy there.  The arrivals of this step are
            // no candidates: the copy re-sent by the rolled-back sender can carry the same identity (counter, times) with
            // a different payload and, the deliveries of one flush being unordered, may sit in front of the anti-message.
            // (the receive times of eight records are fetched together: on a hot LP the region holds a thousand records,
            // and one dependent load per record made a single anti-message cost hundreds of microseconds)
            uint32_t b = c.consumed0; bool found = false;
            while (b != head && !found) {
                const uint32_t nb = b - head < 8u ? b - head : 8u;
                double ts8[8];
#pragma unroll
                for (uint32_t u = 0; u < 8u; ++u) ts8[u] = u < nb ? pg[(b - 1u - u) & mask].recv_ts : neg_inf();
#pragma unroll
                for (uint32_t u = 0; u < 8u; ++u) {
                    if (found || u >= nb || ts8[u] != e.recv_ts) continue;
                    const Event m = pg[(b - 1u - u) & mask];
                    if (!(m.ctr & (kAntiBit | kDeadBit)) && same_message(m, e)) { found = true; b = b - 1u - u; }
                }
                if (!found) b -= nb;
            }
            if (!found) { fail_checked<M>(d, acc, TW_ERR_ANTI_UNMATCHED, l, c.head_prev); c.live = false; return; }
            const uint32_t rel = b - head;
            if (rel < rb) rb = rel;                   // twin already processed: roll back including it
            pg[b & mask].ctr |= kDeadBit;
            pg[a & mask].ctr = e.ctr | kDeadBit;
            acc.annih++;
        } else if (nproc > 0 && !(e.recv_ts > c.lvt)) {
            // possible straggler: walk back over processed entries whose key is greater
            uint32_t r = rb < nproc ? rb : nproc;
            while (r > 0 && key_less(e, pg[(head + r - 1) & mask])) --r;
            if (r < nproc) { if (r < rb) rb = r; acc.strag++; }
        }
    }
    c.rb = rb;
}

// Compaction under memory pressure: annihilated pending entries (tombstones) whose time lies far ahead would
// otherwise sit in the ring until GVT reaches them.  Gather them next to the processed region, slide the processed
// entries (and their saved states) over them and release the slots through `head`.  Returns the new head.
template <class M>
TW_NOINLINE uint32_t compact_slow(const Dev<M> &d, Acc &acc, uint32_t l, uint32_t head, uint32_t nproc, uint32_t consumed)
{
    using Event = typename M::Event;
    const uint32_t mask = d.mask;
    Event *pg = d.pages + (size_t)l * d.cap;
    typename M::State *sv = d.saved + (size_t)l * d.cap;
    const uint32_t pb = head + nproc;
    uint32_t w = pb;
    for (uint32_t a0 = pb; a0 != consumed; ) {                 // (the flag words of eight records are fetched together)
        const uint32_t nb = consumed - a0 < 8u ? consumed - a0 : 8u;
        uint64_t c8[8];
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) c8[u] = u < nb ? pg[(a0 + u) & mask].ctr : 0ull;
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) {
            if (u >= nb || !(c8[u] & kDeadBit)) continue;
            const uint32_t a = a0 + u;
            if (a != w) { const Event e = pg[a & mask]; pg[a & mask] = pg[w & mask]; pg[w & mask] = e; }
            ++w;
        }
        a0 += nb;
    }
    const uint32_t k = w - pb;
    if (k) {
        for (uint32_t r = nproc; r-- > 0;) {
            pg[(head + r + k) & mask] = pg[(head + r) & mask];
            if constexpr (!M::kReversible) sv[(head + r + k) & mask] = sv[(head + r) & mask];
        }
        head += k; acc.fossil += k;
    }
    return head;
}

// Rare path of the selection: two pending records with equal receive times -> order by the full key.
template <class M>
TW_NOINLINE bool key_less_at(const typename M::Event *pg, uint32_t mask, uint32_t pbeg, uint32_t a, uint32_t b)
{
    const typename M::Event x = pg[(pbeg + a) & mask], y = pg[(pbeg + b) & mask];
    return key_less(x, y);
}

// Minimum by full key over n pending records starting at ring index beg (sequential; the CPU emulation's search
// for long regions and the tie fallback of the cooperative search).
template <class M>
TW_NOINLINE uint32_t min_by_full_key(const typename M::Event *pg, uint32_t mask, uint32_t beg, uint32_t n, double *ts_out)
{
    uint32_t best = 0; typename M::Event be = pg[beg & mask];
    for (uint32_t k = 1; k < n; ++k) {
        const typename M::Event e = pg[(beg + k) & mask];
        if (key_less(e, be)) { be = e; best = k; }
    }
    *ts_out = be.recv_ts;
    return best;
}

#if TW_SIMT
// Cooperative search of a long pending region (more than kMaskBits entries): the whole warp looks for the
// earliest record of the leader's region, 32 records per load (four loads in flight per lane).  All lanes call it with
// the leader's arguments; the result is meaningful on every lane.  When the minimum receive time occurs more than once
// (typically an annihilated message and its anti-message, both waiting to be skipped as no-ops) the order is settled by
// the full key (send time, sender, counter) -- also cooperatively: a second pass over the records of that time.
template <class M>
__device__ __forceinline__ void coop_min(const typename M::Event *lpg, uint32_t mask, uint32_t beg, uint32_t n,
                                         double &ts_out, uint32_t &ix_out)
{
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long kInf = ordered_bits(pos_inf());    // "nothing": an empty region reports +inf
    unsigned long long bo = kInf; uint32_t bix = 0; bool tie = false;
    for (uint32_t k = lane; k < n; k += 128u) {
        unsigned long long o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t kk = k + 32u * (uint32_t)u;
            o[u] = kk < n ? ordered_bits(lpg[(beg + kk) & mask].recv_ts) : kInf;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (o[u] < bo) { bo = o[u]; bix = k + 32u * (uint32_t)u; tie = false; }
            else if (o[u] == bo && o[u] != kInf) tie = true;
        }
    }
    unsigned long long v = bo;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    const bool mine = (bo == v) && v != kInf;
    const unsigned win = __ballot_sync(0xffffffffu, mine);
    const unsigned tw = __ballot_sync(0xffffffffu, mine && tie);
    ts_out = from_ordered(v);
    if (tw == 0u && __popc(win) <= 1) {                       // a unique minimum
        ix_out = __shfl_sync(0xffffffffu, bix, win ? __ffs((int)win) - 1 : 0);
        return;
    }
    // second pass: among the records with that receive time, the smallest (send time, sender, counter)
    unsigned long long bs = ~0ull, bc = ~0ull; uint32_t bsrc = 0xFFFFFFFFu; bix = 0;
    for (uint32_t k = lane; k < n; k += 32u) {
        const typename M::Event *pe = lpg + ((beg + k) & mask);
        if (ordered_bits(pe->recv_ts) != v) continue;
        const unsigned long long so = ordered_bits(pe->send_ts), c = pe->ctr & kCtrMask; const uint32_t sr = pe->src;
        if (so < bs || (so == bs && (sr < bsrc || (sr == bsrc && c < bc)))) { bs = so; bsrc = sr; bc = c; bix = k; }
    }
    unsigned long long ms = bs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, ms, o); ms = t < ms ? t : ms; }
    uint32_t msrc = (bs == ms) ? bsrc : 0xFFFFFFFFu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const uint32_t t = __shfl_xor_sync(0xffffffffu, msrc, o); msrc = t < msrc ? t : msrc; }
    unsigned long long mc = (bs == ms && bsrc == msrc) ? bc : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, mc, o); mc = t < mc ? t : mc; }
    const unsigned w2 = __ballot_sync(0xffffffffu, bs == ms && bsrc == msrc && bc == mc && bs != ~0ull);
    ix_out = __shfl_sync(0xffffffffu, bix, w2 ? __ffs((int)w2) - 1 : 0);    // identical keys (twins): any of them
}

// The slow path of the inbox drain (drain_slow) run by the whole warp for an LP with a long ring: the case analysis is
// the same, arrival by arrival, but the twin of an anti-message is looked for by all 32 lanes, 128 records per round trip,
// newest first -- on a hot LP of a skewed workload one lane's search through a thousand records was the longest thing a
// super-step had to wait for.  All lanes call it with the leader's arguments and compute the same results.
template <class M>
__device__ __forceinline__ void drain_slow_coop(typename M::Event *pg, uint32_t mask, uint32_t head, uint32_t nproc, uint32_t consumed0,
                                                uint32_t consumed, double lvt, uint32_t &rb_out, uint32_t &annih_out, uint32_t &strag_out,
                                                bool &unmatched)
{
    using Event = typename M::Event;
    const uint32_t lane = threadIdx.x & 31u;
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    uint32_t rb = nproc, annih = 0, strag = 0;
    unmatched = false;
    for (uint32_t a = consumed0; a != consumed; ++a) {
        const Event e = pg[a & mask];                           // (the same record on every lane)
        if (e.ctr & kAntiBit) {
            // the positive twin among what arrived in EARLIER super-steps (see drain_slow), newest first
            uint32_t found = kNone;
            for (uint32_t top = consumed0; top != head && found == kNone; ) {
                const uint32_t span = top - head < 128u ? top - head : 128u;
                double ts[4];
#pragma unroll
                for (uint32_t u = 0; u < 4u; ++u) {
                    const uint32_t off = u * 32u + lane;
                    ts[u] = off < span ? pg[(top - 1u - off) & mask].recv_ts : neg_inf();
                }
                uint32_t my = kNone;                            // the newest match among this lane's four records
#pragma unroll
                for (uint32_t u = 0; u < 4u; ++u) {
                    const uint32_t off = u * 32u + lane;
                    if (my != kNone || off >= span || ts[u] != e.recv_ts) continue;
                    const Event m = pg[(top - 1u - off) & mask];
                    if (!(m.ctr & (kAntiBit | kDeadBit)) && same_message(m, e)) my = off;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const uint32_t t = __shfl_xor_sync(0xffffffffu, my, o); my = t < my ? t : my; }
                if (my != kNone) found = top - 1u - my;
                top -= span;
            }
            if (found == kNone) { unmatched = true; break; }
            const uint32_t rel = found - head;
            if (rel < rb) rb = rel;                             // twin already processed: roll back including it
            if (lane == 0u) { pg[found & mask].ctr |= kDeadBit; pg[a & mask].ctr = e.ctr | kDeadBit; }
            __syncwarp();
            annih++;
        } else if (nproc > 0 && !(e.recv_ts > lvt)) {
            // possible straggler: walk back over processed entries whose key is greater
            uint32_t r = rb < nproc ? rb : nproc;
            while (r > 0 && key_less(e, pg[(head + r - 1) & mask])) --r;
            if (r < nproc) { if (r < rb) rb = r; strag++; }
        }
    }
    rb_out = rb; annih_out = annih; strag_out = strag;
}

// broadcast a plain struct from one lane of the warp (32 bits at a time)
template <class T>
__device__ __forceinline__ T shfl_pod(const T &v, int src)
{
    static_assert(sizeof(T) % 4 == 0, "shfl_pod moves 32-bit words");
    T r; const uint32_t *in = reinterpret_cast<const uint32_t *>(&v); uint32_t *out = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 4; ++i) out[i] = __shfl_sync(0xffffffffu, in[i], src);
    return r;
}

// Cooperative bulk selection for a long pending region (hot LPs of a skewed workload hold a thousand pending events and
// execute dozens of them per super-step; one whole-region search per event made that quadratic).  The warp selects up to
// `want` (<= 32) of the earliest in-window records of the leader's region in ONE pass and moves them, in order, to the
// front of the region, where the owner then executes them one after the other without any further search.
//   pass 1   every lane keeps the four smallest keys of its strided share of the region.  A key is (receive time
//            quantised to 32 bits over [gvt, limit]) << 32 | position: one 64-bit min/max per comparison.
//   merge    four bitonic sort + merge rounds over the lanes leave the 32 smallest of those 128 keys, sorted, one per lane.
//   validity everything below `bound` -- the smallest key a lane had to drop, or that fell out of a merge -- is certainly
//            among them; the sorted prefix below it is exact.  Two neighbours with the same time quantum (or the bound in
//            the quantum of the last one) need the full key: the prefix ends before them; if it is empty, the caller does
//            one exact whole-region search (coop_min) and comes back.
//   move     the m selected records go to positions 0..m-1; the records they displace go to the vacated positions.
// All lanes call it with the leader's arguments.  Returns m; rest_out = lower bound of the receive times of everything
// that is still pending behind those m records, in_window_left = whether any of it lies inside the window.
template <class M>
TW_NOINLINE uint32_t coop_select(typename M::Event *lpg, uint32_t mask, uint32_t beg, uint32_t n, uint64_t gbits,
                                                uint64_t lbits, uint32_t want, double &rest_out, bool &in_window_left)
{
    using Event = typename M::Event;
    const uint32_t lane = threadIdx.x & 31u;
    constexpr unsigned long long kNone = ~0ull;
    uint32_t qs = 0;                                            // (lbits - gbits) >> qs < 2^32 - 1
    { const uint64_t R = lbits - gbits; const uint32_t len = bitlen64(R); qs = len > 32u ? len - 32u : 0u; if ((R >> qs) >= 0xFFFFFFFFull) ++qs; }
    unsigned long long c0 = kNone, c1 = kNone, c2 = kNone, c3 = kNone, bound = kNone;   // this lane's four smallest keys; smallest dropped
    uint64_t oow = ~0ull;                                       // earliest record beyond the window (exact bits)
    for (uint32_t k = lane; k < n; k += 256u) {                 // (eight loads in flight per lane: the pass is latency-bound)
        uint64_t tb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t kk = k + 32u * (uint32_t)u;
            tb[u] = kk < n ? d2bits(lpg[(beg + kk) & mask].recv_ts) : ~0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t kk = k + 32u * (uint32_t)u;
            const bool in = tb[u] <= lbits;
            oow = (!in && tb[u] < oow) ? tb[u] : oow;
            unsigned long long cur = in ? (((tb[u] - gbits) >> qs) << 32) | kk : kNone;
            unsigned long long lo;
            lo = cur < c0 ? cur : c0; cur = cur < c0 ? c0 : cur; c0 = lo;
            lo = cur < c1 ? cur : c1; cur = cur < c1 ? c1 : cur; c1 = lo;
            lo = cur < c2 ? cur : c2; cur = cur < c2 ? c2 : cur; c2 = lo;
            lo = cur < c3 ? cur : c3; cur = cur < c3 ? c3 : cur; c3 = lo;
            bound = cur < bound ? cur : bound;
        }
    }
    // the 32 smallest of the 4 x 32 kept keys, ascending over the lanes
    unsigned long long best = kNone;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        unsigned long long v = r == 0 ? c0 : r == 1 ? c1 : r == 2 ? c2 : c3;
#pragma unroll
        for (uint32_t kk = 2; kk <= 32u; kk <<= 1)               // bitonic sort across the lanes
#pragma unroll
            for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, (int)j);
                const bool take_min = ((lane & j) == 0u) == ((lane & kk) == 0u);
                v = take_min ? (o < v ? o : v) : (o > v ? o : v);
            }
        if (r == 0) { best = v; continue; }
        // merge: lane i keeps the smaller of best[i] and v[31 - i]; what falls out bounds the rest
        const unsigned long long o = __shfl_sync(0xffffffffu, v, (int)(31u - lane));
        const unsigned long long out = o < best ? best : o;
        best = o < best ? o : best;
        bound = out < bound ? out : bound;
#pragma unroll
        for (uint32_t j = 16; j > 0; j >>= 1) {                 // the result is bitonic: five more stages sort it
            const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, best, (int)j);
            best = ((lane & j) == 0u) ? (o2 < best ? o2 : best) : (o2 > best ? o2 : best);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, bound, o); bound = t < bound ? t : bound; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const uint64_t t = __shfl_xor_sync(0xffffffffu, oow, o); oow = t < oow ? t : oow; }
    // the exact prefix: below the bound, and no two neighbours (nor the last one and the bound) in one time quantum
    // One kind of tie is harmless and common: an annihilated message and its anti-message, both waiting to be skipped as
    // no-ops, have the very same receive time.  A run of neighbours in one quantum may stay in the prefix if ALL of them
    // are equal in time and ALL are annihilated: their mutual order is irrelevant.  Any other run ends the prefix where it
    // starts: an entry of the same quantum with an earlier exact time may sit anywhere inside it, and even one live record
    // among no-ops of its own time must not be misplaced -- the processed region has to stay sorted by the full key among
    // live records, or the straggler test (which walks back until it meets a smaller key) would stop at a no-op in front of
    // a live record it ought to roll back (found by the warp-level emulator on a workload where every time is tied).
    const unsigned long long nxt = __shfl_sync(0xffffffffu, best, (int)((lane + 1u) & 31u));
    uint64_t my_ctr = 0; double my_ts = 0.0;
    if (best != kNone) { const Event *pe = lpg + ((beg + (uint32_t)best) & mask); my_ctr = pe->ctr; my_ts = pe->recv_ts; }
    const bool dead = (my_ctr & kDeadBit) != 0ull;
    const double nxt_ts = __shfl_sync(0xffffffffu, my_ts, (int)((lane + 1u) & 31u));
    const bool same_next = lane < 31u && best != kNone && nxt != kNone && (best >> 32) == (nxt >> 32);
    const unsigned samem = __ballot_sync(0xffffffffu, same_next);
    const unsigned diffm = __ballot_sync(0xffffffffu, same_next && my_ts != nxt_ts);     // one quantum, two times
    const unsigned livem = __ballot_sync(0xffffffffu, best != kNone && !dead);
    // this lane's run: lanes [run_s, run_e], linked by set bits of samem
    const unsigned below = ~samem & ((1u << lane) - 1u);
    const uint32_t run_s = bitlen64((uint64_t)below);
    const unsigned above = ~samem >> lane;                      // (bit 31 of samem is never set, so `above` is never 0)
    const uint32_t run_e = lane + (uint32_t)__ffs((int)above) - 1u;
    const unsigned runm = (run_e >= 31u ? 0xFFFFFFFFu : ((1u << (run_e + 1u)) - 1u)) & ~((1u << run_s) - 1u);
    const bool bad_run = run_e > run_s && ((diffm & runm & ~(1u << run_e)) != 0u || (livem & runm) != 0u);
    const unsigned badrm = __ballot_sync(0xffffffffu, bad_run);
    const uint32_t run_limit = badrm ? (uint32_t)__ffs((int)badrm) - 1u : 32u;             // first lane of the first bad run
    const bool ok = best != kNone && lane < want && lane < run_limit && best < bound && (best >> 32) != (bound >> 32);
    const unsigned okm = __ballot_sync(0xffffffffu, ok);
    const uint32_t m = okm == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~okm) - 1u;          // leading run of ok lanes
    // what stays pending behind the prefix: the next key (or the bound), and the records beyond the window
    unsigned long long restk = bound;
    { const unsigned long long bm = __shfl_sync(0xffffffffu, best, (int)(m & 31u)); if (m < 32u && bm < restk) restk = bm; }
    in_window_left = restk != kNone;
    double rest = pos_inf();
    if (restk != kNone) rest = bits2d(gbits + ((restk >> 32) << qs));
    if (oow != ~0ull) rest = dmin(rest, bits2d(oow));
    rest_out = rest;
    if (m == 0u) return 0u;
    // move: selected record i -> position i; the records displaced from the first m positions -> the vacated positions
    const bool sel = lane < m;
    const uint32_t pos = (uint32_t)best;
    unsigned lowsel = (sel && pos < m) ? (1u << pos) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lowsel |= __shfl_xor_sync(0xffffffffu, lowsel, o);
    const bool disp = lane < m && !((lowsel >> lane) & 1u);
    const unsigned dmask = __ballot_sync(0xffffffffu, disp);
    const unsigned hmask = __ballot_sync(0xffffffffu, sel && pos >= m);     // as many holes as displaced records
    Event rs, rd;
    if (sel) rs = lpg[(beg + pos) & mask];
    if (disp) rd = lpg[(beg + lane) & mask];
    uint32_t hl = 0;                                            // the lane that selected this lane's hole
    if (disp) {
        uint32_t k = (uint32_t)__popc(dmask & ((1u << lane) - 1u));
        unsigned hmm = hmask;
        while (k--) hmm &= hmm - 1u;
        hl = (uint32_t)__ffs((int)hmm) - 1u;
    }
    const uint32_t hp = __shfl_sync(0xffffffffu, pos, (int)(hl & 31u));
    __syncwarp();
    if (sel) lpg[(beg + lane) & mask] = rs;
    if (disp) lpg[(beg + hp) & mask] = rd;
    __syncwarp();
    return m;
}
#endif

// The super-step of LP `l` (local index).  `gvt` is the GVT of this step, `wend` the end of the optimistic window,
// `step` the super-step number.  Must be called by every lane of the warp (lanes with l >= n_local pass through idle).
// kBulk: compile the cooperative bulk selection for long pending regions in (coop_select).  The engine keeps two builds
// of k_step and runs the one with it only while the partition has such regions: the extra state costs the common path
// registers (c4 6.0 -> 5.3 G ev/s when it is always compiled in).
template <class M, bool kBulk = true>
TW_HD void lp_step(const Dev<M> &d, Acc &acc, WorkQ<M> &q, uint32_t l, double gvt, double wend, uint32_t step)
{
    const uint32_t parity = step & 1u;
    using Event = typename M::Event;
    using State = typename M::State;
    using Work = typename M::Work;
    const uint32_t lp = d.lp_base + l;
    const uint32_t par_prev = parity ^ 1u;
    const uint32_t mask = d.mask;
    bool live = l < d.n_local;
    LPHdr h; h.head = h.np_hp = h.consumed = h.rollbacks = 0; h.lvt = neg_inf(); h.min_pend = pos_inf();
    State st;
    uint32_t head, nproc, consumed0, consumed;
    uint32_t narr = 0;
    // (the inbox word is read past L1: senders change it with atomics, which L1 does not see -- the closed-group kernel
    // runs many super-steps in one launch)
    if (live) { h = d.hdr[l]; narr = in_count(load_inbox_word(&d.in[l]), par_prev); }
    uint32_t head_prev = hdr_head_prev(h);

    // idle LP: no arrivals and nothing pending inside the window -> only the GVT contribution
    if (live && narr == 0 && !(h.min_pend <= wend)) {
        acc.pend_min = dmin(acc.pend_min, h.min_pend);
        if (head_prev != h.head) d.hdr[l].np_hp = hdr_pack(hdr_nproc(h), h.head, h.head);
        live = false;
    }

    Event *pg = d.pages + (size_t)(live ? l : 0) * d.cap;
    State *sv = d.saved + (size_t)(live ? l : 0) * d.cap;
    head = h.head; nproc = hdr_nproc(h);
    consumed0 = h.consumed; consumed = consumed0 + narr;
    if (live) {
        if (narr) atomic_add_u64(&d.in[l].word, 0ull - ((unsigned long long)narr << (16u * par_prev)));   // consume them
        if (consumed - head_prev > d.cap) { acc.fail(TW_ERR_QUEUE_OVERFLOW, lp); live = false; }
        head_prev = head;
    }
    if (live) st = d.state[l];

    const uint32_t hdr_consumed_guess = consumed - head; (void)hdr_consumed_guess;
    TW_T0();
    // ---- 1. lazy fossil collection, strict `< gvt`: four independent probes at the quartiles of the processed region
    // (Every fourth step unless the processed region is about to throttle optimism: the probes are cold DRAM reads
    // that the whole warp waits for, so the lanes of a warp had better take them in the same step.)
    if (live && nproc >= d.fossil_thresh && ((step & 3u) == 0u || 2u * nproc >= d.proc_limit + d.fossil_thresh)) {
        const uint32_t n = nproc;
        const uint32_t q1 = n >> 2, q2 = n >> 1, q3 = n - (n >> 2), q4 = n;      // candidate cut counts
        const double t1 = q1 ? pg[(head + q1 - 1) & mask].recv_ts : neg_inf();
        const double t2 = q2 ? pg[(head + q2 - 1) & mask].recv_ts : neg_inf();
        const double t3 = pg[(head + q3 - 1) & mask].recv_ts;
        const double t4 = pg[(head + q4 - 1) & mask].recv_ts;
        uint32_t cut = 0;
        if (q1 && t1 < gvt) cut = q1;
        if (q2 && t2 < gvt) cut = q2;
        if (t3 < gvt) cut = q3;
        if (t4 < gvt) cut = q4;
        head += cut; nproc -= cut; acc.fossil += cut;
    }
    TW_RECONVERGE();

    TW_TICK(0);
    // ---- 2. drain the inbox, fast path: arrivals are the last narr entries of the region.  Anything that
    // needs the processed region (anti-message, possible straggler) takes the slow path.
    bool slow = false;
    if (live) {
        for (uint32_t a0 = consumed0; a0 != consumed; ) {
            double ts[4]; uint64_t ct[4];
            const uint32_t left = consumed - a0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {                       // independent loads: one exposed latency per four arrivals
                if ((uint32_t)u < left) { const Event *pe = pg + ((a0 + (uint32_t)u) & mask); ts[u] = pe->recv_ts; ct[u] = pe->ctr; }
                else { ts[u] = pos_inf(); ct[u] = 0; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!(ts[u] >= gvt)) { fail_checked<M>(d, acc, TW_ERR_OUTDATED_MESSAGE, l, head_prev); live = false; }
                if ((ct[u] & kAntiBit) || (nproc > 0 && !(ts[u] > h.lvt))) slow = true;
            }
            a0 += left < 4u ? left : 4u;
        }
        if (!live) slow = false;
    }
    TW_TICK(1);
    uint32_t rb = nproc;                                         // processed position to roll back to (nproc = none)
#if TW_SIMT
    if constexpr (kBulk) {
        // an LP with a long ring: the slow path runs with all 32 lanes (drain_slow_coop)
        for (unsigned sm = __ballot_sync(0xffffffffu, slow && consumed0 - head > 8u * kMaskBits); sm; sm &= sm - 1u) {
            const int leader = __ffs(sm) - 1;
            Event *lpg = reinterpret_cast<Event *>(__shfl_sync(0xffffffffu, (unsigned long long)pg, leader));
            uint32_t rb_ = 0, annih_ = 0, strag_ = 0; bool unm_ = false;
            drain_slow_coop<M>(lpg, mask, __shfl_sync(0xffffffffu, head, leader), __shfl_sync(0xffffffffu, nproc, leader),
                               __shfl_sync(0xffffffffu, consumed0, leader), __shfl_sync(0xffffffffu, consumed, leader),
                               __shfl_sync(0xffffffffu, h.lvt, leader), rb_, annih_, strag_, unm_);
            if ((int)(threadIdx.x & 31u) == leader) {
                rb = rb_; acc.annih += annih_; acc.strag += strag_; slow = false;
                if (unm_) { fail_checked<M>(d, acc, TW_ERR_ANTI_UNMATCHED, l, head_prev); live = false; }
            }
        }
    }
#endif
    if (slow) {
        StepCtx<M> c;                                            // lives in local memory only on this rare path
        c.lvt = h.lvt; c.head_prev = head_prev; c.head = head; c.nproc = nproc; c.consumed0 = consumed0; c.consumed = consumed; c.rb = nproc; c.live = live;
        Acc sacc; sacc.init();                                   // separate accumulators: `acc` must stay in registers
        drain_slow<M>(d, sacc, c, l);
        acc.merge(sacc);
        rb = c.rb; live = c.live;
    }
    TW_RECONVERGE();

    TW_TICK(2);
    // ---- 3. rollback to processed position rb: restore the saved state (rollbackLog, Internal/UndoableLog.hs:93-118),
    // un-process the inputs (rollbackInputMessages, Internal/InputMessageQueue.hs:190-201) and queue the anti-messages
    // of the undone executions, regenerated from the saved pre-states.  Warp-uniform loop: one undone entry per lane
    // and iteration.
    {
        const uint32_t rend = live ? nproc : 0u;
        const uint32_t rlow = (live && rb < nproc) ? rb : rend;
        if (rlow < rend) { acc.rollbacks++; h.rollbacks++; }
        if constexpr (M::kReversible) {
            // reverse computation, newest first: undoing entry r leaves the state before r, which is also what its
            // anti-messages are regenerated from
            uint32_t rcur = rend;
            while (warp_any(rcur > rlow)) {
                bool has = false;
                Work w;
                if (rcur > rlow) {
                    rcur--;
                    const Event e = pg[(head + rcur) & mask];
                    if (!(e.ctr & (kAntiBit | kNoopBit))) {        // a no-op when processed changed and sent nothing
                        M::reverse(d.mp, lp, st, e);
                        w = M::make_work(lp, st, e, true); has = true;
                        acc.rb_events++;
                    }
                }
                q.push(d, acc, parity, has, w);
            }
        } else {
            uint32_t rcur = rlow;
            while (warp_any(rcur < rend)) {
                bool has = false;
                if constexpr (M::kSplit) {
                    Work w;
                    if (rcur < rend) {
                        const uint32_t slot = (head + rcur) & mask;
                        const Event e = pg[slot];
                        if (!(e.ctr & (kAntiBit | kNoopBit))) {        // a no-op when processed sent nothing
                            const State s0 = sv[slot];
                            w = M::make_work(lp, s0, e, true); has = true;
                            acc.rb_events++;
                        }
                        rcur++;
                    }
                    q.push(d, acc, parity, has, w);
                } else {
                    Event out[M::kMaxSends]; int n_out = 0;
                    if (rcur < rend) {
                        const uint32_t slot = (head + rcur) & mask;
                        const Event e = pg[slot];
                        if (!(e.ctr & (kAntiBit | kNoopBit))) {
                            State s0 = sv[slot];
                            M::handle(d.mp, lp, s0, e, out, n_out);    // replay on the saved pre-state
                            acc.rb_events++;
                        }
                        rcur++;
                    }
#pragma unroll
                    for (int k = 0; k < M::kMaxSends; ++k) {
                        has = k < n_out;
                        if (has) out[k].ctr |= kAntiBit;
                        q.push(d, acc, parity, has, out[k]);
                    }
                }
            }
            if (rlow < rend) st = sv[(head + rlow) & mask];
        }
        if (rlow < rend) {
            nproc = rlow;
            h.lvt = rlow > 0 ? pg[(head + rlow - 1) & mask].recv_ts : neg_inf();
        }
        // (a ring that is half full is not necessarily full of tombstones -- a hot LP of a skewed workload simply has that
        // many pending events --, and the pass reads every pending record: every 16th step, and at once above seven
        // eighths.  Not rarer: tombstones that stay are "executed" as no-ops and use up the LP's batch -- c5 needed 576
        // instead of 545 super-steps with the pass only above three quarters)
        TW_TICK(3);
        if (live && consumed - head >= (d.cap >> 1) && ((step & 15u) == 0u || consumed - head >= d.cap - (d.cap >> 3))) {
            Acc sacc; sacc.init();
            head = compact_slow<M>(d, sacc, l, head, nproc, consumed);
            acc.merge(sacc);
        }
    }
    TW_RECONVERGE();

    TW_TICK(4);
    // ---- 4./5. select and execute in key order, at most max_batch events.
    // Selection (light lanes: regions of at most kMaskBits entries, non-negative times): ONE pass over the pending
    // region sorts the entries inside the optimistic window by a packed 32-bit key -- the receive time quantised to 27
    // bits over [gvt, limit] (the bit pattern of a non-negative double is monotone in its value, so this is integer
    // arithmetic only) and the entry's 5-bit index -- with a branch-free min/max insertion network that keeps the
    // kCand smallest.  Two entries in one quantum (or more in-window entries than the list holds at a quantum
    // boundary) send the lane to the exact search that long regions use anyway.
    // Execution: one event per lane and iteration.  The executed entry moves to ring index pbeg + done (the
    // processed region grows over the front of the pending region); the pending entry that sat there takes the
    // vacated slot (`mv` remembers where to, for candidates still in the list).  The lane does the state transition
    // and queues the Work item.
    const uint32_t pbeg = head + nproc;                      // first pending ring index
    const uint32_t nreg = live ? consumed - pbeg : 0u;       // pending entries (arrivals included)
    const double limit = wend;
    const uint64_t gbits = d2bits(gvt), lbits = d2bits(limit);
    const bool exact_only = (gbits >> 63) != 0ull || !(gvt <= limit);   // negative times: always the exact search
    uint32_t qshift = 0;
    if (!exact_only) {
        const uint64_t R = lbits - gbits;                   // kernel-uniform: (R >> qshift) < 2^kQuantBits - 1
        const uint32_t len = bitlen64(R);
        qshift = len > kQuantBits ? len - kQuantBits : 0u;
        if ((R >> qshift) >= ((1ull << kQuantBits) - 1ull)) ++qshift;
    }
    bool heavy = live && (nreg > kMaskBits || exact_only);  // exact search over the whole region, no candidate list
    // long regions on the GPU: up to 32 events at a time are selected and lined up by the whole warp (coop_select)
    const bool bulk = kBulk && TW_SIMT && heavy && !exact_only;
    // (tells the host which build of k_step to launch next: a few dozen entries over the list's reach -- a c4 run has some
    // hundred such LPs in every step -- are handled well enough by the per-event search; hundreds of entries are not)
    if (live && nreg > 8u * kMaskBits) {
        acc.long_lanes++;
        acc.max_occ = consumed - head > acc.max_occ ? consumed - head : acc.max_occ;   // (only worth tracking where rings fill up)
    }
    bool exact_once = false;                                 // bulk lane: the front of the region is tied, one exact search
    double bulk_rest = pos_inf();                            // bulk lane: lower bound of what the last selection left behind
    double rest_min = pos_inf();
    uint32_t done = 0;                                       // events executed in this step
    bool stopped = !live;                                    // no further execution by this lane in this step
    bool searched = !heavy;                                  // heavy lanes: rest_min covers the remaining region
    const uint32_t maxb = d.max_batch;
    uint32_t c[kCand];                                       // sorted candidate keys, kNoCand = empty
#pragma unroll
    for (int i = 0; i < kCand; ++i) c[i] = kNoCand;
    uint32_t drop = kNoCand;                                 // smallest in-window key that did not fit into the list
    uint32_t nleft = 0;                                      // in-window entries of the current round not yet executed
    uint32_t &presel = nleft;                                // bulk lane (no candidate list): selected records lined up at the
                                                             // front of the region, not yet executed -- the same register
    uint32_t rbase = 0, mv = 0;                              // round base (= done at the scan); displaced-entry map
    bool need_scan = live && !heavy;
    const char *pgb = reinterpret_cast<const char *>(pg);
    for (;;) {                                               // rounds: one scan, then up to kCand executions per lane
        if (!warp_any(need_scan || (heavy && live && !searched))) break;
        if (need_scan) {
            rbase = done; mv = 0; drop = kNoCand; nleft = 0;
            const uint32_t rbeg = pbeg + done, rn = nreg - done;
            uint64_t rest_bits = ~0ull;
            for (uint32_t k0 = 0; k0 < rn; k0 += kScanBatch) {
                uint64_t tb[kScanBatch];
#pragma unroll
                for (int u = 0; u < kScanBatch; ++u) {       // independent loads first: one exposed latency per batch
                    const uint32_t k = k0 + u;
                    const uint32_t off = ((rbeg + k) & mask) * (uint32_t)sizeof(Event);
                    tb[u] = k < rn ? *reinterpret_cast<const uint64_t *>(pgb + off) : ~0ull;   // recv_ts bits
                }
#pragma unroll
                for (int u = 0; u < kScanBatch; ++u) {
                    const uint64_t bits = tb[u];
                    const bool in = bits <= lbits;
                    uint32_t cur = in ? ((uint32_t)((bits - gbits) >> qshift) << 5) | (k0 + u) : kNoCand;
                    nleft += in ? 1u : 0u;
                    rest_bits = (!in && bits < rest_bits) ? bits : rest_bits;
#pragma unroll
                    for (int i = 0; i < kCand; ++i) {
                        const uint32_t lo = cur < c[i] ? cur : c[i];
                        cur = cur < c[i] ? c[i] : cur;
                        c[i] = lo;
                    }
                    drop = cur < drop ? cur : drop;
                }
            }
            if (rest_bits != ~0ull) rest_min = dmin(rest_min, bits2d(rest_bits));
            // Two or more in-window entries in one time quantum: their order needs the exact key (receive time, send time,
            // sender, counter).  The list is sorted by quantum, so such entries are neighbours: an exact insertion sort of
            // the (at most kCand) candidates settles every run of equal quanta -- MachRep2 sends three messages with one
            // receive time per failure (tests/MachRep2DistributedReproducible.hs:133-135), so for that model this is the
            // common case, not the exception.  An annihilated message and its anti-message have identical keys: their
            // mutual order does not matter (both are skipped as no-ops).
            bool tie = false;
            bool any_run = false;
#pragma unroll
            for (int i = 0; i + 1 < kCand; ++i) any_run = any_run || (c[i + 1] != kNoCand && (c[i] >> 5) == (c[i + 1] >> 5));
            if (any_run) {
#pragma unroll
                for (int i = 1; i < kCand; ++i) {
                    bool moving = true;
#pragma unroll
                    for (int j = i; j > 0; --j) {
                        moving = moving && c[j] != kNoCand && (c[j - 1] >> 5) == (c[j] >> 5) &&
                                 key_less_at<M>(pg, mask, rbeg, c[j] & 31u, c[j - 1] & 31u);
                        const uint32_t lo = moving ? c[j] : c[j - 1], hi = moving ? c[j - 1] : c[j];
                        c[j - 1] = lo; c[j] = hi;
                    }
                }
            }
            // the last candidate shares its quantum with an entry that did not fit into the list: `drop` is only the
            // first of possibly several such entries -> the exact search over the whole region
            if (drop != kNoCand && (c[kCand - 1] >> 5) == (drop >> 5)) tie = true;
            if (tie) { heavy = true; searched = false; }
            need_scan = false;
        }
        TW_RECONVERGE();
#if TW_SIMT
        // bulk lanes with nothing lined up: the next selection
        if constexpr (kBulk)
        for (unsigned bm = __ballot_sync(0xffffffffu, bulk && live && !searched && !exact_once && presel == 0u && !stopped && done < maxb); bm; bm &= bm - 1u) {
            const int leader = __ffs(bm) - 1;
            Event *lpg = reinterpret_cast<Event *>(__shfl_sync(0xffffffffu, (unsigned long long)pg, leader));
            const uint32_t lbeg = __shfl_sync(0xffffffffu, pbeg + done, leader);
            const uint32_t ln = __shfl_sync(0xffffffffu, nreg - done, leader);
            const uint32_t lwant = __shfl_sync(0xffffffffu, maxb - done < 32u ? maxb - done : 32u, leader);
            // (GVT and the window are the leader's: in the closed-group kernel every group of the warp has its own)
            const uint64_t lg = __shfl_sync(0xffffffffu, (unsigned long long)gbits, leader), ll = __shfl_sync(0xffffffffu, (unsigned long long)lbits, leader);
            double rest_; bool left_;
            const uint32_t m = coop_select<M>(lpg, mask, lbeg, ln, lg, ll, lwant, rest_, left_);
#if defined(TW_PHASE_TIMERS) && defined(__CUDA_ARCH__)
            if ((int)(threadIdx.x & 31u) == leader) { atomicAdd(&d.ctl->t_phase[6], 1ull); atomicAdd(&d.ctl->t_phase[7], (unsigned long long)m); }
#endif
            // A model whose transition cannot fail and needs no saved state (split + reversible: PHOLD) has its lined-up
            // records executed right here by the whole warp: lane j fetches record j (one coalesced load), every lane runs
            // the leader's short sequential chain of state transitions (same instructions on all lanes, record k comes by
            // shuffle from lane k) and keeps the state before ITS record, and then all lanes queue their Work items at once
            // -- instead of one warp-wide loop iteration per event with a single busy lane.
            uint32_t m_exec = 0; bool thr = false;
            if constexpr (M::kSplit && M::kReversible) {
                if (m) {
                    const uint32_t lane = threadIdx.x & 31u;
                    const uint32_t llp = __shfl_sync(0xffffffffu, lp, leader);
                    State cs = shfl_pod(st, leader);
                    double clvt = __shfl_sync(0xffffffffu, h.lvt, leader);
                    uint32_t cnproc = __shfl_sync(0xffffffffu, nproc, leader);
                    const double lgvt = __shfl_sync(0xffffffffu, gvt, leader);
                    Event e; e.recv_ts = 0.0; e.ctr = 0;
                    if (lane < m) e = lpg[(lbeg + lane) & mask];
                    State pre = cs; bool bad_order = false;
                    for (uint32_t k = 0; k < m; ++k) {
                        const Event ek = shfl_pod(e, (int)k);
                        if (ek.recv_ts < clvt) { bad_order = true; break; }
                        if (cnproc >= d.proc_limit && !(ek.recv_ts <= lgvt)) { thr = true; break; }   // isEventOverflow analogue
                        if (lane == k) pre = cs;
                        if (!(ek.ctr & (kAntiBit | kDeadBit))) M::advance(d.mp, llp, cs, ek);
                        clvt = ek.recv_ts; ++cnproc; ++m_exec;
                    }
                    const bool mine = lane < m_exec;
                    const bool noop = (e.ctr & (kAntiBit | kDeadBit)) != 0ull;
                    if (mine && noop && !(e.ctr & kNoopBit)) lpg[(lbeg + lane) & mask].ctr = e.ctr | kNoopBit;
                    Work w;
                    const bool has = mine && !noop;
                    if (has) { w = M::make_work(llp, pre, e, false); acc.exec++; }
                    q.push(d, acc, parity, has, w);
                    if ((int)lane == leader) {
                        st = cs; h.lvt = clvt; nproc = cnproc; done += m_exec;
                        if (bad_order) { fail_checked<M>(d, acc, TW_ERR_TIME_ORDER, l, head_prev); live = false; stopped = true; searched = true; }
                        if (thr) stopped = true;
                    }
                }
            }
            if ((int)(threadIdx.x & 31u) == leader) {
                presel = m - m_exec; bulk_rest = rest_;
                if (m == 0u) {
                    if (left_ && !(rest_ > limit)) exact_once = true;            // a tie at the front: settle one event exactly
                    else { rest_min = dmin(rest_min, rest_); searched = true; }  // nothing (more) inside the window
                }
            }
        }
#endif

        for (;;) {                                           // iterations within the round
            const bool can_exec = !stopped && done < maxb;
            // a bulk lane that cannot go on reports what it leaves pending: the next lined-up record and the rest
            if (bulk && live && !searched && !exact_once && !can_exec) {
                if (presel) bulk_rest = dmin(bulk_rest, pg[(pbeg + done) & mask].recv_ts);
                rest_min = dmin(rest_min, bulk_rest); searched = true;
            }
            // a heavy lane keeps searching until it knows the earliest record it leaves pending (its GVT contribution)
            const bool lined = bulk && !exact_once;           // executes from the front of its region, no search
            bool want = heavy ? (live && !searched && (!lined || presel > 0u)) : (can_exec && c[0] != kNoCand);
            if (!warp_any(want)) break;
            uint32_t bix = 0; double bts = pos_inf();
            // heavy lanes: earliest record of [pbeg + done, consumed), found by the whole warp
#if TW_SIMT
            for (unsigned hm = __ballot_sync(0xffffffffu, want && heavy && !lined); hm; hm &= hm - 1u) {
                const int leader = __ffs(hm) - 1;
                const Event *lpg = reinterpret_cast<const Event *>(__shfl_sync(0xffffffffu, (unsigned long long)pg, leader));
                const uint32_t lbeg = __shfl_sync(0xffffffffu, pbeg + done, leader);
                const uint32_t ln = __shfl_sync(0xffffffffu, nreg - done, leader);
                double ts_; uint32_t ix_;
                coop_min<M>(lpg, mask, lbeg, ln, ts_, ix_);
                if ((int)(threadIdx.x & 31u) == leader) { bts = ts_; bix = ix_; }
            }
#else
            if (want && heavy && nreg > done) bix = min_by_full_key<M>(pg, mask, pbeg + done, nreg - done, &bts);
#endif
            // isEventOverflow analogue (Internal/Event.hs:289-306): with proc_limit processed-but-uncommitted entries
            // only safe events (<= GVT, fossil-collected next) proceed
            if (want && lined) {
                bix = done;                                   // the next lined-up record; throttled like any other
                if (nproc >= d.proc_limit && !(pg[(pbeg + bix) & mask].recv_ts <= gvt)) { stopped = true; want = false; }
            } else if (want && heavy) {
                bix += done;
                const bool throttled = nproc >= d.proc_limit && !(bts <= gvt);
                if (!(bts <= limit) || !can_exec || throttled) { rest_min = dmin(rest_min, bts); searched = true; want = false; }
            } else if (want) {
                bix = c[0] & 31u;
                const uint32_t rd = done - rbase;            // executions of this round so far
                while (bix < rd) bix = (mv >> (5u * bix)) & 31u;     // displaced since the scan
                bix += rbase;
                if (nproc >= d.proc_limit) {
                    bts = pg[(pbeg + bix) & mask].recv_ts;
                    if (!(bts <= gvt)) { stopped = true; want = false; }
                }
            }
            TW_RECONVERGE();

            bool has = false;
            Work w;
            Event outs[M::kSplit ? 1 : M::kMaxSends]; int n_out = 0;
            if (want) {
                const uint32_t pi = done;                    // relative index of the slot being filled
                Event e = pg[(pbeg + bix) & mask];
                const Event x = pg[(pbeg + pi) & mask];      // its current occupant (loaded together with e)
                if (e.recv_ts < h.lvt) { fail_checked<M>(d, acc, TW_ERR_TIME_ORDER, l, head_prev); live = false; stopped = true; searched = true; }
                else {
                    const State before = st;
                    const bool noop = (e.ctr & (kAntiBit | kDeadBit)) != 0;
                    int status = H_OK;
                    if (!noop) {
                        if constexpr (M::kSplit) status = M::advance(d.mp, lp, st, e);
                        else status = M::handle(d.mp, lp, st, e, outs, n_out);
                    }
                    if (status == H_OK) {
                        if (bix != pi) pg[(pbeg + bix) & mask] = x;     // displaced pending entry takes the vacated slot
                        if (!heavy) {
                            mv |= (bix - rbase) << (5u * (pi - rbase));
#pragma unroll
                            for (int i = 0; i + 1 < kCand; ++i) c[i] = c[i + 1];
                            c[kCand - 1] = kNoCand;
                            nleft--;
                        }
                        if (noop) e.ctr |= kNoopBit;
                        if (bix != pi || noop) pg[(pbeg + pi) & mask] = e;
                        if constexpr (!M::kReversible) sv[(pbeg + pi) & mask] = before;
                        nproc++; h.lvt = e.recv_ts; done++;
                        if (lined) presel--;
                        exact_once = false;
                        if (!noop) {
                            acc.exec++;
                            if constexpr (M::kSplit) { w = M::make_work(lp, before, e, false); has = true; }
                        }
                    } else {
                        st = before; n_out = 0;              // SimulationRetry: stay pending, wait for the rollback
                        acc.retry++; stopped = true;
                    }
                }
            }
            if constexpr (M::kSplit) q.push(d, acc, parity, has, w);
            else {
#pragma unroll
                for (int k = 0; k < M::kMaxSends; ++k) q.push(d, acc, parity, k < n_out, outs[k]);
            }
        }
        // the list is used up but more in-window entries were seen: next round
        need_scan = !heavy && !stopped && done < maxb && c[0] == kNoCand && nleft > 0;
    }
    TW_TICK(5);
    // what stays pending inside the window (batch limit, retry, throttle) still bounds GVT: the earliest such entry
    if (live && !heavy && nleft > 0) {
        if (c[0] != kNoCand) {                               // the list's head is its exact minimum (ties were ruled out)
            uint32_t ix = c[0] & 31u;
            const uint32_t rd = done - rbase;
            while (ix < rd) ix = (mv >> (5u * ix)) & 31u;
            rest_min = dmin(rest_min, pg[(pbeg + rbase + ix) & mask].recv_ts);
        } else {
            // only entries that did not fit into the list are left: `drop` names the earliest quantum among them, but
            // not necessarily the earliest entry inside that quantum -- take the quantum's lower bound (still >= gvt)
            rest_min = dmin(rest_min, bits2d(gbits + ((uint64_t)(drop >> 5) << qshift)));
        }
    }
    TW_RECONVERGE();
    // (the caller drains the work queue: a warp that runs several tiles of LPs carries a partly filled queue over)

    // ---- 6. write back, GVT contribution
    if (live) {
        h.head = head; h.np_hp = hdr_pack(nproc, head, head_prev); h.consumed = consumed; h.min_pend = rest_min;
        d.hdr[l] = h;
        d.state[l] = st;
        acc.pend_min = dmin(acc.pend_min, rest_min);
    }
}

#if TW_SIMT
// The warp-local super-step loop of a closed-group model (k_run_groups in engine.cu; tests/emu runs the same function with
// 32 host threads): the warp owns groups [g0, g0 + gpw) of M::kGroup LPs, one lane = one LP, and runs their super-steps back
// to back, every group under its own GVT.  Returns whether this lane owns an LP; graw = the GVT of the lane's group at
// the end (whatever the external bound `cap` says), rounds = super-steps run (always even: the arrival parity of the
// inboxes stays aligned with the engine's step counter `step0`).
template <class M>
__device__ __forceinline__ bool group_rounds(const Dev<M> &d, Acc &acc, WorkQ<M> &q, uint32_t g0, uint32_t gpw, uint32_t max_rounds,
                                             uint32_t step0, double cap, double &graw_out, uint32_t &rounds_out)
{
    constexpr uint32_t G = M::kGroup;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_groups = d.n_local / G;
    const uint32_t grp = g0 + lane / G;
    const bool mine = lane < gpw * G && grp < n_groups;
    const uint32_t l = mine ? g0 * G + lane : 0xFFFFFFFFu;
    const uint32_t base_lane = lane - lane % G;                            // first lane of my group
    double graw = mine ? from_ordered(d.grp_gvt[grp]) : pos_inf();         // the group's own minimum, whatever the bound says
    uint32_t rounds = 0, stalled = 0;
    q.l0 = g0 * G;
    for (;;) {
        const double gvt = graw < cap ? graw : cap;
        if ((rounds & 1u) == 0u) {                                         // a launch ends after an even number of rounds
            // a group beyond the stop time is done; one that sits at the Time Server's bound waits for the next
            // GlobalTimeMessage (the host returns and comes back)
            const bool active = mine && gvt <= d.stop && !(graw > cap);
            if (!__any_sync(0xffffffffu, active) || rounds >= max_rounds) break;
        }
        double wend = d.stop;
        if (d.horizon >= 0.0) { const double w = gvt + d.horizon; wend = w < wend ? w : wend; }
        acc.pend_min = pos_inf(); acc.emit_min = pos_inf();
        q.gmin[lane] = ~0ull;
        __syncwarp();
        const uint32_t step = step0 + rounds;
        lp_step<M, false>(d, acc, q, l, gvt, wend, step);                 // every group of the warp takes part in every round
        q.drain(d, acc, step & 1u);
        __syncwarp();
        // the group's GVT: minimum over its lanes of what stays pending, and over the messages its LPs sent
        const unsigned long long m = ordered_bits(acc.pend_min);
        unsigned long long gm = mine ? q.gmin[lane / G] : ~0ull;
#pragma unroll
        for (uint32_t k = 0; k < G; ++k) {
            const unsigned long long t = __shfl_sync(0xffffffffu, m, (int)((base_lane + k) & 31u));
            gm = t < gm ? t : gm;
        }
        ++rounds;
        if (mine) {
            const double g = from_ordered(gm);
            // a group that does not move for a very long time awaits a rollback that never comes (Internal/Event.hs:809-812)
            stalled = (g == graw && g <= d.stop && !(g > cap)) ? stalled + 1u : 0u;
            if (stalled > 200000u) acc.fail(TW_ERR_DEADLOCK, d.lp_base + l);
            graw = g;
        }
        if (__any_sync(0xffffffffu, acc.err != 0u)) break;
    }
    graw_out = graw; rounds_out = rounds;
    return mine;
}
#else
template <class M>   // host pass of nvcc / scalar emulation: never called
inline bool group_rounds(const Dev<M> &, Acc &, WorkQ<M> &, uint32_t, uint32_t, uint32_t, uint32_t, double, double &g, uint32_t &r) { g = 0.0; r = 0; return false; }
#endif

// Host-injected / cross-GPU routed message -> destination inbox (thread per message).
template <class M>
TW_HD void route_event(const Dev<M> &d, Acc &acc, uint32_t parity, const tw_event &w)
{
    if (w.dst >= d.n_global) { acc.fail(TW_ERR_BAD_DESTINATION, w.dst); return; }
    const uint32_t owner = d.world > 1 ? w.dst / d.n_per : 0u;
    if (owner != d.rank) return;                                  // another partition's LP
    if (w.recv_ts < w.send_ts) { acc.fail(TW_ERR_SEND_AFTER_RECV, w.src); return; }
    const uint32_t l = w.dst - d.lp_base;
    const typename M::Event ev = M::from_wire(w);
    const uint32_t idx = in_total(atomic_add_u64(&d.in[l].word, in_add(parity)));
    d.pages[(size_t)l * d.cap + (idx & d.mask)] = ev;
    acc.emit_min = dmin(acc.emit_min, ev.recv_ts);
}

}  // namespace twb
