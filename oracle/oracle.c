/*
 * oracle.c -- TEST INFRASTRUCTURE.  Sequential discrete-event ground truth.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (libtwb200.so) never does.
 *
 * What it restates.  A Time Warp simulation must commit exactly the schedule a
 * sequential simulator would execute (the reference enforces this through rollback:
 * Simulation/Aivika/Distributed/Optimistic/Internal/InputMessageQueue.hs:120-162 and
 * Internal/Event.hs:196-250).  This file is that sequential simulator for the two
 * model families of BASELINE.json:
 *   - the machine-repair model of tests/MachRep2DistributedReproducible.hs:125-229
 *     (slaveModel / masterModel, constants :61-74), replicated R times;
 *   - the machine-repair model of tests/MachRep1.hs:52-93 (two repair persons, no contention; the `machine`
 *     process as three self-scheduled events), replicated R times;
 *   - PHOLD (not in the reference; defined in SURVEY.md App. C / DESIGN.md) written
 *     against the reference's message API semantics (Optimistic/Message.hs:48-69:
 *     send time = current time, receive time chosen by the sender).
 *
 * Total order of events at one logical process: (recv_ts, send_ts, src_lp, src_ctr),
 * consistent with the reference wherever the reference is deterministic (same sender =>
 * sequence-number order, Internal/TransientMessageQueue.hs:63-78).
 * Stop time is inclusive (Internal/Event.hs:228-229).
 *
 * PARITY UNPINNED: the reference has no golden vectors for this path and cannot be
 * built in this image (no GHC); see DESIGN.md.  The only reference-pinned check is
 * "about 0.6 and 0.67" (tests/MachRep2DistributedReproducible.hs:19-22).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle.c -o liboracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_math.h"

/* ------------------------------------------------------------------ events + heap */

typedef struct {
    double   recv_ts;
    double   send_ts;
    uint64_t ctr;
    uint32_t src;
    uint32_t dst;
    uint32_t type;
    uint32_t pad;
    double   value;
} or_event;

static inline int ev_less(const or_event *a, const or_event *b)
{
    if (a->recv_ts != b->recv_ts) return a->recv_ts < b->recv_ts;
    if (a->send_ts != b->send_ts) return a->send_ts < b->send_ts;
    if (a->src != b->src) return a->src < b->src;
    if (a->ctr != b->ctr) return a->ctr < b->ctr;
    return a->dst < b->dst;
}

typedef struct { or_event *a; size_t n, cap; } or_heap;

static int heap_push(or_heap *h, const or_event *e)
{
    if (h->n == h->cap) {
        size_t nc = h->cap ? h->cap * 2 : 1024;
        or_event *na = (or_event *)realloc(h->a, nc * sizeof(or_event));
        if (!na) return -1;
        h->a = na; h->cap = nc;
    }
    size_t i = h->n++;
    while (i > 0) {
        size_t p = (i - 1) >> 1;
        if (!ev_less(e, &h->a[p])) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = *e;
    return 0;
}

static void heap_pop(or_heap *h, or_event *out)
{
    *out = h->a[0];
    or_event last = h->a[--h->n];
    size_t i = 0, n = h->n;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && ev_less(&h->a[c + 1], &h->a[c])) c++;
        if (!ev_less(&h->a[c], &last)) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (n) h->a[i] = last;
}

/* optional committed-trace sink: (dst, src, ctr, recv bits) per executed event */
typedef struct {
    uint32_t dst, src;
    uint64_t ctr;
    uint64_t recv_bits;
} or_trace_rec;

/* ------------------------------------------------------------------ PHOLD */

typedef struct {
    uint32_t n_lps;       /* N */
    uint32_t population;  /* P initial events per LP */
    double   lookahead;   /* L */
    double   mean;        /* mean of the exponential part of the delay */
    double   stop_time;   /* T, inclusive */
    uint32_t remote_pct;  /* 0..100 */
    uint32_t n_hot;       /* size of the hot set (0 = uniform remote targets) */
    uint32_t hot_pct;     /* share of remote messages aimed at the hot set, 0..100 */
    uint32_t pad;
    uint64_t seed;
} or_phold_params;

typedef struct {
    uint64_t committed;   /* executed events (sequential => all committed) */
    uint64_t hash_sum;    /* sum over LPs of the per-LP trace hash (wrapping) */
    uint64_t remote;      /* events whose destination differs from the source */
    double   last_time;   /* time of the last executed event */
} or_summary;

/* One PHOLD event execution at LP i, state (rng_ctr, send_ctr); returns the new event. */
static inline void phold_handle(const or_phold_params *p, uint32_t i, uint64_t lpkey,
                                uint64_t *rng_ctr, uint64_t *send_ctr, double now, or_event *out)
{
    uint64_t d1 = or_draw(lpkey, *rng_ctr + 0);
    uint64_t d2 = or_draw(lpkey, *rng_ctr + 1);
    uint64_t d3 = or_draw(lpkey, *rng_ctr + 2);
    uint64_t d4 = or_draw(lpkey, *rng_ctr + 3);
    double delay = p->lookahead + or_expo(or_u01(d1), p->mean);
    uint32_t dst = i;
    uint64_t N = p->n_lps;
    if (N > 1 && (((d2 >> 32) * 100ULL) >> 32) < p->remote_pct) {
        if (p->n_hot > 0 && (((d4 >> 32) * 100ULL) >> 32) < p->hot_pct) {
            uint64_t k = or_mulhi(d3, p->n_hot);
            dst = (uint32_t)((k * N) / p->n_hot);
        } else {
            uint64_t k = or_mulhi(d3, N - 1);
            dst = (uint32_t)(k + (k >= i ? 1 : 0));
        }
    }
    out->recv_ts = now + delay;
    out->send_ts = now;
    out->ctr = *send_ctr;
    out->src = i;
    out->dst = dst;
    out->type = 0; out->pad = 0; out->value = 0.0;
    *rng_ctr += 4;
    *send_ctr += 1;
}

/* Fill `out` (n_lps * population records) with the initial event population. */
int oracle_phold_initial_events(const or_phold_params *p, or_event *out)
{
    for (uint32_t i = 0; i < p->n_lps; ++i) {
        uint64_t key = or_lpkey(p->seed, i);
        for (uint32_t j = 0; j < p->population; ++j) {
            or_event *e = &out[(size_t)i * p->population + j];
            e->recv_ts = or_expo(or_u01(or_draw(key, j)), p->mean);
            e->send_ts = 0.0;
            e->ctr = j; e->src = i; e->dst = i;
            e->type = 0; e->pad = 0; e->value = 0.0;
        }
    }
    return 0;
}

/*
 * Sequential PHOLD run.  n_exec[i] / hash[i]: per-LP executed-event count and
 * order-dependent trace hash.  trace (optional, capacity trace_cap) receives the
 * global committed sequence; *trace_len is the number of executed events.
 */
int oracle_phold_run(const or_phold_params *p, uint64_t *n_exec, uint64_t *hash,
                     or_summary *sum, or_trace_rec *trace, uint64_t trace_cap, uint64_t *trace_len)
{
    uint32_t N = p->n_lps;
    or_heap h = {0, 0, 0};
    uint64_t *rng = (uint64_t *)malloc(sizeof(uint64_t) * N);
    uint64_t *snd = (uint64_t *)malloc(sizeof(uint64_t) * N);
    uint64_t *key = (uint64_t *)malloc(sizeof(uint64_t) * N);
    if (!rng || !snd || !key) return -1;
    for (uint32_t i = 0; i < N; ++i) {
        key[i] = or_lpkey(p->seed, i);
        rng[i] = p->population; snd[i] = p->population;
        n_exec[i] = 0; hash[i] = or_fin(key[i] + 1ULL);
        for (uint32_t j = 0; j < p->population; ++j) {
            or_event e;
            e.recv_ts = or_expo(or_u01(or_draw(key[i], j)), p->mean);
            e.send_ts = 0.0; e.ctr = j; e.src = i; e.dst = i; e.type = 0; e.pad = 0; e.value = 0.0;
            if (heap_push(&h, &e)) return -1;
        }
    }
    uint64_t count = 0, remote = 0; double last = 0.0;
    while (h.n > 0 && h.a[0].recv_ts <= p->stop_time) {
        or_event e, ne;
        heap_pop(&h, &e);
        uint32_t i = e.dst;
        if (trace && count < trace_cap) {
            trace[count].dst = i; trace[count].src = e.src; trace[count].ctr = e.ctr;
            trace[count].recv_bits = or_bits(e.recv_ts);
        }
        hash[i] = or_trace(hash[i], e.recv_ts, e.src, e.ctr);
        n_exec[i] += 1;
        phold_handle(p, i, key[i], &rng[i], &snd[i], e.recv_ts, &ne);
        if (ne.dst != i) remote++;
        if (heap_push(&h, &ne)) return -1;
        last = e.recv_ts;
        count++;
    }
    if (sum) {
        sum->committed = count; sum->remote = remote; sum->last_time = last;
        uint64_t s = 0; for (uint32_t i = 0; i < N; ++i) s += hash[i];
        sum->hash_sum = s;
    }
    if (trace_len) *trace_len = count;
    free(rng); free(snd); free(key); free(h.a);
    return 0;
}

/* ------------------------------------------------------------------ MachRep2 */

enum {
    MR_START = 0,
    MR_TOTAL_UP_TIME_CHANGE = 1, MR_NREP_CHANGE = 2, MR_NIMMED_REP_CHANGE = 3,
    MR_REPAIR_PERSON_COUNT = 4, MR_REQUEST_REPAIR_PERSON = 5, MR_RELEASE_REPAIR_PERSON = 6,
    MR_TOTAL_UP_TIME_CHANGE_RESP = 7, MR_NREP_CHANGE_RESP = 8, MR_NIMMED_REP_CHANGE_RESP = 9,
    MR_REPAIR_PERSON_COUNT_RESP = 10, MR_REQUEST_REPAIR_PERSON_RESP = 11, MR_RELEASE_REPAIR_PERSON_RESP = 12
};

typedef struct {
    uint32_t n_replicas;       /* R groups of (master, slave 1, slave 2) */
    uint32_t pad;
    double   mean_up_time;     /* 1.0  tests/MachRep2DistributedReproducible.hs:61 */
    double   mean_repair_time; /* 0.5  :62 */
    double   delta;            /* 1e-6 :71 */
    double   stop_time;        /* 1000 :65 */
    int64_t  seed1, seed2;     /* 456, 789 :74 (replica 0) */
    int64_t  seed_stride;      /* replica r uses seed + r * stride */
} or_machrep_params;

typedef struct {
    double   total_up_time;
    int64_t  n_rep;
    int64_t  n_immed_rep;
    double   up_time_prop;     /* totalUpTime / (2 * t)   :214-217 */
    double   immed_prop;       /* nImmedRep / nRep        :219-224 */
    uint64_t n_exec[3];        /* executed messages per LP (START excluded) */
    uint64_t hash[3];
} or_machrep_result;

typedef struct {
    /* master (masterModel :165-229) */
    double  total_up_time; int64_t n_rep, n_immed; int32_t count; /* resourceCount, max 1 */
    uint32_t wait[4]; uint32_t wait_len;
    /* slaves (slaveModel :125-162) */
    or_stdgen gen[2];
    uint64_t send_ctr[3];
    uint64_t n_exec[3], hash[3];
} mr_replica;

static int mr_send(or_heap *h, mr_replica *s, uint32_t base, uint32_t from, uint32_t to,
                   double now, double recv, uint32_t type, double value)
{
    or_event e;
    e.recv_ts = recv; e.send_ts = now; e.ctr = s->send_ctr[from]++; e.src = base + from; e.dst = base + to;
    e.type = type; e.pad = 0; e.value = value;
    return heap_push(h, &e);
}

/* `machine` :130-135 */
static int mr_machine(or_heap *h, mr_replica *s, const or_machrep_params *p, uint32_t base, uint32_t me, double t)
{
    double x = or_stdrandom_double(&s->gen[me - 1]);
    double up = or_expo(x, p->mean_up_time);
    double at = (t + p->delta) + up;
    if (mr_send(h, s, base, me, 0, t, at, MR_TOTAL_UP_TIME_CHANGE, up)) return -1;
    if (mr_send(h, s, base, me, 0, t, at, MR_NREP_CHANGE, 1.0)) return -1;
    if (mr_send(h, s, base, me, 0, t, at, MR_REPAIR_PERSON_COUNT, 0.0)) return -1;
    return 0;
}

static int machrep_one(const or_machrep_params *p, uint32_t r, or_machrep_result *res, or_heap *h,
                       or_trace_rec *trace, uint64_t trace_cap, uint64_t *trace_len)
{
    mr_replica s; memset(&s, 0, sizeof s);
    uint32_t base = 3 * r;
    s.count = 1;
    s.gen[0] = or_mkstdgen(p->seed1 + (int64_t)r * p->seed_stride);
    s.gen[1] = or_mkstdgen(p->seed2 + (int64_t)r * p->seed_stride);
    for (int k = 0; k < 3; ++k) s.hash[k] = or_fin(or_lpkey(0, base + k) + 1ULL);
    h->n = 0;
    for (uint32_t me = 1; me <= 2; ++me) {      /* runEventInStartTime machine  :154 */
        or_event e; e.recv_ts = 0.0; e.send_ts = 0.0; e.ctr = s.send_ctr[me]++; e.src = base + me; e.dst = base + me;
        e.type = MR_START; e.pad = 0; e.value = 0.0;
        if (heap_push(h, &e)) return -1;
    }
    uint64_t tl = trace_len ? *trace_len : 0;
    while (h->n > 0 && h->a[0].recv_ts <= p->stop_time) {
        or_event e; heap_pop(h, &e);
        uint32_t me = e.dst - base, from = e.src - base;
        double t = e.recv_ts, td = t + p->delta;
        if (trace) {
            if (tl < trace_cap) { trace[tl].dst = e.dst; trace[tl].src = e.src; trace[tl].ctr = e.ctr; trace[tl].recv_bits = or_bits(t); }
            tl++;
        }
        s.hash[me] = or_trace(s.hash[me], t, e.src, e.ctr);
        if (e.type != MR_START) s.n_exec[me]++;
        int rc = 0;
        switch (e.type) {
        case MR_START:
        case MR_RELEASE_REPAIR_PERSON_RESP:               /* :150-154 */
            rc = mr_machine(h, &s, p, base, me, t); break;
        case MR_REPAIR_PERSON_COUNT_RESP:                 /* :137-142 */
            if ((int64_t)e.value == 1) rc = mr_send(h, &s, base, me, 0, t, td, MR_NIMMED_REP_CHANGE, 1.0);
            if (!rc) rc = mr_send(h, &s, base, me, 0, t, td, MR_REQUEST_REPAIR_PERSON, 0.0);
            break;
        case MR_REQUEST_REPAIR_PERSON_RESP: {             /* :144-148 */
            double x = or_stdrandom_double(&s.gen[me - 1]);
            double rt = or_expo(x, p->mean_repair_time);
            rc = mr_send(h, &s, base, me, 0, t, td + rt, MR_RELEASE_REPAIR_PERSON, 0.0);
            break; }
        case MR_TOTAL_UP_TIME_CHANGE:                     /* :176-180 */
            s.total_up_time += e.value;
            rc = mr_send(h, &s, base, 0, from, t, td, MR_TOTAL_UP_TIME_CHANGE_RESP, 0.0); break;
        case MR_NREP_CHANGE:                              /* :182-186 */
            s.n_rep += (int64_t)e.value;
            rc = mr_send(h, &s, base, 0, from, t, td, MR_NREP_CHANGE_RESP, 0.0); break;
        case MR_NIMMED_REP_CHANGE:                        /* :188-192 */
            s.n_immed += (int64_t)e.value;
            rc = mr_send(h, &s, base, 0, from, t, td, MR_NIMMED_REP_CHANGE_RESP, 0.0); break;
        case MR_REPAIR_PERSON_COUNT:                      /* :194-198 */
            rc = mr_send(h, &s, base, 0, from, t, td, MR_REPAIR_PERSON_COUNT_RESP, (double)s.count); break;
        case MR_REQUEST_REPAIR_PERSON:                    /* :200-206 requestResource */
            if (s.count == 0) { if (s.wait_len >= 4) return -2; s.wait[s.wait_len++] = from; }
            else { s.count -= 1; rc = mr_send(h, &s, base, 0, from, t, td, MR_REQUEST_REPAIR_PERSON_RESP, 0.0); }
            break;
        case MR_RELEASE_REPAIR_PERSON: {                  /* :208-212 releaseResourceWithinEvent */
            int32_t c1 = s.count + 1;
            if (c1 > 1) return -3;                        /* sequential run can never over-release */
            uint32_t waiter = 0; int resumed = 0;
            if (s.wait_len == 0) s.count = c1;
            else { waiter = s.wait[0]; for (uint32_t k = 1; k < s.wait_len; ++k) s.wait[k - 1] = s.wait[k]; s.wait_len--; resumed = 1; }
            rc = mr_send(h, &s, base, 0, from, t, td, MR_RELEASE_REPAIR_PERSON_RESP, 0.0);
            if (!rc && resumed) rc = mr_send(h, &s, base, 0, waiter, t, td, MR_REQUEST_REPAIR_PERSON_RESP, 0.0);
            break; }
        default: break;                                    /* *_RESP with no subscriber */
        }
        if (rc) return -1;
    }
    if (trace_len) *trace_len = tl;
    res->total_up_time = s.total_up_time; res->n_rep = s.n_rep; res->n_immed_rep = s.n_immed;
    res->up_time_prop = s.total_up_time / (2.0 * p->stop_time);
    res->immed_prop = (double)s.n_immed / (double)s.n_rep;
    for (int k = 0; k < 3; ++k) { res->n_exec[k] = s.n_exec[k]; res->hash[k] = s.hash[k]; }
    return 0;
}

/* Run replicas [r0, r1); res has (r1 - r0) entries.  Replicas are closed groups. */
int oracle_machrep_run(const or_machrep_params *p, uint32_t r0, uint32_t r1, or_machrep_result *res,
                       or_trace_rec *trace, uint64_t trace_cap, uint64_t *trace_len)
{
    or_heap h = {0, 0, 0};
    if (trace_len) *trace_len = 0;
    for (uint32_t r = r0; r < r1; ++r) {
        int rc = machrep_one(p, r, &res[r - r0], &h, trace, trace_cap, trace_len);
        if (rc) { free(h.a); return rc; }
    }
    free(h.a);
    return 0;
}

/* ------------------------------------------------------------------ MachRep1 (tests/MachRep1.hs) */

enum { M1_START = 0, M1_UP_END = 1, M1_REPAIR_END = 2, M1_TOTAL_UP_TIME_CHANGE = 3 };

static int machrep1_one(const or_machrep_params *p, uint32_t r, or_machrep_result *res, or_heap *h)
{
    uint32_t base = 3 * r;
    double total_up_time = 0.0; int64_t n_msgs = 0;          /* masterModel :75-93 */
    or_stdgen gen[2];
    uint64_t send_ctr[3] = {0, 0, 0}, n_exec[3] = {0, 0, 0}, hash[3];
    gen[0] = or_mkstdgen(p->seed1 + (int64_t)r * p->seed_stride);
    gen[1] = or_mkstdgen(p->seed2 + (int64_t)r * p->seed_stride);
    for (int k = 0; k < 3; ++k) hash[k] = or_fin(or_lpkey(0, base + k) + 1ULL);
    h->n = 0;
    for (uint32_t me = 1; me <= 2; ++me) {                   /* runProcessInStartTime machine :68 */
        or_event e; e.recv_ts = 0.0; e.send_ts = 0.0; e.ctr = send_ctr[me]++; e.src = base + me; e.dst = base + me;
        e.type = M1_START; e.pad = 0; e.value = 0.0;
        if (heap_push(h, &e)) return -1;
    }
    while (h->n > 0 && h->a[0].recv_ts <= p->stop_time) {
        or_event e; heap_pop(h, &e);
        uint32_t me = e.dst - base;
        double t = e.recv_ts;
        hash[me] = or_trace(hash[me], t, e.src, e.ctr);
        if (e.type != M1_START) n_exec[me]++;
        or_event o; o.send_ts = t; o.src = base + me; o.pad = 0;
        switch (e.type) {
        case M1_START:
        case M1_REPAIR_END: {                                /* :55-58 upTime <- randomExponential meanUpTime; holdProcess upTime */
            double up = or_expo(or_stdrandom_double(&gen[me - 1]), p->mean_up_time);
            o.recv_ts = t + up; o.ctr = send_ctr[me]++; o.dst = base + me; o.type = M1_UP_END; o.value = up;
            if (heap_push(h, &o)) return -1;
            break; }
        case M1_UP_END: {                                    /* :60-66 sendMessage masterId (TotalUpTimeChange upTime); repair */
            o.recv_ts = t; o.ctr = send_ctr[me]++; o.dst = base; o.type = M1_TOTAL_UP_TIME_CHANGE; o.value = e.value;
            if (heap_push(h, &o)) return -1;
            double rt = or_expo(or_stdrandom_double(&gen[me - 1]), p->mean_repair_time);
            o.recv_ts = t + rt; o.ctr = send_ctr[me]++; o.dst = base + me; o.type = M1_REPAIR_END; o.value = 0.0;
            if (heap_push(h, &o)) return -1;
            break; }
        case M1_TOTAL_UP_TIME_CHANGE:                        /* :84-86 modifyRef totalUpTime (+ x) */
            if (me == 0) { total_up_time += e.value; n_msgs++; }
            break;
        default: break;
        }
    }
    res->total_up_time = total_up_time; res->n_rep = n_msgs; res->n_immed_rep = 0;
    res->up_time_prop = total_up_time / (2.0 * p->stop_time);      /* :88-91 with n = 2 */
    res->immed_prop = 0.0;
    for (int k = 0; k < 3; ++k) { res->n_exec[k] = n_exec[k]; res->hash[k] = hash[k]; }
    return 0;
}

int oracle_machrep1_run(const or_machrep_params *p, uint32_t r0, uint32_t r1, or_machrep_result *res)
{
    or_heap h = {0, 0, 0};
    for (uint32_t r = r0; r < r1; ++r) {
        int rc = machrep1_one(p, r, &res[r - r0], &h);
        if (rc) { free(h.a); return rc; }
    }
    free(h.a);
    return 0;
}

/* ------------------------------------------------------------------ scalar probes for tests */

double   oracle_log(double x) { return or_log(x); }
double   oracle_u01(uint64_t d) { return or_u01(d); }
uint64_t oracle_draw(uint64_t seed, uint32_t lp, uint64_t ctr) { return or_draw(or_lpkey(seed, lp), ctr); }
uint64_t oracle_trace(uint64_t h, double t, uint32_t src, uint64_t ctr) { return or_trace(h, t, src, ctr); }
void oracle_stdgen_doubles(int64_t seed, uint32_t n, double *out, int32_t *state)
{
    or_stdgen g = or_mkstdgen(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = or_stdrandom_double(&g);
    if (state) { state[0] = g.s1; state[1] = g.s2; }
}
void oracle_stdgen_next(int64_t seed, uint32_t n, int32_t *out)
{
    or_stdgen g = or_mkstdgen(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = or_stdnext(&g);
}
uint32_t oracle_sizeof_event(void) { return (uint32_t)sizeof(or_event); }
