// tw_dsl.cuh -- building blocks for compiled model handlers (device + host): SURVEY.md 8(f) rank 3.
//
// A model written against the reference's API uses, per event, a handful of primitives of aivika-transformers /
// aivika-distributed.  Their device forms live here so that a new model is a struct with a State, an Event payload and
// one `handle` function composed from them (tw_models.cuh: MachRepModel and MachRep1Model are written this way):
//
//   Outbox          enqueueMessage / sendMessage        Simulation/Aivika/Distributed/Optimistic/Message.hs:41-69
//                   (send time = current time; sendMessage => receive time = current time; the send counter comes
//                   from the LP's saved state, so identities are rollback-invariant)
//   FcfsResource    newFCFSResource / requestResource / releaseResourceWithinEvent / resourceCount of
//                   aivika-transformers with the FCFS wait list of Optimistic/QueueStrategy.hs:24-45; over-release is
//                   the SimulationRetry of Internal/Event.hs:247-249, 768-813
//   random_exponential   randomRefExponential over an undoable StdGen Ref (tests/MachRep2Reproducible.hs:58-67,
//                   tests/MachRep2DistributedReproducible.hs:81-86); Generator.hs:52: -log x * mu
// Undoable Refs (Optimistic/Ref/Base/Strict.hs:25-48) need no primitive: every field of a model's State is one, because
// the engine saves or reverses the whole State per executed event.
#pragma once
#include "tw_math.cuh"

namespace twb {

enum HandlerStatus { H_OK = 0, H_RETRY = 1, H_BAD = 2 };

namespace dsl {

template <class Event>
struct Outbox {
    Event *out; int &n; uint64_t &ctr; uint32_t self; double now;
    TW_HD Outbox(Event *o, int &n_, uint64_t &ctr_, uint32_t self_, double now_) : out(o), n(n_), ctr(ctr_), self(self_), now(now_) {}
    // enqueueMessage pid t a: the sender chooses the receive time
    TW_HD void enqueueMessage(uint32_t to, double at, uint32_t type, double value)
    {
        Event &o = out[n++];
        o.recv_ts = at; o.send_ts = now; o.ctr = ctr++; o.src = self; o.dst = to; o.type = type; o.pad = 0; o.value = value;
    }
    // sendMessage pid a = enqueueMessage pid now a
    TW_HD void sendMessage(uint32_t to, uint32_t type, double value) { enqueueMessage(to, now, type, value); }
};

template <int N>
struct FcfsResource {                      // 4 + N + 1 bytes
    int32_t count;                         // resourceCount
    uint8_t wait[N];                       // FCFS wait list (who is waiting, in arrival order)
    uint8_t wait_len;
    TW_HD void init(int32_t c) { count = c; wait_len = 0; for (int i = 0; i < N; ++i) wait[i] = 0; }
    // requestResource: granted at once (count > 0) or queued.  H_RETRY: the wait list is full, which only a stale
    // optimistic state can cause.
    TW_HD int request(uint8_t who, bool &granted)
    {
        granted = false;
        if (count == 0) {
            if (wait_len >= N) return H_RETRY;
            wait[wait_len++] = who;
        } else { count -= 1; granted = true; }
        return H_OK;
    }
    // releaseResourceWithinEvent: count' = count + 1 > max => SimulationRetry; otherwise hand the unit to the first
    // waiter (count unchanged) or store count'.
    TW_HD int release(int32_t max_count, bool &resumed, uint8_t &waiter)
    {
        resumed = false; waiter = 0;
        const int32_t c1 = count + 1;
        if (c1 > max_count) return H_RETRY;
        if (wait_len == 0) count = c1;
        else {
            waiter = wait[0]; resumed = true;
            for (int i = 0; i + 1 < N; ++i) wait[i] = wait[i + 1];
            wait[N - 1] = 0; wait_len--;
        }
        return H_OK;
    }
};

TW_HD double random_exponential(StdGen &g, double mu) { return expo_variate(std_random_double(g), mu); }

}  // namespace dsl
}  // namespace twb
