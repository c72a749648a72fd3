// dio.hpp -- C++ host-side mirror of the reference's DIO surface over the C ABI (include/twb200.h).
//
// The reference is compiled Haskell; GHC is not available in this image, so the host side above the C ABI is
// provided in C++ (header-only) with the reference's own names, argument meaning and error behaviour
// (paths relative to /root/reference/Simulation/Aivika/Distributed/Optimistic/):
//   DIOParams, defaultDIOParams      Internal/DIO.hs:70-110, 196-217
//   Specs                            aivika-transformers; tests/MachRep2DistributedReproducible.hs:64-68
//   runDIO                           Internal/DIO.hs:327-328
//   enqueueMessage / sendMessage     Message.hs:41-69
//   timeServer, TimeServerParams     Internal/TimeServer.hs:47-72, 485
//   LogicalProcessState              State.hs:27-54
//   SimulationAbort                  Internal/Event.hs:476-484 (`error` / SimulationAbort become an exception)
// The Haskell twin (foreign import ccall) is in ../hs/ and INTEGRATION.md.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/twb200.h"

namespace aivika::distributed {

struct SimulationAbort : std::runtime_error {
    int code;
    SimulationAbort(int c, const std::string &m) : std::runtime_error(std::string(tw_strerror(c)) + ": " + m), code(c) {}
};

struct DIOParams {
    std::optional<double> dioTimeHorizon;                 // Nothing = unbounded optimism
    int dioUndoableLogSizeThreshold = 10000000;
    int dioOutputMessageQueueSizeThreshold = 10000;
    int dioTransientMessageQueueSizeThreshold = 5000;
    int dioSyncTimeout = 60000000;
    // engine knobs without a reference counterpart
    uint32_t queueCapacity = 128, maxEventsPerStep = 4, stepsPerLaunch = 16, procLimit = 0, fossilThreshold = 0;
};
inline const DIOParams defaultDIOParams{};

struct Specs { double spcStartTime = 0.0, spcStopTime = 1000.0, spcDT = 1.0; };

struct TimeServerParams { int tsTimeSyncDelay = 100000; };
inline const TimeServerParams defaultTimeServerParams{};

struct PHOLD {
    uint32_t n_lps; uint32_t population = 16, remote_pct = 10; double lookahead = 0.1, mean = 1.0; uint64_t seed = 1;
    uint32_t n_hot = 0, hot_pct = 0;
    void fill(tw_params &p) const
    {
        p.model = TW_MODEL_PHOLD; p.n_lps = n_lps; p.phold = tw_phold_params{population, remote_pct, n_hot, hot_pct, lookahead, mean, seed};
    }
};
struct MachRep2 {          // tests/MachRep2DistributedReproducible.hs, replicated
    uint32_t n_replicas = 1; double meanUpTime = 1.0, meanRepairTime = 0.5, delta = 1e-6; int64_t seed1 = 456, seed2 = 789, seed_stride = 1000;
    void fill(tw_params &p) const
    {
        p.model = TW_MODEL_MACHREP2; p.n_lps = 3 * n_replicas; p.machrep = tw_machrep_params{meanUpTime, meanRepairTime, delta, seed1, seed2, seed_stride};
    }
};
struct MachRep1 {          // tests/MachRep1.hs (two repair persons), replicated; third model of the handler table
    uint32_t n_replicas = 1; double meanUpTime = 1.0, meanRepairTime = 0.5; int64_t seed1 = 456, seed2 = 789, seed_stride = 1000;
    void fill(tw_params &p) const
    {
        p.model = TW_MODEL_MACHREP1; p.n_lps = 3 * n_replicas; p.machrep = tw_machrep_params{meanUpTime, meanRepairTime, 0.0, seed1, seed2, seed_stride};
    }
};

struct LogicalProcessState {
    uint32_t lpStateId; double lpStateLocalTime, lpStateEventQueueTime; uint32_t lpStateEventQueueSize, lpStateLogSize,
        lpStateInputMessageCount, lpStateOutputMessageCount; uint64_t lpStateRollbackCount;
};

class DIO;
// The Time Server role: quorum + global virtual time (reduced on the device every super-step).
class TimeServer {
  public:
    explicit TimeServer(int quorum, TimeServerParams ps = defaultTimeServerParams) : quorum_(quorum), params_(ps) {}
    std::optional<double> globalTime() const;
    int quorum() const { return quorum_; }
  private:
    friend class DIO;
    int quorum_; TimeServerParams params_; DIO *dio_ = nullptr;
};
inline TimeServer timeServer(int n, TimeServerParams ps = defaultTimeServerParams) { return TimeServer(n, ps); }

// A batch of logical processes on one GPU partition.
class DIO {
  public:
    template <class Model>
    DIO(const Model &m, const DIOParams &ps, const Specs &specs, uint32_t rank = 0, uint32_t world = 1, int device = -1,
        const uint8_t *nccl_id = nullptr)
    {
        tw_params p{};
        m.fill(p);
        p.rank = rank; p.world = world; p.device = device;
        p.queue_capacity = ps.queueCapacity; p.max_events_per_step = ps.maxEventsPerStep; p.steps_per_launch = ps.stepsPerLaunch;
        p.proc_limit = ps.procLimit; p.fossil_threshold = ps.fossilThreshold;
        p.start_time = specs.spcStartTime; p.stop_time = specs.spcStopTime;
        p.time_horizon = ps.dioTimeHorizon ? *ps.dioTimeHorizon : -1.0;
        if (nccl_id) for (int i = 0; i < 128; ++i) p.nccl_id[i] = nccl_id[i];
        params_ = p;
        int rc = tw_create(&p, &h_);
        if (rc != TW_OK) throw SimulationAbort(rc, tw_last_error(nullptr));
        tw_partition(h_, &lp_base_, &n_local_);
    }
    DIO(const DIO &) = delete;
    DIO &operator=(const DIO &) = delete;
    ~DIO() { if (h_) tw_destroy(h_); }

    std::vector<tw_event> initialEvents() const
    {
        const size_t per = params_.model == TW_MODEL_PHOLD ? params_.phold.population : 1;
        std::vector<tw_event> ev(n_local_ * per + 1);
        size_t n = 0;
        check(tw_model_initial_events_range(&params_, lp_base_, lp_base_ + n_local_, ev.data(), ev.size(), &n));
        ev.resize(n);
        return ev;
    }
    void enqueueMessages(const std::vector<tw_event> &ev) { check(tw_enqueue_messages(h_, ev.data(), ev.size())); }
    // runEventInStartTime on every LP, on the device (== enqueueMessages(initialEvents()) without the host buffer)
    void runEventInStartTime() { check(tw_populate(h_)); }
    // enqueueMessage pid t a
    void enqueueMessage(uint32_t pid, double t, uint32_t src, double send_time, uint64_t ctr = 0, uint32_t type = 0, double value = 0.0)
    {
        tw_event e{t, send_time, ctr, src, pid, type, 0, value};
        check(tw_enqueue_message(h_, &e));
    }
    // sendMessage pid a = enqueueMessage pid now a
    void sendMessage(uint32_t pid, uint32_t src, uint32_t type = 0, double value = 0.0) { const double now = globalTime(); enqueueMessage(pid, now, src, now, 0, type, value); }
    tw_run_stats runEventInStopTime(std::optional<double> until = std::nullopt)
    {
        tw_run_stats st{};
        check(tw_run(h_, until ? *until : params_.stop_time, &st));
        return st;
    }
    double globalTime() const { double g = 0; check(tw_gvt(h_, &g)); return g; }
    // the LP side of the Time Server protocol (Internal/TimeServer.hs): registerDIO, LocalTimeMessage, GlobalTimeMessage
    void attachTimeServer(bool on = true) { check(tw_attach_time_server(h_, on ? 1 : 0)); }
    double localTime() const { double t = 0; check(tw_local_time(h_, &t)); return t; }
    bool globalTimeMessage(double t) { int ok = 0; check(tw_global_time_message(h_, t, &ok)); return ok != 0; }
    LogicalProcessState logicalProcessState(uint32_t lp) const
    {
        tw_lp_state s{}; check(tw_lp_state_get(h_, lp, &s));
        return {s.lp, s.local_time, s.event_queue_time, s.event_queue_size, s.log_size, s.input_message_count, s.output_message_count, s.rollback_count};
    }
    void lpStats(std::vector<uint64_t> &n_exec, std::vector<uint64_t> &hash) const
    {
        n_exec.assign(n_local_, 0); hash.assign(n_local_, 0);
        check(tw_read_lp_stats(h_, n_exec.data(), hash.data(), n_local_));
    }
    std::vector<tw_machrep_result> machRepResults() const
    {
        const uint32_t first = ((lp_base_ + 2) / 3) * 3, end = lp_base_ + n_local_;
        std::vector<tw_machrep_result> r(first < end ? (end - first + 2) / 3 : 0);
        if (!r.empty()) check(tw_read_machrep(h_, r.data(), r.size()));
        return r;
    }
    void reset() { check(tw_reset(h_)); }
    uint32_t lpBase() const { return lp_base_; }
    uint32_t localCount() const { return n_local_; }
    void attach(TimeServer &ts) { ts.dio_ = this; }

  private:
    void check(int rc) const { if (rc != TW_OK) throw SimulationAbort(rc, tw_last_error(h_)); }
    tw_engine *h_ = nullptr; tw_params params_{}; uint32_t lp_base_ = 0, n_local_ = 0;
};

inline std::optional<double> TimeServer::globalTime() const { if (!dio_) return std::nullopt; return dio_->globalTime(); }

// runDIO m ps timeServerId: returns the LP batch; `simulation` loads the start-time messages and runs to the stop time.
template <class Model>
inline std::pair<std::unique_ptr<DIO>, tw_run_stats> runDIO(const Model &m, const DIOParams &ps, TimeServer &ts, const Specs &specs)
{
    auto dio = std::make_unique<DIO>(m, ps, specs);
    dio->attach(ts);
    dio->enqueueMessages(dio->initialEvents());
    tw_run_stats st = dio->runEventInStopTime();
    return {std::move(dio), st};
}

}  // namespace aivika::distributed
