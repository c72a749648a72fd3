/*
 * twb200.h -- C ABI of the B200-native Time Warp engine (libtwb200.so).
 *
 * This is the drop-in boundary for the logical-process hot path of
 * dsorokin/aivika-distributed.  The reference has no FFI (it is 100 % Haskell), so each
 * entry point below names the reference interface it stands in for; the Haskell
 * `foreign import ccall` stubs a maintainer would add are in INTEGRATION.md and
 * aivika-distributed_b200/hs/.  All paths are relative to /root/reference/.
 *
 * Conventions: every function returns an int status (TW_OK == 0); the engine handle is
 * single-owner and not thread-safe (one LP is mutated by exactly one process in the
 * reference too: Simulation/Aivika/Distributed/Optimistic/Internal/DIO.hs:427-476 and the
 * queueBusy guard, Internal/Event.hs:199-208); pointer arguments are HOST pointers borrowed
 * for the duration of the call; outputs are caller-allocated.  There is no CPU fallback:
 * tw_create fails with TW_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef TWB200_H
#define TWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tw_engine tw_engine;

/* ---- status codes: the reference's `error` strings become distinct codes ---- */
enum {
    TW_OK = 0,
    TW_ERR_INVALID_ARG        = 1,
    TW_ERR_CUDA               = 2,   /* CUDA/NCCL failure or no usable device */
    TW_ERR_SEND_AFTER_RECV    = 3,   /* "The Send time cannot be greater than the Receive message time" Internal/OutputMessageQueue.hs:67-68 */
    TW_ERR_OUTDATED_MESSAGE   = 4,   /* "Received the outdated message" Internal/Event.hs:370-374 */
    TW_ERR_QUEUE_OVERFLOW     = 5,   /* fixed-capacity analogue of isEventOverflow, Internal/Event.hs:289-306 */
    TW_ERR_ANTI_UNMATCHED     = 6,   /* "Cannot annihilate another message" Internal/InputMessageQueue.hs:273-274 */
    TW_ERR_DEADLOCK           = 7,   /* "Detected a deadlock when retrying the computations" Internal/Event.hs:809-812 */
    TW_ERR_TIME_ORDER         = 8,   /* "The time value is too small" Internal/Event.hs:223-227 */
    TW_ERR_BAD_DESTINATION    = 9,   /* receiver id outside [0, n_lps) */
    TW_ERR_MODEL              = 10,  /* a model registered as closed groups sent a message outside the sender's group */
    TW_ERR_OUTBOX_OVERFLOW    = 11,  /* cross-GPU staging ring full: raise tw_params.outbox_capacity */
    TW_ERR_NOT_SUPPORTED      = 12,
    TW_ERR_ABORTED            = 13   /* AbortSimulationMessage received: SimulationAbort, Internal/Event.hs:476-484 */
};

enum { TW_MODEL_PHOLD = 1, TW_MODEL_MACHREP2 = 2, TW_MODEL_MACHREP1 = 3 };

/* tw_params.flags */
enum {
    TW_FLAG_SKIP_OUTDATED   = 1,     /* dioAllowSkippingOutdatedMessage (Internal/DIO.hs:96-98; reference default True, :206): a
                                        host-injected message older than GVT is counted and dropped instead of raising
                                        TW_ERR_OUTDATED_MESSAGE (Internal/Event.hs:370-374) */
    TW_FLAG_NO_GROUP_KERNEL = 2      /* closed-group models (MachRep1/2): one grid-wide super-step per launch like PHOLD, instead
                                        of the warp-local super-step loop (same results; for comparison and tests) */
};

/* MachRep2 message kinds == the message data types of tests/MachRep2DistributedReproducible.hs:88-104 */
enum {
    TW_MR_START = 0,
    TW_MR_TOTAL_UP_TIME_CHANGE = 1, TW_MR_NREP_CHANGE = 2, TW_MR_NIMMED_REP_CHANGE = 3,
    TW_MR_REPAIR_PERSON_COUNT = 4, TW_MR_REQUEST_REPAIR_PERSON = 5, TW_MR_RELEASE_REPAIR_PERSON = 6,
    TW_MR_TOTAL_UP_TIME_CHANGE_RESP = 7, TW_MR_NREP_CHANGE_RESP = 8, TW_MR_NIMMED_REP_CHANGE_RESP = 9,
    TW_MR_REPAIR_PERSON_COUNT_RESP = 10, TW_MR_REQUEST_REPAIR_PERSON_RESP = 11, TW_MR_RELEASE_REPAIR_PERSON_RESP = 12
};

/* MachRep1 event kinds (tests/MachRep1.hs:52-72): the machine process as self-scheduled events + TotalUpTimeChange (:46-50) */
enum { TW_M1_START = 0, TW_M1_UP_END = 1, TW_M1_REPAIR_END = 2, TW_M1_TOTAL_UP_TIME_CHANGE = 3 };

/*
 * One timestamped message at the boundary (48 bytes).  Replaces `Message`
 * (Internal/Message.hs:37-54): sequenceNo -> ctr, sendTime, receiveTime, senderId -> src,
 * receiverId -> dst, antiToggle -> bit 63 of ctr, messageData -> (type, value).
 */
typedef struct tw_event {
    double   recv_ts;
    double   send_ts;
    uint64_t ctr;
    uint32_t src;
    uint32_t dst;
    uint32_t type;
    uint32_t pad;
    double   value;
} tw_event;

typedef struct tw_phold_params {
    uint32_t population;   /* initial events per LP */
    uint32_t remote_pct;   /* 0..100 */
    uint32_t n_hot;        /* hot-set size for skewed fan-out, 0 = uniform */
    uint32_t hot_pct;      /* share of remote messages sent to the hot set */
    double   lookahead;
    double   mean;
    uint64_t seed;
} tw_phold_params;

/* constants of tests/MachRep2DistributedReproducible.hs:61-74 */
typedef struct tw_machrep_params {
    double  mean_up_time;      /* 1.0 */
    double  mean_repair_time;  /* 0.5 */
    double  delta;             /* 1e-6 */
    int64_t seed1, seed2;      /* 456, 789 */
    int64_t seed_stride;       /* replica r uses seed + r*stride */
} tw_machrep_params;

/*
 * Engine parameters.  Hot-path fields of DIOParams (Internal/DIO.hs:75-84, defaults :200-204)
 * plus the Specs start/stop times and the partition description.
 */
typedef struct tw_params {
    uint32_t model;               /* TW_MODEL_* */
    uint32_t n_lps;               /* global LP count (MachRep1/MachRep2: 3 * replicas) */
    uint32_t rank, world;         /* this partition / number of partitions (GPUs); LP i lives on i / ceil(n_lps/world) */
    int32_t  device;              /* CUDA device ordinal, -1 = current */
    uint32_t queue_capacity;      /* per-LP ring slots: a power of two, 8..32768 */
    uint32_t max_events_per_step; /* optimistic batch per LP per super-step, 1..64 (0 = 8) */
    uint32_t steps_per_launch;    /* super-steps per graph launch between host GVT checks */
    double   start_time;          /* spcStartTime */
    double   stop_time;           /* spcStopTime (inclusive, Internal/Event.hs:228-229) */
    double   time_horizon;        /* dioTimeHorizon; < 0 or NaN == Nothing (unbounded optimism) */
    uint32_t proc_limit;          /* processed-but-uncommitted entries per LP before optimism stalls (0 = capacity/4): the
                                     per-LP form of dioUndoableLogSizeThreshold / dioOutputMessageQueueSizeThreshold
                                     (isEventOverflow, Internal/Event.hs:289-306) */
    uint32_t fossil_threshold;    /* lazy fossil collection trigger (0 = capacity/8) */
    uint32_t outbox_capacity;     /* multi-GPU: records per staging ring (one ring per sender GPU and step parity in the
                                     receiver's memory); 0 = max(8192, 2 * LPs per GPU) */
    uint32_t flags;               /* TW_FLAG_* */
    tw_phold_params   phold;
    tw_machrep_params machrep;
    uint8_t  nccl_id[128];        /* ncclUniqueId shared by all ranks when world > 1 (tw_nccl_unique_id) */
    uint32_t sync_timeout_ms;     /* dioSyncTimeout (Internal/DIO.hs:88-89, default 60 s): how long a GPU waits for the GVT mailbox
                                     entry of a peer before the run fails with TW_ERR_DEADLOCK; 0 = 20000 */
    uint32_t reserved;
} tw_params;

/* Result of tw_run: counters the reference never kept next to its lpStateRollbackCount. */
typedef struct tw_run_stats {
    double   gvt;                 /* global virtual time at return */
    double   device_ms;           /* CUDA-event time of the run on this device */
    uint64_t steps;               /* super-steps executed */
    uint64_t committed;           /* committed handler executions on this partition */
    uint64_t executed;            /* executed incl. rolled back */
    uint64_t rolled_back;         /* executions undone */
    uint64_t rollbacks;           /* rollback episodes (lpStateRollbackCount summed) */
    uint64_t stragglers;
    uint64_t anti_sent;
    uint64_t annihilated;
    uint64_t retries;             /* SimulationRetry stalls */
    uint64_t sent_local, sent_remote;
    uint64_t kernel_launches;     /* launches of this library's kernels during the run */
    uint64_t trace_hash_sum;      /* sum over local LPs of the per-LP committed-trace hash */
} tw_run_stats;

/* Mirrors LogicalProcessState (Simulation/Aivika/Distributed/Optimistic/State.hs:27-54). */
typedef struct tw_lp_state {
    uint32_t lp;
    double   local_time;               /* lpStateLocalTime  */
    double   event_queue_time;         /* lpStateEventQueueTime (min pending) */
    uint32_t event_queue_size;         /* lpStateEventQueueSize (pending) */
    uint32_t log_size;                 /* lpStateLogSize (processed, uncommitted) */
    uint32_t input_message_count;      /* lpStateInputMessageCount */
    uint32_t output_message_count;     /* lpStateOutputMessageCount (regenerated by replay: processed entries) */
    uint64_t rollback_count;           /* lpStateRollbackCount: rollback episodes of this LP */
    uint64_t executed;                 /* committed handler executions of this LP */
    uint64_t trace_hash;
} tw_lp_state;

/* Per-replica MachRep2 result: the pair returned by masterModel (:214-229).  MachRep1 (tests/MachRep1.hs:88-91) fills
 * total_up_time, up_time_prop and n_rep (= TotalUpTimeChange messages received); the other fields are 0. */
typedef struct tw_machrep_result {
    double  total_up_time;
    int64_t n_rep;
    int64_t n_immed_rep;
    double  up_time_prop;
    double  immed_prop;
} tw_machrep_result;

/* runDIO / newEventQueue (Internal/DIO.hs:327-328, Internal/Event.hs:92-116): allocate all LP state on the device. */
int tw_create(const tw_params *params, tw_engine **out);
/* terminateDIO / unregisterDIO (Internal/DIO.hs:246-305). */
int tw_destroy(tw_engine *e);
/* Back to start_time with empty queues (a new runSimulation on the same LPs). */
int tw_reset(tw_engine *e);

/* enqueueMessage pid t a issued from runEventInStartTime (Optimistic/Message.hs:48-69): host-side injection. */
int tw_enqueue_message(tw_engine *e, const tw_event *ev);
/* QueueMessageBulk (Internal/Message.hs:113-116, Internal/Event.hs:382-405): n messages from a host buffer.
 * Messages whose dst is owned by another partition are ignored (each rank passes the same buffer or its own slice). */
int tw_enqueue_messages(tw_engine *e, const tw_event *evs, size_t n);

/* runEventInStartTime (the body every reference program runs on each LP at the start time, e.g. `machine` in
 * tests/MachRep2DistributedReproducible.hs:154, tests/MachRep1.hs:68; for PHOLD the initial population): every LP of this
 * partition creates its model's own start-time events ON THE DEVICE.  Same events as tw_model_initial_events_range +
 * tw_enqueue_messages, without the host buffer and the PCIe copy.  Collective over the ranks of a multi-GPU engine. */
int tw_populate(tw_engine *e);

/* runEventInStopTime (Internal/Event.hs:124-129, 743-752): advance until GVT passes until_time
 * (the reference returns when its GVT copy equals the stop time, :717-730). */
int tw_run(tw_engine *e, double until_time, tw_run_stats *stats);

/* GlobalTimeMessage (Internal/TimeServer.hs:384-403): current GVT. */
int tw_gvt(tw_engine *e, double *gvt);

/* ProvideLogicalProcessStateMessage / sendState (Internal/Event.hs:468-475, 910-942). lp is a global id owned by this rank. */
int tw_lp_state_get(tw_engine *e, uint32_t lp, tw_lp_state *out);

/* Per-LP committed-event counts and trace hashes for the LPs of this partition (n = local LP count). */
int tw_read_lp_stats(tw_engine *e, uint64_t *n_exec, uint64_t *trace_hash, size_t n);
/* MachRep1/MachRep2: results of the replicas whose master lives on this partition. */
int tw_read_machrep(tw_engine *e, tw_machrep_result *out, size_t n_replicas);

/*
 * Monitoring: dioSimulationMonitoringAction (Internal/DIO.hs:453-476) + TimeServerState
 * (Optimistic/State.hs:57-68).  The callback is invoked once per inspection of the device control block (every
 * steps_per_launch super-steps) on a worker thread owned by the engine -- never on the thread that launches the
 * super-steps --, in order; tw_run returns after the last invocation has finished.
 */
typedef struct tw_monitor_state {
    double   gvt;                 /* tsStateGlobalVirtualTime */
    uint64_t steps;               /* super-steps since tw_create/tw_reset */
    uint64_t executed, rolled_back, rollbacks, anti_sent, annihilated, sent_local, sent_remote;
    uint32_t n_lps_local;         /* tsStateLogicalProcesses (count on this partition) */
    uint32_t rank;
} tw_monitor_state;
typedef void (*tw_monitor_fn)(const tw_monitor_state *state, void *user);
int tw_set_monitor(tw_engine *e, tw_monitor_fn fn, void *user);

/*
 * The Time Server role across partitions: the LP side of the protocol of Internal/TimeServer.hs, so
 * that a device partition can sit under an external Time Server next to other partitions or CPU logical processes.
 *   tw_attach_time_server(e, 1)  registerDIO / RegisterLogicalProcessMessage (Internal/DIO.hs:267-283): from now on GVT
 *                                is bounded by the last accepted GlobalTimeMessage (initially the start time); resets the run
 *   tw_local_time                the answer to ComputeLocalTimeMessage: LocalTimeMessage pid t (Internal/Event.hs:425-436,
 *                                getLocalTime :496-514)
 *   tw_global_time_message       GlobalTimeMessage t -> updateGlobalTime (Internal/Event.hs:437-449, 528-540): a value above
 *                                the local time is ignored (*accepted = 0), otherwise fossil collection and termination use it
 * tw_run on an attached engine returns when GVT has passed until_time or sits at the bound (stats->gvt tells which).
 */
int tw_attach_time_server(tw_engine *e, int on);
int tw_local_time(tw_engine *e, double *local_time);
int tw_global_time_message(tw_engine *e, double global_time, int *accepted);

/* Messages dropped under TW_FLAG_SKIP_OUTDATED since tw_create / tw_reset (logOutdatedMessage, Internal/Event.hs:372). */
int tw_skipped_outdated(tw_engine *e, uint64_t *count);
/* Largest number of ring slots (retained entries + arrivals) an LP with a long pending region (> 256 entries: the hot LPs
 * of a skewed workload) had in use since tw_create / tw_reset; 0 when no LP ever held that many.  The fixed ring is this
 * engine's form of the memory thresholds of isEventOverflow (Internal/Event.hs:289-306): compare with queue_capacity. */
int tw_max_ring_occupancy(tw_engine *e, uint32_t *slots);

/* Partition geometry: first global LP id and LP count owned by this rank. */
int tw_partition(const tw_engine *e, uint32_t *lp_base, uint32_t *n_local);

/* Synthetic model inputs (host-side, bit-identical to what the device handler would draw):
 * PHOLD initial population (n_lps*population events) / MachRep2 START events (2 per replica).
 * Writes at most cap events, returns the count in *n. */
int tw_model_initial_events(const tw_params *params, tw_event *out, size_t cap, size_t *n);
/* Same, restricted to the start-time messages addressed to LPs lp_begin <= lp < lp_end (a rank's slice). */
int tw_model_initial_events_range(const tw_params *params, uint32_t lp_begin, uint32_t lp_end,
                                  tw_event *out, size_t cap, size_t *n);

/*
 * The reference's wire format at the boundary (csrc/tw_wire.cpp).  `Message`,
 * `LogicalProcessMessage` and `TimeServerMessage` (Internal/Message.hs:37-54, 113-137, 140-156) are encoded exactly as their
 * GHC.Generics `Binary` instances lay them out, so that a device partition can exchange bytes with Cloud Haskell processes:
 * receive QueueMessage / QueueMessageBulk from CPU logical processes (relayed by the inbox process, Internal/DIO.hs:366-370),
 * and sit under a stock `timeServer` (ComputeLocalTimeMessage in, LocalTimeMessage out, GlobalTimeMessage in).
 * Byte-level parity is UNPINNED against GHC (no toolchain here): see the header of tw_wire.cpp.
 */
typedef struct tw_wire_ctx tw_wire_ctx;
/* DP.ProcessId of one logical process: NodeId (the EndPointAddress bytes, e.g. "127.0.0.1:8088:0") + LocalProcessId. */
typedef struct tw_wire_pid {
    const uint8_t *node_address;
    uint32_t address_len;
    int32_t  lpid_unique, lpid_counter;
} tw_wire_pid;
/* How the (type, value) pair of a tw_event maps onto the Haskell message data type of that kind. */
enum {
    TW_WIRE_PAYLOAD_NONE           = 0,   /* constructor without fields */
    TW_WIRE_PAYLOAD_SRC_PID        = 1,   /* C DP.ProcessId           (tests/MachRep2DistributedReproducible.hs:89 ...Resp: the sender's id) */
    TW_WIRE_PAYLOAD_SRC_PID_DOUBLE = 2,   /* C (DP.ProcessId, Double) (:88 TotalUpTimeChange) */
    TW_WIRE_PAYLOAD_SRC_PID_INT    = 3,   /* C (DP.ProcessId, Int)    (:91 NRepChange, :94 NImmedRepChange, :98 RepairPersonCountResp) */
    TW_WIRE_PAYLOAD_DOUBLE         = 4,   /* C Double                 (tests/MachRep1.hs:46 TotalUpTimeChange) */
    TW_WIRE_PAYLOAD_INT            = 5,   /* C Int */
    TW_WIRE_PAYLOAD_SRC_INT_CTR    = 6    /* C Int Word64: sender LP index and send counter (hs/PHOLD.hs `Job`) */
};
typedef struct tw_wire_type {
    uint8_t  fingerprint[16];             /* encodeFingerprint (fingerprint (undefined :: T)): supplied by the Haskell side */
    uint32_t payload;                     /* TW_WIRE_PAYLOAD_* */
} tw_wire_type;
/* constructor indices of LogicalProcessMessage (Internal/Message.hs:113-137) and TimeServerMessage (:140-156) */
enum { TW_WIRE_QUEUE_MESSAGE = 0, TW_WIRE_QUEUE_MESSAGE_BULK = 1, TW_WIRE_COMPUTE_LOCAL_TIME = 4, TW_WIRE_GLOBAL_TIME = 5,
       TW_WIRE_ABORT_SIMULATION = 9 };
enum { TW_WIRE_TS_REGISTER = 0, TW_WIRE_TS_UNREGISTER = 1, TW_WIRE_TS_TERMINATE = 2, TW_WIRE_TS_COMPUTE_LOCAL_TIME_ACK = 3,
       TW_WIRE_TS_REQUEST_GLOBAL_TIME = 4, TW_WIRE_TS_LOCAL_TIME = 5, TW_WIRE_TS_PROVIDE_STATE = 6 };

/* pids[lp] for every LP id that can occur as sender or receiver; types[kind] for every tw_event.type in use. */
int tw_wire_create(const tw_wire_pid *pids, uint32_t n_lps, const tw_wire_type *types, uint32_t n_types, tw_wire_ctx **out);
int tw_wire_destroy(tw_wire_ctx *ctx);
/* `encode (m :: Message)`.  Encoders return TW_ERR_INVALID_ARG with *len = bytes needed when cap is too small. */
int tw_wire_encode_message(const tw_wire_ctx *ctx, const tw_event *ev, uint8_t *buf, size_t cap, size_t *len);
int tw_wire_decode_message(const tw_wire_ctx *ctx, const uint8_t *buf, size_t len, tw_event *ev, size_t *consumed);
/* `encode (m :: LogicalProcessMessage)`: kind = TW_WIRE_QUEUE_MESSAGE (n = 1), _BULK (n events), _COMPUTE_LOCAL_TIME, _GLOBAL_TIME
 * (time), _ABORT_SIMULATION.  The decoder returns TW_ERR_INVALID_ARG with *n = events needed when evs_cap is too small. */
int tw_wire_encode_lp_message(const tw_wire_ctx *ctx, uint32_t kind, const tw_event *evs, size_t n, double time,
                              uint8_t *buf, size_t cap, size_t *len);
int tw_wire_decode_lp_message(const tw_wire_ctx *ctx, const uint8_t *buf, size_t len, uint32_t *kind, tw_event *evs, size_t evs_cap,
                              size_t *n, double *time);
/* `encode (m :: TimeServerMessage)`: kind = TW_WIRE_TS_*, lp = the logical process named in it, time for LocalTimeMessage. */
int tw_wire_encode_ts_message(const tw_wire_ctx *ctx, uint32_t kind, uint32_t lp, double time, uint8_t *buf, size_t cap, size_t *len);
int tw_wire_decode_ts_message(const tw_wire_ctx *ctx, const uint8_t *buf, size_t len, uint32_t *kind, uint32_t *lp, double *time);
/* processChannelMessage (Internal/Event.hs:358-449) for a device partition registered under the ProcessId of LP self_lp:
 * QueueMessage / QueueMessageBulk -> tw_enqueue_messages; GlobalTimeMessage t -> tw_global_time_message;
 * ComputeLocalTimeMessage -> the replies ComputeLocalTimeAcknowledgementMessage pid and LocalTimeMessage pid t
 * (Internal/DIO.hs:560-566, 655-664), each preceded by its length as a little-endian uint32; AbortSimulationMessage ->
 * TW_ERR_ABORTED. */
int tw_wire_handle_lp_message(tw_engine *e, const tw_wire_ctx *ctx, uint32_t self_lp, const uint8_t *buf, size_t len,
                              uint8_t *replies, size_t cap, size_t *replies_len, uint32_t *n_replies);

/* ncclGetUniqueId for multi-GPU runs; rank 0 calls it and distributes the 128 bytes. */
int tw_nccl_unique_id(uint8_t out[128]);

const char *tw_last_error(const tw_engine *e);
const char *tw_strerror(int code);
/* Library identification: "twb200 <version> sm_100a". */
const char *tw_version(void);
/* Number of kernels this library has launched in this process (all engines). */
uint64_t tw_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
