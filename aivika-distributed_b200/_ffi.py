"""ctypes view of include/twb200.h (the C ABI of libtwb200.so).

This is the binding a Python caller uses; the Haskell `foreign import ccall` twin is in
aivika-distributed_b200/hs/ and INTEGRATION.md.  The library is loaded from the in-tree build
(aivika-distributed_b200/csrc/libtwb200.so); there is no fallback of any kind: if it is missing,
`load()` raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtwb200.so")

TW_OK = 0
TW_MODEL_PHOLD = 1
TW_MODEL_MACHREP2 = 2
TW_MODEL_MACHREP1 = 3
TW_FLAG_SKIP_OUTDATED = 1
TW_FLAG_NO_GROUP_KERNEL = 2

ERROR_NAMES = {
    0: "TW_OK", 1: "TW_ERR_INVALID_ARG", 2: "TW_ERR_CUDA", 3: "TW_ERR_SEND_AFTER_RECV",
    4: "TW_ERR_OUTDATED_MESSAGE", 5: "TW_ERR_QUEUE_OVERFLOW", 6: "TW_ERR_ANTI_UNMATCHED",
    7: "TW_ERR_DEADLOCK", 8: "TW_ERR_TIME_ORDER", 9: "TW_ERR_BAD_DESTINATION", 10: "TW_ERR_MODEL",
    11: "TW_ERR_OUTBOX_OVERFLOW", 12: "TW_ERR_NOT_SUPPORTED", 13: "TW_ERR_ABORTED",
}

# every symbol include/twb200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "tw_create", "tw_destroy", "tw_reset", "tw_enqueue_message", "tw_enqueue_messages", "tw_run", "tw_gvt",
    "tw_lp_state_get", "tw_read_lp_stats", "tw_read_machrep", "tw_partition", "tw_set_monitor", "tw_model_initial_events", "tw_model_initial_events_range", "tw_populate",
    "tw_attach_time_server", "tw_local_time", "tw_global_time_message", "tw_skipped_outdated", "tw_max_ring_occupancy",
    "tw_wire_create", "tw_wire_destroy", "tw_wire_encode_message", "tw_wire_decode_message", "tw_wire_encode_lp_message",
    "tw_wire_decode_lp_message", "tw_wire_encode_ts_message", "tw_wire_decode_ts_message", "tw_wire_handle_lp_message",
    "tw_nccl_unique_id", "tw_last_error", "tw_strerror", "tw_version", "tw_kernel_launch_count",
]


class tw_event(C.Structure):
    _fields_ = [("recv_ts", C.c_double), ("send_ts", C.c_double), ("ctr", C.c_uint64), ("src", C.c_uint32),
                ("dst", C.c_uint32), ("type", C.c_uint32), ("pad", C.c_uint32), ("value", C.c_double)]


class tw_phold_params(C.Structure):
    _fields_ = [("population", C.c_uint32), ("remote_pct", C.c_uint32), ("n_hot", C.c_uint32), ("hot_pct", C.c_uint32),
                ("lookahead", C.c_double), ("mean", C.c_double), ("seed", C.c_uint64)]


class tw_machrep_params(C.Structure):
    _fields_ = [("mean_up_time", C.c_double), ("mean_repair_time", C.c_double), ("delta", C.c_double),
                ("seed1", C.c_int64), ("seed2", C.c_int64), ("seed_stride", C.c_int64)]


class tw_params(C.Structure):
    _fields_ = [("model", C.c_uint32), ("n_lps", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32),
                ("device", C.c_int32), ("queue_capacity", C.c_uint32), ("max_events_per_step", C.c_uint32),
                ("steps_per_launch", C.c_uint32), ("start_time", C.c_double), ("stop_time", C.c_double),
                ("time_horizon", C.c_double), ("proc_limit", C.c_uint32), ("fossil_threshold", C.c_uint32),
                ("outbox_capacity", C.c_uint32), ("flags", C.c_uint32), ("phold", tw_phold_params),
                ("machrep", tw_machrep_params), ("nccl_id", C.c_uint8 * 128),
                ("sync_timeout_ms", C.c_uint32), ("reserved", C.c_uint32)]


class tw_run_stats(C.Structure):
    _fields_ = [("gvt", C.c_double), ("device_ms", C.c_double), ("steps", C.c_uint64), ("committed", C.c_uint64),
                ("executed", C.c_uint64), ("rolled_back", C.c_uint64), ("rollbacks", C.c_uint64),
                ("stragglers", C.c_uint64), ("anti_sent", C.c_uint64), ("annihilated", C.c_uint64),
                ("retries", C.c_uint64), ("sent_local", C.c_uint64), ("sent_remote", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("trace_hash_sum", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class tw_lp_state(C.Structure):
    _fields_ = [("lp", C.c_uint32), ("local_time", C.c_double), ("event_queue_time", C.c_double),
                ("event_queue_size", C.c_uint32), ("log_size", C.c_uint32), ("input_message_count", C.c_uint32),
                ("output_message_count", C.c_uint32), ("rollback_count", C.c_uint64), ("executed", C.c_uint64),
                ("trace_hash", C.c_uint64)]


class tw_machrep_result(C.Structure):
    _fields_ = [("total_up_time", C.c_double), ("n_rep", C.c_int64), ("n_immed_rep", C.c_int64),
                ("up_time_prop", C.c_double), ("immed_prop", C.c_double)]


class tw_monitor_state(C.Structure):
    _fields_ = [("gvt", C.c_double), ("steps", C.c_uint64), ("executed", C.c_uint64), ("rolled_back", C.c_uint64),
                ("rollbacks", C.c_uint64), ("anti_sent", C.c_uint64), ("annihilated", C.c_uint64),
                ("sent_local", C.c_uint64), ("sent_remote", C.c_uint64), ("n_lps_local", C.c_uint32), ("rank", C.c_uint32)]


class tw_wire_pid(C.Structure):
    _fields_ = [("node_address", C.c_char_p), ("address_len", C.c_uint32), ("lpid_unique", C.c_int32), ("lpid_counter", C.c_int32)]


class tw_wire_type(C.Structure):
    _fields_ = [("fingerprint", C.c_uint8 * 16), ("payload", C.c_uint32)]


MONITOR_FN = C.CFUNCTYPE(None, C.POINTER(tw_monitor_state), C.c_void_p)

_lib = None


def load(path=None):
    """Load libtwb200.so and declare the prototypes.  Raises OSError when the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("TWB200_LIB") or LIB_PATH     # TWB200_LIB: tuning builds of the same CUDA library
    if not os.path.exists(p):
        raise OSError(f"{p} not found: build it with `make -C aivika-distributed_b200/csrc` "
                      f"(or __graft_entry__.build()); this engine has no CPU fallback")
    lib = C.CDLL(p)
    H = C.c_void_p
    lib.tw_create.argtypes = [C.POINTER(tw_params), C.POINTER(H)]; lib.tw_create.restype = C.c_int
    lib.tw_destroy.argtypes = [H]; lib.tw_destroy.restype = C.c_int
    lib.tw_reset.argtypes = [H]; lib.tw_reset.restype = C.c_int
    lib.tw_enqueue_message.argtypes = [H, C.POINTER(tw_event)]; lib.tw_enqueue_message.restype = C.c_int
    lib.tw_enqueue_messages.argtypes = [H, C.c_void_p, C.c_size_t]; lib.tw_enqueue_messages.restype = C.c_int
    lib.tw_run.argtypes = [H, C.c_double, C.POINTER(tw_run_stats)]; lib.tw_run.restype = C.c_int
    lib.tw_gvt.argtypes = [H, C.POINTER(C.c_double)]; lib.tw_gvt.restype = C.c_int
    lib.tw_lp_state_get.argtypes = [H, C.c_uint32, C.POINTER(tw_lp_state)]; lib.tw_lp_state_get.restype = C.c_int
    lib.tw_read_lp_stats.argtypes = [H, C.c_void_p, C.c_void_p, C.c_size_t]; lib.tw_read_lp_stats.restype = C.c_int
    lib.tw_read_machrep.argtypes = [H, C.c_void_p, C.c_size_t]; lib.tw_read_machrep.restype = C.c_int
    lib.tw_set_monitor.argtypes = [H, MONITOR_FN, C.c_void_p]; lib.tw_set_monitor.restype = C.c_int
    lib.tw_attach_time_server.argtypes = [H, C.c_int]; lib.tw_attach_time_server.restype = C.c_int
    lib.tw_local_time.argtypes = [H, C.POINTER(C.c_double)]; lib.tw_local_time.restype = C.c_int
    lib.tw_global_time_message.argtypes = [H, C.c_double, C.POINTER(C.c_int)]; lib.tw_global_time_message.restype = C.c_int
    lib.tw_skipped_outdated.argtypes = [H, C.POINTER(C.c_uint64)]; lib.tw_skipped_outdated.restype = C.c_int
    lib.tw_max_ring_occupancy.argtypes = [H, C.POINTER(C.c_uint32)]; lib.tw_max_ring_occupancy.restype = C.c_int
    lib.tw_partition.argtypes = [H, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]; lib.tw_partition.restype = C.c_int
    lib.tw_populate.argtypes = [C.c_void_p]
    lib.tw_populate.restype = C.c_int
    lib.tw_model_initial_events.argtypes = [C.POINTER(tw_params), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.tw_model_initial_events.restype = C.c_int
    lib.tw_model_initial_events_range.argtypes = [C.POINTER(tw_params), C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t,
                                                  C.POINTER(C.c_size_t)]
    lib.tw_model_initial_events_range.restype = C.c_int
    U8P, SZP = C.c_void_p, C.POINTER(C.c_size_t)
    lib.tw_wire_create.argtypes = [C.POINTER(tw_wire_pid), C.c_uint32, C.POINTER(tw_wire_type), C.c_uint32, C.POINTER(H)]
    lib.tw_wire_destroy.argtypes = [H]
    lib.tw_wire_encode_message.argtypes = [H, C.POINTER(tw_event), U8P, C.c_size_t, SZP]
    lib.tw_wire_decode_message.argtypes = [H, U8P, C.c_size_t, C.POINTER(tw_event), SZP]
    lib.tw_wire_encode_lp_message.argtypes = [H, C.c_uint32, C.c_void_p, C.c_size_t, C.c_double, U8P, C.c_size_t, SZP]
    lib.tw_wire_decode_lp_message.argtypes = [H, U8P, C.c_size_t, C.POINTER(C.c_uint32), C.c_void_p, C.c_size_t, SZP, C.POINTER(C.c_double)]
    lib.tw_wire_encode_ts_message.argtypes = [H, C.c_uint32, C.c_uint32, C.c_double, U8P, C.c_size_t, SZP]
    lib.tw_wire_decode_ts_message.argtypes = [H, U8P, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    lib.tw_wire_handle_lp_message.argtypes = [H, H, C.c_uint32, U8P, C.c_size_t, U8P, C.c_size_t, SZP, C.POINTER(C.c_uint32)]
    for f in ("tw_wire_create", "tw_wire_destroy", "tw_wire_encode_message", "tw_wire_decode_message", "tw_wire_encode_lp_message",
              "tw_wire_decode_lp_message", "tw_wire_encode_ts_message", "tw_wire_decode_ts_message", "tw_wire_handle_lp_message"):
        getattr(lib, f).restype = C.c_int
    lib.tw_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8 * 128)]; lib.tw_nccl_unique_id.restype = C.c_int
    lib.tw_last_error.argtypes = [H]; lib.tw_last_error.restype = C.c_char_p
    lib.tw_strerror.argtypes = [C.c_int]; lib.tw_strerror.restype = C.c_char_p
    lib.tw_version.argtypes = []; lib.tw_version.restype = C.c_char_p
    lib.tw_kernel_launch_count.argtypes = []; lib.tw_kernel_launch_count.restype = C.c_uint64
    if path is None:
        _lib = lib
    return lib
