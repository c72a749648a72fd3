"""Diagnostic: time k_step with phases switched off (tw_params.flags bits 8..10).  Results of such runs are meaningless;
only the device time per launch is of interest.  Usage: python tools/ablate.py [c3|c4]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
n, rp = ((1 << 20), 50) if wl == "c4" else ((1 << 16), 10)
lib = H.ffi.load()
for ab, name in ((0, "full"), (1, "no execute/emit (header + scan + GVT reduction only)"), (8, "no message delivery"),
                 (16, "no record / saved-state stores"), (32, "no handler math"), (24, "no delivery, no stores"), (56, "no delivery, no stores, no handler math")):
    p = H.phold_params(n, 1000.0, remote_pct=rp, horizon=0.2, steps_per_launch=16 if ab == 1 else 1)
    p.flags = ab << 8
    h = C.c_void_p()
    assert lib.tw_create(C.byref(p), C.byref(h)) == 0
    cap = n * 16
    buf = (H.ffi.tw_event * cap)(); cnt = C.c_size_t()
    lib.tw_model_initial_events(C.byref(p), buf, cap, C.byref(cnt))
    assert lib.tw_enqueue_messages(h, buf, cnt.value) == 0
    st = H.ffi.tw_run_stats()
    if ab == 0:
        rc = lib.tw_run(h, 0.8, C.byref(st))             # ~10 real super-steps
        per = st.device_ms * 1e3 / max(1, st.steps)
    else:
        rc = lib.tw_run(h, 1000.0 if ab == 1 else 0.8, C.byref(st))      # timing-only: 8 graph launches x 16 steps (ab=1) or the first ~10 steps
        per = st.device_ms * 1e3 / 128 if ab == 1 else st.device_ms * 1e3 / max(1, st.steps)
    print(f"{wl} ablate={ab} {name}: {per:.1f} us per launch (rc={rc}, steps={st.steps}, executed={st.executed})", flush=True)
    lib.tw_destroy(h)
