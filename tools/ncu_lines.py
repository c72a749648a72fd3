"""Per-source-line summary of an ncu report (instructions executed, stall samples): which lines of tw_core.cuh / engine.cu a
kernel spends its instructions and its waiting on.  usage: python tools/ncu_lines.py report.ncu-rep [top N] [kernel regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []
cur_file = None
hdr = None
for rec in csv.reader(io.StringIO(out)):
    if not rec:
        continue
    if rec[0] == "File Path":
        cur_file = rec[1].split("/")[-1]; continue
    if rec[0] == "Function Name":
        continue
    if rec[0] == "Line No":
        hdr = rec; continue
    if hdr is None or rec[0] == "":
        continue          # SASS rows (their metrics are already summed into the source row)
    d = dict(zip(hdr, rec))
    try:
        rows.append((cur_file, int(rec[0]), rec[1].strip()[:110], int(d["Instructions Executed"]), int(d["# Samples"]),
                     int(d.get("stall_long_sb", 0) or 0), int(d["Thread Instructions Executed"])))
    except Exception:
        pass
tot_i = sum(r[3] for r in rows) or 1; tot_s = sum(r[4] for r in rows) or 1
print(f"total warp instructions {tot_i}, stall samples {tot_s}, avg active threads {sum(r[6] for r in rows) / tot_i:.1f}")
print("--- by instructions")
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print(f"{100 * r[3] / tot_i:5.1f}% inst {100 * r[4] / tot_s:5.1f}% samp  lanes {r[6] / max(1, r[3]):4.1f}  {r[0]}:{r[1]}  {r[2]}")
print("--- by stall samples")
for r in sorted(rows, key=lambda r: -r[4])[:top]:
    print(f"{100 * r[4] / tot_s:5.1f}% samp {100 * r[3] / tot_i:5.1f}% inst (long_sb {100 * r[5] / max(1, r[4]):3.0f}%)  {r[0]}:{r[1]}  {r[2]}")

# coarse phases of lp_step / k_step (line ranges of the current sources are looked up by marker comments)
import re, os
src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "aivika-distributed_b200", "csrc")
def line_of(fname, marker):
    for i, l in enumerate(open(os.path.join(src_dir, fname)), 1):
        if marker in l:
            return i
    return None
marks = [("tw_core.cuh", "TW_HD void lp_step(", "lp_step: header / idle"), ("tw_core.cuh", "---- 1. lazy fossil collection", "1 fossil"),
         ("tw_core.cuh", "---- 2. drain the inbox", "2 drain arrivals"), ("tw_core.cuh", "---- 3. rollback to processed position", "3 rollback"),
         ("tw_core.cuh", "---- 4./5. select and execute", "4 selection scan"), ("tw_core.cuh", "for (;;) {                                           // iterations within the round", "5 execute loop"),
         ("tw_core.cuh", "// what stays pending inside the window", "6 write back"), ("tw_core.cuh", "// Host-injected / cross-GPU routed", None)]
bounds = [(f, line_of(f, m), name) for f, m, name in marks]
phase = {}
def classify(f, ln):
    if f == "tw_core.cuh":
        for (f0, a, name), (f1, b, _) in zip(bounds, bounds[1:]):
            if a and b and a <= ln < b:
                return name
        wq0, wq1 = line_of("tw_core.cuh", "struct WorkQ {"), line_of("tw_core.cuh", "struct StepCtx {")
        if wq0 and wq1 and wq0 <= ln < wq1:
            return "work queue (push / run / complete)"
        return "tw_core other (emit_classify, drain_slow, helpers)"
    if f in ("tw_math.cuh", "tw_models.cuh", "tw_dsl.cuh"):
        return "model handler + math"
    if f == "engine.cu":
        return "k_step wrapper + block_finish"
    return "intrinsics / other"
for r in rows:
    k = classify(r[0], r[1]); a = phase.setdefault(k, [0, 0, 0]); a[0] += r[3]; a[1] += r[4]; a[2] += r[6]
print("--- by phase")
for k, a in sorted(phase.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * a[0] / tot_i:5.1f}% inst {100 * a[1] / tot_s:5.1f}% samp  lanes {a[2] / max(1, a[0]):4.1f}  {k}")
