"""A short randomized run of tools/stress_emu.py (emulated device functions vs the sequential oracle over random PHOLD /
MachRep2 / MachRep1 configurations) as a regression net; the long runs of the round are summarised in
profiles/r1_tuning.md."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_short_randomized_stress():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_emu.py"), "20261020", "12", "fast"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    last = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0 and "fails 0" in last, r.stdout[-2000:]
    assert int(last.split()[1]) >= 20          # enough cases really ran
