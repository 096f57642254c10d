"""Single-gate sweep: one gate per HBM pass (fusion off), timed with CUDA events on the handle's stream.
Usage: python tools/sweep_1q.py [n] [reps] ; prints GB/s per target for H, RZ, CNOT(t, t+1 mod n), CZ."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from quaru_b200 import Register, operation as op

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reg = Register.new([False] * n)
reg.set_option(1, 0)
reg.apply_all(op.hadamard(0))
reg.sync()
B = 32.0 * (1 << n)
rows = []
for name, mk, frac in [("H", lambda t: op.hadamard(t), 1.0), ("RZ", lambda t: op.rz(0.37, t), 1.0),
                       ("CNOT", lambda t: op.cnot(t, (t + 1) % n), 0.5), ("CZ", lambda t: op.cz(t, (t + 1) % n), 0.25)]:
    for t in range(n):
        g = mk(t)
        reg.apply(g); reg.sync()
        best = 1e9
        for _ in range(reps):
            reg.timer_start(); reg.apply(g); ms = reg.timer_stop()
            best = min(best, ms)
        rows.append({"gate": name, "target": t, "ms": best, "eff_GBps": B / best / 1e6, "phys_GBps": frac * B / best / 1e6})
        print(f"{name:5s} t={t:2d} {best:8.3f} ms  eff {B / best / 1e6:8.1f} GB/s  phys {frac * B / best / 1e6:8.1f} GB/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"n": n, "rows": rows}, open(f"gpurun_out/sweep_1q_n{n}.json", "w"))
hs = [r["phys_GBps"] for r in rows if r["gate"] == "H"]
print("H: min %.1f median %.1f max %.1f GB/s" % (min(hs), float(np.median(hs)), max(hs)))
