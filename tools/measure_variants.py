"""One GPU call that compares the tile-kernel variants: for each QB_OPT_FORCE_PATH value it first checks the result against
the default kernel bit for bit on a small register (same arithmetic in every variant), then times the fused random-layer
circuit and the QFT.  Usage: python tools/measure_variants.py [n=30] [depth=20] [variants=0,5,6,7,8,9,10]
Run it a second time with QB_TILE_RTAB=1 in the environment to add the register-table op to the persistent kernels.
Writes gpurun_out/variants.json (variants_rtab.json with QB_TILE_RTAB) when the directory exists."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quaru_b200 import Operation, Register  # noqa: E402
from tests import circuits  # noqa: E402
from tests.util import random_state  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 5, 6, 7, 8, 9, 10, 11]


def small_check(variant):
    ops = [Operation(m, t) for m, t in circuits.mixed_circuit(16, 150, 3) + circuits.qft_circuit(16)]
    s = random_state(16, 1)
    regs = []
    for force in (0, variant):
        r = Register.new([False] * 16)
        r.set_state(s)
        r.set_option(4, force)
        r.apply_batch(ops)
        regs.append(r)
    return regs[0].max_abs_diff(regs[1])


rows = []
for v in variants:
    row = {"force_path": v, "max_abs_diff_vs_default": small_check(v) if v else 0.0}
    for kind, raw in (("layers", circuits.random_circuit(n, depth, seed=0)), ("qft", circuits.qft_circuit(n))):
        ops = [Operation(m, t) for m, t in raw]
        reg = Register.new([False] * n)
        reg.set_option(4, v)
        reg.apply_batch(ops[:min(len(ops), 90)])          # warm-up: first launches, function attributes
        reg.sync()
        best = None
        for _ in range(3):
            p0 = reg.stat(2)
            reg.timer_start()
            reg.apply_batch(ops)
            ms = reg.timer_stop()
            passes = reg.stat(2) - p0
            best = ms if best is None else min(best, ms)
        row[kind] = {"gates": len(ops), "passes": passes, "ms": round(best, 2), "ms_per_pass": round(best / passes, 3),
                     "gates_per_s": round(len(ops) / best * 1e3, 1)}
        reg.close()
    rows.append(row)
    print(json.dumps(row), flush=True)
if os.path.isdir("gpurun_out"):
    name = "variants_rtab.json" if os.environ.get("QB_TILE_RTAB") else "variants.json"
    json.dump({"n": n, "depth": depth, "rtab": bool(os.environ.get("QB_TILE_RTAB")), "rows": rows}, open(os.path.join("gpurun_out", name), "w"), indent=1)
