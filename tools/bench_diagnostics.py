"""Hardware diagnostics appended to the bench line (bench.py, N = 1): what the tile-pass variants and the planner switches
cost on THIS GPU, so that every bench run also answers "which kernel should be the default".

    python tools/bench_diagnostics.py [n=30] [budget_s=110]       # prints ONE JSON object on the last stdout line

Every configuration runs in its own process (an experimental kernel that faults cannot take the others or the caller
down) under its own timeout:
  * force_path 0 (default kernel) with the planner / encoder look-ahead on (default) and off (QB_PLAN_LOOKAHEAD=0), and with
    SWAP gates executed instead of absorbed into the qubit names (QB_SWAP_RELABEL=0);
  * force_path 5..11: the persistent tile kernels of qb_tile2_kernel.cuh (coefficients / op records in shared memory,
    L2 prefetch of the next tile, op records fetched one op ahead), with and without the register-table op (QB_TILE_RTAB=1).
Per configuration: the result at n = 24 is compared with the default kernel's (max |diff|, 0.0 = bit-identical), then the
bench circuit family is timed at n qubits: 20 random layers, the QFT, and two skeleton circuits (H on 4 / on 12 qubits =
one pass of one / three rounds) that separate the streaming cost of a pass from the cost of its ops.
Nothing here touches oracle/ (numpy circuits from bench.build_circuit).
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = [  # (label, force_path, extra environment); most informative first, the time budget may cut the tail
    ("default", 0, {}),
    ("default_no_lookahead", 0, {"QB_PLAN_LOOKAHEAD": "0"}),
    ("default_no_swap_relabel", 0, {"QB_SWAP_RELABEL": "0"}),
    ("default_io_low3", 0, {"QB_TILE_IO_LOW": "3"}),
    ("persist_mats_ops", 6, {}), ("persist_mats_ops_rtab", 6, {"QB_TILE_RTAB": "1"}),
    ("persist_mats_ops_ahead", 10, {}), ("persist_mats_ops_pf", 8, {}), ("persist_mats_ops_pf_ahead", 11, {}),
    ("persist_mats_ops_pf_ahead_rtab", 11, {"QB_TILE_RTAB": "1"}),
    ("persist_mats", 5, {}), ("persist_mats_pf", 7, {}), ("persist_mats_ahead", 9, {}),
]


def one(n, force):
    import numpy as np

    import bench
    from quaru_b200 import Operation, Register

    def circuit(kind, nq):
        if kind == "layers":
            return bench.build_circuit(nq, layers=20, seed=0, qft=False)
        if kind == "qft":
            return bench.build_circuit(nq, layers=0, seed=0, qft=True)
        if kind == "h4":
            return [(bench._h(), [q]) for q in (5, 6, 7, 8)]
        return [(bench._h(), [q]) for q in range(12)]

    row = {"force_path": force}
    # correctness first: same circuit through the default kernel and through this configuration, 4096 tiles
    nc = min(24, n)
    ops = [Operation(m, t) for m, t in bench.build_circuit(nc, layers=6, seed=1, qft=True)]
    regs = []
    for f in (0, force):
        r = Register.new([False] * nc)
        r.set_option(4, f)
        r.apply_batch(ops)
        regs.append(r)
    row["max_abs_diff_vs_default_n24"] = regs[0].max_abs_diff(regs[1])
    row["norm2_n24"] = regs[1].norm2()
    for r in regs:
        r.close()
    reg = Register.new([False] * n)
    reg.set_option(4, force)
    for kind in ("layers", "qft", "h4", "h12") if n >= 12 else ("layers", "qft"):
        ops = [Operation(m, t) for m, t in circuit(kind, n)]
        reg.apply_batch(ops[:min(len(ops), 100)])   # first launches, function attributes
        reg.sync()
        best, passes, rounds = None, 0, 0
        for _ in range(3):
            p0, r0 = reg.stat(2), reg.stat(9)
            reg.timer_start()
            reg.apply_batch(ops)
            ms = reg.timer_stop()
            passes, rounds = reg.stat(2) - p0, reg.stat(9) - r0
            best = ms if best is None else min(best, ms)
        row[kind] = {"gates": len(ops), "passes": passes, "rounds": rounds, "ms": round(best, 3),
                     "ms_per_pass": round(best / max(passes, 1), 3)}
    row["norm2_end"] = reg.norm2()
    reg.close()
    print("DIAG " + json.dumps(row), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return one(int(sys.argv[2]), int(sys.argv[3]))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 110.0
    t_end = time.time() + budget
    out = {"n": n, "unit": "ms", "note": "best of 3, CUDA events on the library stream; one process per configuration", "rows": []}
    for label, force, env in CONFIGS:
        left = t_end - time.time()
        if left < 5:
            out["rows"].append({"label": label, "skipped": "time budget"})
            continue
        e = dict(os.environ)
        for k in ("QB_PLAN_LOOKAHEAD", "QB_TILE_RTAB", "QB_OPT_FORCE_PATH", "QB_SWAP_RELABEL", "QB_TILE_IO_LOW"):
            e.pop(k, None)
        e.update(env)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(n), str(force)], env=e, cwd=ROOT,
                               capture_output=True, text=True, timeout=min(30.0, left))
            rows = [ln[5:] for ln in p.stdout.splitlines() if ln.startswith("DIAG ")]
            if rows:
                row = json.loads(rows[-1])
            else:
                row = {"error": (p.stderr or p.stdout).strip().splitlines()[-1][:200] if (p.stderr or p.stdout).strip() else f"exit {p.returncode}"}
        except subprocess.TimeoutExpired:
            row = {"error": "timeout"}
        except Exception as ex:   # diagnostics never take the caller down
            row = {"error": repr(ex)[:200]}
        row["label"] = label
        if env:
            row["env"] = env
        out["rows"].append(row)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
