"""Hardware diagnostics appended to the bench line (bench.py, N = 1): what the tile-pass kernel variants and the planner
switches cost on THIS GPU, so that every bench run also answers "which configuration should be the default".

    python tools/bench_diagnostics.py [n=30] [budget_s=100]       # prints ONE JSON object on the last stdout line

Every configuration runs in its own process (a faulting variant cannot take the others or the caller down) under its own
timeout:
  * `interpreter`: the interpreter kernel k2r_tile_pass (QB_JIT=0) -- the round-1 kernel;
  * `jit`: the specialised kernels (K2j, qb_jit_gen.cu), default shape: one CTA per tile, 128 registers, 2 CTAs / SM;
  * `jit_pipelined`: one persistent CTA per SM, a producer warp streaming tiles into a ring of three shared-memory buffers with
    bulk-async copies (cp.async.bulk + mbarrier), two consumer groups of 256 threads, bulk-async stores;
  * `jit_3cta` (80 registers, 3 CTAs / SM), `jit_persistent`, `jit_persistent_prefetch`: the other kernel shapes;
  * `jit_lc4`: the planner allowed 256-byte contiguous segments (QB_PLAN_MIN_LC=4: eight arbitrary high tile bits instead of 7);
  * `jit_no_lookahead`, `jit_no_swap_relabel`: planner switches.
Per configuration: the result at n = 24 is compared with the interpreter kernel's (max |diff|), then the bench circuit
family is timed at n qubits: 20 random layers, the QFT, and two skeleton circuits (H on 4 / on 12 qubits = one pass of one /
three rounds) that separate the streaming cost of a pass from the cost of its ops.
Nothing here touches oracle/ (numpy circuits from bench.build_circuit).
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = [  # (label, jit mode, jit variant, extra environment); most informative first, the time budget may cut the tail
    ("jit", 2, 2, {}),
    ("interpreter", 0, 2, {}),
    ("jit_no_peephole", 2, 2, {"QB_PEEPHOLE": "0"}),
    ("jit_prefetch_ring", 2, 2 | 64, {}),
    ("jit_lc4", 2, 2, {"QB_PLAN_MIN_LC": "4"}),
    ("jit_3cta", 2, 3, {}),
    ("jit_persistent", 2, 2 | 16, {}),
    ("jit_persistent_prefetch", 2, 2 | 16 | 32, {}),
    ("jit_no_lookahead", 2, 2, {"QB_PLAN_LOOKAHEAD": "0"}),
    ("jit_no_swap_relabel", 2, 2, {"QB_SWAP_RELABEL": "0"}),
]


def one(n, jit, variant):
    import bench
    from quaru_b200 import Operation, Register, _ffi

    def circuit(kind, nq):
        if kind == "layers":
            return bench.build_circuit(nq, layers=20, seed=0, qft=False)
        if kind == "qft":
            return bench.build_circuit(nq, layers=0, seed=0, qft=True)
        if kind == "h4":
            return [(bench._h(), [q]) for q in (5, 6, 7, 8)]
        return [(bench._h(), [q]) for q in range(12)]

    def configure(r, mode):
        r.set_option(_ffi.QB_OPT_JIT, mode)
        r.set_option(_ffi.QB_OPT_JIT_MIN_QUBITS, 12)
        r.set_option(_ffi.QB_OPT_JIT_VARIANT, variant)

    row = {"jit": jit, "variant": variant}
    # correctness first: same circuit through the interpreter kernel and through this configuration, 4096 tiles
    nc = min(24, n)
    ops = [Operation(m, t) for m, t in bench.build_circuit(nc, layers=6, seed=1, qft=True)]
    regs = []
    for mode in (0, jit):
        r = Register.new([False] * nc)
        configure(r, 1 if mode else 0)
        if mode:                                  # hand every program to the compiler pool first (a throw-away application), then
            r.apply_batch(ops)                    # run the checked application with every kernel ready
            r.sync()
            _ffi.lib().qb_jit_drain()
            r.reset(0)
            configure(r, mode)
        r.apply_batch(ops)
        regs.append(r)
    row["max_abs_diff_vs_interpreter_n24"] = regs[0].max_abs_diff(regs[1])
    row["norm2_n24"] = regs[1].norm2()
    for r in regs:
        r.close()
    reg = Register.new([False] * n)
    configure(reg, jit)
    for kind in ("layers", "qft", "h4", "h12"):
        ops = [Operation(m, t) for m, t in circuit(kind, n)]
        if jit:
            reg.set_option(_ffi.QB_OPT_JIT, 1)    # asynchronous while the programs are first seen: the pool compiles them in parallel
        for i in range(4):                        # SWAP relabelling alternates between two placements: two applications see every
            reg.apply_batch(ops)                  # program; then wait for the compiler and load every kernel
            reg.sync()
            if i == 1:
                _ffi.lib().qb_jit_drain()
        if jit:
            reg.set_option(_ffi.QB_OPT_JIT, jit)
        best, passes, rounds = None, 0, 0
        for _ in range(4):
            p0, r0 = reg.stat(2), reg.stat(9)
            reg.timer_start()
            reg.apply_batch(ops)
            ms = reg.timer_stop()
            passes, rounds = reg.stat(2) - p0, reg.stat(9) - r0
            best = ms if best is None else min(best, ms)
        row[kind] = {"gates": len(ops), "passes": passes, "rounds": rounds, "ms": round(best, 3),
                     "ms_per_pass": round(best / max(passes, 1), 3)}
    row["norm2_end"] = reg.norm2()
    row["jit_launches"] = reg.stat(_ffi.QB_STAT_JIT_LAUNCHES)
    row["compile_ms_total"] = reg.stat(_ffi.QB_STAT_JIT_COMPILE_MS)
    reg.close()
    print("DIAG " + json.dumps(row), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
    t_end = time.time() + budget
    out = {"n": n, "unit": "ms", "note": "best of 4, CUDA events on the library stream; one process per configuration", "rows": []}
    for label, jit, variant, env in CONFIGS:
        left = t_end - time.time()
        if left < 8:
            out["rows"].append({"label": label, "skipped": "time budget"})
            continue
        e = dict(os.environ)
        for k in ("QB_PLAN_LOOKAHEAD", "QB_SWAP_RELABEL", "QB_TILE_IO_LOW", "QB_PLAN_MIN_LC", "QB_JIT", "QB_JIT_VARIANT", "QB_JIT_MIN_QUBITS", "QB_PEEPHOLE"):
            e.pop(k, None)
        e.update(env)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(n), str(jit), str(variant)], env=e, cwd=ROOT,
                               capture_output=True, text=True, timeout=min(40.0, left))
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
