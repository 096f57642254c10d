#!/usr/bin/env python
"""bench.py -- gates/s and effective HBM GB/s of the gate-application hot path on synthetic circuits.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--qubits n] [--workload c4|c3]

Headline workload (BASELINE.json `metric`: "33-qubit random circuit, 1/2/4/8 B200"; configs[3]):
textbook QFT on 33 qubits (H + controlled-phase cascade + swaps) followed by 20 random layers (per qubit H or
RZ(theta); random perfect matching of CNOT / CZ) -- 1557 gates on a 137.4 GB Complex<f64> state.  The SAME 33-qubit
circuit runs at every N ("scaling": "strong"): with N GPUs the state is sharded on its top log2(N) qubits, one process
per GPU.  One STEP is one application of the whole circuit to the evolving state through the C ABI.

`value` = gates/s, whole job, state resident in HBM, device-timed with CUDA events on the library's stream
(max over ranks).  `e2e` = the same, measured by wall clock through the public per-gate API
(Register.apply with HOST matrices, host->device transfer of every gate inside the timed region, and a
device->host read of a measurement probability at the end of each step).

Extra keys: `parity_check` (a 20-qubit circuit through the same code path -- sharding, scheduler and specialised kernels
included -- against the CPU oracle, outside every timed region; the bench exits non-zero if it fails), `sharded_target`
(N > 1: the north-star configuration -- 35-qubit random circuit, depth 50, at N = 4 / 8, 34 qubits at N = 2 -- with its
HBM + NVLink roofline fraction), `c3` (N = 1: BASELINE config 3, 28 qubits, depth 200), `jit` (compiler statistics).

--impl reference times the reference's CPU algorithm (oracle/quaru_oracle.c, a C restatement of
src/register.rs:140-191 -- the Rust reference cannot be compiled in this image) on ONE host core (the reference is
single-threaded, src/register.rs:158), on a bounded sample (same circuit family at the largest n that finishes in
seconds), converted to the metric's unit; the all-cores OpenMP figure is reported beside it.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_LAYERS = 20


# --------------------------------------------------------------------------------------------------
# synthetic circuit (numpy only; identical for the engine and the CPU baseline)
# --------------------------------------------------------------------------------------------------
def _h():
    return (np.sqrt(0.5) * np.array([[1.0, 1.0], [1.0, -1.0]])).astype(np.complex128)


def _rz(th):
    return np.array([[np.exp(-0.5j * th), 0], [0, np.exp(0.5j * th)]], dtype=np.complex128)


_CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)   # targets [t, c]
_CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)
_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def _cphase(angle):
    return np.diag([1, 1, 1, np.exp(1j * angle)]).astype(np.complex128)                            # targets [t, c]


def build_circuit(n, layers=N_LAYERS, seed=0, qft=True):
    """[(matrix, targets)] in Quaru's convention: matrix bit j <-> targets[j] (src/register.rs:168-173)."""
    rng = np.random.default_rng(seed)
    ops = []
    if qft:
        for t in range(n - 1, -1, -1):
            ops.append((_h(), [t]))
            for c in range(t - 1, -1, -1):
                ops.append((_cphase(2.0 * np.pi / (1 << (t - c + 1))), [t, c]))
        for a in range(n // 2):
            ops.append((_SWAP, [a, n - 1 - a]))
    for _ in range(layers):
        for q in range(n):
            if rng.random() < 0.5:
                ops.append((_h(), [q]))
            else:
                ops.append((_rz(float(rng.uniform(0.0, 2.0 * np.pi))), [q]))
        perm = [int(x) for x in rng.permutation(n)]
        for a, b in zip(perm[0::2], perm[1::2]):
            ops.append((_CNOT, [b, a]) if rng.random() < 0.5 else (_CZ, [b, a]))
    return ops


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        pw = [float(s[6]) for s in self.samples if s[6].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "power_w_max": max(pw) if pw else None, "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (port of the reference's algorithm) on a bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_baseline(n_workload, threads, budget_s=20.0, n_sample=None):
    """Times try_apply_fast (register.rs:140-191 restated in C) on the same circuit family at a reduced n and
    converts amplitude-updates/s to gates/s at n_workload.  Returns the cpu_baseline object."""
    from oracle import oracle as orc
    orc.build()
    if n_sample is None:
        n_sample = 24          # ~15 s single-threaded on the GPU box's host (10-30 s of CPU work; bounded by budget_s)
    ops = build_circuit(n_sample, layers=2, seed=0, qft=False)
    reg = orc.OracleRegister.new([False] * n_sample)
    t0 = time.perf_counter()
    done = 0
    for m, t in ops:
        reg.try_apply_fast((m, t), threads=threads)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    gps_sample = done / dt
    gps_scaled = gps_sample / float(1 << (n_workload - n_sample))
    return {"value": gps_scaled, "unit": "gates/s", "cores": threads, "kind": "port",
            "sample": f"{done} gates of the random H/RZ/CNOT/CZ layers at n={n_sample} in {dt:.1f} s "
                      f"({gps_sample:.2f} gates/s there), scaled by 2^-{n_workload - n_sample} to n={n_workload} "
                      f"(cost per gate is linear in 2^n); oracle/quaru_oracle.c try_apply_fast, {threads} thread(s); "
                      f"the Rust reference itself is single-threaded and limited to n<=30",
            "n_sample": n_sample, "gates_per_s_at_sample": gps_sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.qubits or 33
    vals, step_ms = [], []
    cb = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cb = cpu_baseline(n, 1, budget_s=6.0)          # the reference is single-threaded (src/register.rs:158)
        if i >= args.warmup:
            vals.append(cb["value"])
            step_ms.append((time.perf_counter() - t0) * 1e3)
    v = float(np.mean(vals)) if vals else cb["value"]
    cb["value"] = v
    try:    # what the same port does with every host thread (OpenMP over the groups): extra information, not the headline
        allc = cpu_baseline(n, os.cpu_count() or 1, budget_s=6.0)
        cb["all_cores"] = {"value": allc["value"], "cores": allc["cores"], "sample": allc["sample"]}
    except Exception as e:
        cb["all_cores"] = {"error": repr(e)[:120]}
    g = int(np.log2(args.gpus))
    # same workload description as our arm; a step here is the bounded CPU sample of it described in cpu_baseline.sample
    # (ms_per_step is the wall time of that sample, not of the 33-qubit circuit, which would take days on the host)
    line = {"impl": "reference", "metric": "gates/s", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)) if step_ms else None,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, "c4", g), "qubits": n,
                       "step": "bounded CPU sample of the workload, scaled (see cpu_baseline.sample)"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(n, kind, g):
    size = f"{n} qubits, Complex<f64> state {16 * (1 << n) / 1e9:.1f} GB ({16 * (1 << (n - g)) / 1e9:.1f} GB per GPU)"
    if kind == "c3":
        return f"random H/RZ/CNOT/CZ circuit, depth 200, {size}"
    if kind == "c5":
        return f"random H/RZ/CNOT/CZ circuit, depth 50, {size}"
    return f"QFT({n}) + {N_LAYERS} random H/RZ/CNOT/CZ layers, {size}"


def circuit_of(kind, n):
    if kind == "c3":
        return build_circuit(n, layers=200, seed=0, qft=False)
    if kind == "c5":
        return build_circuit(n, layers=50, seed=0, qft=False)
    return build_circuit(n, seed=0)


# --------------------------------------------------------------------------------------------------
# after the measurement: what the kernel variants cost on this GPU (outside every timed region)
# --------------------------------------------------------------------------------------------------
def diagnostics(n, budget_s=100.0):
    """tools/bench_diagnostics.py in a child process: a faulting variant can neither disturb the numbers above, which are
    final by now, nor keep the line from printing."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_diagnostics.py"), str(n), str(budget_s)],
                           cwd=ROOT, capture_output=True, text=True, timeout=budget_s + 40.0)
        last = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        return json.loads(last[-1]) if last else {"error": (p.stderr.strip().splitlines() or [f"exit {p.returncode}"])[-1][:200]}
    except Exception as e:
        return {"error": repr(e)[:200]}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class Job:
    """One process of the bench: torch.distributed plumbing + register construction."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: quaru_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        import __graft_entry__ as ge
        if self.rank == 0:
            ge.build(verbose=False)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            dist.barrier()
        self.g = int(np.log2(self.world))

    def register(self, n):
        import ctypes

        from quaru_b200 import Register, _ffi
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return Register.new([False] * n)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")      # a ncclUniqueId is single-use: one per register
        if self.rank == 0:
            buf = ctypes.create_string_buffer(128)
            if _ffi.lib().qb_nccl_unique_id(buf) != 0:
                raise SystemExit("qb_nccl_unique_id failed: " + _ffi.last_error()[0])
            uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        return Register.sharded(n, 0, self.rank, self.world, bytes(uid.cpu().tolist()), device=self.local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def warm_up(job, reg, ops, warmup):
    """At least W untimed steps, and as many more (at most 10 in all) as the library needs to settle on this circuit: the first
    application of a gate list is planned by the streaming planner and runs through the interpreter kernel while NVRTC
    compiles its specialised kernels in the background; the second time a list is seen the beam-search planner looks for a
    shorter plan (new pass programs, compiled in turn); SWAP relabelling makes consecutive steps alternate between two gate
    lists.  After every step the compiler is waited for (qb_jit_drain); warm-up ends once two consecutive steps (one of each
    list) have compiled nothing new and run no pass through the interpreter.  Returns the number of warm-up steps done."""
    from quaru_b200 import _ffi
    quiet, done = 0, 0
    while done < 10 and (done < max(warmup, 3) or quiet < 2):
        c0, m0 = reg.stat(_ffi.QB_STAT_JIT_COMPILES), reg.stat(_ffi.QB_STAT_JIT_MISSES)
        reg.apply_batch(ops)
        reg.sync()
        _ffi.lib().qb_jit_drain()
        done += 1
        settled = reg.stat(_ffi.QB_STAT_JIT_COMPILES) == c0 and reg.stat(_ffi.QB_STAT_JIT_MISSES) == m0
        if job.world > 1:   # every rank must leave the loop after the same step
            settled = job.maxr(0.0 if settled else 1.0) == 0.0
        quiet = quiet + 1 if settled else 0
    job.barrier()
    return done


def timed_steps(job, reg, ops, steps, clock_index=None):
    """Exactly `steps` applications, device-timed on the library's stream between barriers; every step ends with a
    non-blocking flush so that consecutive steps launch the same pass programs."""
    from quaru_b200 import _ffi
    S = _ffi
    keys = {"launches": S.QB_STAT_KERNEL_LAUNCHES, "passes": S.QB_STAT_PASSES, "bytes": S.QB_STAT_BYTES_MOVED,
            "tile_launches": S.QB_STAT_TILE_LAUNCHES, "tile_ns": S.QB_STAT_TILE_NS, "swap_bytes": S.QB_STAT_SWAP_BYTES,
            "swaps": S.QB_STAT_SWAPS, "jit_launches": S.QB_STAT_JIT_LAUNCHES, "jit_misses": S.QB_STAT_JIT_MISSES,
            "remote_passes": S.QB_STAT_REMOTE_PASSES, "rounds": S.QB_STAT_ROUNDS, "fused_exchanges": S.QB_STAT_FUSED_EXCHANGES}
    before = {k: reg.stat(v) for k, v in keys.items()}
    sampler = ClockSampler(clock_index) if clock_index is not None else None
    if sampler:
        sampler.__enter__()
    job.barrier()
    reg.timer_start()
    for _ in range(steps):
        reg.apply_batch(ops)
        reg.flush()
    ms = reg.timer_stop()
    job.barrier()
    if sampler:
        sampler.__exit__()
    out = {k: reg.stat(v) - before[k] for k, v in keys.items()}
    out["ms"] = job.maxr(ms)
    out["clocks"] = sampler.summary() if sampler else None
    return out


def parity_check(job, n=20):
    """A 20-qubit circuit (QFT + 2 random layers) through the same code path as the timed runs -- this process's sharding,
    the default scheduler, specialised kernels forced on (QB_OPT_JIT = 2, so a compile failure is an error) -- against the CPU
    oracle: every rank compares its slice of the final state and three measurement outcomes under a shared `u` stream.
    The oracle is the checker here, outside every timed region (what tests/sharded_worker.py::check does)."""
    from oracle.oracle import OracleRegister
    from quaru_b200 import Operation, _ffi
    ops_np = build_circuit(n, layers=2, seed=5)
    reg = job.register(n)
    reg.set_option(_ffi.QB_OPT_JIT, 2)
    reg.set_option(_ffi.QB_OPT_JIT_MIN_QUBITS, 12)
    reg.set_option(_ffi.QB_OPT_PEEPHOLE, 1)      # on by default only from 28 local qubits: what the timed registers run with
    ora = OracleRegister.new([False] * n)
    reg.apply_batch([Operation(m, t) for m, t in ops_np])
    for m, t in ops_np:
        ora.apply((m, t))
    rng = np.random.default_rng(99)
    outcomes_equal = True
    for t in (n - 1, 0, n // 2):
        u = float(rng.random())
        outcomes_equal = outcomes_equal and (reg.measure(t, u=u) == ora.measure(t, u))
    nl = n - job.g
    lo = job.rank << nl
    mine = reg.read(0, 1 << n)[lo:lo + (1 << nl)]
    d = np.abs(mine - ora.state[lo:lo + (1 << nl)])
    res = {"qubits": n, "gates": len(ops_np), "max_abs_diff": job.maxr(float(d.max())),
           "l2_diff": float(np.sqrt(job.sumr(float((d * d).sum())))),
           "outcomes_equal": job.maxr(0.0 if outcomes_equal else 1.0) == 0.0,
           "jit_launches": reg.stat(_ffi.QB_STAT_JIT_LAUNCHES), "exchanges": reg.stat(_ffi.QB_STAT_SWAPS),
           "remote_passes": reg.stat(_ffi.QB_STAT_REMOTE_PASSES),
           "peephole_rewrites": reg.stat(_ffi.QB_STAT_PEEPHOLE_REWRITES),
           "tolerance": {"max_abs_diff": 1e-12, "l2_diff": 1e-10}}
    res["ok"] = bool(res["max_abs_diff"] <= 1e-12 and res["l2_diff"] <= 1e-10 and res["outcomes_equal"])
    reg.close()
    return res


def nvlink_roofline(m, steps, per_gpu_bytes, hbm_peak):
    """Roofline of one step on one GPU: every pass streams the shard once (HBM) and every exchange / remote pass sends
    swap_bytes over NVLink.  Both NVLink peaks: 900 GB/s per direction (nominal) and 770 GB/s (measured peer
    copy, B200_PROFILING.md)."""
    step_s = m["ms"] / steps / 1e3
    hbm_s = (m["passes"] / steps) * per_gpu_bytes / (hbm_peak * 1e9)
    sb = m["swap_bytes"] / steps
    out = {"swap_bytes_sent_per_gpu_per_step": sb, "exchanges_per_step": m["swaps"] / steps,
           "exchanges_carried_by_a_pass_per_step": m["fused_exchanges"] / steps,
           "remote_passes_per_step": m["remote_passes"] / steps, "passes_per_step": m["passes"] / steps,
           "roofline_hbm_s": hbm_s, "step_s": step_s}
    for peak in (900.0, 770.0):
        nv_s = sb / (peak * 1e9)
        out[f"frac_of_roofline_{int(peak)}"] = (hbm_s + nv_s) / step_s
        out[f"roofline_nvlink_s_{int(peak)}"] = nv_s
    out["frac_of_roofline"] = out["frac_of_roofline_900"]
    return out


def run_ours(args):
    job = Job(args)
    world, rank, g = job.world, job.rank, job.g
    from quaru_b200 import Operation, _ffi

    kind = args.workload
    n = args.qubits or (28 if kind == "c3" else 33)
    ops_np = circuit_of(kind, n)
    ops = [Operation(m, t) for m, t in ops_np]
    n_gates = len(ops)
    peak, peak_src = measured_peaks()
    B_gate = 32.0 * (1 << n)           # algorithmic bytes per gate, whole state
    per_gpu_bytes = B_gate / world

    # ---- parity first: nothing is timed on a path that does not reproduce the oracle ----------------
    parity = None
    if not args.no_parity_check:
        parity = parity_check(job)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"error": "parity_check failed", "parity_check": parity}), flush=True)
            raise SystemExit(3)

    reg = job.register(n)
    reg.set_option(_ffi.QB_OPT_PROFILE, 1)
    # ---- device-timed: state resident, one C-ABI batch call per step --------------------------
    warmup_done = warm_up(job, reg, ops, args.warmup)
    m = timed_steps(job, reg, ops, args.steps, clock_index=job.local_rank)
    ms = m["ms"]
    value = n_gates * args.steps / (ms / 1e3)

    # ---- end to end: per-gate public API, host matrices, wall clock ------------------------------
    job.barrier()
    h2d = sum(mm.nbytes + 4 * len(t) for mm, t in ops_np)
    t0 = time.perf_counter()
    probe = 0.0
    for _ in range(args.steps):
        for o in ops:
            reg.apply(o)
        probe = reg.probabilities(n - 1)[0]
    job.barrier()
    e2e_s = job.maxr(time.perf_counter() - t0)
    e2e_value = n_gates * args.steps / e2e_s

    # ---- unfused single-qubit gate (the >= 70 %-of-roofline target of BASELINE.md): H on spread targets --
    reg.set_option(_ffi.QB_OPT_FUSION, 0)
    targets_1q = sorted(set([0, 1, 4, (n - g) // 2, n - g - 2, n - g - 1]))
    for t in targets_1q:
        reg.apply(Operation(_h(), [t]))
    reg.sync()
    job.barrier()
    reg.timer_start()
    reps = 2
    for _ in range(reps):
        for t in targets_1q:
            reg.apply(Operation(_h(), [t]))
    ms_1q = job.maxr(reg.timer_stop()) / (reps * len(targets_1q))
    reg.set_option(_ffi.QB_OPT_FUSION, 1)
    norm2 = reg.norm2()

    roof = None
    if m["tile_launches"] > 0 and m["tile_ns"] > 0:
        avg_ms = m["tile_ns"] / 1e6 / m["tile_launches"]
        phys = per_gpu_bytes / avg_ms / 1e6          # GB/s per GPU: one read + one write of the shard per launch
        jit_share = m["jit_launches"] / max(1, m["tile_launches"])
        roof = {"kernel": "qb_jit_pass (K2j: specialised fused tile pass, one NVRTC-compiled kernel per pass program)" if jit_share > 0.5
                else "k2r_tile_pass (fused tile pass, interpreter kernel)", "bound": "hbm",
                "achieved": phys, "peak": peak, "unit": "GB/s", "frac": phys / peak, "traffic": None,
                "peak_source": peak_src, "launches": m["tile_launches"], "avg_launch_ms": avg_ms,
                "jit_launch_share": jit_share, "algorithmic_bytes_per_launch": per_gpu_bytes,
                "note": "algorithmic = physical bytes of one sweep (32*2^n/P: every amplitude read once and written once) per launch; "
                        "a launch applies many gates -- see effective_GBps for gates x 32*2^n / time"}
    roof_1q = {"kernel": "k1_apply_1q (one dense 1-qubit gate per sweep, fusion off)", "bound": "hbm",
               "achieved": per_gpu_bytes / ms_1q / 1e6, "peak": peak, "unit": "GB/s",
               "frac": per_gpu_bytes / ms_1q / 1e6 / peak, "frac_of_nominal_8TBps": per_gpu_bytes / ms_1q / 1e6 / 8000.0,
               "traffic": None, "targets": targets_1q, "avg_launch_ms": ms_1q, "peak_source": peak_src}
    if roof is None:
        roof = roof_1q
    try:   # DRAM traffic of the dominant kernel: this round's ncu --set full capture of the same kernel (profiles/README.md)
        cap = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        roof["traffic_capture"] = cap
        roof["traffic"] = cap["traffic_over_algorithmic"] * per_gpu_bytes
        roof["traffic_note"] = ("per launch: dram__bytes_read.sum + dram__bytes_write.sum of the capture (n = %d), scaled by the "
                                "algorithmic-byte ratio (measured ratio %.4f)" % (cap.get("n", 28), cap["traffic_over_algorithmic"]))
    except Exception:
        pass

    jit_stats = {"launches_timed": m["jit_launches"], "misses_timed": m["jit_misses"],
                 "programs_compiled": reg.stat(_ffi.QB_STAT_JIT_COMPILES), "compile_ms_total": reg.stat(_ffi.QB_STAT_JIT_COMPILE_MS),
                 "failed": reg.stat(_ffi.QB_STAT_JIT_FAILED),
                 "note": "specialised kernels are compiled during warm-up (background NVRTC, interpreter kernel meanwhile) and "
                         "cached by program hash; misses_timed = passes of the timed region that ran through the interpreter"}
    line = {
        "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup_done, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, kind, g), "qubits": n, "gates_per_step": n_gates,
                   "sharding": f"top {g} qubits over {world} GPU(s)",
                   "l2_policy": "inputs (>= 17 GB per GPU) are far larger than the 126 MB L2; no flush needed",
                   "fusion": "on (tile passes)", "seed": 0,
                   "peephole": "on (H ; CX -> CZ ; H and friends on every flushed gate list, QB_OPT_PEEPHOLE; the operator is "
                               "unchanged, see parity_check; QB_PEEPHOLE=0 python bench.py measures without it)",
                   "peephole_rewrites_total": reg.stat(_ffi.QB_STAT_PEEPHOLE_REWRITES)},
        "effective_GBps": value * B_gate / 1e9,
        "physical_GBps": m["bytes"] * world / (ms / 1e3) / 1e9,
        "passes_per_step": m["passes"] / args.steps,
        "rounds_per_step": m["rounds"] / args.steps,
        "gpu_launches": int(m["launches"]),
        "roofline": roof,
        "roofline_1q_unfused": roof_1q,
        "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 16,
                "seconds": e2e_s, "probe_p0": probe},
        "clocks": m["clocks"],
        "norm2_after": norm2,
        "jit": jit_stats,
        "parity_check": parity,
    }
    if world > 1:
        line["nvlink"] = nvlink_roofline(m, args.steps, per_gpu_bytes, peak)
        line["nvlink"]["transport"] = ("peer-memory kernel (CUDA IPC)" if reg.stat(_ffi.QB_STAT_P2P_EXCHANGES) > 0 or
                                       reg.stat(_ffi.QB_STAT_REMOTE_PASSES) > 0 else "NCCL send/recv + staging")
    reg.close()    # frees the state before anything below allocates its own

    # ---- the north-star sharded configuration (N > 1): 35 qubits at N = 4 / 8, 34 at N = 2 (35 does not fit 2 GPUs) -------
    if world > 1 and not args.no_target and kind == "c4":
        nt = 34 if world == 2 else 35
        try:
            tops_np = circuit_of("c5", nt)
            tops = [Operation(mm, t) for mm, t in tops_np]
            treg = job.register(nt)
            treg.set_option(_ffi.QB_OPT_PROFILE, 1)
            twarm = warm_up(job, treg, tops, 3)
            tm = timed_steps(job, treg, tops, args.target_steps)
            tb = 32.0 * (1 << nt) / world
            tgt = {"workload": workload_name(nt, "c5", g), "qubits": nt, "gates_per_step": len(tops), "steps": args.target_steps,
                   "warmup": twarm, "value": len(tops) * args.target_steps / (tm["ms"] / 1e3), "unit": "gates/s",
                   "ms_per_step": tm["ms"] / args.target_steps,
                   "effective_GBps": len(tops) * args.target_steps / (tm["ms"] / 1e3) * 32.0 * (1 << nt) / 1e9,
                   "tile_pass_ms": tm["tile_ns"] / 1e6 / max(1, tm["tile_launches"]),
                   "jit_misses_timed": tm["jit_misses"], "norm2_after": treg.norm2(),
                   "nvlink": nvlink_roofline(tm, args.target_steps, tb, peak)}
            tgt["frac_of_roofline"] = tgt["nvlink"]["frac_of_roofline"]
            treg.close()
            line["sharded_target"] = tgt
        except Exception as e:
            line["sharded_target"] = {"error": repr(e)[:300]}

    if rank == 0:
        if world == 1 and not args.no_c3 and kind == "c4":
            # BASELINE config 3: 28 qubits, depth 200 (8400 gates), no QFT -- reported beside the headline, same measurement rules
            try:
                cops_np = circuit_of("c3", 28)
                cops = [Operation(mm, t) for mm, t in cops_np]
                creg = job.register(28)
                creg.set_option(_ffi.QB_OPT_PROFILE, 1)
                cwarm = warm_up(job, creg, cops, 3)
                cm = timed_steps(job, creg, cops, 3)
                cb = 32.0 * (1 << 28)
                cavg = cm["tile_ns"] / 1e6 / max(1, cm["tile_launches"])
                line["c3"] = {"workload": workload_name(28, "c3", 0), "gates_per_step": len(cops), "steps": 3, "warmup": cwarm,
                              "value": len(cops) * 3 / (cm["ms"] / 1e3), "unit": "gates/s", "ms_per_step": cm["ms"] / 3,
                              "effective_GBps": len(cops) * 3 / (cm["ms"] / 1e3) * cb / 1e9,
                              "passes_per_step": cm["passes"] / 3, "tile_pass_ms": cavg,
                              "tile_pass_GBps": cb / cavg / 1e6 if cavg else None, "tile_pass_frac_of_hbm": cb / cavg / 1e6 / peak if cavg else None,
                              "jit_misses_timed": cm["jit_misses"], "norm2_after": creg.norm2()}
                creg.close()
            except Exception as e:
                line["c3"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(n, 1, budget_s=25.0)
            except Exception as e:  # the oracle is test infrastructure; never let it take the bench down
                line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        if world == 1 and not args.no_diagnostics:
            line["diagnostics"] = diagnostics(min(30, n))
        print(json.dumps(line), flush=True)
    if world > 1:
        job.dist.barrier()
        job.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the qubit count (default 33; 28 for --workload c3)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c3", "c5"],
                    help="c4 (default): QFT + 20 layers; c3: depth-200 random circuit (28 qubits); c5: depth-50 random circuit")
    ap.add_argument("--target-steps", type=int, default=1, help="timed steps of the 35-qubit sharded target run (N > 1)")
    ap.add_argument("--no-target", action="store_true", help="skip the 35-qubit sharded target run at N > 1")
    ap.add_argument("--no-c3", action="store_true", help="skip the 28-qubit depth-200 line at N = 1")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-diagnostics", action="store_true", help="skip the kernel-variant table appended at N = 1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
