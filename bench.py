#!/usr/bin/env python
"""bench.py -- gates/s and effective HBM GB/s of the gate-application hot path on synthetic circuits.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--qubits n]

Workload (BASELINE.json configs[3]): textbook QFT on n qubits (H + controlled-phase
cascade + swaps) followed by 20 random layers (per qubit H or RZ(theta); random perfect matching of CNOT / CZ).
At N=1 n = 33 (137.4 GB of Complex<f64> amplitudes, the largest state one B200 holds); with N GPUs the state is
sharded on its top log2(N) qubits with n = 33 + log2(N) (weak scaling: 2^33 amplitudes per GPU).  One STEP is
one application of the whole circuit to the evolving state through the C ABI.

`value` = gates/s, whole job, state resident in HBM, device-timed with CUDA events on the library's stream
(max over ranks).  `e2e` = the same, measured by wall clock through the public per-gate API
(Register.apply with HOST matrices, host->device transfer of every gate inside the timed region, and a
device->host read of a measurement probability at the end of each step).

--impl reference times the reference's CPU algorithm (oracle/quaru_oracle.c, a C restatement of
src/register.rs:140-191 -- the Rust reference cannot be compiled in this image) on the host cores, on a
bounded sample (same circuit family at the largest n that finishes in seconds), converted to the metric's unit.
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
    # the allocation-faithful variant (per-group vectors + matrix clone, register.rs:155-188) on a smaller sample
    faithful = None
    try:
        nf = min(n_sample, 18)
        regf = orc.OracleRegister.new([False] * nf)
        opsf = build_circuit(nf, layers=1, seed=0, qft=False)[:12]
        tf = time.perf_counter()
        for m, t in opsf:
            regf.try_apply_fast((m, t), faithful=True)
        faithful = len(opsf) / (time.perf_counter() - tf) / float(1 << (n_workload - nf))
    except Exception:
        pass
    return {"value": gps_scaled, "unit": "gates/s", "cores": threads, "kind": "port",
            "sample": f"{done} gates of the random H/RZ/CNOT/CZ layers at n={n_sample} in {dt:.1f} s "
                      f"({gps_sample:.2f} gates/s there), scaled by 2^-{n_workload - n_sample} to n={n_workload} "
                      f"(cost per gate is linear in 2^n); oracle/quaru_oracle.c try_apply_fast, {threads} OpenMP thread(s); "
                      f"the Rust reference itself is single-threaded and limited to n<=30",
            "n_sample": n_sample, "gates_per_s_at_sample": gps_sample,
            "faithful_variant_value": faithful}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.qubits or (33 + int(np.log2(args.gpus)))
    threads = os.cpu_count() or 1
    vals, step_ms = [], []
    cb = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cb = cpu_baseline(n, threads, budget_s=6.0)
        if i >= args.warmup:
            vals.append(cb["value"])
            step_ms.append((time.perf_counter() - t0) * 1e3)
    v = float(np.mean(vals)) if vals else cb["value"]
    cb["value"] = v
    g = int(np.log2(args.gpus))
    # same workload description as our arm; a step here is the bounded CPU sample of it described in cpu_baseline.sample
    # (ms_per_step is the wall time of that sample, not of the 33-qubit circuit, which would take days on the host)
    line = {"impl": "reference", "metric": "gates/s", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)) if step_ms else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"QFT({n}) + {N_LAYERS} random H/RZ/CNOT/CZ layers, {n} qubits, Complex<f64> state "
                                   f"{16 * (1 << n) / 1e9:.1f} GB ({16 * (1 << (n - g)) / 1e9:.1f} GB per GPU)",
                       "qubits": n, "step": "bounded CPU sample of the workload, scaled (see cpu_baseline.sample)"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# after the measurement: what the tile-kernel variants and planner switches cost on this GPU (outside every timed region)
# --------------------------------------------------------------------------------------------------
def diagnostics(n, budget_s=110.0):
    """tools/bench_diagnostics.py in a child process (each configuration in a grandchild of its own): a faulting
    experimental kernel can neither disturb the numbers above, which are final by now, nor keep the line from printing."""
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
def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quaru_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if rank == 0:
        ge.build(verbose=False)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from quaru_b200 import Operation, Register, _ffi

    g = int(np.log2(world))
    n = args.qubits or (33 + g)
    ops_np = build_circuit(n, seed=0)
    ops = [Operation(m, t) for m, t in ops_np]
    n_gates = len(ops)

    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes
            buf = ctypes.create_string_buffer(128)
            rc = _ffi.lib().qb_nccl_unique_id(buf)
            if rc != 0:
                raise SystemExit("qb_nccl_unique_id failed: " + _ffi.last_error()[0])
            uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        reg = Register.sharded(n, 0, rank, world, bytes(uid.cpu().tolist()), device=local_rank)
    else:
        reg = Register.new([False] * n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    reg.set_option(_ffi.QB_OPT_PROFILE, 1)
    B_gate = 32.0 * (1 << n)           # algorithmic bytes per gate, whole state
    # ---- device-timed: state resident, one C-ABI batch call per step --------------------------
    for _ in range(args.warmup):
        reg.apply_batch(ops)
        reg.sync()
    barrier()
    l0, p0, b0 = reg.stat(_ffi.QB_STAT_KERNEL_LAUNCHES), reg.stat(_ffi.QB_STAT_PASSES), reg.stat(_ffi.QB_STAT_BYTES_MOVED)
    tl0, tn0 = reg.stat(_ffi.QB_STAT_TILE_LAUNCHES), reg.stat(_ffi.QB_STAT_TILE_NS)
    sb0, sw0 = reg.stat(_ffi.QB_STAT_SWAP_BYTES), reg.stat(_ffi.QB_STAT_SWAPS)
    with ClockSampler(local_rank) as clk:
        barrier()
        reg.timer_start()
        for _ in range(args.steps):
            reg.apply_batch(ops)
        ms = reg.timer_stop()
        barrier()
    ms = maxr(ms)
    launches = reg.stat(_ffi.QB_STAT_KERNEL_LAUNCHES) - l0
    passes = reg.stat(_ffi.QB_STAT_PASSES) - p0
    phys_bytes = reg.stat(_ffi.QB_STAT_BYTES_MOVED) - b0
    tile_launches = reg.stat(_ffi.QB_STAT_TILE_LAUNCHES) - tl0
    tile_ms = (reg.stat(_ffi.QB_STAT_TILE_NS) - tn0) / 1e6
    swap_bytes_step = (reg.stat(_ffi.QB_STAT_SWAP_BYTES) - sb0) / args.steps
    swaps_step = (reg.stat(_ffi.QB_STAT_SWAPS) - sw0) / args.steps
    value = n_gates * args.steps / (ms / 1e3)
    clocks = clk.summary()

    # ---- end to end: per-gate public API, host matrices, wall clock ------------------------------
    barrier()
    h2d = sum(m.nbytes + 4 * len(t) for m, t in ops_np)
    t0 = time.perf_counter()
    probe = 0.0
    for _ in range(args.steps):
        for o in ops:
            reg.apply(o)
        probe = reg.probabilities(n - 1)[0]
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e_value = n_gates * args.steps / e2e_s

    # ---- unfused single-qubit gate (the >= 70 %-of-roofline target of BASELINE.md): H on spread targets --
    reg.set_option(_ffi.QB_OPT_FUSION, 0)
    h = Operation(_h(), [0])
    targets_1q = sorted(set([0, 1, 4, (n - g) // 2, n - g - 2, n - g - 1]))
    for t in targets_1q:
        reg.apply(Operation(_h(), [t]))
    reg.sync()
    barrier()
    reg.timer_start()
    reps = 2
    for _ in range(reps):
        for t in targets_1q:
            reg.apply(Operation(_h(), [t]))
    ms_1q = maxr(reg.timer_stop()) / (reps * len(targets_1q))
    reg.set_option(_ffi.QB_OPT_FUSION, 1)
    norm2 = reg.norm2()

    peak, peak_src = measured_peaks()
    per_gpu_bytes = B_gate / world
    gates_per_tile_launch = None
    roof = None
    if tile_launches > 0 and tile_ms > 0:
        avg_ms = tile_ms / tile_launches
        phys = per_gpu_bytes / avg_ms / 1e6          # GB/s per GPU: one read + one write of the shard per launch
        roof = {"kernel": "k2_tile_pass (fused multi-gate tile pass)", "bound": "hbm",
                "achieved": phys, "peak": peak, "unit": "GB/s", "frac": phys / peak, "traffic": None,
                "peak_source": peak_src, "launches": tile_launches, "avg_launch_ms": avg_ms,
                "algorithmic_bytes_per_launch": per_gpu_bytes,
                "note": "physical bytes of one sweep (32*2^n/P) per launch; a launch applies many gates, so the pass "
                        "is shared-memory/FP64-bound, not HBM-bound -- see effective_GBps for gates x 32*2^n / time"}
    roof_1q = {"kernel": "k1_apply_1q (one dense 1-qubit gate per sweep, fusion off)", "bound": "hbm",
               "achieved": per_gpu_bytes / ms_1q / 1e6, "peak": peak, "unit": "GB/s",
               "frac": per_gpu_bytes / ms_1q / 1e6 / peak, "frac_of_nominal_8TBps": per_gpu_bytes / ms_1q / 1e6 / 8000.0,
               "traffic": None, "targets": targets_1q, "avg_launch_ms": ms_1q, "peak_source": peak_src}
    if roof is None:
        roof = roof_1q
    try:   # DRAM traffic of the dominant kernel from the committed ncu capture (n = 28; see profiles/README.md)
        cap = json.load(open(os.path.join(ROOT, "profiles", "r1c_traffic.json")))
        roof["traffic_capture"] = cap
        roof["traffic"] = cap["traffic_over_algorithmic"] * per_gpu_bytes
        roof["traffic_note"] = "per launch, scaled from the n = 28 ncu capture by the algorithmic-byte ratio (measured ratio %.4f)" % cap["traffic_over_algorithmic"]
    except Exception:
        pass

    line = {
        "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"QFT({n}) + {N_LAYERS} random H/RZ/CNOT/CZ layers, {n} qubits, Complex<f64> state "
                               f"{16 * (1 << n) / 1e9:.1f} GB ({16 * (1 << (n - g)) / 1e9:.1f} GB per GPU)",
                   "qubits": n, "gates_per_step": n_gates, "sharding": f"top {g} qubits over {world} GPU(s)",
                   "l2_policy": "inputs (>= 17 GB per GPU) are far larger than the 126 MB L2; no flush needed",
                   "fusion": "on (tile passes)", "seed": 0},
        "effective_GBps": value * B_gate / 1e9,
        "physical_GBps": phys_bytes * world / (ms / 1e3) / 1e9,
        "passes_per_step": passes / args.steps,
        "gpu_launches": int(launches),
        "roofline": roof,
        "roofline_1q_unfused": roof_1q,
        "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 16,
                "seconds": e2e_s, "probe_p0": probe},
        "clocks": clocks,
        "norm2_after": norm2,
    }
    if world > 1:
        # roofline of one step on one GPU: every pass streams the shard once (HBM) and every exchange sends
        # swap_bytes over NVLink (measured peer copy 770 GB/s per direction, B200_PROFILING.md; 900 nominal)
        nv_peak = 770.0
        hbm_s = (passes / args.steps) * per_gpu_bytes / (peak * 1e9)
        nv_s = swap_bytes_step / (nv_peak * 1e9)
        line["nvlink"] = {"swap_bytes_sent_per_gpu_per_step": swap_bytes_step, "swaps_per_step": swaps_step,
                          "transport": "peer-memory kernel (CUDA IPC)" if reg.stat(_ffi.QB_STAT_P2P_EXCHANGES) > 0 else "NCCL send/recv + staging",
                          "p2p_exchanges_total": reg.stat(_ffi.QB_STAT_P2P_EXCHANGES), "exchanges_total": reg.stat(_ffi.QB_STAT_SWAPS),
                          "remote_passes_total": reg.stat(_ffi.QB_STAT_REMOTE_PASSES),
                          "scheduler": ("remote tile passes (QB_SHARD_REMOTE=1)" if os.environ.get("QB_SHARD_REMOTE") == "1" else
                                        "first scheduler, cut at every exchange (QB_SHARD_REORDER=0)" if os.environ.get("QB_SHARD_REORDER") == "0"
                                        else "commutation-aware (default)"),
                          "peak_GBps_per_direction": nv_peak, "roofline_s_per_step": hbm_s + nv_s,
                          "roofline_hbm_s": hbm_s, "roofline_nvlink_s": nv_s,
                          "frac_of_roofline": (hbm_s + nv_s) / (ms / args.steps / 1e3)}
    wd = None
    if world > 1 and not args.no_diagnostics:
        # The other two shard schedulers (the first one, which every multi-GPU number of round 1 was measured with, and the
        # experimental exchange-free one), one warm-up and one timed step each, AFTER every number above is final.
        # The remote one has never run on hardware: a watchdog (armed until the line has been printed) prints the line as it stands
        # and ends the process if anything hangs; every rank has one, so torchrun still sees exit code 0.
        sched_diag = {"note": "gates/s of one step after one warm-up step; outside every timed region", "default": value}
        line["scheduler_diagnostics"] = sched_diag

        def bail():
            if rank == 0:
                sched_diag.setdefault("error", "watchdog: the process did not finish after the last entry above")
                print(json.dumps(line), flush=True)
            os._exit(0)
        wd = threading.Timer(150.0, bail)
        wd.daemon = True
        wd.start()
        for label, var, val in (("first_scheduler", "QB_SHARD_REORDER", "0"), ("remote", "QB_SHARD_REMOTE", "1")):
            os.environ[var] = val        # read by the library at every flush (getenv)
            try:
                sw0, sb0, rp0, p0 = (reg.stat(_ffi.QB_STAT_SWAPS), reg.stat(_ffi.QB_STAT_SWAP_BYTES),
                                     reg.stat(_ffi.QB_STAT_REMOTE_PASSES), reg.stat(_ffi.QB_STAT_PASSES))
                reg.apply_batch(ops)
                reg.sync()
                barrier()
                reg.timer_start()
                reg.apply_batch(ops)
                ms_x = maxr(reg.timer_stop())
                sched_diag[label] = {"gates_per_s": n_gates / (ms_x / 1e3), "ms_per_step": ms_x,
                                     "passes_2_steps": reg.stat(_ffi.QB_STAT_PASSES) - p0,
                                     "exchanges_2_steps": reg.stat(_ffi.QB_STAT_SWAPS) - sw0,
                                     "remote_passes_2_steps": reg.stat(_ffi.QB_STAT_REMOTE_PASSES) - rp0,
                                     "nvlink_bytes_sent_2_steps": reg.stat(_ffi.QB_STAT_SWAP_BYTES) - sb0,
                                     "norm2": reg.norm2()}
            except Exception as e:
                # an engine error on this rank (its peers may be waiting in a collective: their watchdogs end them)
                sched_diag[label] = {"error": repr(e)[:200]}
                if rank == 0:
                    print(json.dumps(line), flush=True)
                os._exit(0)
            finally:
                os.environ.pop(var, None)
    reg.close()    # frees the state (137 GB) before the diagnostics processes allocate theirs
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(n, 1, budget_s=25.0)
            except Exception as e:  # the oracle is test infrastructure; never let it take the bench down
                line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        if world == 1 and not args.no_diagnostics:
            line["diagnostics"] = diagnostics(min(30, n))
        print(json.dumps(line), flush=True)
    if wd is not None:
        wd.cancel()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the qubit count (default 33 + log2(gpus))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-diagnostics", action="store_true", help="skip the kernel-variant table appended at N = 1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
