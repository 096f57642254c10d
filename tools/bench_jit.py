"""Timing of the specialised tile kernels against the interpreter kernel.
Usage: bench_jit.py [n=30] [layers|qft|bench] [variants=0,2,3,18,50]   (variant 0 = interpreter; see QB_OPT_JIT_VARIANT)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from quaru_b200 import Operation, Register, _ffi
from tests import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
kind = sys.argv[2] if len(sys.argv) > 2 else "layers"
variants = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,2,3,18,50").split(",")]
raw = circuits.qft_circuit(n) if kind == "qft" else bench.build_circuit(n) if kind == "bench" else circuits.random_circuit(n, 20, seed=0)
ops = [Operation(m, t) for m, t in raw]
reg = Register.new([False] * n)
reg.set_option(_ffi.QB_OPT_PROFILE, 1)
B = 32.0 * (1 << n)
for v in variants:
    reg.set_option(_ffi.QB_OPT_JIT, 2 if v else 0)
    if v:
        reg.set_option(_ffi.QB_OPT_JIT_VARIANT, v)
    t0 = time.time()
    reg.apply_batch(ops); reg.sync()
    warm = time.time() - t0
    p0, j0, tn0, tl0 = reg.stat(2), reg.stat(12), reg.stat(8), reg.stat(7)
    reg.timer_start()
    for _ in range(2):
        reg.apply_batch(ops); reg.flush()
    ms = reg.timer_stop() / 2
    passes = (reg.stat(2) - p0) / 2
    tile_ms = (reg.stat(8) - tn0) / 1e6 / max(1, reg.stat(7) - tl0)
    print(f"{kind} n={n} variant={v}: {len(ops)} gates in {ms:.1f} ms = {len(ops) / ms * 1e3:.1f} gates/s, {passes:.1f} passes, "
          f"{tile_ms:.2f} ms/tile pass = {B / tile_ms / 1e6:.0f} GB/s physical, jit launches {(reg.stat(12) - j0) // 2}, "
          f"first application {warm:.1f} s, compile total {reg.stat(15)} ms, norm2={reg.norm2():.12f}", flush=True)
