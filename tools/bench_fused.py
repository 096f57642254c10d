"""Fused vs unfused timing of the C3-style random layered circuit.
Usage: bench_fused.py [n] [depth] [tile_bits] [fused|both] [layers|qft] [force_path]   (see QB_OPT_FORCE_PATH)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quaru_b200 import Register, Operation
from tests import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tb = int(sys.argv[3]) if len(sys.argv) > 3 else 12
kind = sys.argv[5] if len(sys.argv) > 5 else 'layers'
raw = circuits.qft_circuit(n) if kind == 'qft' else circuits.random_circuit(n, depth, seed=0)
ops = [Operation(m, t) for m, t in raw]
depth = 1 if kind == 'qft' else depth
reg = Register.new([False] * n)
reg.set_option(2, tb)
if len(sys.argv) > 6:
    reg.set_option(4, int(sys.argv[6]))
B = 32.0 * (1 << n)
modes = (1,) if (len(sys.argv) > 4 and sys.argv[4] == 'fused') else (1, 0)
for fusion in modes:
    reg.set_option(1, fusion)
    reg.apply_batch(ops[:len(ops) // depth]); reg.sync()
    p0, l0, b0, r0 = reg.stat(2), reg.stat(1), reg.stat(4), reg.stat(9)
    t0 = time.time()
    reg.timer_start(); reg.apply_batch(ops); ms = reg.timer_stop()
    wall = (time.time() - t0) * 1e3
    passes = reg.stat(2) - p0
    print(f"{kind} n={n} depth={depth} T={tb} fusion={fusion}: {len(ops)} gates in {ms:.1f} ms (wall {wall:.1f}) = {len(ops) / ms * 1e3:.1f} gates/s, "
          f"{passes} passes ({ms / passes:.2f} ms/pass, {reg.stat(9) - r0} rounds), effective {len(ops) * B / ms / 1e6:.0f} GB/s, "
          f"physical {(reg.stat(4) - b0) / ms / 1e6:.0f} GB/s, norm2={reg.norm2():.12f}", flush=True)
