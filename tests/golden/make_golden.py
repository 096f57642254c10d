#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the CPU oracle (oracle/quaru_oracle.c).

The reference holds no numeric golden vectors for the hot path and cannot be executed here (Rust, no
toolchain), so the fixtures are outputs of the oracle -- which is itself pinned by every known-answer / property test
of the reference (tests/test_oracle_reference_suite.py).  They freeze the BASELINE configs that fit a file:
  grover_t64      examples/grovers --target 64 (n = 7, 8 iterations): pre-measurement state + the 7 outcomes
  shor15_a7       find_period_slow(15, 7) (12 qubits): pre-measurement state + the 8 outcomes + the period
  layers_n10      6 random H/RZ/CNOT/CZ layers (config C3 at n = 10) on |0..0>: state, 4 outcomes, collapsed state
  qft_n8          textbook QFT (config C4's front half at n = 8) on basis state 37
Run from the repository root:  python tests/golden/make_golden.py
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests import reference_suite as rs

HERE = os.path.dirname(os.path.abspath(__file__))
U_SEED = 20261017


def main():
    out = {}
    # --- Grover, target 64 -------------------------------------------------------------------
    rng = np.random.default_rng(U_SEED)
    pre, bits = rs.grover_full(rs.OracleBackend(), rng)
    out["grover_t64_state"] = pre
    out["grover_t64_bits"] = np.array(bits, dtype=np.uint8)
    # --- Shor 15, a = 7 -----------------------------------------------------------------------
    number, a = 15, 7
    n = math.ceil(math.log2(number + 1))
    reg = OracleRegister.new([False] * (3 * n))
    reg.apply(og.hadamard_transform(list(range(2 * n))))
    reg.apply(og.not_(2 * n))
    for i in range(2 * n):
        reg.apply(og.to_controlled(og.shor_u_gate(list(range(2 * n, 3 * n)), number, a, i), i))
    reg.apply(og.qft(2 * n))
    out["shor15_a7_state"] = reg.state.copy()
    rng = np.random.default_rng(U_SEED + 1)
    us = rng.random(2 * n)
    res = 0
    outcomes = []
    for i in range(2 * n):
        m = reg.measure(i, float(us[i]))
        outcomes.append(m)
        res |= int(m) << i
    out["shor15_a7_us"] = us
    out["shor15_a7_outcomes"] = np.array(outcomes, dtype=np.uint8)
    out["shor15_a7_period"] = np.array([og.limit_denominator(res, 2 ** (2 * n) - 1, number - 1)[1]])
    # --- random layers, n = 10 -----------------------------------------------------------------
    n = 10
    reg = OracleRegister.new([False] * n)
    for g in circuits.random_circuit(n, 6, seed=10):
        reg.apply(g)
    out["layers_n10_state"] = reg.state.copy()
    rng = np.random.default_rng(U_SEED + 2)
    us = rng.random(4)
    outs = [reg.measure(t, float(u)) for t, u in zip([9, 0, 4, 7], us)]
    out["layers_n10_us"] = us
    out["layers_n10_outcomes"] = np.array(outs, dtype=np.uint8)
    out["layers_n10_collapsed"] = reg.state.copy()
    # --- QFT, n = 8, on |37> --------------------------------------------------------------------
    reg = OracleRegister.from_int(8, 37)
    for g in circuits.qft_circuit(8):
        reg.apply(g)
    out["qft_n8_state"] = reg.state.copy()
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
