"""CPU coverage of the fused path's HOST logic: the pass programs produced by the planner + round encoder
(qb_rounds_debug) are replayed with a NumPy emulator of the kernel's op semantics (tests/rounds_emulator.py)
and must reproduce the CPU oracle.  Also checks structural invariants of the programs (register / thread bit
partition, coalescing constraints of the HBM-facing rounds, host-computed swizzles) and that fusion happens."""
import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests import rounds_emulator as em
from tests.util import assert_close, random_state


def _replay(n, ops, seed):
    s0 = random_state(n, seed)
    ora = OracleRegister(s0.copy(), n)
    for o in ops:
        ora.apply(o)
    passes = em.dump_programs(n, ops)
    st = s0.copy()
    seen = []
    stats = {"prog": 0, "gates": 0, "rounds": 0, "ops": 0, "codes": set()}
    for p in passes:
        seen += p[1]
        if p[0] == "gates":
            stats["gates"] += 1
            tmp = OracleRegister(st, n)
            for tag in p[1]:
                tmp.apply(ops[tag])
            st = tmp.state
        else:
            stats["prog"] += 1
            stats["rounds"] += len(p[3]["rounds"])
            stats["ops"] += len(p[3]["ops"])
            stats["codes"] |= {o["code"] for o in p[3]["ops"]}
            st = em.emulate_program(st, n, p[2], p[3])
    return st, ora.state, seen, stats


@pytest.mark.parametrize("n,depth,seed", [(12, 8, 1), (13, 6, 2), (14, 6, 3), (15, 4, 4)])
def test_random_layers_programs(n, depth, seed):
    ops = circuits.random_circuit(n, depth, seed=seed)
    got, want, seen, st = _replay(n, ops, 100 + seed)
    assert sorted(seen) == list(range(len(ops))), "every gate must be scheduled exactly once"
    assert_close(got, want, f"layers n={n}")
    assert st["prog"] >= 1 and st["prog"] + st["gates"] < len(ops) / 4, "gates should be fused into few passes"
    assert em.RC_LAYER in st["codes"] or em.RC_H in st["codes"]
    assert st["ops"] < len(ops), "host-side merging (layers, phase tables) should shrink the op count"


@pytest.mark.parametrize("n", [12, 14, 16])
def test_qft_programs(n):
    ops = circuits.qft_circuit(n)
    got, want, seen, st = _replay(n, ops, 200 + n)
    assert sorted(seen) == list(range(len(ops)))
    assert_close(got, want, f"qft n={n}")
    assert {em.RC_TAB_ACC, em.RC_PH_FLUSH} & {c - (c - em.RC_PH_FLUSH) % 4 if em.RC_PH_FLUSH <= c < em.RC_PH_FLUSH + 4 else c for c in st["codes"]}, \
        "controlled-phase cascades should become accumulated chains"
    assert st["ops"] < 0.8 * len(ops)


@pytest.mark.parametrize("n,seed", [(12, 5), (13, 6), (14, 7), (15, 8)])
def test_mixed_programs(n, seed):
    ops = circuits.mixed_circuit(n, 120, seed)
    got, want, seen, st = _replay(n, ops, 300 + seed)
    identities = [i for i, (m, t) in enumerate(ops) if np.array_equal(np.asarray(m), np.eye(len(m)))]
    assert sorted(seen + identities) == list(range(len(ops)))
    assert_close(got, want, f"mixed n={n}")


def test_controls_on_register_thread_and_outer_bits():
    """A dense gate whose controls land on register bits, thread bits and bits outside the tile at once, plus
    diagonals with mixed bit sources (n = 15: three outer bits)."""
    n = 15
    rng = np.random.default_rng(9)
    ops = []
    for _ in range(40):
        qs = [int(x) for x in rng.permutation(n)]
        nc = int(rng.integers(1, 5))
        m = np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]
        big = (m, [qs[0]])
        for c in qs[1:1 + nc]:
            big = og.to_controlled(big, c)
        ops.append(big)
        ops.append(og.hadamard(qs[5]))
        ops.append(og.cnx(qs[6:8], qs[8]))
        d = np.exp(1j * rng.uniform(0, 6.28, size=8))
        ops.append((np.diag(d), qs[9:12]))
        ops.append(og.cphase(float(rng.uniform(0, 6.28)), qs[12], qs[13]))
    got, want, seen, st = _replay(n, ops, 17)
    assert_close(got, want)
    assert st["prog"] >= 1


def test_program_dump_handles_small_registers_and_single_gates():
    # fewer than 12 qubits: no register-blocked programs, everything is reported as plain gates
    passes = em.dump_programs(8, circuits.random_circuit(8, 3, seed=1))
    assert passes and all(p[0] == "gates" for p in passes)
    # a lone one-qubit gate goes to the streaming kernel, not to a tile program
    passes = em.dump_programs(13, [og.hadamard(3)])
    assert [p[0] for p in passes] == ["gates"]


def test_interleaved_phases_on_register_bits_become_per_bit_chains():
    """156 RZ gates alternating over four qubits that are register bits of the same round: diagonal ops commute, so the
    encoder regroups them per bit (one accumulated chain + flush each) instead of overflowing the program and falling
    back to the shared-memory kernel."""
    n = 14
    rng = np.random.default_rng(0)
    ops = [og.hadamard(q) for q in (8, 9, 10, 11)]
    ops += [og.rz(float(rng.uniform(0, 6)), 8 + i % 4) for i in range(156)]
    passes = em.dump_programs(n, ops)
    assert [p[0] for p in passes] == ["prog"] and len(passes[0][1]) == 160
    assert len(passes[0][3]["ops"]) <= 16
    got, want, seen, st = _replay(n, ops, 5)
    assert_close(got, want)


def test_short_phase_groups_stay_in_place_for_layer_packing():
    """RZ then H on each of four register bits: the lone phases must not be deferred in between the Hadamards, which the
    layer packing wants adjacent (one RC_LAYER instead of four ops)."""
    n = 13
    ops = [og.rz(0.3 + q, q) for q in (8, 9, 10, 11)] + [og.hadamard(q) for q in (8, 9, 10, 11)]
    passes = em.dump_programs(n, ops)
    codes = [o["code"] for p in passes if p[0] == "prog" for o in p[3]["ops"]]
    assert codes.count(em.RC_LAYER) == 1 and not any(em.RC_H <= c < em.RC_H + 4 for c in codes)
    got, want, seen, st = _replay(n, ops, 6)
    assert_close(got, want)


@pytest.mark.parametrize("kind,n,seed", [("layers", 14, 11), ("mixed", 15, 12), ("qft", 16, 13), ("mixed", 13, 14), ("layers", 16, 15)])
def test_thorough_planning_and_encoding_forced_on_small_registers(kind, n, seed, monkeypatch):
    """Big registers (>= 26 / 28 local qubits) plan passes with a look-ahead over the tile bits and choose rounds by trying
    three strategies (plain scan, low bits first, roll-out search).  QB_PLAN_LOOKAHEAD=1 forces all of that on at sizes
    the emulator can replay; the programs must still reproduce the oracle and never need more rounds than the plain scan."""
    ops = {"layers": circuits.random_circuit(n, 8, seed=seed), "qft": circuits.qft_circuit(n),
           "mixed": circuits.mixed_circuit(n, 150, seed)}[kind]
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "0")
    _, _, _, plain = _replay(n, ops, 400 + seed)
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "1")
    got, want, seen, st = _replay(n, ops, 400 + seed)
    identities = [i for i, (m, t) in enumerate(ops) if np.array_equal(np.asarray(m), np.eye(len(m)))]
    assert sorted(seen + identities) == list(range(len(ops)))
    assert_close(got, want, f"thorough {kind} n={n}")
    assert st["prog"] + st["gates"] <= plain["prog"] + plain["gates"] + 1
    if st["prog"] + st["gates"] == plain["prog"] + plain["gates"]:
        assert st["rounds"] <= plain["rounds"], (st["rounds"], plain["rounds"])


def test_unconditional_phases_on_one_bit_are_multiplied_on_the_host():
    """Five RZ / Z / S gates on the same qubit between two Hadamards commute and have no controls: the encoder multiplies
    them into one relative phase (one op), instead of an accumulated chain; Z Z cancels completely."""
    n = 13
    q = 9
    ops = [og.hadamard(q), og.rz(0.3, q), og.pauli_z(q), og.rz(-1.1, q), og.phase(q), og.pauli_z(q), og.hadamard(q)]
    passes = em.dump_programs(n, ops)
    codes = [o["code"] for p in passes if p[0] == "prog" for o in p[3]["ops"]]
    assert codes.count(em.RC_PH_R) + codes.count(em.RC_PH_NEG) == 1
    assert not any(em.RC_PH_FLUSH <= c < em.RC_PH_FLUSH + 4 for c in codes)
    got, want, seen, st = _replay(n, ops, 31)
    assert_close(got, want)
    ops = [og.hadamard(q), og.pauli_z(q), og.pauli_z(q), og.hadamard(q), og.hadamard(3)]
    passes = em.dump_programs(n, ops)
    codes = [o["code"] for p in passes if p[0] == "prog" for o in p[3]["ops"]]
    assert em.RC_PH_R not in codes and em.RC_PH_NEG not in codes
    got, want, seen, st = _replay(n, ops, 32)
    assert_close(got, want)


@pytest.mark.parametrize("kind,n,seed", [("layers", 14, 21), ("qft", 16, 22), ("mixed", 15, 23), ("qft_layers", 16, 24)])
def test_second_tier_plans_replay_to_the_oracle(kind, n, seed, monkeypatch):
    """The beam-search planner (used when a gate list comes back; it is allowed more gates per pass than the streaming planner,
    counting on the encoder's merging of phase cascades) through the encoder and the program emulator: every gate exactly
    once, amplitudes as the oracle's."""
    monkeypatch.setenv("QB_PLAN_DEBUG_BEAM", "4")
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "1")
    ops = {"layers": circuits.random_circuit(n, 8, seed=seed), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 150, seed),
           "qft_layers": circuits.qft_circuit(n) + circuits.random_circuit(n, 4, seed=seed)}[kind]
    got, want, seen, st = _replay(n, ops, 500 + seed)
    identities = [i for i, (m, t) in enumerate(ops) if np.array_equal(np.asarray(m), np.eye(len(m)))]
    assert sorted(seen + identities) == list(range(len(ops)))
    assert_close(got, want, f"second-tier plan {kind} n={n}")
