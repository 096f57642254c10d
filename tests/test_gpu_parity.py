"""GPU parity tests proper: the CUDA engine (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerance from BASELINE.json north_star: amplitudes within 1e-12 absolute and
L2 <= 1e-10 (f64); index/permutation logic, basis states and measured outcomes bit-exact."""
import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests.util import assert_close, random_matrix, random_state, tree_probabilities

pytestmark = pytest.mark.gpu

FORCE_GENERIC, FORCE_TILE, FORCE_STREAM = 1, 2, 3
FORCE_GENERIC_OUT_OF_PLACE = 5
OPT_FUSION, OPT_TILE_BITS, OPT_FORCE = 1, 2, 4


def _pair(engine, n, seed, fusion=0, force=0):
    s = random_state(n, seed)
    reg = engine.Register.new([False] * n)
    reg.set_option(OPT_FUSION, fusion)
    if force:
        reg.set_option(OPT_FORCE, force)
    reg.set_state(s)
    return reg, OracleRegister(s.copy(), n)


def _both(engine, reg, ora, matrix, targets):
    reg.apply(engine.Operation(matrix, targets))
    ora.apply((matrix, targets))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 9, 12, 15])
def test_dense_1q_every_target(engine, n):
    """register.rs:140-191 with k=1 on every target (streaming kernel K1, all three access modes)."""
    reg, ora = _pair(engine, n, 100 + n)
    for t in range(n):
        _both(engine, reg, ora, random_matrix(1, 1000 * n + t), [t])
    assert_close(reg.state, ora.state, f"n={n}")


@pytest.mark.parametrize("n", [2, 3, 5, 8, 13])
def test_controlled_1q_all_pairs(engine, n):
    """cnot (operation.rs:372) and to_controlled (operation.rs:406) on every (control, target) pair."""
    reg, ora = _pair(engine, n, 200 + n)
    for c in range(n):
        for t in range(n):
            if c == t:
                continue
            _both(engine, reg, ora, *og.cnot(c, t))
            _both(engine, reg, ora, *og.to_controlled((random_matrix(1, 7 * c + t, unitary=True), [t]), c))
    assert_close(reg.state, ora.state, f"n={n}")


@pytest.mark.parametrize("n", [3, 4, 6, 7, 10])
def test_cnx_cnz(engine, n):
    """cnx / cnz (operation.rs:491, 520) with 1..5 controls in scrambled positions."""
    rng = np.random.default_rng(n)
    reg, ora = _pair(engine, n, 300 + n)
    for nc in range(1, min(n, 6)):
        for _ in range(4):
            qs = [int(x) for x in rng.permutation(n)[:nc + 1]]
            _both(engine, reg, ora, *og.cnx(qs[1:], qs[0]))
            _both(engine, reg, ora, *og.cnz(qs[1:], qs[0]))
    assert_close(reg.state, ora.state, f"n={n}")


@pytest.mark.parametrize("n", [1, 2, 4, 7, 11, 14])
def test_diagonal_gates(engine, n):
    """RZ, phase, pauli_z, CZ, controlled phase on every target; wide diagonals (Grover oracle / diffuser,
    examples/grovers.rs:181-218)."""
    reg, ora = _pair(engine, n, 400 + n)
    rng = np.random.default_rng(n)
    for t in range(n):
        _both(engine, reg, ora, *og.rz(float(rng.uniform(0, 6.28)), t))
        _both(engine, reg, ora, *og.phase(t))
        _both(engine, reg, ora, *og.pauli_z(t))
        for c in range(n):
            if c != t:
                _both(engine, reg, ora, *og.cz(c, t))
                _both(engine, reg, ora, *og.cphase(float(rng.uniform(0, 6.28)), c, t))
    if n <= 11:
        _both(engine, reg, ora, *og.grover_oracle(n, 1, list(range(1 << n))))
        _both(engine, reg, ora, *og.grover_diffusion(n))
        k = min(n, 5)
        qs = [int(x) for x in rng.permutation(n)[:k]]
        d = np.exp(1j * rng.uniform(0, 6.28, size=1 << k))
        _both(engine, reg, ora, np.diag(d), qs)
    assert_close(reg.state, ora.state, f"n={n}")


@pytest.mark.parametrize("n,k", [(2, 2), (3, 2), (3, 3), (5, 2), (5, 4), (6, 6), (8, 3), (8, 5), (10, 7), (12, 2), (12, 6)])
def test_dense_k_scrambled_targets(engine, n, k):
    """General k-qubit dense, NON-unitary matrices, targets in arbitrary order (register.rs:163-176)."""
    rng = np.random.default_rng(17 * n + k)
    reg, ora = _pair(engine, n, 500 + 13 * n + k)
    for i in range(4):
        qs = [int(x) for x in rng.permutation(n)[:k]]
        _both(engine, reg, ora, random_matrix(k, 31 * n + 7 * k + i), qs)
    assert_close(reg.state, ora.state, f"n={n} k={k}")


@pytest.mark.parametrize("n", [1, 3, 6, 11])
@pytest.mark.parametrize("force", [FORCE_GENERIC, FORCE_GENERIC_OUT_OF_PLACE])
def test_generic_kernel_matches_specialised(engine, n, force):
    """The generic kernel forced for every gate shape -- in place (K3i, the default for dense gates the tile path does not
    take) and out of place (K3, the k > 13 path) -- gives the oracle's result too."""
    rng = np.random.default_rng(n)
    reg, ora = _pair(engine, n, 600 + n, force=force)
    for t in range(n):
        _both(engine, reg, ora, random_matrix(1, t), [t])
        _both(engine, reg, ora, *og.rz(0.3 + t, t))
        if n > 1:
            c = (t + 1 + int(rng.integers(n - 1))) % n
            _both(engine, reg, ora, *og.cnot(c, t))
            _both(engine, reg, ora, *og.cz(c, t))
    assert_close(reg.state, ora.state, f"n={n}")


@pytest.mark.parametrize("n,k", [(14, 9), (14, 12), (13, 13), (18, 3), (20, 5)])
def test_wide_dense_gates_in_place(engine, n, k):
    """Dense k-qubit gates beyond the tile path go through the in-place kernel K3i (no second state buffer): scattered and
    low/high targets, with and without controls, chunks of several groups (k < 11) and of one group (k >= 11)."""
    rng = np.random.default_rng(100 * n + k)
    reg, ora = _pair(engine, n, 900 + n)
    for i in range(1 if k >= 12 else 3):
        qs = [int(x) for x in rng.permutation(n)[:k]]
        m = random_matrix(k, 17 * n + k + i)
        _both(engine, reg, ora, m, qs)
        if k <= 9 and i == 0:
            rest = [q for q in range(n) if q not in qs]
            op = og.to_controlled(og.to_controlled((m, qs), rest[0]), rest[-1])
            _both(engine, reg, ora, *op)
    assert_close(reg.state, ora.state, f"in-place dense gate n={n} k={k}")


def test_shor_and_swap_gates(engine):
    """to_quantum_gate (operation.rs:386) incl. the non-permutation u_gate of examples/shors.rs:79-92,
    its controlled form, hadamard_transform, qft and swap."""
    n = 9
    reg, ora = _pair(engine, n, 700)
    _both(engine, reg, ora, *og.hadamard_transform([0, 1, 2, 3, 4, 5]))
    _both(engine, reg, ora, *og.to_controlled(og.shor_u_gate([5, 6, 7, 8], 15, 7, 1), 2))
    _both(engine, reg, ora, *og.to_controlled(og.shor_u_gate([5, 6, 7, 8], 13, 6, 0), 0))   # not a permutation
    _both(engine, reg, ora, *og.qft(6))
    _both(engine, reg, ora, *og.swap(0, 8))
    _both(engine, reg, ora, *og.swap(3, 1))
    assert_close(reg.state, ora.state)


def test_init_states_bit_exact(engine):
    """Register::new / from_int (register.rs:41-63) and try_from_qubits (:81-102) are bit-exact."""
    for n in [1, 2, 5, 10]:
        for k in [0, 1, (1 << n) - 1, (1 << n) // 3]:
            want = OracleRegister.from_int(n, k).state
            assert (engine.Register.from_int(n, k).state == want).all()
    bits = [True, False, True, True, False, False, True]
    assert (engine.Register.new(bits).state == OracleRegister.new(bits).state).all()
    rng = np.random.default_rng(5)
    for n in [1, 2, 3, 8, 12]:
        qs = []
        for _ in range(n):
            v = rng.normal(size=2) + 1j * rng.normal(size=2)
            qs.append((v / np.linalg.norm(v)).reshape(2, 1))
        got = engine.Register.try_from_qubits(qs).state
        want = OracleRegister.try_from_qubits(qs).state
        assert (got == want).all(), f"product state n={n} not bit-exact"


def test_from_qubits_errors(engine):
    good = np.array([[1.0], [0.0]], dtype=complex)
    bad = np.array([[1.0], [0.5]], dtype=complex)
    with pytest.raises(engine.OperationError) as e:
        engine.Register.try_from_qubits([good, good, bad])
    assert e.value == engine.OperationError("InvalidQubit", 2)
    with pytest.raises(engine.OperationError) as e:
        engine.Register.try_from_qubits([good, np.array([[1.0], [0.0], [0.0]])])
    assert e.value == engine.OperationError("InvalidQubit", 1)
    with pytest.raises(RuntimeError, match="Incorrect input qubits"):
        engine.Register.from_qubits([bad])


def test_apply_errors_match_reference_order(engine):
    """register.rs:146-153: smallest duplicate first, then first out-of-range target in `targets` order."""
    reg = engine.Register.new([False, False])
    with pytest.raises(engine.OperationError) as e:
        reg.try_apply(engine.operation.cnot(1, 2))
    assert e.value == engine.OperationError("InvalidTarget", 2)
    with pytest.raises(engine.OperationError) as e:
        reg.try_apply(engine.operation.cnot(1, 1))
    assert e.value == engine.OperationError("InvalidTarget", 1)
    reg.try_apply(engine.operation.cnot(0, 1))
    reg5 = engine.Register.new([False] * 5)
    m = np.eye(8, dtype=complex)
    for targets in [[7, 9, 9], [9, 7, 3], [6, 5, 6], [0, 4, 5]]:
        ora = OracleRegister.new([False] * 5)
        with pytest.raises(Exception) as eo:
            ora.try_apply_fast((m, targets))
        with pytest.raises(engine.OperationError) as eg:
            reg5.try_apply(engine.Operation(m, targets))
        assert (eg.value.kind, eg.value.payload) == (eo.value.kind, eo.value.payload)
    with pytest.raises(RuntimeError, match="Could not apply operation"):
        reg5.apply(engine.operation.hadamard(5))
    with pytest.raises(engine.OperationError) as e:
        reg5.try_measure(5)
    assert e.value == engine.OperationError("InvalidTarget", 5)
    with pytest.raises(RuntimeError, match="Could not measure bit"):
        reg5.measure(17)
    with pytest.raises(engine.OperationError) as e:
        reg5.try_apply_all(engine.operation.cnot(0, 1))
    assert e.value == engine.OperationError("InvalidArity", 2)
    # the failed calls must not have touched the state
    assert (reg5.state == OracleRegister.new([False] * 5).state).all()


@pytest.mark.parametrize("n", [1, 2, 5, 9, 14, 15, 18])
def test_probabilities_bit_exact_tree(engine, n):
    """qb_probabilities is a balanced adjacent-pair tree: bit-identical to the numpy restatement, and
    within rounding of the reference's sequential sum (register.rs:303-315)."""
    reg, ora = _pair(engine, n, 800 + n)
    for t in sorted({0, 1, n // 2, n - 1} & set(range(n))):
        got = reg.probabilities(t)
        assert got == tree_probabilities(ora.state, t), f"n={n} t={t}"
        seq = ora.probabilities(t)
        assert abs(got[0] - seq[0]) < 1e-13 and abs(got[1] - seq[1]) < 1e-13
    assert reg.norm2() == tree_probabilities(ora.state, None)[0]


@pytest.mark.parametrize("n", [1, 2, 3, 6, 10, 15])
def test_measure_outcomes_and_collapse(engine, n):
    """try_measure (register.rs:297-342) under a shared seeded uniform stream: outcomes identical,
    collapsed + renormalised states within tolerance; chained over every qubit."""
    rng = np.random.default_rng(900 + n)
    reg, ora = _pair(engine, n, 900 + n)
    for t in rng.permutation(n):
        u = float(rng.random())
        assert reg.measure(int(t), u=u) == ora.measure(int(t), u)
        assert_close(reg.state, ora.state, f"n={n} after measuring {t}")
    # fully collapsed: a basis state with unit norm
    assert abs(reg.norm2() - 1.0) < 1e-12


def test_measure_deterministic_edges(engine):
    """u > p0 is strict (register.rs:321): on |0> even u = 1-eps gives false; on |1> (p0 = 0) the smallest
    positive u gives true."""
    reg = engine.Register.new([False, True])
    assert reg.measure(0, u=np.nextafter(1.0, 0.0)) is False
    assert reg.measure(1, u=5e-324) is True
    assert (reg.state == OracleRegister.new([False, True]).state).all()


def test_apply_all_matches_per_qubit_and_reference(engine):
    """try_apply_all (register.rs:267-281): literal kron oracle for small n, per-qubit for larger."""
    for n in [1, 2, 5, 8, 12]:
        reg, ora = _pair(engine, n, 1000 + n)
        m = random_matrix(1, n, unitary=True)
        reg.apply_all(engine.Operation(m, [0]))
        ora.apply_all((m, [0]), literal=(n <= 8))
        assert_close(reg.state, ora.state, f"n={n}")


def test_clone_eq_and_state_io(engine):
    reg, ora = _pair(engine, 10, 1100)
    c = reg.clone()
    assert c == reg
    _both(engine, reg, ora, *og.hadamard(3))
    assert not (c == reg)
    assert c.max_abs_diff(reg) > 1e-3
    part = reg.read(100, 57)
    assert (part == reg.state[100:157]).all()
    reg.set_state(np.zeros(8, dtype=complex), offset=16)
    s = reg.state
    assert (s[16:24] == 0).all() and s[15] != 0
    with pytest.raises(IndexError):
        reg.read(1000, 100)
    with pytest.raises(IndexError):
        engine.Register.from_int(3, 8)


@pytest.mark.parametrize("n", [20, 22])
def test_larger_registers(engine, n):
    """A layer of H/RZ on every qubit + CNOT/CZ on a random matching at n = 20, 22 vs the oracle."""
    rng = np.random.default_rng(n)
    reg, ora = _pair(engine, n, 1200 + n)
    for q in range(n):
        g = og.hadamard(q) if rng.random() < 0.5 else og.rz(float(rng.uniform(0, 6.28)), q)
        _both(engine, reg, ora, *g)
    perm = [int(x) for x in rng.permutation(n)]
    for a, b in zip(perm[0::2], perm[1::2]):
        _both(engine, reg, ora, *(og.cnot(a, b) if rng.random() < 0.5 else og.cz(a, b)))
    assert_close(reg.state, ora.state, f"n={n}")
    u = float(rng.random())
    assert reg.measure(n - 1, u=u) == ora.measure(n - 1, u)
    assert_close(reg.state, ora.state, f"n={n} collapsed")


@pytest.mark.parametrize("fusion", [0, 1])
def test_apply_controlled_equals_dense_controlled_matrix(engine, fusion):
    """qb_apply_controlled == the reference's dense construction (to_controlled / cnx / cnz, operation.rs:406-536)."""
    rng = np.random.default_rng(4)
    for n in [3, 6, 9, 13]:
        reg, ora = _pair(engine, n, 1300 + n, fusion=fusion)
        for _ in range(12):
            nc = int(rng.integers(1, min(n, 6)))
            k = int(rng.integers(1, min(n - nc, 2) + 1))
            qs = [int(x) for x in rng.permutation(n)[:nc + k]]
            targets, controls = qs[:k], qs[k:]
            m = random_matrix(k, int(rng.integers(1 << 30)), unitary=True)
            reg.apply_controlled(engine.Operation(m, targets), controls)
            big = (m, targets)
            for c in controls:
                big = og.to_controlled(big, c)
            ora.apply(big)
        # many controls: Z controlled by all other qubits == cnz
        reg.apply_controlled(engine.operation.pauli_z(0), list(range(1, n)))
        ora.apply(og.cnz(list(range(1, n)), 0))
        assert_close(reg.state, ora.state, f"n={n}")
    reg5 = engine.Register.new([False] * 5)
    with pytest.raises(RuntimeError, match="Could not apply operation"):
        reg5.apply_controlled(engine.operation.not_(1), [1])            # control == target
    with pytest.raises(RuntimeError, match="Could not apply operation"):
        reg5.apply_controlled(engine.operation.not_(1), [7])
