"""The reference's own tests for the hot path, ported once and run against two backends:
the CPU oracle (tests/test_oracle_reference_suite.py -- this is what PINS the oracle) and the CUDA engine
(tests/test_gpu_reference_suite.py).  Sources (relative to /root/reference):
  tests/register_operation_tests.rs, src/register.rs:437-445, src/operation.rs:565-594,
  examples/grovers.rs:267-312, examples/shors.rs:375-504, examples/quantum_teleportation.rs:142-179,
  examples/superdense_coding.rs:83-97, src/math.rs:91-122.
The reference draws measurement randomness from an unseeded thread_rng and uses proptest; here every random
choice comes from a seeded numpy generator and proptest's ranges are enumerated exhaustively where small.
"""
import itertools
import math

import numpy as np

from oracle import gates as og
from oracle.oracle import OperationError as OracleOperationError
from oracle.oracle import OracleRegister


# ------------------------------------------------------------------------------------------------
# backends
# ------------------------------------------------------------------------------------------------
class OracleBackend:
    name = "oracle"
    OperationError = OracleOperationError

    def op(self, g):
        return g

    def new(self, bits):
        return _OracleReg(OracleRegister.new(bits))

    def from_int(self, n, k):
        return _OracleReg(OracleRegister.from_int(n, k))

    def try_from_qubits(self, qs):
        return _OracleReg(OracleRegister.try_from_qubits(qs))


class _OracleReg:
    def __init__(self, r):
        self.r = r

    def apply(self, g):
        self.r.apply(g)
        return self

    def try_apply(self, g):
        self.r.try_apply_fast(g)
        return self

    def apply_all(self, g):
        self.r.apply_all(g)
        return self

    def measure(self, t, u):
        return self.r.measure(t, u)

    def apply_circuit(self, ops, measurement_targets, conditional_ops, us):      # register.rs:345-377
        return self.r.apply_circuit(ops, measurement_targets, conditional_ops, us)

    def state(self):
        return self.r.state.copy()

    def set_state(self, s):
        self.r.state[:] = np.asarray(s).reshape(-1)

    def size(self):
        return self.r.size()

    def clone(self):
        return _OracleReg(self.r.clone())

    def __eq__(self, other):
        return self.r == other.r

    __hash__ = None


class EngineBackend:
    name = "engine"

    def __init__(self, engine, fusion=1):
        self.e = engine
        self.fusion = fusion
        self.OperationError = engine.OperationError

    def op(self, g):
        return self.e.Operation(g[0], g[1])

    def _wrap(self, r):
        r.set_option(1, self.fusion)
        return _EngineReg(self, r)

    def new(self, bits):
        return self._wrap(self.e.Register.new(bits))

    def from_int(self, n, k):
        return self._wrap(self.e.Register.from_int(n, k))

    def try_from_qubits(self, qs):
        return self._wrap(self.e.Register.try_from_qubits(qs))


class _EngineReg:
    def __init__(self, b, r):
        self.b, self.r = b, r

    def apply(self, g):
        self.r.apply(self.b.op(g))
        return self

    def try_apply(self, g):
        self.r.try_apply(self.b.op(g))
        return self

    def apply_all(self, g):
        self.r.apply_all(self.b.op(g))
        return self

    def measure(self, t, u):
        return self.r.measure(t, u=u)

    def apply_circuit(self, ops, measurement_targets, conditional_ops, us):      # register.rs:345-377
        circ = self.b.e.QuantumCircuit()                                         # the builder of operation.rs:276-321
        for g in ops:
            circ.add_operation(self.b.op(g))
        for t in measurement_targets:
            circ.add_measurement(t)
        for g, conds in conditional_ops:
            circ.add_conditional_operation(self.b.op(g), conds)
        return self.r.apply_circuit(circ, us)

    def state(self):
        return self.r.state

    def set_state(self, s):
        self.r.set_state(s)

    def size(self):
        return self.r.size()

    def clone(self):
        return _EngineReg(self.b, self.r.clone())

    def __eq__(self, other):
        return self.r == other.r

    __hash__ = None


def equal_qubits(a, b):
    """math.rs:21-29"""
    return bool((np.abs(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)) < 1e-8).all())


# ------------------------------------------------------------------------------------------------
# tests/register_operation_tests.rs
# ------------------------------------------------------------------------------------------------
def t_measure_on_zero_state_gives_false(B, rng):                      # :13-19
    assert B.new([False]).measure(0, rng.random()) is False


def t_from_qubits_test(B, rng):                                       # :25-40
    f = 1.0 / math.sqrt(2.0)
    reg_qubits = B.try_from_qubits([np.array([[1.0], [0.0]], dtype=complex), np.array([[0.0], [1.0]], dtype=complex),
                                    np.array([[f], [f]], dtype=complex)])
    reg = B.new([False, True, False])
    reg.apply(og.hadamard(2))
    assert reg == reg_qubits


def t_not_on_arbitrary_qubit(B, rng):                                 # :43-50
    for i in range(3):
        reg = B.new([False, False, False])
        reg.apply(og.not_(i))
        assert reg.measure(i, rng.random()) is True


def t_measure_basis_states_gives_same_result(B, rng):                 # :52-96 (4 and 8 qubits)
    for n, cases in [(4, list(itertools.product([False, True], repeat=4))),
                     (8, [tuple(bool(x) for x in rng.integers(0, 2, size=8)) for _ in range(24)])]:
        for bits in cases:
            reg = B.new(list(bits))
            got = [reg.measure(q, rng.random()) for q in range(n)]
            assert got == list(bits)


def t_hadamard_hadamard_retains_original_state(B, rng):               # :98-107
    for i in range(3):
        reg = B.new([False, False, False])
        expected = reg.clone()
        h = og.hadamard(i)
        reg.apply(h).apply(h)
        assert reg == expected


def t_any_entangled_bit_measure_same(B, rng):                         # :109-125
    for i in range(6):
        for j in range(6):
            jj = (j + 1) % 6 if i == j else j
            reg = B.new([False] * 6)
            reg.apply(og.hadamard(i))
            reg.apply(og.cnot(i, jj))
            assert reg.measure(i, rng.random()) == reg.measure(jj, rng.random())


def t_first_bell_state_measure_equal(B, rng):                         # :127-140
    for i in range(5):
        reg = B.new([False] * 6)
        reg.apply(og.hadamard(i))
        reg.apply(og.cnot(i, i + 1))
        assert reg.measure(i, rng.random()) == reg.measure(i + 1, rng.random())


def t_arbitrary_unary_applied_twice_gives_equal(B, rng):              # :142-152
    for i in range(6):
        for g in (og.hadamard(i), og.not_(i), og.identity(i)):
            reg = B.new([False] * 6)
            expected = reg.clone()
            reg.apply(g)
            reg.apply(g)
            assert reg == expected


def t_arbitrary_binary_applied_twice_gives_equal(B, rng):             # :154-163
    for i in range(6):
        for j in range(6):
            if i == j:
                continue
            for g in (og.cnot(i, j), og.swap(i, j)):
                reg = B.new([False] * 6)
                expected = reg.clone()
                reg.apply(g)
                reg.apply(g)
                assert reg == expected


def t_swap_single_true_qubit(B, rng):                                 # :165-181
    for i in range(5):
        for j in range(5):
            if i == j:
                continue
            qubits = [False] * 6
            qubits[i] = True
            reg = B.new(qubits)
            reg.apply(og.swap(i, j))
            for k in range(5):
                assert reg.measure(k, rng.random()) == (k == j)


def t_toffoli_test(B, rng):                                           # :183-208
    for n in range(2, 7):
        for init in itertools.product([False, True], repeat=6):
            reg = B.new(list(init))
            init_target = init[n - 1]
            reg.apply(og.cnx(list(range(n - 1)), n - 1))
            control_measure = all(reg.measure(i, rng.random()) for i in range(n - 1))   # short-circuits like Rust's all()
            target_measure = reg.measure(n - 1, rng.random())
            assert (target_measure != init_target) if control_measure else (target_measure == init_target)


def t_apply_all_test(B, rng):                                         # :210-219
    for n in range(2, 7):
        reg1 = B.new([False] * n)
        reg2 = B.new([False] * n)
        for i in range(reg1.size()):
            reg1.apply(og.hadamard(i))
        reg2.apply_all(og.hadamard(0))
        assert reg1 == reg2


def t_apply_all_panics_if_not_unary(B, rng):                          # :221-226
    for g in (og.cnot(0, 1), og.swap(2, 4)):
        reg = B.new([False] * 6)
        try:
            reg.apply_all(g)
        except RuntimeError as e:
            assert "Could not apply operation" in str(e)
        else:
            raise AssertionError("apply_all accepted a binary operation")


# ------------------------------------------------------------------------------------------------
# src/register.rs:437-445
# ------------------------------------------------------------------------------------------------
def t_invalid_target_returns_error(B, rng):
    reg = B.new([False, False])
    try:
        reg.try_apply(og.cnot(1, 2))
        raise AssertionError("no error")
    except B.OperationError as e:
        assert (e.kind, e.payload) == ("InvalidTarget", 2)
    try:
        reg.try_apply(og.cnot(1, 1))
        raise AssertionError("no error")
    except B.OperationError:
        pass
    reg.try_apply(og.cnot(0, 1))


# ------------------------------------------------------------------------------------------------
# examples/grovers.rs:267-312
# ------------------------------------------------------------------------------------------------
def _grover_regsize(target):
    regsize = math.ceil(math.log2(target + 1)) if target + 1 > 1 else 0
    return max(regsize, 2)


def t_grover_oracle_operation(B, rng):                                # :272-286
    for target in range(20):
        regsize = _grover_regsize(target)
        reg = B.new([False] * regsize)
        reg.apply_all(og.hadamard(0))
        reg.apply(og.grover_oracle(regsize, target, list(range(20))))
        st = reg.state()
        assert np.signbit(st[target].real)
        others = np.delete(st, target)
        assert not np.signbit(others.real).any()


def t_grover_diffuser_operation(B, rng):                              # :292-311
    for target in range(20):
        regsize = _grover_regsize(target)
        reg = B.new([False] * regsize)
        reg.apply_all(og.hadamard(0))
        reg.apply(og.grover_oracle(regsize, target, list(range(20))))
        reg.apply_all(og.hadamard(0))                                  # diffuser(), grovers.rs:172-176
        reg.apply(og.grover_diffusion(regsize))
        reg.apply_all(og.hadamard(0))
        st = np.abs(reg.state())
        for i in range(reg.size()):                                    # (sic) the reference loops over reg.size()
            if i != target:
                assert st[i] < st[target]


def grover_full(B, rng, n=7, target=64):
    """grovers_algorithm (grovers.rs:121-168): BASELINE config C1.  Returns (pre-measurement state, bits MSB..LSB)."""
    lst = list(range(1 << n))
    reg = B.new([False] * n)
    reg.apply_all(og.hadamard(0))
    oracle = og.grover_oracle(n, target, lst)
    iters = int(math.floor(math.pi / 4.0 * math.sqrt((1 << n) / 1)))
    for _ in range(iters):
        reg.apply(oracle)
        reg.apply_all(og.hadamard(0))
        reg.apply(og.grover_diffusion(n))
        reg.apply_all(og.hadamard(0))
    pre = reg.state()
    bits = [reg.measure(i, rng.random()) for i in reversed(range(n))]
    return pre, bits


def t_grover_finds_target(B, rng):
    pre, bits = grover_full(B, rng)
    assert np.argmax(np.abs(pre)) == 64
    assert abs(pre[64]) ** 2 > 0.99
    assert sum(b << i for i, b in zip(reversed(range(7)), bits)) == 64


# ------------------------------------------------------------------------------------------------
# examples/shors.rs:375-504
# ------------------------------------------------------------------------------------------------
def find_period_slow(B, rng, number, a):                              # shors.rs:180-234
    n = math.ceil(math.log2(number + 1))
    reg = B.new([False] * (3 * n))
    reg.apply(og.hadamard_transform(list(range(2 * n))))
    reg.apply(og.not_(2 * n))
    targets = list(range(2 * n, 3 * n))
    for i in range(2 * n):
        reg.apply(og.to_controlled(og.shor_u_gate(targets, number, a, i), i))
    reg.apply(og.qft(2 * n))
    res = 0
    for i in range(2 * n):
        res |= int(reg.measure(i, rng.random())) << i
    return og.limit_denominator(res, 2 ** (2 * n) - 1, number - 1)[1]


def find_period_fast(B, rng, number, a):                              # shors.rs:287-361
    n = math.ceil(math.log2(number + 1))
    reg = B.new([False] * (n + 1))
    reg.apply(og.not_(1))
    res, sm = 0, 0.0
    targets = list(range(1, n + 1))
    for i in range(2 * n):
        reg.apply(og.hadamard(0))
        reg.apply(og.to_controlled(og.shor_u_gate(targets, number, a, 2 * n - 1 - i), 0))
        phi = np.exp(1j * (-2.0 * math.pi * sm))
        reg.apply((np.array([[1, 0], [0, phi]], dtype=complex), [0]))
        reg.apply(og.hadamard(0))
        m = reg.measure(0, rng.random())
        sm /= 2.0
        if m:
            reg.apply(og.not_(0))
            res |= 1 << i
            sm += 0.25
    return og.limit_denominator(res, 2 ** (2 * n) - 1, number - 1)[1]


def t_period_finder_working(B, rng, trials=20, need=15):              # :382-421
    for fast in (False, True):
        for number in range(2, 15 if fast else 6):
            for a in range(2, number):
                if math.gcd(number, a) != 1:
                    continue
                period, a_pow = 1, a
                while a_pow != 1:
                    a_pow = a_pow * a % number
                    period += 1
                ok = 0
                for _ in range(trials):
                    r = find_period_fast(B, rng, number, a) if fast else find_period_slow(B, rng, number, a)
                    if period % r == 0:
                        ok += 1
                assert ok >= need, (fast, number, a, ok)


def t_test_u_gate(B, rng, stride=1):                                  # :433-463 (exhaustive for stride 1)
    n = 4
    cases = list(itertools.product(range(1, 1 << n), range(1 << n), range(1 << n), range(1 << n)))
    for modulus, base, i, inp in cases[::stride]:
        reg = B.from_int(n, inp)
        reg.apply(og.shor_u_gate(list(range(n)), modulus, base, i))
        want = B.from_int(n, inp * og.modpow(base, 1 << i, modulus) % modulus)
        assert equal_qubits(reg.state(), want.state())


def _add_operation(a, targets):                                       # :465-483
    m = np.eye(1, dtype=complex)
    for i in range(len(targets)):
        em = np.array([[1, 0], [0, np.exp(1j * (math.pi / 2.0 ** i * a))]], dtype=complex)
        m = np.kron(m, em)
    return m, list(targets)


def t_qft_add(B, rng, limit=50):                                      # :485-503
    n = 5
    qft_gate = og.qft(n)
    for a in range(limit):
        for b in range(limit):
            reg1 = B.from_int(n, b % (1 << n))
            reg1.apply(qft_gate)
            reg1.apply(_add_operation(a, list(range(n))))
            reg2 = B.from_int(n, (a + b) % (1 << n))
            reg2.apply(qft_gate)
            assert equal_qubits(reg1.state(), reg2.state())


# ------------------------------------------------------------------------------------------------
# examples/quantum_teleportation.rs:83-179, examples/superdense_coding.rs:28-97
# ------------------------------------------------------------------------------------------------
def _state_of_qubit(state, n):                                        # get_state_of_qubit
    idx = np.arange(state.shape[0])
    beta = state[((idx >> n) & 1) == 1].sum()
    alpha = state[((idx >> n) & 1) == 0].sum()
    prob = abs(alpha) ** 2 + abs(beta) ** 2
    return np.array([alpha, beta]) / math.sqrt(prob)


def t_quantum_teleportation(B, rng, trials=20):                       # :142-149, protocol :83-99
    for _ in range(trials):
        v = rng.normal(size=2) + 1j * rng.normal(size=2)
        q0 = v / np.linalg.norm(v)
        reg = B.new([False] * 3)
        reg.set_state(np.kron(np.kron([1, 0], [1, 0]), q0))            # kron(kron(|0>,|0>), q0)
        reg.apply(og.hadamard(2))
        reg.apply(og.cnot(2, 1))
        reg.apply(og.cnot(0, 1))
        reg.apply(og.hadamard(0))
        if reg.measure(1, rng.random()):
            reg.apply(og.not_(2))
        if reg.measure(0, rng.random()):
            reg.apply(og.pauli_z(2))
        assert equal_qubits(_state_of_qubit(reg.state(), 2), q0)


def t_qantum_teleportation_circuit(B, rng, trials=20):                # :150-157 (sic), circuit :100-118, driver :60-74
    for _ in range(trials):
        v = rng.normal(size=2) + 1j * rng.normal(size=2)
        q0 = v / np.linalg.norm(v)
        reg = B.new([False] * 3)
        reg.set_state(np.kron(np.kron([1, 0], [1, 0]), q0))
        ops = [og.hadamard(2), og.cnot(2, 1), og.cnot(0, 1), og.hadamard(0)]
        conditional = [(og.not_(2), [(1, True)]), (og.pauli_z(2), [(0, True)])]
        measures = reg.apply_circuit(ops, [1, 0], conditional, [rng.random(), rng.random()])
        assert [t for t, _ in measures] == [1, 0]                      # insertion order (register.rs:369-375)
        assert equal_qubits(_state_of_qubit(reg.state(), 2), q0)


def t_superdense_coding(B, rng):                                      # :83-97
    for a in (False, True):
        for b in (False, True):
            reg = B.new([False, False])
            reg.apply(og.hadamard(0))
            reg.apply(og.cnot(0, 1))
            if b:
                reg.apply(og.not_(0))
            if a:
                reg.apply(og.pauli_z(0))
            reg.apply(og.cnot(0, 1))
            reg.apply(og.hadamard(0))
            assert (reg.measure(0, rng.random()), reg.measure(1, rng.random())) == (a, b)


ALL = [t_measure_on_zero_state_gives_false, t_from_qubits_test, t_not_on_arbitrary_qubit,
       t_measure_basis_states_gives_same_result, t_hadamard_hadamard_retains_original_state,
       t_any_entangled_bit_measure_same, t_first_bell_state_measure_equal, t_arbitrary_unary_applied_twice_gives_equal,
       t_arbitrary_binary_applied_twice_gives_equal, t_swap_single_true_qubit, t_toffoli_test, t_apply_all_test,
       t_apply_all_panics_if_not_unary, t_invalid_target_returns_error, t_grover_oracle_operation,
       t_grover_diffuser_operation, t_grover_finds_target, t_qft_add, t_quantum_teleportation, t_qantum_teleportation_circuit,
       t_superdense_coding]
