// C++ host-side mirror (include/quaru_b200.hpp) exercised with a handful of the reference's own tests
// (tests/register_operation_tests.rs, src/register.rs:437-445, src/operation.rs:565-594).  Built by
// tests/test_gpu_cpp.py with g++ and linked against libquaru_b200.so; `--host-only` runs the part that needs no GPU.
#include <cstdio>
#include <cstring>

#include "quaru_b200.hpp"

using namespace quaru;
namespace op = quaru::operation;

#define REQUIRE(c)                                                         \
    do {                                                                   \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

static int host_only() {
    // sz_matches + toffoli2_equals_cnot (operation.rs:565-594)
    std::vector<Operation> all = {op::identity(0), op::hadamard(0), op::cnot(0, 1), op::swap(0, 1), op::phase(0), op::not_(0),
                                  op::pauli_y(0), op::pauli_z(0), op::cnx({0}, 1), op::cnx({0, 1}, 2), op::cnx({0, 1, 2, 3, 4}, 5)};
    for (auto& o : all) REQUIRE(o.matrix().size() == o.side() * o.side());
    REQUIRE(op::cnx({0}, 1).matrix() == op::cnot(0, 1).matrix());
    REQUIRE((op::cnot(3, 1).targets() == std::vector<size_t>{1, 3}));
    REQUIRE(!Operation::make({1, 0, 0, 1}, {0, 1}).has_value());
    REQUIRE(op::hadamard_transform({0, 1, 2}).side() == 8);
    REQUIRE(op::to_controlled(op::hadamard(1), 4).targets().back() == 4);
    QuantumCircuit c;
    c.add_operation(op::hadamard(0));
    c.add_measurement(0);
    bool threw = false;
    try { c.add_operation(op::not_(0)); } catch (const std::runtime_error&) { threw = true; }
    REQUIRE(threw);
    std::printf("HOST_OK\n");
    return 0;
}

int main(int argc, char** argv) {
    if (host_only()) return 1;
    if (argc > 1 && !std::strcmp(argv[1], "--host-only")) return 0;

    {   // invalid_target_returns_error (register.rs:437-445)
        Register r(std::vector<bool>{false, false});
        try { r.try_apply(op::cnot(1, 2)); REQUIRE(false); }
        catch (const OperationError& e) { REQUIRE(e.kind == OperationError::InvalidTarget && e.payload == 2); }
        try { r.try_apply(op::cnot(1, 1)); REQUIRE(false); } catch (const OperationError&) {}
        r.try_apply(op::cnot(0, 1));
    }
    {   // hadamard_hadamard_retains_original_state + chaining (tests :98-107)
        Register r(std::vector<bool>{false, false, false});
        Register expected = r;
        r.apply(op::hadamard(1)).apply(op::hadamard(1));
        REQUIRE(r == expected);
    }
    {   // first_bell_state_measure_equal (tests :127-140)
        for (size_t i = 0; i < 5; ++i) {
            Register r(std::vector<bool>(6, false));
            r.seed(17 + i);
            r.apply(op::hadamard(i)).apply(op::cnot(i, i + 1));
            REQUIRE(r.measure(i) == r.measure(i + 1));
        }
    }
    {   // toffoli truth table on basis states (tests :183-208)
        for (unsigned bits = 0; bits < 8; ++bits) {
            Register r(std::vector<bool>{bool(bits & 1), bool(bits & 2), bool(bits & 4)});
            r.apply(op::cnx({0, 1}, 2));
            const bool want = ((bits & 3) == 3) ? !(bits & 4) : bool(bits & 4);
            REQUIRE(r.try_measure_with(2, 0.5) == want);
        }
    }
    {   // apply_all_test (tests :210-219) + apply_all panics on a binary operation
        Register a(std::vector<bool>(5, false)), b(std::vector<bool>(5, false));
        for (size_t i = 0; i < 5; ++i) a.apply(op::hadamard(i));
        b.apply_all(op::hadamard(0));
        REQUIRE(a == b);
        bool threw = false;
        try { b.apply_all(op::cnot(0, 1)); } catch (const std::runtime_error& e) { threw = std::string(e.what()) == "Could not apply operation"; }
        REQUIRE(threw);
    }
    {   // from_qubits_test (tests :25-40) and a circuit with a conditional operation
        const double f = std::sqrt(0.5);
        Register q = Register::try_from_qubits({{1, 0}, {0, 1}, {f, f}});
        Register r(std::vector<bool>{false, true, false});
        r.apply(op::hadamard(2));
        REQUIRE(r == q);
        QuantumCircuit c;
        c.add_operation(op::not_(0));
        c.add_measurement(0);
        c.add_conditional_operation(op::not_(1), {{0, true}});
        Register t(std::vector<bool>{false, false});
        auto m = t.apply_circuit(c);
        REQUIRE(m.size() == 1 && m[0].second);
        REQUIRE(std::abs(t.state()[3] - c64(1, 0)) < 1e-12);
    }
    {   // raw state files: save -> load is bit-exact, a missing file is an engine error
        Register a(std::vector<bool>(9, false));
        for (size_t i = 0; i < 9; ++i) a.apply(openqasm::u(0.3 + 0.1 * double(i), 0.7, 1.1, i));
        const std::string path = std::string(argc > 2 ? argv[2] : "/tmp") + "/cpp_state.qbstate";
        a.save_state(path);
        Register b(std::vector<bool>(9, true));
        b.load_state(path);
        REQUIRE(a.state() == b.state());
        std::remove(path.c_str());
        bool threw = false;
        try { b.load_state(path); } catch (const std::runtime_error&) { threw = true; }
        REQUIRE(threw);
    }
    std::printf("CPP_OK\n");
    return 0;
}
