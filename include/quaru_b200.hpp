// quaru_b200.hpp -- header-only C++17 mirror of Quaru's host-side API over the C ABI (quaru_b200.h).
//
// The reference is Rust (no toolchain in this image), so the compiled-language host side is C++:
//   quaru::Operation + gate constructors   <->  src/operation.rs:49-70, 340-536 and openqasm::u (src/openqasm.rs:145-163)
//   quaru::Register                        <->  src/register.rs:23-413
//   quaru::OperationError                  <->  src/register.rs:10-20
//   quaru::QuantumCircuit (builder part)   <->  src/operation.rs:79-109, 271-321
// Same names, argument order, target conventions and error behaviour; Rust panics become
// std::runtime_error with the reference's message.  The state lives on the GPU: `state()` downloads it.
#pragma once

#include <cmath>
#include <complex>
#include <cstdint>
#include <functional>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "quaru_b200.h"

namespace quaru {

using c64 = std::complex<double>;  // math::c64 (src/math.rs:86)

// struct Operation {matrix, targets} (operation.rs:49-52); matrix row-major 2^k x 2^k
class Operation {
  public:
    static std::optional<Operation> make(std::vector<c64> matrix, std::vector<size_t> targets) {  // Operation::new (:60-69)
        const size_t side = size_t(1) << targets.size();
        if (matrix.size() != side * side) return std::nullopt;
        return Operation(std::move(matrix), std::move(targets));
    }
    const std::vector<c64>& matrix() const { return matrix_; }
    const std::vector<size_t>& targets() const { return targets_; }
    size_t arity() const { return targets_.size(); }
    size_t side() const { return size_t(1) << targets_.size(); }
    bool operator==(const Operation& o) const { return targets_ == o.targets_ && matrix_ == o.matrix_; }

    Operation(std::vector<c64> m, std::vector<size_t> t) : matrix_(std::move(m)), targets_(std::move(t)) {}

  private:
    std::vector<c64> matrix_;
    std::vector<size_t> targets_;
};

namespace operation {
inline Operation identity(size_t t) { return Operation({1, 0, 0, 1}, {t}); }                                   // :340
inline Operation hadamard(size_t t) { const double f = std::sqrt(0.5); return Operation({f, f, f, -f}, {t}); }  // :350
inline Operation cnot(size_t control, size_t target) {                                                         // :372
    return Operation({1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0}, {target, control});
}
inline Operation swap(size_t a, size_t b) { return Operation({1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1}, {a, b}); }  // :427
inline Operation phase(size_t t) { return Operation({1, 0, 0, c64(0, 1)}, {t}); }                               // :443
inline Operation not_(size_t t) { return Operation({0, 1, 1, 0}, {t}); }                                        // :458
inline Operation pauli_y(size_t t) { return Operation({0, c64(0, -1), c64(0, 1), 0}, {t}); }                    // :468
inline Operation pauli_z(size_t t) { return Operation({1, 0, 0, -1}, {t}); }                                    // :481

inline std::vector<c64> kron(const std::vector<c64>& a, size_t ra, const std::vector<c64>& b, size_t rb) {
    std::vector<c64> out(ra * rb * ra * rb);
    for (size_t i = 0; i < ra; ++i)
        for (size_t j = 0; j < ra; ++j)
            for (size_t k = 0; k < rb; ++k)
                for (size_t l = 0; l < rb; ++l) out[(i * rb + k) * (ra * rb) + (j * rb + l)] = a[i * ra + j] * b[k * rb + l];
    return out;
}
inline Operation hadamard_transform(const std::vector<size_t>& targets) {                                       // :358
    std::vector<c64> m = hadamard(targets[0]).matrix();
    size_t side = 2;
    for (size_t i = 1; i < targets.size(); ++i) { m = kron(hadamard(targets[i]).matrix(), 2, m, side); side *= 2; }
    return Operation(std::move(m), targets);
}
inline Operation to_quantum_gate(const std::function<size_t(size_t)>& f, const std::vector<size_t>& targets) {   // :386
    const size_t len = size_t(1) << targets.size();
    std::vector<c64> m(len * len, 0.0);
    for (size_t c = 0; c < len; ++c) {
        const size_t v = f(c);
        if (v >= len) throw std::out_of_range("to_quantum_gate: f(c) out of range");   // reference: index panic (math.rs:46)
        m[v * len + c] = 1.0;
    }
    return Operation(std::move(m), targets);
}
inline Operation to_controlled(const Operation& op, size_t control) {                                            // :406
    const size_t old = op.side(), n = 2 * old;
    std::vector<c64> m(n * n, 0.0);
    for (size_t i = 0; i < old; ++i) m[i * n + i] = 1.0;
    for (size_t i = 0; i < old; ++i)
        for (size_t j = 0; j < old; ++j) m[(i + old) * n + (j + old)] = op.matrix()[i * old + j];
    std::vector<size_t> t = op.targets();
    t.push_back(control);
    return Operation(std::move(m), std::move(t));
}
inline Operation cnx(const std::vector<size_t>& controls, size_t target) {                                       // :491
    std::vector<size_t> t{target};
    t.insert(t.end(), controls.begin(), controls.end());
    const size_t n = size_t(1) << t.size();
    std::vector<c64> m(n * n, 0.0);
    for (size_t i = 0; i + 2 < n; ++i) m[i * n + i] = 1.0;
    m[(n - 1) * n + (n - 2)] = 1.0;
    m[(n - 2) * n + (n - 1)] = 1.0;
    return Operation(std::move(m), std::move(t));
}
inline Operation cnz(const std::vector<size_t>& controls, size_t target) {                                       // :520
    std::vector<size_t> t{target};
    t.insert(t.end(), controls.begin(), controls.end());
    const size_t n = size_t(1) << t.size();
    std::vector<c64> m(n * n, 0.0);
    for (size_t i = 0; i + 1 < n; ++i) m[i * n + i] = 1.0;
    m[(n - 1) * n + (n - 1)] = -1.0;
    return Operation(std::move(m), std::move(t));
}
}  // namespace operation

namespace openqasm {
inline Operation u(double theta, double phi, double lambda, size_t target) {                                     // openqasm.rs:145
    const c64 i(0, 1);
    return Operation({std::exp(-i * (phi + lambda) / 2.0) * std::cos(theta / 2.0), -std::exp(-i * (phi - lambda) / 2.0) * std::sin(theta / 2.0),
                      std::exp(i * (phi - lambda) / 2.0) * std::sin(theta / 2.0), std::exp(i * (phi + lambda) / 2.0) * std::cos(theta / 2.0)},
                     {target});
}
}  // namespace openqasm

// enum OperationError (register.rs:10-20)
struct OperationError : std::runtime_error {
    enum Kind { InvalidTarget, InvalidDimensions, InvalidArity, InvalidQubit } kind;
    size_t payload;
    OperationError(Kind k, size_t p) : std::runtime_error(name(k) + "(" + std::to_string(p) + ")"), kind(k), payload(p) {}
    static std::string name(Kind k) {
        switch (k) {
            case InvalidTarget: return "InvalidTarget";
            case InvalidDimensions: return "InvalidDimensions";
            case InvalidArity: return "InvalidArity";
            default: return "InvalidQubit";
        }
    }
};

// A quantum circuit: ordered gates + measurement targets + conditional operations (operation.rs:79-109, 271-321)
class QuantumCircuit {
  public:
    const std::vector<Operation>& get_operations() const { return operations_; }
    const std::vector<size_t>& get_measurement_targets() const { return measurement_targets_; }
    const std::vector<std::pair<Operation, std::vector<std::pair<size_t, bool>>>>& get_conditional_operations() const { return conditional_; }
    void clear_operations() { operations_.clear(); }
    QuantumCircuit& reset_register() { operations_.clear(); measurement_targets_.clear(); conditional_.clear(); return *this; }
    void add_operation(Operation op) {                                                                             // :276
        for (size_t t : op.targets())
            if (measured(t)) throw std::runtime_error("Cannot add operation after measurement");
        if (!conditional_.empty()) throw std::runtime_error("Cannot add operation after conditional operation (yet...)");
        operations_.push_back(std::move(op));
    }
    void add_measurement(size_t target) {                                                                          // :292
        if (measured(target)) throw std::runtime_error("Cannot add measurement after measurement");
        measurement_targets_.push_back(target);
    }
    void add_conditional_operation(Operation op, std::vector<std::pair<size_t, bool>> condition) {                 // :302
        for (size_t t : op.targets())
            if (measured(t)) throw std::runtime_error("Cannot add operation after measurement");
        for (auto& c : condition)
            if (!measured(c.first)) throw std::runtime_error("Condition must be a measurement");
        conditional_.emplace_back(std::move(op), std::move(condition));
    }

  private:
    bool measured(size_t t) const {
        for (size_t m : measurement_targets_) if (m == t) return true;
        return false;
    }
    std::vector<Operation> operations_;
    std::vector<size_t> measurement_targets_;
    std::vector<std::pair<Operation, std::vector<std::pair<size_t, bool>>>> conditional_;
};

// struct Register (register.rs:23-37) -- the amplitudes live in HBM behind the qb_reg handle
class Register {
  public:
    explicit Register(const std::vector<bool>& input_bits) : Register(input_bits.size(), fold(input_bits)) {}     // Register::new (:41)
    static Register from_int(size_t n, size_t k) { return Register(n, k); }                                        // :58
    static Register try_from_qubits(const std::vector<std::vector<c64>>& qubits) {                                 // :81
        std::vector<double> flat;
        for (size_t i = 0; i < qubits.size(); ++i) {
            if (qubits[i].size() != 2) throw OperationError(OperationError::InvalidQubit, i);                      // is_qubit (:120-126)
            for (const c64& v : qubits[i]) { flat.push_back(v.real()); flat.push_back(v.imag()); }
        }
        qb_reg* h = nullptr;
        check(qb_create_product(flat.data(), (uint32_t)qubits.size(), &h));
        return Register(h);
    }
    static Register from_qubits(const std::vector<std::vector<c64>>& qubits) {                                     // :74
        try { return try_from_qubits(qubits); } catch (const OperationError&) { throw std::runtime_error("Incorrect input qubits"); }
    }
    Register(const Register& o) : h_(nullptr), rng_(o.rng_) { check(qb_clone(o.h_, &h_)); }                         // #[derive(Clone)]
    Register(Register&& o) noexcept : h_(o.h_), rng_(std::move(o.rng_)) { o.h_ = nullptr; }
    Register& operator=(Register o) { std::swap(h_, o.h_); std::swap(rng_, o.rng_); return *this; }
    ~Register() { qb_destroy(h_); }

    Register& try_apply_fast(const Operation& op) {                                                                // :140
        std::vector<uint32_t> t(op.targets().begin(), op.targets().end());
        check(qb_apply(h_, reinterpret_cast<const double*>(op.matrix().data()), t.data(), (uint32_t)t.size()));
        return *this;
    }
    Register& try_apply(const Operation& op) { return try_apply_fast(op); }                                        // :193
    Register& apply(const Operation& op) {                                                                         // :134
        try { return try_apply_fast(op); } catch (const OperationError&) { throw std::runtime_error("Could not apply operation"); }
    }
    Register& try_apply_all(const Operation& op) {                                                                 // :267
        if (op.arity() != 1) throw OperationError(OperationError::InvalidArity, op.arity());
        check(qb_apply_all(h_, reinterpret_cast<const double*>(op.matrix().data())));
        return *this;
    }
    Register& apply_all(const Operation& op) {                                                                     // :259
        try { return try_apply_all(op); } catch (const OperationError&) { throw std::runtime_error("Could not apply operation"); }
    }
    void seed(uint64_t s) { rng_.seed(s); }
    bool try_measure(size_t target) { return try_measure_with(target, std::uniform_real_distribution<double>(0.0, 1.0)(rng_)); }  // :297
    bool try_measure_with(size_t target, double u) {
        int32_t out = 0;
        check(qb_measure(h_, (uint32_t)target, u, &out));
        return out != 0;
    }
    bool measure(size_t target) {                                                                                  // :291
        try { return try_measure(target); } catch (const OperationError&) { throw std::runtime_error("Could not measure bit"); }
    }
    std::vector<std::pair<size_t, bool>> measure_circuit(const QuantumCircuit& c) {                                // :371
        std::vector<std::pair<size_t, bool>> res;
        for (size_t t : c.get_measurement_targets()) res.emplace_back(t, measure(t));
        return res;
    }
    std::vector<std::pair<size_t, bool>> apply_circuit(const QuantumCircuit& c) {                                  // :345
        std::vector<std::vector<uint32_t>> tg;
        std::vector<qb_gate> gates;
        for (const Operation& op : c.get_operations()) tg.emplace_back(op.targets().begin(), op.targets().end());
        for (size_t i = 0; i < tg.size(); ++i)
            gates.push_back({reinterpret_cast<const double*>(c.get_operations()[i].matrix().data()), tg[i].data(), (uint32_t)tg[i].size()});
        try { check(qb_apply_batch(h_, gates.data(), gates.size())); } catch (const OperationError&) { throw std::runtime_error("Could not apply operation"); }
        std::vector<std::pair<size_t, bool>> measures;
        if (!c.get_measurement_targets().empty()) measures = measure_circuit(c);
        for (const auto& co : c.get_conditional_operations()) {
            bool all = true;
            for (const auto& cond : co.second) {
                bool found = false;
                for (const auto& m : measures) found = found || m == cond;
                all = all && found;
            }
            if (all) apply(co.first);
        }
        return measures;
    }
    std::vector<c64> state() const {                                                                               // `pub state` (:35)
        std::vector<c64> out(size_t(1) << size());
        check(qb_read_state(h_, 0, out.size(), reinterpret_cast<double*>(out.data())));
        return out;
    }
    void set_state(const std::vector<c64>& s) { check(qb_write_state(h_, 0, s.size(), reinterpret_cast<const double*>(s.data()))); }
    void save_state(const std::string& path) const { check(qb_save_state(h_, path.c_str())); }                     // raw shard file
    void load_state(const std::string& path) { check(qb_load_state(h_, path.c_str())); }
    size_t size() const { return qb_num_qubits(h_); }                                                              // :404
    bool operator==(const Register& o) const {                                                                     // PartialEq (:409-413)
        if (size() != o.size()) return false;
        double d = 0.0;
        check(qb_max_abs_diff(h_, o.h_, &d));
        return d < 1e-8;
    }
    qb_reg* handle() const { return h_; }

  private:
    explicit Register(qb_reg* h) : h_(h), rng_(std::random_device{}()) {}
    Register(size_t n, size_t k) : h_(nullptr), rng_(std::random_device{}()) { check(qb_create((uint32_t)n, k, &h_)); }
    static size_t fold(const std::vector<bool>& bits) {
        size_t k = 0;
        for (size_t q = 0; q < bits.size(); ++q) if (bits[q]) k |= size_t(1) << q;
        return k;
    }
    static void check(int rc) {
        if (rc == QB_OK) return;
        const size_t p = (size_t)qb_last_error_payload();
        switch (rc) {
            case QB_ERR_INVALID_TARGET: throw OperationError(OperationError::InvalidTarget, p);
            case QB_ERR_INVALID_DIMENSIONS: throw OperationError(OperationError::InvalidDimensions, p);
            case QB_ERR_INVALID_ARITY: throw OperationError(OperationError::InvalidArity, p);
            case QB_ERR_INVALID_QUBIT: throw OperationError(OperationError::InvalidQubit, p);
            default: throw std::runtime_error(std::string("quaru_b200: ") + qb_last_error());   // no CPU fallback
        }
    }
    qb_reg* h_;
    std::mt19937_64 rng_;
};

}  // namespace quaru
