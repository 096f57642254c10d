// qb_qasm.cu -- OpenQASM 2.0 front end: source text -> flat list of basic operations.
//
// Reference: src/openqasm.rs drives the register from `OpenQASMProgram::get_basic_operations()` of the
// third-party crate openqasm-parser =^1.0.0 (Cargo.toml:21), which is NOT vendored under /root/reference.
// This file restates that crate's observable contract from the reference's call sites
// (src/openqasm.rs:53-54, 73, 91-137; src/bin/cli.rs:172-221; src/cli_utils/display.rs:165):
//   * program.qregs / program.cregs: name -> size;
//   * get_basic_operations(): Vec<(Option<(creg name, u32)>, BasicOp)> with gate definitions already
//     inlined and register-wide statements already broadcast;
//   * BasicOp::{U(f32, f32, f32, qubit), CX(qubit, qubit), Measure(qubit, cbit), ResetQ(qubit), ResetC(cbit)}.
// Grammar: the OpenQASM 2.0 specification (Cross et al., arXiv:1707.03429, appendix A).
// PARITY UNPINNED: the reference has no test that runs a program and checks amplitudes, so
// whether the crate evaluates angle expressions in f32 throughout or in f64 and rounds at the end cannot be
// determined; here expressions are evaluated in f64 and rounded once to f32 when the U parameters are emitted.
// Extension: `include "qelib1.inc";` is served from a built-in copy of the standard header's definitions.
#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "qb_common.h"

namespace qb {

struct QasmProgram {
    std::vector<std::pair<std::string, uint32_t>> qregs, cregs;
    std::vector<qb_qasm_op> ops;
};

namespace {

enum TokKind { T_ID, T_REAL, T_INT, T_STR, T_SYM, T_EOF };
struct Tok {
    TokKind kind;
    std::string text;
    double num = 0.0;
    int line = 0;
};

struct ParseError {
    std::string msg;
    int line;
};

struct Expr {  // expression tree over gate parameters
    enum Op { NUM, PARAM, ADD, SUB, MUL, DIV, POW, NEG, FUNC } op = NUM;
    double num = 0.0;
    int param = -1;
    std::string func;
    std::shared_ptr<Expr> a, b;
};
typedef std::shared_ptr<Expr> ExprP;

struct GateCall {                 // a statement inside a gate body
    std::string name;             // "U", "CX", or a user gate
    std::vector<ExprP> params;
    std::vector<int> args;        // indices into the enclosing gate's formal qubit list
};
struct GateDef {
    std::vector<std::string> params, qargs;
    std::vector<GateCall> body;
    bool opaque = false;
};

struct Arg {  // a register reference in a top-level statement
    int reg;
    int idx;  // -1: whole register
};

static const char* QELIB1 =
    "gate u3(theta,phi,lambda) q { U(theta,phi,lambda) q; }\n"
    "gate u2(phi,lambda) q { U(pi/2,phi,lambda) q; }\n"
    "gate u1(lambda) q { U(0,0,lambda) q; }\n"
    "gate cx c,t { CX c,t; }\n"
    "gate id a { U(0,0,0) a; }\n"
    "gate u0(gamma) q { U(0,0,0) q; }\n"
    "gate x a { u3(pi,0,pi) a; }\n"
    "gate y a { u3(pi,pi/2,pi/2) a; }\n"
    "gate z a { u1(pi) a; }\n"
    "gate h a { u2(0,pi) a; }\n"
    "gate s a { u1(pi/2) a; }\n"
    "gate sdg a { u1(-pi/2) a; }\n"
    "gate t a { u1(pi/4) a; }\n"
    "gate tdg a { u1(-pi/4) a; }\n"
    "gate rx(theta) a { u3(theta,-pi/2,pi/2) a; }\n"
    "gate ry(theta) a { u3(theta,0,0) a; }\n"
    "gate rz(phi) a { u1(phi) a; }\n"
    "gate cz a,b { h b; cx a,b; h b; }\n"
    "gate cy a,b { sdg b; cx a,b; s b; }\n"
    "gate swap a,b { cx a,b; cx b,a; cx a,b; }\n"
    "gate ch a,b { h b; sdg b; cx a,b; h b; t b; cx a,b; t b; h b; s b; x b; s a; }\n"
    "gate ccx a,b,c { h c; cx b,c; tdg c; cx a,c; t c; cx b,c; tdg c; cx a,c; t b; t c; h c; cx a,b; t a; tdg b; cx a,b; }\n"
    "gate crz(lambda) a,b { u1(lambda/2) b; cx a,b; u1(-lambda/2) b; cx a,b; }\n"
    "gate cu1(lambda) a,b { u1(lambda/2) a; cx a,b; u1(-lambda/2) b; cx a,b; u1(lambda/2) b; }\n"
    "gate cu3(theta,phi,lambda) c,t { u1((lambda-phi)/2) t; cx c,t; u3(-theta/2,0,-(phi+lambda)/2) t; cx c,t; u3(theta/2,phi,0) t; }\n";

class Parser {
  public:
    explicit Parser(QasmProgram* out) : prog_(out) {}

    void parse_source(const std::string& src, bool main_file) {
        std::vector<Tok> saved_toks;
        size_t saved_pos = pos_;
        saved_toks.swap(toks_);
        lex(src);
        pos_ = 0;
        if (main_file) {
            expect_id("OPENQASM");
            Tok v = next();
            if (v.kind != T_REAL && v.kind != T_INT) fail("expected a version number after OPENQASM", v.line);
            if (fabs(v.num - 2.0) > 1e-9) fail("only OPENQASM 2.0 is supported", v.line);
            expect_sym(";");
        }
        while (peek().kind != T_EOF) statement();
        toks_.swap(saved_toks);
        pos_ = saved_pos;
    }

  private:
    QasmProgram* prog_;
    std::vector<Tok> toks_;
    size_t pos_ = 0;
    std::map<std::string, GateDef> gates_;
    std::map<std::string, int> qreg_ix_, creg_ix_;
    int depth_ = 0;

    [[noreturn]] void fail(const std::string& m, int line) { throw ParseError{m, line}; }

    void lex(const std::string& s) {
        int line = 1;
        size_t i = 0;
        while (i < s.size()) {
            const char c = s[i];
            if (c == '\n') { ++line; ++i; continue; }
            if (isspace((unsigned char)c)) { ++i; continue; }
            if (c == '/' && i + 1 < s.size() && s[i + 1] == '/') { while (i < s.size() && s[i] != '\n') ++i; continue; }
            if (isalpha((unsigned char)c) || c == '_') {
                size_t j = i;
                while (j < s.size() && (isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
                toks_.push_back({T_ID, s.substr(i, j - i), 0.0, line});
                i = j;
                continue;
            }
            if (isdigit((unsigned char)c) || (c == '.' && i + 1 < s.size() && isdigit((unsigned char)s[i + 1]))) {
                size_t j = i;
                bool real = false;
                while (j < s.size() && isdigit((unsigned char)s[j])) ++j;
                if (j < s.size() && s[j] == '.') { real = true; ++j; while (j < s.size() && isdigit((unsigned char)s[j])) ++j; }
                if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
                    size_t k = j + 1;
                    if (k < s.size() && (s[k] == '+' || s[k] == '-')) ++k;
                    if (k < s.size() && isdigit((unsigned char)s[k])) { real = true; j = k; while (j < s.size() && isdigit((unsigned char)s[j])) ++j; }
                }
                Tok t{real ? T_REAL : T_INT, s.substr(i, j - i), 0.0, line};
                t.num = strtod(t.text.c_str(), nullptr);
                toks_.push_back(t);
                i = j;
                continue;
            }
            if (c == '"') {
                size_t j = s.find('"', i + 1);
                if (j == std::string::npos) fail("unterminated string", line);
                toks_.push_back({T_STR, s.substr(i + 1, j - i - 1), 0.0, line});
                i = j + 1;
                continue;
            }
            if (c == '-' && i + 1 < s.size() && s[i + 1] == '>') { toks_.push_back({T_SYM, "->", 0.0, line}); i += 2; continue; }
            if (c == '=' && i + 1 < s.size() && s[i + 1] == '=') { toks_.push_back({T_SYM, "==", 0.0, line}); i += 2; continue; }
            if (strchr(";,()[]{}+-*/^", c)) { toks_.push_back({T_SYM, std::string(1, c), 0.0, line}); ++i; continue; }
            fail(std::string("unexpected character '") + c + "'", line);
        }
        toks_.push_back({T_EOF, "", 0.0, line});
    }

    const Tok& peek(size_t k = 0) const { return toks_[std::min(pos_ + k, toks_.size() - 1)]; }
    Tok next() { Tok t = peek(); if (pos_ < toks_.size() - 1) ++pos_; return t; }
    bool is_sym(const char* s, size_t k = 0) const { return peek(k).kind == T_SYM && peek(k).text == s; }
    bool is_id(const char* s) const { return peek().kind == T_ID && peek().text == s; }
    void expect_sym(const char* s) { Tok t = next(); if (t.kind != T_SYM || t.text != s) fail(std::string("expected '") + s + "' but found '" + t.text + "'", t.line); }
    void expect_id(const char* s) { Tok t = next(); if (t.kind != T_ID || t.text != s) fail(std::string("expected '") + s + "'", t.line); }
    std::string ident() { Tok t = next(); if (t.kind != T_ID) fail("expected an identifier but found '" + t.text + "'", t.line); return t.text; }
    uint32_t nninteger() {
        Tok t = next();
        if (t.kind != T_INT) fail("expected a non-negative integer", t.line);
        errno = 0;
        const unsigned long long v = strtoull(t.text.c_str(), nullptr, 10);
        if (errno == ERANGE || v > 0xffffffffull) fail("integer '" + t.text + "' does not fit 32 bits", t.line);
        return (uint32_t)v;
    }

    // ---- expressions ---------------------------------------------------------------------------
    // Recursion guards: the tree depth bounds the C++ stack of this parser, of eval() and of ~Expr (hostile or
    // machine-generated input: 10^5 nested parentheses, unary-minus chains, 1+1+1+... sums).
    static constexpr int kMaxExprNesting = 200;   // parentheses / function calls / unary chains
    static constexpr int kMaxExprNodes = 4096;    // operator nodes of one expression (left-deep chains)
    int expr_nesting_ = 0, expr_nodes_ = 0;
    struct NestGuard {
        Parser* p;
        NestGuard(Parser* p_, int line) : p(p_) { if (++p->expr_nesting_ > kMaxExprNesting) { --p->expr_nesting_; p->fail("expression nested too deeply", line); } }
        ~NestGuard() { --p->expr_nesting_; }
    };
    ExprP expr(const std::vector<std::string>* params) {
        const bool outermost = expr_nesting_ == 0;
        if (outermost) expr_nodes_ = 0;
        return additive(params);
    }
    ExprP mk(Expr::Op op, ExprP a, ExprP b = nullptr) {
        if (++expr_nodes_ > kMaxExprNodes) fail("expression too long", peek().line);
        auto e = std::make_shared<Expr>(); e->op = op; e->a = a; e->b = b; return e;
    }
    ExprP additive(const std::vector<std::string>* p) {
        ExprP l = multiplicative(p);
        while (is_sym("+") || is_sym("-")) { const bool add = next().text == "+"; l = mk(add ? Expr::ADD : Expr::SUB, l, multiplicative(p)); }
        return l;
    }
    ExprP multiplicative(const std::vector<std::string>* p) {
        ExprP l = unary(p);
        while (is_sym("*") || is_sym("/")) { const bool mul = next().text == "*"; l = mk(mul ? Expr::MUL : Expr::DIV, l, unary(p)); }
        return l;
    }
    ExprP unary(const std::vector<std::string>* p) {
        NestGuard g(this, peek().line);
        if (is_sym("-")) { next(); return mk(Expr::NEG, unary(p)); }
        if (is_sym("+")) { next(); return unary(p); }
        return power(p);
    }
    ExprP power(const std::vector<std::string>* p) {
        ExprP b = atom(p);
        if (is_sym("^")) { next(); return mk(Expr::POW, b, unary(p)); }  // right associative
        return b;
    }
    ExprP atom(const std::vector<std::string>* params) {
        Tok t = next();
        auto e = std::make_shared<Expr>();
        if (t.kind == T_REAL || t.kind == T_INT) { e->num = t.num; return e; }
        if (t.kind == T_SYM && t.text == "(") { ExprP in = expr(params); expect_sym(")"); return in; }
        if (t.kind == T_ID) {
            if (t.text == "pi") { e->num = M_PI; return e; }
            static const char* funcs[] = {"sin", "cos", "tan", "exp", "ln", "sqrt"};
            for (const char* f : funcs)
                if (t.text == f) { expect_sym("("); e->op = Expr::FUNC; e->func = f; e->a = expr(params); expect_sym(")"); return e; }
            if (params)
                for (size_t i = 0; i < params->size(); ++i)
                    if ((*params)[i] == t.text) { e->op = Expr::PARAM; e->param = (int)i; return e; }
            fail("unknown identifier '" + t.text + "' in expression", t.line);
        }
        fail("unexpected '" + t.text + "' in expression", t.line);
    }
    static double eval(const ExprP& e, const std::vector<double>& pv) {
        switch (e->op) {
            case Expr::NUM: return e->num;
            case Expr::PARAM: return pv[e->param];
            case Expr::ADD: return eval(e->a, pv) + eval(e->b, pv);
            case Expr::SUB: return eval(e->a, pv) - eval(e->b, pv);
            case Expr::MUL: return eval(e->a, pv) * eval(e->b, pv);
            case Expr::DIV: return eval(e->a, pv) / eval(e->b, pv);
            case Expr::POW: return pow(eval(e->a, pv), eval(e->b, pv));
            case Expr::NEG: return -eval(e->a, pv);
            default: {
                const double x = eval(e->a, pv);
                if (e->func == "sin") return sin(x);
                if (e->func == "cos") return cos(x);
                if (e->func == "tan") return tan(x);
                if (e->func == "exp") return exp(x);
                if (e->func == "ln") return log(x);
                return sqrt(x);
            }
        }
    }

    // ---- statements ------------------------------------------------------------------------------
    void statement() {
        const Tok& t = peek();
        if (t.kind != T_ID) fail("unexpected '" + t.text + "'", t.line);
        if (t.text == "include") {
            next();
            Tok f = next();
            if (f.kind != T_STR) fail("expected a file name after include", f.line);
            expect_sym(";");
            if (f.text != "qelib1.inc") fail("only the standard header \"qelib1.inc\" can be included", f.line);
            if (++depth_ > 4) fail("include nesting too deep", f.line);
            parse_source(QELIB1, false);
            --depth_;
            return;
        }
        if (t.text == "qreg" || t.text == "creg") {
            const bool q = next().text == "qreg";
            const int line = peek().line;
            std::string name = ident();
            expect_sym("[");
            const uint32_t size = nninteger();
            expect_sym("]");
            expect_sym(";");
            if (qreg_ix_.count(name) || creg_ix_.count(name) || gates_.count(name)) fail("'" + name + "' is already declared", line);
            if (size == 0) fail("register '" + name + "' has size 0", line);
            if (q) { qreg_ix_[name] = (int)prog_->qregs.size(); prog_->qregs.push_back({name, size}); }
            else { creg_ix_[name] = (int)prog_->cregs.size(); prog_->cregs.push_back({name, size}); }
            return;
        }
        if (t.text == "gate") { gate_decl(false); return; }
        if (t.text == "opaque") { gate_decl(true); return; }
        if (t.text == "barrier") { next(); while (!is_sym(";")) { if (peek().kind == T_EOF) fail("unterminated barrier", peek().line); next(); } next(); return; }
        if (t.text == "if") {
            next();
            expect_sym("(");
            const int line = peek().line;
            std::string c = ident();
            if (!creg_ix_.count(c)) fail("undeclared classical register '" + c + "' in condition", line);
            expect_sym("==");
            const uint32_t v = nninteger();
            expect_sym(")");
            qop(creg_ix_[c], v);
            return;
        }
        qop(-1, 0);
    }

    void gate_decl(bool opaque) {
        next();
        const int line = peek().line;
        std::string name = ident();
        if (gates_.count(name) || qreg_ix_.count(name) || creg_ix_.count(name) || name == "U" || name == "CX")
            fail("'" + name + "' is already declared", line);
        GateDef def;
        def.opaque = opaque;
        if (is_sym("(")) {
            next();
            if (!is_sym(")")) { def.params.push_back(ident()); while (is_sym(",")) { next(); def.params.push_back(ident()); } }
            expect_sym(")");
        }
        def.qargs.push_back(ident());
        while (is_sym(",")) { next(); def.qargs.push_back(ident()); }
        {   // formal names must be distinct (a repeated qubit argument would alias two operands of a CX in the body)
            std::vector<std::string> names = def.params;
            names.insert(names.end(), def.qargs.begin(), def.qargs.end());
            std::sort(names.begin(), names.end());
            const auto dup = std::adjacent_find(names.begin(), names.end());
            if (dup != names.end()) fail("gate '" + name + "': formal name '" + *dup + "' is used twice", peek().line);
        }
        if (opaque) { expect_sym(";"); gates_[name] = def; return; }
        expect_sym("{");
        while (!is_sym("}")) {
            const Tok& t = peek();
            if (t.kind == T_EOF) fail("unterminated gate body", t.line);
            if (t.kind == T_ID && t.text == "barrier") {
                next();
                while (!is_sym(";")) { if (peek().kind == T_EOF) fail("unterminated barrier", peek().line); next(); }
                next();
                continue;
            }
            GateCall call;
            const int cl = t.line;
            call.name = ident();
            if (call.name != "U" && call.name != "CX" && !gates_.count(call.name)) fail("gate '" + call.name + "' is not defined", cl);
            if (is_sym("(")) {
                next();
                if (!is_sym(")")) { call.params.push_back(expr(&def.params)); while (is_sym(",")) { next(); call.params.push_back(expr(&def.params)); } }
                expect_sym(")");
            }
            for (;;) {
                const int al = peek().line;
                std::string a = ident();
                int ix = -1;
                for (size_t i = 0; i < def.qargs.size(); ++i) if (def.qargs[i] == a) ix = (int)i;
                if (ix < 0) fail("'" + a + "' is not a qubit argument of gate '" + name + "'", al);
                call.args.push_back(ix);
                if (is_sym(",")) { next(); continue; }
                break;
            }
            expect_sym(";");
            check_arity(call.name, call.params.size(), call.args.size(), cl);
            def.body.push_back(call);
        }
        next();
        gates_[name] = def;
    }

    void check_arity(const std::string& name, size_t np, size_t nq, int line) {
        size_t wp, wq;
        if (name == "U") { wp = 3; wq = 1; }
        else if (name == "CX") { wp = 0; wq = 2; }
        else { const GateDef& d = gates_[name]; wp = d.params.size(); wq = d.qargs.size(); }
        if (np != wp) fail("gate '" + name + "' takes " + std::to_string(wp) + " parameter(s), " + std::to_string(np) + " given", line);
        if (nq != wq) fail("gate '" + name + "' acts on " + std::to_string(wq) + " qubit(s), " + std::to_string(nq) + " given", line);
    }

    Arg argument(bool quantum, bool allow_classical = false) {
        const int line = peek().line;
        std::string name = ident();
        Arg a{-1, -1};
        if (quantum && qreg_ix_.count(name)) a.reg = qreg_ix_[name];
        else if ((!quantum || allow_classical) && creg_ix_.count(name)) a.reg = creg_ix_[name] | (quantum ? 0x40000000 : 0);
        else fail(std::string("undeclared ") + (quantum ? "quantum" : "classical") + " register '" + name + "'", line);
        if (is_sym("[")) {
            next();
            a.idx = (int)nninteger();
            expect_sym("]");
            const bool cl = (a.reg & 0x40000000) || !quantum;
            const uint32_t size = cl ? prog_->cregs[a.reg & 0x3fffffff].second : prog_->qregs[a.reg].second;
            if ((uint32_t)a.idx >= size) fail("index " + std::to_string(a.idx) + " out of range for register '" + name + "'", line);
        }
        return a;
    }

    static constexpr size_t kMaxOps = size_t(1) << 24;   // nested gate definitions can expand exponentially
    void emit(qb_qasm_op op, int cond_reg, uint32_t cond_val) {
        if (prog_->ops.size() >= kMaxOps) fail("program expands to more than 2^24 basic operations", 0);
        op.cond_creg = cond_reg;
        op.cond_val = cond_val;
        prog_->ops.push_back(op);
    }

    // expand one gate application on concrete qubits
    void apply_gate(const std::string& name, const std::vector<double>& pv, const std::vector<std::pair<int, int>>& qubits, int cond_reg,
                    uint32_t cond_val, int line, int depth) {
        if (depth > 64) fail("gate expansion too deep (recursive definition?)", line);
        qb_qasm_op op;
        memset(&op, 0, sizeof(op));
        if (name == "U") {
            op.kind = QB_QASM_U;
            for (int i = 0; i < 3; ++i) op.p[i] = (float)pv[i];   // BasicOp::U carries f32 parameters
            op.reg_a = qubits[0].first; op.idx_a = qubits[0].second;
            emit(op, cond_reg, cond_val);
            return;
        }
        if (name == "CX") {
            if (qubits[0] == qubits[1]) fail("CX control and target are the same qubit", line);
            op.kind = QB_QASM_CX;
            op.reg_a = qubits[0].first; op.idx_a = qubits[0].second;
            op.reg_b = qubits[1].first; op.idx_b = qubits[1].second;
            emit(op, cond_reg, cond_val);
            return;
        }
        const GateDef& d = gates_[name];
        if (d.opaque) fail("opaque gate '" + name + "' has no definition to simulate", line);
        for (const GateCall& c : d.body) {
            std::vector<double> cp;
            for (const ExprP& e : c.params) cp.push_back(eval(e, pv));
            std::vector<std::pair<int, int>> cq;
            for (int ix : c.args) cq.push_back(qubits[ix]);
            for (size_t i = 0; i < cq.size(); ++i)
                for (size_t j = i + 1; j < cq.size(); ++j)
                    if (cq[i] == cq[j]) fail("gate '" + c.name + "' applied to the same qubit twice", line);
            apply_gate(c.name, cp, cq, cond_reg, cond_val, line, depth + 1);
        }
    }

    void qop(int cond_reg, uint32_t cond_val) {
        const Tok t = peek();
        if (t.kind != T_ID) fail("unexpected '" + t.text + "'", t.line);
        qb_qasm_op op;
        memset(&op, 0, sizeof(op));
        if (t.text == "measure") {
            next();
            Arg q = argument(true);
            expect_sym("->");
            Arg c = argument(false);
            expect_sym(";");
            const uint32_t qs = prog_->qregs[q.reg].second, cs = prog_->cregs[c.reg].second;
            if ((q.idx < 0) != (c.idx < 0)) fail("measure needs a qubit and a bit, or two registers", t.line);
            if (q.idx < 0 && qs != cs) fail("measure between registers of different size", t.line);
            const uint32_t cnt = q.idx < 0 ? qs : 1;
            for (uint32_t i = 0; i < cnt; ++i) {
                op.kind = QB_QASM_MEASURE;
                op.reg_a = q.reg; op.idx_a = q.idx < 0 ? (int)i : q.idx;
                op.reg_b = c.reg; op.idx_b = c.idx < 0 ? (int)i : c.idx;
                emit(op, cond_reg, cond_val);
            }
            return;
        }
        if (t.text == "reset") {
            next();
            Arg a = argument(true, true);
            expect_sym(";");
            const bool classical = a.reg & 0x40000000;
            const int reg = a.reg & 0x3fffffff;
            const uint32_t size = classical ? prog_->cregs[reg].second : prog_->qregs[reg].second;
            const uint32_t cnt = a.idx < 0 ? size : 1;
            for (uint32_t i = 0; i < cnt; ++i) {
                op.kind = classical ? QB_QASM_RESET_C : QB_QASM_RESET_Q;
                op.reg_a = reg; op.idx_a = a.idx < 0 ? (int)i : a.idx;
                emit(op, cond_reg, cond_val);
            }
            return;
        }
        // gate application
        std::string name = ident();
        if (name != "U" && name != "CX" && !gates_.count(name)) fail("gate '" + name + "' is not defined", t.line);
        std::vector<double> pv;
        if (is_sym("(")) {
            next();
            if (!is_sym(")")) { pv.push_back(eval(expr(nullptr), {})); while (is_sym(",")) { next(); pv.push_back(eval(expr(nullptr), {})); } }
            expect_sym(")");
        }
        std::vector<Arg> args;
        args.push_back(argument(true));
        while (is_sym(",")) { next(); args.push_back(argument(true)); }
        expect_sym(";");
        check_arity(name, pv.size(), args.size(), t.line);
        // broadcast over whole registers (all whole-register arguments must have the same size)
        uint32_t bsize = 0;
        for (const Arg& a : args)
            if (a.idx < 0) {
                const uint32_t s = prog_->qregs[a.reg].second;
                if (bsize && s != bsize) fail("register arguments of different sizes", t.line);
                bsize = s;
            }
        const uint32_t reps = bsize ? bsize : 1;
        for (uint32_t i = 0; i < reps; ++i) {
            std::vector<std::pair<int, int>> qs;
            for (const Arg& a : args) qs.push_back({a.reg, a.idx < 0 ? (int)i : a.idx});
            for (size_t x = 0; x < qs.size(); ++x)
                for (size_t y = x + 1; y < qs.size(); ++y)
                    if (qs[x] == qs[y]) fail("gate '" + name + "' applied to the same qubit twice", t.line);
            apply_gate(name, pv, qs, cond_reg, cond_val, t.line, 0);
        }
    }
};

}  // namespace

int qasm_parse(const char* source, QasmProgram** out) {
    std::unique_ptr<QasmProgram> prog(new QasmProgram());
    try {
        Parser p(prog.get());
        p.parse_source(source, true);
    } catch (const ParseError& e) {
        return set_error(QB_ERR_QASM, e.line, "OpenQASM error at line %d: %s", e.line, e.msg.c_str());
    } catch (const std::exception& e) {
        return set_error(QB_ERR_QASM, -1, "OpenQASM error: %s", e.what());
    }
    *out = prog.release();
    return QB_OK;
}

}  // namespace qb

using namespace qb;

extern "C" {

int qb_qasm_parse(const char* source, qb_qasm** out) try {
    if (!source || !out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    *out = nullptr;
    QasmProgram* p = nullptr;
    int rc = qasm_parse(source, &p);
    if (rc != QB_OK) return rc;
    *out = reinterpret_cast<qb_qasm*>(p);
    return QB_OK;
} QB_CATCH_ALL

void qb_qasm_free(qb_qasm* prog) { delete reinterpret_cast<QasmProgram*>(prog); }

#define QP(p) reinterpret_cast<const QasmProgram*>(p)
uint32_t qb_qasm_num_qregs(const qb_qasm* p) { return p ? (uint32_t)QP(p)->qregs.size() : 0; }
uint32_t qb_qasm_num_cregs(const qb_qasm* p) { return p ? (uint32_t)QP(p)->cregs.size() : 0; }
const char* qb_qasm_qreg_name(const qb_qasm* p, uint32_t i) { return p && i < QP(p)->qregs.size() ? QP(p)->qregs[i].first.c_str() : nullptr; }
const char* qb_qasm_creg_name(const qb_qasm* p, uint32_t i) { return p && i < QP(p)->cregs.size() ? QP(p)->cregs[i].first.c_str() : nullptr; }
uint32_t qb_qasm_qreg_size(const qb_qasm* p, uint32_t i) { return p && i < QP(p)->qregs.size() ? QP(p)->qregs[i].second : 0; }
uint32_t qb_qasm_creg_size(const qb_qasm* p, uint32_t i) { return p && i < QP(p)->cregs.size() ? QP(p)->cregs[i].second : 0; }
uint64_t qb_qasm_num_ops(const qb_qasm* p) { return p ? QP(p)->ops.size() : 0; }
const qb_qasm_op* qb_qasm_ops(const qb_qasm* p) { return p && !QP(p)->ops.empty() ? QP(p)->ops.data() : nullptr; }
#undef QP

}  // extern "C"
