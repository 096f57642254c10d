/*
 * quaru_oracle.c -- CPU restatement of Quaru's Register hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.  The product
 * (quaru_b200/) never links, imports or calls it.
 *
 * What it restates (paths relative to /root/reference):
 *   src/register.rs:41-63    Register::new / from_int
 *   src/register.rs:81-126   try_from_qubits / is_one / is_qubit
 *   src/register.rs:140-191  try_apply_fast   (the normative apply algorithm)
 *   src/register.rs:267-281  try_apply_all
 *   src/register.rs:297-342  try_measure      (uniform draw passed in as `u`)
 *   src/register.rs:409-429  PartialEq (1e-8) / get_duplicate
 *
 * Third-party arithmetic that is NOT vendored under /root/reference and is
 * therefore restated from its published source:
 *   ndarray ^0.15.6 (Cargo.toml:10)  Array2<c64>.dot(Array1<c64>) on contiguous
 *       rows = numeric_util::unrolled_dot: eight independent accumulators
 *       p0..p7 over full blocks of 8, folded as sum += (p0+p4), (p1+p5),
 *       (p2+p6), (p3+p7), then the <8 tail added sequentially.  See
 *       orc_unrolled_dot().
 *   num-complex ^0.4 (Cargo.toml:11)  (a+bi)(c+di) = (ac-bd) + (ad+bc)i with
 *       no fused multiply-add; Complex / f64 divides both parts; norm_sqr =
 *       re*re + im*im.  Build with -ffp-contract=off (see Makefile).
 *   rand ^0.8.5 thread_rng().gen::<f64>() (register.rs:317-318) is unseeded in
 *       the reference; here the draw is the caller-supplied `u`.
 *
 * Pinning: the reference has NO numeric golden vectors for this path
 * (SURVEY.md section 4 / 8c).  What it does hold are known-answer / property
 * tests (tests/register_operation_tests.rs, src/register.rs:437-445,
 * examples/{grovers,shors,quantum_teleportation,superdense_coding}.rs test
 * modules).  tests/test_oracle_reference_suite.py runs every one of those
 * against this oracle; that is the pin.  The Rust reference itself cannot be
 * compiled in this image (no cargo/rustc), so there is no oracle/_ref.
 *
 * Layout: state = 2^n interleaved (re,im) doubles; amplitude index bit q is
 * qubit q (register.rs:47-50 folds kron(b_k, acc)); matrix row-major, matrix
 * index bit j <-> targets[j] (register.rs:168-173).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } oc64;

/* OperationError mirror (register.rs:10-20) */
#define ORC_OK 0
#define ORC_ERR_INVALID_TARGET (-1)
#define ORC_ERR_INVALID_DIMENSIONS (-2)
#define ORC_ERR_INVALID_ARITY (-3)
#define ORC_ERR_INVALID_QUBIT (-4)
#define ORC_ERR_NOMEM (-100)

static inline oc64 c_mul(oc64 a, oc64 b) {
    /* num-complex Mul: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re) */
    oc64 r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r;
}
static inline oc64 c_add(oc64 a, oc64 b) { oc64 r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }

/* ndarray 0.15 numeric_util::unrolled_dot restated for Complex<f64>. */
static inline oc64 orc_unrolled_dot(const oc64* xs, const oc64* ys, size_t len) {
    oc64 z = {0.0, 0.0};
    oc64 sum = z, p0 = z, p1 = z, p2 = z, p3 = z, p4 = z, p5 = z, p6 = z, p7 = z;
    while (len >= 8) {
        p0 = c_add(p0, c_mul(xs[0], ys[0])); p1 = c_add(p1, c_mul(xs[1], ys[1]));
        p2 = c_add(p2, c_mul(xs[2], ys[2])); p3 = c_add(p3, c_mul(xs[3], ys[3]));
        p4 = c_add(p4, c_mul(xs[4], ys[4])); p5 = c_add(p5, c_mul(xs[5], ys[5]));
        p6 = c_add(p6, c_mul(xs[6], ys[6])); p7 = c_add(p7, c_mul(xs[7], ys[7]));
        xs += 8; ys += 8; len -= 8;
    }
    sum = c_add(sum, c_add(p0, p4));
    sum = c_add(sum, c_add(p1, p5));
    sum = c_add(sum, c_add(p2, p6));
    sum = c_add(sum, c_add(p3, p7));
    for (size_t i = 0; i < len; ++i) sum = c_add(sum, c_mul(xs[i], ys[i]));
    return sum;
}

static int cmp_u32(const void* a, const void* b) {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return (x > y) - (x < y);
}

/* register.rs:418-429 get_duplicate: sort a copy, first equal neighbour. */
static int orc_get_duplicate(const uint32_t* t, uint32_t k, uint32_t* dup) {
    if (k < 2) return 0;
    uint32_t* c = (uint32_t*)malloc(sizeof(uint32_t) * k);
    memcpy(c, t, sizeof(uint32_t) * k);
    qsort(c, k, sizeof(uint32_t), cmp_u32);
    int found = 0;
    for (uint32_t i = 1; i < k; ++i)
        if (c[i] == c[i - 1]) { *dup = c[i]; found = 1; break; }
    free(c);
    return found;
}

/* register.rs:141-153: validation, in the reference's order. */
int orc_validate_apply(uint32_t n, const uint32_t* targets, uint32_t k, int64_t* payload) {
    /* :143 `(rows, cols) == (k, k)` can only hold for a 2^k matrix when 2^k == k: never. */
    uint32_t dup;
    if (orc_get_duplicate(targets, k, &dup)) { if (payload) *payload = dup; return ORC_ERR_INVALID_TARGET; }
    for (uint32_t j = 0; j < k; ++j)
        if (targets[j] >= n) { if (payload) *payload = targets[j]; return ORC_ERR_INVALID_TARGET; }
    return ORC_OK;
}

/* deposit the bits of `g` into the positions where `mask` is 0 (low to high) */
static inline uint64_t deposit_zero_bits(uint64_t g, const uint32_t* sorted_t, uint32_t k) {
    for (uint32_t j = 0; j < k; ++j) {
        uint64_t lo = g & ((1ull << sorted_t[j]) - 1);
        g = ((g >> sorted_t[j]) << (sorted_t[j] + 1)) | lo;
    }
    return g;
}

/*
 * register.rs:140-191 try_apply_fast.  For every base index whose target bits
 * are all zero: gather 2^k amplitudes (matrix bit j <-> targets[j]), res =
 * M.dot(arr) (row-wise unrolled_dot), scatter into a CLONE of the state.
 * `threads` > 1 parallelises the group loop with OpenMP (the reference is
 * single-threaded; groups are disjoint so the result is identical).
 */
int orc_try_apply_fast(oc64* state, uint32_t n, const oc64* M, const uint32_t* targets, uint32_t k,
                       int threads, int64_t* payload) {
    int rc = orc_validate_apply(n, targets, k, payload);
    if (rc != ORC_OK) return rc;
    const uint64_t N = 1ull << n;
    const uint64_t dim = 1ull << k;
    oc64* new_state = (oc64*)malloc(sizeof(oc64) * N);           /* :155 clone */
    if (!new_state) return ORC_ERR_NOMEM;
    memcpy(new_state, state, sizeof(oc64) * N);
    uint32_t* st = (uint32_t*)malloc(sizeof(uint32_t) * (k ? k : 1));
    memcpy(st, targets, sizeof(uint32_t) * k);
    qsort(st, k, sizeof(uint32_t), cmp_u32);
    uint64_t* offs = (uint64_t*)malloc(sizeof(uint64_t) * dim);  /* :163-176 permutations - index */
    for (uint64_t i = 0; i < dim; ++i) {
        uint64_t o = 0;
        for (uint32_t j = 0; j < k; ++j) if ((i >> j) & 1) o |= 1ull << targets[j];
        offs[i] = o;
    }
    const int64_t groups = (int64_t)(N >> k);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
#endif
    {
        oc64* arr = (oc64*)malloc(sizeof(oc64) * dim);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t g = 0; g < groups; ++g) {
            const uint64_t index = deposit_zero_bits((uint64_t)g, st, k);  /* :156-157 filter, ascending */
            for (uint64_t i = 0; i < dim; ++i) arr[i] = state[index | offs[i]];          /* :178-181 */
            for (uint64_t r = 0; r < dim; ++r)                                            /* :183 */
                new_state[index | offs[r]] = orc_unrolled_dot(M + r * dim, arr, dim);    /* :185-187 */
        }
        free(arr);
    }
    memcpy(state, new_state, sizeof(oc64) * N);                  /* :189 */
    free(new_state); free(st); free(offs);
    return ORC_OK;
}

/*
 * Same arithmetic, but mirroring the reference's allocation pattern for
 * timing (BASELINE.md "ref_faithful"): scan all 2^n indices with the filter
 * closure (:156-157), per group a fresh permutations vector, a fresh arr, and
 * a fresh copy of the matrix (op.matrix() clones, operation.rs:326-328).
 */
int orc_try_apply_fast_faithful(oc64* state, uint32_t n, const oc64* M, const uint32_t* targets,
                                uint32_t k, int64_t* payload) {
    int rc = orc_validate_apply(n, targets, k, payload);
    if (rc != ORC_OK) return rc;
    const uint64_t N = 1ull << n, dim = 1ull << k;
    oc64* new_state = (oc64*)malloc(sizeof(oc64) * N);
    if (!new_state) return ORC_ERR_NOMEM;
    memcpy(new_state, state, sizeof(oc64) * N);
    for (uint64_t x = 0; x < N; ++x) {
        uint32_t* tclone = (uint32_t*)malloc(sizeof(uint32_t) * (k ? k : 1)); /* op.targets() clone per index */
        memcpy(tclone, targets, sizeof(uint32_t) * k);
        int all0 = 1;
        for (uint32_t j = 0; j < k; ++j) if ((x >> tclone[j]) & 1) { all0 = 0; break; }
        free(tclone);
        if (!all0) continue;
        uint64_t* perm = (uint64_t*)malloc(sizeof(uint64_t) * dim);
        for (uint64_t i = 0; i < dim; ++i) {
            uint64_t p = x;
            for (uint32_t j = 0; j < k; ++j) if ((i >> j) & 1) p |= 1ull << targets[j];
            perm[i] = p;
        }
        oc64* arr = (oc64*)calloc(dim, sizeof(oc64));
        for (uint64_t i = 0; i < dim; ++i) arr[i] = state[perm[i]];
        oc64* mclone = (oc64*)malloc(sizeof(oc64) * dim * dim);
        memcpy(mclone, M, sizeof(oc64) * dim * dim);
        oc64* res = (oc64*)malloc(sizeof(oc64) * dim);
        for (uint64_t r = 0; r < dim; ++r) res[r] = orc_unrolled_dot(mclone + r * dim, arr, dim);
        for (uint64_t i = 0; i < dim; ++i) new_state[perm[i]] = res[i];
        free(res); free(mclone); free(arr); free(perm);
    }
    memcpy(state, new_state, sizeof(oc64) * N);
    free(new_state);
    return ORC_OK;
}

/*
 * register.rs:267-281 try_apply_all.  The reference folds kron(acc, m) n times
 * into a 2^n x 2^n matrix and multiplies the state by it.  `literal` != 0 does
 * exactly that (only sensible for n <= 12: 16*4^n bytes); otherwise the same
 * operator is applied as the 2x2 gate on every qubit 0..n-1 (equal up to
 * rounding; pinned by apply_all_test, tests/register_operation_tests.rs:211-219).
 * The arity check (:269-271) is done by the caller, who owns the Operation.
 */
int orc_try_apply_all(oc64* state, uint32_t n, const oc64* m2, int literal) {
    const uint64_t N = 1ull << n;
    if (!literal) {
        for (uint32_t q = 0; q < n; ++q) {
            int rc = orc_try_apply_fast(state, n, m2, &q, 1, 1, NULL);
            if (rc != ORC_OK) return rc;
        }
        return ORC_OK;
    }
    /* kron(acc, m): result[(i*2+a),(j*2+b)] = acc[i][j]*m[a][b] -> last factor is the LSB */
    oc64* acc = (oc64*)malloc(sizeof(oc64));
    acc[0].re = 1.0; acc[0].im = 0.0;
    uint64_t d = 1;
    for (uint32_t s = 0; s < n; ++s) {
        uint64_t nd = d * 2;
        oc64* nx = (oc64*)malloc(sizeof(oc64) * nd * nd);
        if (!nx) { free(acc); return ORC_ERR_NOMEM; }
        for (uint64_t i = 0; i < d; ++i)
            for (uint64_t j = 0; j < d; ++j)
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b)
                        nx[(i * 2 + a) * nd + (j * 2 + b)] = c_mul(acc[i * d + j], m2[a * 2 + b]);
        free(acc); acc = nx; d = nd;
    }
    oc64* out = (oc64*)malloc(sizeof(oc64) * N);
    for (uint64_t r = 0; r < N; ++r) out[r] = orc_unrolled_dot(acc + r * N, state, N);
    memcpy(state, out, sizeof(oc64) * N);
    free(out); free(acc);
    return ORC_OK;
}

/* register.rs:303-315: sequential, index-order probability sums. */
int orc_probabilities(const oc64* state, uint32_t n, uint32_t target, double* p0, double* p1, int64_t* payload) {
    if (target >= n) { if (payload) *payload = target; return ORC_ERR_INVALID_TARGET; }
    const uint64_t N = 1ull << n;
    double prob_1 = 0.0, prob_0 = 0.0;
    for (uint64_t i = 0; i < N; ++i) {
        double prob = state[i].re * state[i].re + state[i].im * state[i].im;  /* norm_sqr */
        if ((i >> target) & 1) prob_1 += prob; else prob_0 += prob;
    }
    *p0 = prob_0; *p1 = prob_1;
    return ORC_OK;
}

/* register.rs:297-342 try_measure with the uniform draw supplied by the caller. */
int orc_try_measure(oc64* state, uint32_t n, uint32_t target, double u, int32_t* outcome, int64_t* payload) {
    double prob_0, prob_1;
    int rc = orc_probabilities(state, n, target, &prob_0, &prob_1, payload);
    if (rc != ORC_OK) return rc;
    const uint64_t N = 1ull << n;
    const int res = u > prob_0;                                   /* :321 strict */
    const double total_prob = res ? prob_1 : prob_0;              /* :323 */
    const double s = sqrt(total_prob);
    for (uint64_t i = 0; i < N; ++i) {
        if ((int)((i >> target) & 1) != res) { state[i].re = 0.0; state[i].im = 0.0; }   /* :328 */
        else { state[i].re /= s; state[i].im /= s; }                                     /* :337 */
    }
    *outcome = res;
    return ORC_OK;
}

/* register.rs:41-56 / :58-63: basis state; bits[q] is qubit q -> index sum b_q 2^q. */
void orc_init_basis(oc64* state, uint32_t n, uint64_t index) {
    memset(state, 0, sizeof(oc64) << n);
    state[index].re = 1.0;
}

/* register.rs:107-126 is_one / is_qubit for one 2-entry qubit. */
int orc_is_qubit(const oc64* q2) {
    double total_prob = -1.0;
    for (int i = 0; i < 2; ++i) total_prob += q2[i].re * q2[i].re + q2[i].im * q2[i].im;
    return fabs(total_prob) < 0.000001;
}

/* register.rs:81-102 try_from_qubits: validate, then fold kron(q_i, acc). */
int orc_try_from_qubits(oc64* state, uint32_t n, const oc64* qubits /* n x 2 */, int64_t* payload) {
    for (uint32_t i = 0; i < n; ++i)
        if (!orc_is_qubit(qubits + 2 * i)) { if (payload) *payload = i; return ORC_ERR_INVALID_QUBIT; }
    state[0].re = 1.0; state[0].im = 0.0;
    uint64_t len = 1;
    for (uint32_t i = 0; i < n; ++i) {
        /* kron(b, a)[x*len + y] = b[x]*a[y]; write the upper half first so a[] is still intact */
        for (uint64_t y = 0; y < len; ++y) state[len + y] = c_mul(qubits[2 * i + 1], state[y]);
        for (uint64_t y = 0; y < len; ++y) state[y] = c_mul(qubits[2 * i + 0], state[y]);
        len *= 2;
    }
    return ORC_OK;
}

/* register.rs:409-413 PartialEq: all |a_i - b_i| < 1e-8. */
int orc_registers_equal(const oc64* a, const oc64* b, uint32_t n) {
    const uint64_t N = 1ull << n;
    for (uint64_t i = 0; i < N; ++i) {
        double dr = a[i].re - b[i].re, di = a[i].im - b[i].im;
        if (!(hypot(dr, di) < 1e-8)) return 0;
    }
    return 1;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
