/*
 * qsim_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (C + OpenMP, matrix-free) of the Quromorphic.jl state-vector hot path:
 *   /root/reference/src/Quantum/QSim.jl   (state, u2!, controlled!, swap!, mp, measure_x/y/z)
 *   /root/reference/src/Quantum/QInfo.jl  (partial_trace of rho = psi psi')
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (quromorphic.jl_b200/csrc) never links or calls it.
 *
 * Parity pinning: Julia is not available in this environment, so the reference cannot be compiled or run by its own
 * runtime (no oracle/_ref).  Its SOURCE is executed all the same: oracle/jl_subset.py, a generic interpreter for the subset of
 * Julia that src/Quantum/*.jl and test/*.jl are written in, runs QSim.jl and QInfo.jl where they lie — the reference's own test
 * files pass on it (126 + 67 assertions, tests/test_ref_exec.py) — and the golden vectors it produces from the reference's
 * functions (tests/golden/ref_exec_golden.json, tests/golden/make_ref_exec_golden.py) are what this file is held to at 1e-12
 * (tests/test_golden.py).  Besides, this restatement is pinned (tests/test_oracle.py) against
 *   (1) every known-answer assertion of test/test_QSim.jl and test/test_QInfo.jl that touches
 *       this path, and
 *   (2) oracle/literal.py — a numpy transcription of the reference's literal algorithm
 *       (kron(id(r), kron(U, id(l))) then a dense matrix-vector product), for every gate kind
 *       and every (c, t) pair on small registers, and
 *   (3) the committed golden vectors tests/golden/qsim_golden.json (written by tests/golden/make_golden.py
 *       from that literal transcription; tests/test_golden.py), which agree with the executed-source vectors to 2e-15.
 * Calls without a reference counterpart of their own: the all-qubit expectation call is held to the executed reference's
 * per-qubit measure_z of the golden cases, psi -> reduced rho to its partial_trace(s*s', ...) of the same cases, fused/batched
 * application to gate-by-gate application.  "Parity unpinned" in the strict sense remains only the
 * sampler — there this file is the definition.
 *
 * Conventions: qubit arguments are the reference's own 1-BASED indices (qubit t <-> bit t-1 of
 * the 0-based amplitude index, QSim.jl:56-58, :304).  States are interleaved (re, im) doubles.
 * 2x2 matrices are Julia Matrix{ComplexF64} memory: column-major interleaved
 * (U11, U21, U12, U22).  Build with -ffp-contract=off so every product/sum is a separate
 * IEEE operation (the sampler's CDF must be reproducible bit for bit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

static inline cplx cmul(cplx a, cplx b) {
    cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}
static inline cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* statevector(n, m): QSim.jl:33-37 */
void orc_statevector(double* s, int n, uint64_t m) {
    uint64_t N = 1ull << n;
    memset(s, 0, (size_t)N * 16);
    s[2 * m] = 1.0;
}

/* u2!(s, t, U): QSim.jl:52-63.  op = I_{2^(q-t)} (x) U (x) I_{2^(t-1)}; the dense product
 * reduces, for every index pair differing in bit t-1, to
 *   a0' = U11 a0 + U12 a1,  a1' = U21 a0 + U22 a1          (all other operator entries are 0).
 * Returns 1 ("Qubit index out of range", QSim.jl:54) when t > q; like the reference it does
 * not test t >= 1 — callers must. */
int orc_u2(double* s, int n, int t, const double* U) {
    if (t > n) return 1;
    if (t < 1) return 2;
    cplx u11 = { U[0], U[1] }, u21 = { U[2], U[3] }, u12 = { U[4], U[5] }, u22 = { U[6], U[7] };
    cplx* a = (cplx*)s;
    uint64_t mask = 1ull << (t - 1), half = 1ull << (n - 1);
#pragma omp parallel for schedule(static) if (n >= 14)
    for (uint64_t j = 0; j < half; ++j) {
        uint64_t i0 = ((j & ~(mask - 1)) << 1) | (j & (mask - 1));
        uint64_t i1 = i0 | mask;
        cplx a0 = a[i0], a1 = a[i1];
        a[i0] = cadd(cmul(u11, a0), cmul(u12, a1));
        a[i1] = cadd(cmul(u21, a0), cmul(u22, a1));
    }
    return 0;
}

/* The reference's effective (control, target) for controlled!(s, c, t, V): QSim.jl:181-197.
 * U2 is placed directly above `mid` in kron(left, kron(U2, kron(mid, right))), so it acts on
 * bits (b-2, b-1) = qubits (b-1, b) with b = max(c, t); c < t puts the control on the LOW of
 * the two (kron(id(1),P0)+kron(V,P1), :191), c > t on the HIGH (:193). */
void orc_effective_ct(int c, int t, int* ce, int* te) {
    int b = c > t ? c : t;
    if (c < t) { *ce = b - 1; *te = b; } else { *ce = b; *te = b - 1; }
}

/* textbook controlled-V on 1-based qubits (ce, te) */
static void controlled_textbook(cplx* a, int n, int ce, int te, const double* V) {
    cplx v11 = { V[0], V[1] }, v21 = { V[2], V[3] }, v12 = { V[4], V[5] }, v22 = { V[6], V[7] };
    uint64_t cm = 1ull << (ce - 1), tm = 1ull << (te - 1), half = 1ull << (n - 1);
#pragma omp parallel for schedule(static) if (n >= 14)
    for (uint64_t j = 0; j < half; ++j) {
        uint64_t i0 = ((j & ~(tm - 1)) << 1) | (j & (tm - 1));
        if (!(i0 & cm)) continue;
        uint64_t i1 = i0 | tm;
        cplx a0 = a[i0], a1 = a[i1];
        a[i0] = cadd(cmul(v11, a0), cmul(v12, a1));
        a[i1] = cadd(cmul(v21, a0), cmul(v22, a1));
    }
}

/* controlled!(s, c, t, V): QSim.jl:175-203, INCLUDING the non-adjacent behaviour.
 * Check order as the reference: range (:177) -> 1, then c == t (:178) -> 3.
 * quirk = 1: reference semantics; quirk = 0: the intended textbook gate on (c, t). */
int orc_controlled(double* s, int n, int c, int t, const double* V, int quirk) {
    if (!(c <= n && t <= n)) return 1;
    if (c == t) return 3;
    if (c < 1 || t < 1) return 2;
    int ce = c, te = t;
    if (quirk) orc_effective_ct(c, t, &ce, &te);
    controlled_textbook((cplx*)s, n, ce, te, V);
    return 0;
}

static const double X_MAT[8] = { 0, 0, 1, 0, 1, 0, 0, 0 };

/* swap!(s, q1, q2): QSim.jl:262-275 — range check, no-op when equal, else three cnot!. */
int orc_swap(double* s, int n, int q1, int q2, int quirk) {
    if (!(q1 <= n && q2 <= n)) return 1;
    if (q1 < 1 || q2 < 1) return 2;
    if (q1 == q2) return 0;
    orc_controlled(s, n, q1, q2, X_MAT, quirk);
    orc_controlled(s, n, q2, q1, X_MAT, quirk);
    orc_controlled(s, n, q1, q2, X_MAT, quirk);
    return 0;
}

/* abs2 exactly as two products and one sum (matches the device's __dmul_rn/__dadd_rn) */
static inline double abs2(cplx a) { return a.re * a.re + a.im * a.im; }

/* mp(s) = abs2.(s): QSim.jl:293 */
void orc_mp(const double* s, int n, double* out) {
    const cplx* a = (const cplx*)s;
    uint64_t N = 1ull << n;
#pragma omp parallel for schedule(static) if (n >= 14)
    for (uint64_t i = 0; i < N; ++i) out[i] = abs2(a[i]);
}

/* measure_z(s, t): QSim.jl:297-318, literal: left-to-right accumulation over i = 0..2^n-1 of
 * the probabilities strictly greater than `threshold` (reference: 1e-10).  threshold < 0
 * keeps everything. */
int orc_measure_z_seq(const double* s, int n, int t, double threshold, double* out2) {
    if (t > n) return 1;
    if (t < 1) return 2;
    const cplx* a = (const cplx*)s;
    uint64_t N = 1ull << n, mask = 1ull << (t - 1);
    double p0 = 0.0, p1 = 0.0;
    for (uint64_t i = 0; i < N; ++i) {
        double p = abs2(a[i]);
        if (p > threshold) { if ((i & mask) == 0) p0 += p; else p1 += p; }
    }
    out2[0] = p0; out2[1] = p1;
    return 0;
}

/* pairwise tree over [lo, hi) of the thresholded probabilities with bit `mask` == want */
static double pair_sum(const cplx* a, uint64_t lo, uint64_t hi, uint64_t mask, uint64_t want,
                       double threshold) {
    if (hi - lo <= 256) {
        double acc = 0.0;
        for (uint64_t i = lo; i < hi; ++i) {
            double p = abs2(a[i]);
            if (p > threshold && (i & mask) == want) acc += p;
        }
        return acc;
    }
    uint64_t mid = lo + ((hi - lo) >> 1);
    return pair_sum(a, lo, mid, mask, want, threshold) + pair_sum(a, mid, hi, mask, want, threshold);
}

/* measure_z with a fixed pairwise summation tree: same set of terms as QSim.jl:306-315,
 * rounding error O(log N) instead of O(N) — the truth used at large n. */
int orc_measure_z(const double* s, int n, int t, double threshold, double* out2) {
    if (t > n) return 1;
    if (t < 1) return 2;
    const cplx* a = (const cplx*)s;
    uint64_t N = 1ull << n, mask = 1ull << (t - 1);
    int nch = 1; while ((nch < 256) && ((N / (uint64_t)nch) > 65536)) nch <<= 1;
    double part0[256], part1[256];
    uint64_t per = N / (uint64_t)nch;
#pragma omp parallel for schedule(static) if (n >= 18)
    for (int c = 0; c < nch; ++c) {
        part0[c] = pair_sum(a, per * c, per * (c + 1), mask, 0, threshold);
        part1[c] = pair_sum(a, per * c, per * (c + 1), mask, mask, threshold);
    }
    for (int w = 1; w < nch; w <<= 1)
        for (int c = 0; c + w < nch; c += 2 * w) { part0[c] += part0[c + w]; part1[c] += part1[c + w]; }
    out2[0] = part0[0]; out2[1] = part1[0];
    return 0;
}

/* all-qubit measure_z in one sweep (the reservoir readout): out[2(q-1)] = p0, out[2(q-1)+1] = p1.
 * Per-thread blocks are summed in index order, blocks combined by a pairwise tree. */
void orc_expect_z_all(const double* s, int n, double threshold, double* out) {
    const cplx* a = (const cplx*)s;
    uint64_t N = 1ull << n;
    uint64_t blk = N < 4096 ? N : 4096;
    uint64_t nb = N / blk;
    double* part = (double*)calloc((size_t)nb * (size_t)(n + 1), sizeof(double));
#pragma omp parallel for schedule(static) if (n >= 14)
    for (uint64_t b = 0; b < nb; ++b) {
        double acc[65]; for (int q = 0; q <= n; ++q) acc[q] = 0.0;
        for (uint64_t i = b * blk; i < (b + 1) * blk; ++i) {
            double p = abs2(a[i]);
            if (p > threshold) {
                acc[n] += p;
                for (int q = 0; q < n; ++q) if ((i >> q) & 1) acc[q] += p;
            }
        }
        memcpy(part + b * (size_t)(n + 1), acc, sizeof(double) * (size_t)(n + 1));
    }
    for (uint64_t w = 1; w < nb; w <<= 1)
        for (uint64_t b = 0; b + w < nb; b += 2 * w)
            for (int q = 0; q <= n; ++q) part[b * (size_t)(n + 1) + q] += part[(b + w) * (size_t)(n + 1) + q];
    for (int q = 0; q < n; ++q) { out[2 * q + 1] = part[q]; out[2 * q] = part[n] - part[q]; }
    free(part);
}

/* measure_x(s,t) = measure_z(h!(copy(s), t), t), measure_y(s,t) = measure_z(rx!(copy(s), t, pi/2), t):
 * QSim.jl:320-330.  basis: 1 = X, 2 = Y.  The caller passes the rotation matrix `R` it would have
 * used (H, or rx_gate(pi/2)) so that the matrix bits are the host's own; seq selects the literal
 * left-to-right sum. */
int orc_measure_rot(const double* s, int n, int t, const double* R, double threshold, int seq,
                    double* out2) {
    if (t > n) return 1;
    if (t < 1) return 2;
    uint64_t N = 1ull << n;
    double* c = (double*)malloc((size_t)N * 16);
    memcpy(c, s, (size_t)N * 16);
    orc_u2(c, n, t, R);
    int rc = seq ? orc_measure_z_seq(c, n, t, threshold, out2) : orc_measure_z(c, n, t, threshold, out2);
    free(c);
    return rc;
}

/* statevector_to_density_matrix + partial_trace without forming rho.
 * QSim.jl:39-41 (rho = s s') and QInfo.jl:381-405: rho index = a*dB + b, A = high factor.
 *   subsystem == 1 (keep_low = 1): rho_B[i,j] = sum_a psi[a dB + i] conj(psi[a dB + j])   (:385-391)
 *   otherwise      (keep_low = 0): rho_A[i,j] = sum_b psi[i dB + b] conj(psi[j dB + b])   (:395-403)
 * The reference first symmetrises (rho + rho')/2 (:383), a no-op for rho = psi psi' up to
 * rounding; we symmetrise the result instead.  out is column-major interleaved d x d. */
void orc_reduced_density(const double* s, int n, int keep_low, int k, double* out) {
    const cplx* psi = (const cplx*)s;
    uint64_t d = 1ull << k, rest = 1ull << (n - k);
    cplx* r = (cplx*)calloc((size_t)(d * d), sizeof(cplx));
#pragma omp parallel for schedule(static) collapse(2) if (n >= 12)
    for (uint64_t j = 0; j < d; ++j)
        for (uint64_t i = 0; i < d; ++i) {
            /* The reference adds these terms left to right in Float64 (QInfo.jl:387-389, :398-400); with 2^26 terms per
             * entry (30 qubits, k = 4) that order alone carries ~1e-11 of rounding, more than the 1e-12 the parity bar
             * asks of the engine.  The checker therefore accumulates in x87 extended precision (64-bit mantissa): the
             * products are still IEEE doubles, the SUM is the exact one to ~1e-19 per term. */
            long double sr = 0.0L, si = 0.0L;
            for (uint64_t x = 0; x < rest; ++x) {
                cplx p = keep_low ? psi[x * d + i] : psi[i * rest + x];
                cplx q = keep_low ? psi[x * d + j] : psi[j * rest + x];
                sr += (long double)(p.re * q.re) + (long double)(p.im * q.im);       /* p * conj(q) */
                si += (long double)(p.im * q.re) - (long double)(p.re * q.im);
            }
            r[j * d + i].re = (double)sr; r[j * d + i].im = (double)si;
        }
    cplx* o = (cplx*)out;
    for (uint64_t j = 0; j < d; ++j)
        for (uint64_t i = 0; i < d; ++i) {
            o[j * d + i].re = 0.5 * (r[j * d + i].re + r[i * d + j].re);
            o[j * d + i].im = 0.5 * (r[j * d + i].im - r[i * d + j].im);
        }
    free(r);
}

/* ---- sampler (no reference counterpart: "parity unpinned"; this file DEFINES it) --------------
 * Shot s draws u_s = (splitmix64(seed + s) >> 11) * 2^-53 and returns the first index whose
 * cumulative probability exceeds u_s * total.  The cumulative sum is taken over a fixed 4-level
 * structure of 4096-ary "leaves".  A leaf sum is defined the way a 32-lane warp computes it:
 * lane l adds elements l, l+32, l+64, ... left to right, then the 32 lane sums are combined by an
 * xor butterfly (offsets 16, 8, 4, 2, 1; IEEE addition is commutative, so every lane ends with the
 * same bits).  Level k+1 applies the same rule to the level-k sums.  The search descends level by
 * level with plain left-to-right partial sums, subtracting the prefix it skipped.  Every
 * operation is a single IEEE double add/multiply in a fixed order, so the GPU kernel
 * (quromorphic.jl_b200/csrc/kernels.cuh k_leafsum / k_sample) reproduces it bit for bit. */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
uint64_t orc_splitmix64(uint64_t x) { return splitmix64(x); }

#define LEAF 4096ull
/* leaf sum over elements [lo, lo + LEAF) of either abs2(amplitudes) or a double array; elements
 * at or beyond n count as 0 */
static double leaf_sum(const cplx* a, const double* v, uint64_t lo, uint64_t n) {
    double lane[32];
    for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (uint64_t j = 0; j < LEAF / 32; ++j) {
            uint64_t i = lo + 32 * j + (uint64_t)l;
            double x = 0.0;
            if (i < n) x = a ? abs2(a[i]) : v[i];
            acc += x;
        }
        lane[l] = acc;
    }
    for (int o = 16; o > 0; o >>= 1) {
        double nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ o];
        memcpy(lane, nxt, sizeof(lane));
    }
    return lane[0];
}

void orc_sample(const double* s, int n, uint64_t seed, int nshots, uint64_t* out) {
    const cplx* a = (const cplx*)s;
    uint64_t N = 1ull << n;
    uint64_t n1 = (N + LEAF - 1) / LEAF, n2 = (n1 + LEAF - 1) / LEAF, n3 = (n2 + LEAF - 1) / LEAF;
    double* s1 = (double*)malloc(sizeof(double) * n1);
    double* s2 = (double*)malloc(sizeof(double) * n2);
    double* s3 = (double*)malloc(sizeof(double) * n3);
#pragma omp parallel for schedule(static) if (n >= 16)
    for (uint64_t c = 0; c < n1; ++c) s1[c] = leaf_sum(a, 0, c * LEAF, N);
    for (uint64_t c = 0; c < n2; ++c) s2[c] = leaf_sum(0, s1, c * LEAF, n1);
    for (uint64_t c = 0; c < n3; ++c) s3[c] = leaf_sum(0, s2, c * LEAF, n2);
    double total = 0.0;
    for (uint64_t c = 0; c < n3; ++c) total += s3[c];
    for (int sh = 0; sh < nshots; ++sh) {
        double u = (double)(splitmix64(seed + (uint64_t)sh) >> 11) * 0x1.0p-53;
        double r = u * total;
        uint64_t i3 = 0; double acc = 0.0;
        for (; i3 + 1 < n3; ++i3) { double t = acc + s3[i3]; if (t > r) break; acc = t; }
        r -= acc; acc = 0.0;
        uint64_t i2 = i3 * LEAF, e2 = (i3 + 1) * LEAF < n2 ? (i3 + 1) * LEAF : n2;
        for (; i2 + 1 < e2; ++i2) { double t = acc + s2[i2]; if (t > r) break; acc = t; }
        r -= acc; acc = 0.0;
        uint64_t i1 = i2 * LEAF, e1 = (i2 + 1) * LEAF < n1 ? (i2 + 1) * LEAF : n1;
        for (; i1 + 1 < e1; ++i1) { double t = acc + s1[i1]; if (t > r) break; acc = t; }
        r -= acc; acc = 0.0;
        uint64_t i0 = i1 * LEAF, e0 = (i1 + 1) * LEAF < N ? (i1 + 1) * LEAF : N;
        for (; i0 + 1 < e0; ++i0) { double t = acc + abs2(a[i0]); if (t > r) break; acc = t; }
        out[sh] = i0;
    }
    free(s1); free(s2); free(s3);
}

/* ---- circuit runner (for the timed CPU baseline and large-n parity) ---------------------------
 * Gates in the product's wire format (include/qm_b200.h qm_gate: kind, q0, q1, flags, m[8]) but
 * with 0-BASED TEXTBOOK bits — i.e. after the host shim has applied t -> t-1 and the
 * reference's control map, exactly what the engine receives. */
typedef struct { int32_t kind, q0, q1, flags; double m[8]; } orc_gate;

int orc_apply_circuit(double* s, int n, const orc_gate* g, int ng) {
    for (int i = 0; i < ng; ++i) {
        int rc = 0;
        switch (g[i].kind) {
        case 0: rc = orc_u2(s, n, g[i].q0 + 1, g[i].m); break;
        case 1: rc = orc_controlled(s, n, g[i].q0 + 1, g[i].q1 + 1, g[i].m, 0); break;
        case 2: rc = orc_swap(s, n, g[i].q0 + 1, g[i].q1 + 1, 0); break;
        default: rc = 9;
        }
        if (rc) return rc;
    }
    return 0;
}

/* norm^2 by pairwise tree */
double orc_norm2(const double* s, int n) {
    return pair_sum((const cplx*)s, 0, 1ull << n, 0, 0, -1.0);
}

/* seeded Haar-like test state: splitmix64-driven Box-Muller, then normalised.  Used by tests and
 * by bench.py to create identical inputs on both sides without shipping 16 GiB files. */
void orc_random_state(double* s, int n, uint64_t seed) {
    uint64_t N = 1ull << n;
#pragma omp parallel for schedule(static) if (n >= 14)
    for (uint64_t i = 0; i < N; ++i) {
        double u1 = ((double)(splitmix64(seed ^ (4 * i + 1)) >> 11) + 1.0) * 0x1.0p-53;
        double u2 = (double)(splitmix64(seed ^ (4 * i + 2)) >> 11) * 0x1.0p-53;
        double r = sqrt(-2.0 * log(u1));
        s[2 * i] = r * cos(6.283185307179586 * u2);
        s[2 * i + 1] = r * sin(6.283185307179586 * u2);
    }
    double nrm = sqrt(orc_norm2(s, n));
#pragma omp parallel for schedule(static) if (n >= 14)
    for (uint64_t i = 0; i < 2 * N; ++i) s[i] /= nrm;
}
