// lifting.hpp — host-side derivation of the in-place "lifting" form of 2x2 gates (see fused_tma.cuh).
//
//   real orthogonal block M = [m11 m12; m21 m22] acting on a real pair (x, y):
//        x += t1*y ;   y = s1*x + sg1*y ;   x = t2*y + sg2*x
//   with s1 = m21, sg1 = sign(m22), sg2 = sign(m11),
//        t1 = (m22 - sg1)/m21 = -sg1*m21/(1 + |m22|),  t2 = (m11 - sg2)/m21 = -sg2*m21/(1 + |m11|)
//   (the second forms have no cancellation; they use m21^2 = 1 - m22^2 = 1 - m11^2, i.e. orthogonality).
// Every derivation is verified by reconstructing the 2x2 block from the three shears; a matrix that
// is not orthogonal/unitary to a few ulp is reported as not liftable and its pass takes the general
// (v1) kernel, which applies the matrix literally.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "planner.hpp"

namespace qm {

enum { LK_ROT = 0, LK_ROTX = 1, LK_PHASE = 2, LK_PERMX = 3, LK_NONE = -1 };

// Reconstruction tolerance.  Products of several gates (the planner pre-multiplies runs of 1-qubit gates) are
// unitary only to a few 1e-16 per factor; 2e-14 keeps them on the lifting path and is still 50x inside the
// 1e-12 parity bound.  Anything further from unitary is applied literally by the general kernel.
static const double kLiftTol = 2.0e-14;

inline int lift_sg(const uint32_t mk[2]) { return (mk[0] ? 1 : 0) | (mk[1] ? 2 : 0); }

// coefficients {t1, s1, t2}, sign masks {m1, m2} (0 or 0x80000000)
inline bool lift3_from(double m11, double m12, double m21, double m22, double out[3], uint32_t mk[2]) {
    const double s1 = (m22 < 0 || (m22 == 0 && signbit(m22))) ? -1.0 : 1.0;
    const double s2 = (m11 < 0 || (m11 == 0 && signbit(m11))) ? -1.0 : 1.0;
    const double t1 = -s1 * m21 / (1.0 + fabs(m22));
    const double t2 = -s2 * m21 / (1.0 + fabs(m11));
    const double r21 = m21, r22 = m21 * t1 + s1;
    const double r11 = t2 * r21 + s2, r12 = t2 * r22 + s2 * t1;
    if (!(fabs(r11 - m11) <= kLiftTol && fabs(r12 - m12) <= kLiftTol && fabs(r21 - m21) <= kLiftTol &&
          fabs(r22 - m22) <= kLiftTol))
        return false;
    out[0] = t1; out[1] = m21; out[2] = t2;
    mk[0] = s1 < 0 ? 0x80000000u : 0u; mk[1] = s2 < 0 ? 0x80000000u : 0u;
    return true;
}

struct LiftCoef {      // exactly the 64-byte c[8] block of a V2Sub / of one per-state slot entry
    double c[6];
    uint32_t mk[4];
};
static_assert(sizeof(LiftCoef) == 64, "LiftCoef layout");

// Classify a 2x2 (Julia order U11 U21 U12 U22, re/im) that is to be applied as ONE in-place sub.
// Returns LK_* or LK_NONE when the matrix needs the phase*rot*phase decomposition (or is not unitary).
inline int lift_classify(const double* m, LiftCoef& L) {
    memset(&L, 0, sizeof(L));
    const bool all_real = m[1] == 0 && m[3] == 0 && m[5] == 0 && m[7] == 0;
    if (all_real && m[0] == 0 && m[6] == 0 && m[2] == 1 && m[4] == 1) return LK_PERMX;
    if (mat_is_diag(m)) {
        // d0 = (m0, m1), d1 = (m6, m7): complex scale = rotation [dx -dy; dy dx] of (re, im)
        if (lift3_from(m[0], -m[1], m[1], m[0], L.c, L.mk) && lift3_from(m[6], -m[7], m[7], m[6], L.c + 3, L.mk + 2))
            return LK_PHASE;
        return LK_NONE;
    }
    if (all_real) {      // (a0.x, a1.x) and (a0.y, a1.y) transform by [u11 u12; u21 u22]
        if (lift3_from(m[0], m[4], m[2], m[6], L.c, L.mk)) return LK_ROT;
        return LK_NONE;
    }
    if (m[1] == 0 && m[7] == 0 && m[2] == 0 && m[4] == 0) {
        // u11 = p, u21 = i s, u12 = i r, u22 = q:  (a0.x, a1.y) -> [p -r; s q] ; (a0.y, a1.x) -> [p r; -s q]
        // the kernel has the rx-like sign variants only (both signs equal); anything else decomposes
        if (lift3_from(m[0], -m[5], m[3], m[6], L.c, L.mk) && L.mk[0] == L.mk[1]) return LK_ROTX;
        memset(&L, 0, sizeof(L));
        return LK_NONE;
    }
    return LK_NONE;
}

// U = diag(l0, l1) * [c -s; s c] * diag(1, r1), c = |u11|, s = |u21| (both >= 0).
// Writes the three factors as matrices in application order: out[0] = right phase, out[1] = rotation,
// out[2] = left phase.  Returns false when U is not unitary to rounding.
inline bool decompose_unitary(const double* m, double out[3][8]) {
    const double u11r = m[0], u11i = m[1], u21r = m[2], u21i = m[3], u12r = m[4], u12i = m[5], u22r = m[6], u22i = m[7];
    const double c = hypot(u11r, u11i), s = hypot(u21r, u21i);
    if (fabs(c * c + s * s - 1.0) > 4e-14) return false;
    double l0r = 1, l0i = 0, l1r = 1, l1i = 0, r1r, r1i;
    if (c > 0) { l0r = u11r / c; l0i = u11i / c; }
    if (s > 0) { l1r = u21r / s; l1i = u21i / s; }
    if (c >= s) {        // r1 = u22 / (l1 * c) = u22 * conj(l1) / c
        r1r = (u22r * l1r + u22i * l1i) / c; r1i = (u22i * l1r - u22r * l1i) / c;
    } else {             // r1 = -u12 / (l0 * s) = -u12 * conj(l0) / s
        r1r = -(u12r * l0r + u12i * l0i) / s; r1i = -(u12i * l0r - u12r * l0i) / s;
    }
    // renormalise the phases (they are unit modulus up to rounding) and rebuild to check
    auto unit = [](double& x, double& y) { const double h = hypot(x, y); if (h > 0) { x /= h; y /= h; } };
    unit(l0r, l0i); unit(l1r, l1i); unit(r1r, r1i);
    double R[8] = { 1, 0, 0, 0, 0, 0, r1r, r1i };
    double Y[8] = { c, 0, s, 0, -s, 0, c, 0 };
    double Lm[8] = { l0r, l0i, 0, 0, 0, 0, l1r, l1i };
    double T[8], W[8];
    mat_mul(Y, R, T); mat_mul(Lm, T, W);
    for (int i = 0; i < 8; ++i) if (fabs(W[i] - m[i]) > kLiftTol) return false;
    memcpy(out[0], R, sizeof(R)); memcpy(out[1], Y, sizeof(Y)); memcpy(out[2], Lm, sizeof(Lm));
    return true;
}

// After lower_gates: make every (non-slot) gate a single in-place sub where possible.  General
// unitaries are replaced by their three factors; `liftable` is cleared on gates that have no
// lifting form at all (non-unitary), which sends their pass to the v1 kernel.
// A factor of the decomposition that is the identity up to rounding (the right phase of rz * ry is exactly 1 in exact
// arithmetic and 1 +- a few 1e-17 in floating point) is dropped.  The test has a tolerance on purpose: whether a record
// exists must depend on the STRUCTURE of the circuit, not on how an angle happened to round — the compiled pass kernels are
// keyed by the record sequence (jit_codegen.hpp).  1e-14 is 100x inside the 1e-12 parity bound.
inline bool mat_near_identity(const double* m) {
    return mat_is_diag(m) && fabs(m[0] - 1.0) <= 1e-14 && fabs(m[1]) <= 1e-14 && fabs(m[6] - 1.0) <= 1e-14 && fabs(m[7]) <= 1e-14;
}

inline void prepare_for_lifting(std::vector<LGate>& gates) {
    std::vector<LGate> out; out.reserve(gates.size() + 16);
    for (const LGate& g : gates) {
        if (g.slot >= 0) { out.push_back(g); continue; }
        LiftCoef L;
        if (lift_classify(g.m, L) != LK_NONE) { LGate h = g; h.liftable = true; out.push_back(h); continue; }
        double f[3][8];
        if (decompose_unitary(g.m, f)) {
            bool ok = true; LiftCoef t;
            for (int i = 0; i < 3; ++i) if (!mat_near_identity(f[i]) && lift_classify(f[i], t) == LK_NONE) ok = false;
            if (ok) {
                for (int i = 0; i < 3; ++i) {
                    if (mat_near_identity(f[i])) continue;
                    LGate h = g; memcpy(h.m, f[i], sizeof(h.m)); h.diag = mat_is_diag(h.m); h.liftable = true;
                    out.push_back(h);
                }
                continue;
            }
        }
        LGate h = g; h.liftable = false; out.push_back(h);
    }
    gates.swap(out);
}

}  // namespace qm

namespace qm {
// principal square root of a unit-modulus complex number (real part >= 0): half of a phase
inline void unit_sqrt(double x, double y, double& rx, double& ry) {
    const double phi = atan2(y, x);
    rx = cos(0.5 * phi); ry = sin(0.5 * phi);
}
}  // namespace qm
