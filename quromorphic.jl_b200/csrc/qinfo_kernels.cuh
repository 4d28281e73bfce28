// qinfo_kernels.cuh — SURVEY.md §8f N4: the QInfo measures of MANY small density matrices on the device.
//
// The reference computes every measure from a Hermitian eigen-decomposition of a 2^k x 2^k matrix on the host
// (QInfo.jl:24-28 matrix_log2, :47-50 von_neumann_entropy, :199-204 renyi_entropy, :332-357 max/min entropy,
// :289-294 fidelity, :311-316 trace_distance) — one LAPACK call per matrix.  A batched reservoir (config C2) has
// thousands of reduced matrices per step, so here ONE WARP diagonalises one matrix (d <= 16) in shared memory with the
// cyclic complex Jacobi method (a few sweeps of plane rotations; converges quadratically, eigenvalues to ~1e-15) and
// evaluates the measure in place; the batch is one launch and one device-to-host copy of `count` doubles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qm {

static const int kHermMaxD = 16;
enum { HM_VON_NEUMANN = 0, HM_RENYI = 1, HM_MIN_ENTROPY = 2, HM_MAX_ENTROPY = 3, HM_TRACE_DISTANCE = 4, HM_FIDELITY = 5, HM_EIGVALS = 6 };

struct HermSmem {
    double2 A[kHermMaxD][kHermMaxD + 1];
    double2 V[kHermMaxD][kHermMaxD + 1];
    double2 B[kHermMaxD][kHermMaxD + 1];
    double2 T[kHermMaxD][kHermMaxD + 1];
    double  lam[kHermMaxD];
};

__device__ __forceinline__ double2 zmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 zconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// A (Hermitian, d x d, in shared memory) -> diagonal; V <- eigenvectors in columns when want_vecs.  One warp.
__device__ void jacobi_warp(double2 (*A)[kHermMaxD + 1], double2 (*V)[kHermMaxD + 1], int d, bool want_vecs) {
    const int lane = threadIdx.x & 31;
    if (want_vecs) for (int e = lane; e < d * d; e += 32) V[e / d][e % d] = make_double2((e / d) == (e % d) ? 1.0 : 0.0, 0.0);
    // symmetrise as the reference does: rho <- (rho + rho') / 2
    for (int e = lane; e < d * d; e += 32) {
        const int i = e / d, j = e % d;
        if (i <= j) {
            const double2 a = A[i][j], b = A[j][i];
            const double2 h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
            A[i][j] = h; A[j][i] = zconj(h);
        }
    }
    __syncwarp();
    double fro = 0.0;
    for (int e = lane; e < d * d; e += 32) { const double2 a = A[e / d][e % d]; fro += a.x * a.x + a.y * a.y; }
    fro = warp_sum(fro);
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0;
        for (int e = lane; e < d * d; e += 32) { const int i = e / d, j = e % d; if (i < j) { const double2 a = A[i][j]; off += a.x * a.x + a.y * a.y; } }
        off = warp_sum(off);
        if (off <= 1e-34 * fro || off < 1e-300) break;
        for (int p = 0; p < d - 1; ++p)
            for (int q = p + 1; q < d; ++q) {
                const double2 apq = A[p][q];
                const double m = hypot(apq.x, apq.y);
                if (m * m <= 1e-40 * fro || m < 1e-300) continue;                  // uniform: every lane reads the same entry
                const double2 e = make_double2(apq.x / m, apq.y / m);            // e^{i phi}
                const double tau = (A[q][q].x - A[p][p].x) / (2.0 * m);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                const double2 se = make_double2(s * e.x, s * e.y), sec = zconj(se);
                __syncwarp();
                if (lane < d) {                                                  // columns p, q:  A <- A R,  V <- V R
                    const double2 kp = A[lane][p], kq = A[lane][q];
                    const double2 x = zmul(sec, kq), y = zmul(se, kp);
                    A[lane][p] = make_double2(c * kp.x - x.x, c * kp.y - x.y);
                    A[lane][q] = make_double2(y.x + c * kq.x, y.y + c * kq.y);
                    if (want_vecs) {
                        const double2 vp = V[lane][p], vq = V[lane][q];
                        const double2 x2 = zmul(sec, vq), y2 = zmul(se, vp);
                        V[lane][p] = make_double2(c * vp.x - x2.x, c * vp.y - x2.y);
                        V[lane][q] = make_double2(y2.x + c * vq.x, y2.y + c * vq.y);
                    }
                }
                __syncwarp();
                if (lane < d) {                                                  // rows p, q:  A <- R' A
                    const double2 pk = A[p][lane], qk = A[q][lane];
                    const double2 x = zmul(se, qk), y = zmul(sec, pk);
                    A[p][lane] = make_double2(c * pk.x - x.x, c * pk.y - x.y);
                    A[q][lane] = make_double2(y.x + c * qk.x, y.y + c * qk.y);
                }
                __syncwarp();
                if (lane == 0) {
                    A[p][q] = make_double2(0.0, 0.0); A[q][p] = make_double2(0.0, 0.0);
                    A[p][p].y = 0.0; A[q][q].y = 0.0;
                }
                __syncwarp();
            }
    }
    __syncwarp();
}

// one warp per item.  mats: [count][d*d] complex, column-major (Julia Matrix memory) — a Hermitian matrix reads the same
// either way round up to conjugation, which leaves the eigenvalues alone; fidelity takes its second matrix from matsB.
// out: [count] (measures) or [count][d] ascending eigenvalues (HM_EIGVALS).
__global__ void __launch_bounds__(32)
k_herm_measure(const double2* __restrict__ mats, const double2* __restrict__ matsB, int d, int count, int kind, double alpha,
               double* __restrict__ out) {
    __shared__ HermSmem S;
    const int item = blockIdx.x, lane = threadIdx.x;
    if (item >= count) return;
    const double2* M = mats + (size_t)item * d * d;
    for (int e = lane; e < d * d; e += 32) S.A[e % d][e / d] = M[e];          // column-major: element e is (row e % d, column e / d)
    __syncwarp();
    if (kind == HM_FIDELITY) {
        // (Tr sqrt(sqrt(rho) sigma sqrt(rho)))^2, QInfo.jl:289-294
        jacobi_warp(S.A, S.V, d, true);
        if (lane < d) S.lam[lane] = sqrt(fmax(S.A[lane][lane].x, 0.0));
        const double2* Mb = matsB + (size_t)item * d * d;
        for (int e = lane; e < d * d; e += 32) {
            const int i = e % d, j = e / d;
            const double2 a = Mb[e], b = Mb[(size_t)i * d + j];              // (i, j) and (j, i): symmetrise sigma
            S.B[i][j] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        }
        __syncwarp();
        // T = sqrt(rho) = V diag(sqrt lam) V'
        for (int e = lane; e < d * d; e += 32) {
            const int i = e / d, j = e % d;
            double2 acc = make_double2(0.0, 0.0);
            for (int l = 0; l < d; ++l) { const double2 z = zmul(S.V[i][l], zconj(S.V[j][l])); acc.x += S.lam[l] * z.x; acc.y += S.lam[l] * z.y; }
            S.T[i][j] = acc;
        }
        __syncwarp();
        // A = T B (scratch in V), then A = (T B) T
        for (int e = lane; e < d * d; e += 32) {
            const int i = e / d, j = e % d;
            double2 acc = make_double2(0.0, 0.0);
            for (int l = 0; l < d; ++l) { const double2 z = zmul(S.T[i][l], S.B[l][j]); acc.x += z.x; acc.y += z.y; }
            S.V[i][j] = acc;
        }
        __syncwarp();
        for (int e = lane; e < d * d; e += 32) {
            const int i = e / d, j = e % d;
            double2 acc = make_double2(0.0, 0.0);
            for (int l = 0; l < d; ++l) { const double2 z = zmul(S.V[i][l], S.T[l][j]); acc.x += z.x; acc.y += z.y; }
            S.A[i][j] = acc;
        }
        __syncwarp();
        jacobi_warp(S.A, S.V, d, false);
        double v = lane < d ? sqrt(fmax(S.A[lane][lane].x, 0.0)) : 0.0;
        v = warp_sum(v);
        if (lane == 0) out[item] = v * v;
        return;
    }
    jacobi_warp(S.A, S.V, d, false);
    const double lam = lane < d ? S.A[lane][lane].x : 0.0;
    if (kind == HM_EIGVALS) {
        if (lane < d) S.lam[lane] = lam;
        __syncwarp();
        if (lane < d) {                         // rank of this eigenvalue (ties broken by index): ascending order
            int r = 0;
            for (int j = 0; j < d; ++j) { const double o = S.lam[j]; if (o < lam || (o == lam && j < lane)) ++r; }
            out[(size_t)item * d + r] = lam;
        }
        return;
    }
    double v = 0.0;
    if (lane < d) {
        switch (kind) {
        case HM_VON_NEUMANN: v = -lam * log2(fmax(lam, 1e-10)); break;                              // QInfo.jl:24-28, :47-50
        case HM_RENYI: v = lam >= 0.0 ? pow(lam, alpha) : pow(-lam, alpha) * cospi(alpha); break;     // Re(lam^alpha), QInfo.jl:202
        case HM_MAX_ENTROPY: v = lam > 1e-10 ? 1.0 : 0.0; break;                                     // QInfo.jl:334-336
        case HM_TRACE_DISTANCE: v = 0.5 * fabs(lam); break;                                          // QInfo.jl:315
        default: break;
        }
    }
    if (kind == HM_MIN_ENTROPY) {                                                                     // QInfo.jl:355-356
        double mx = lane < d ? lam : -1e300;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) out[item] = -log2(mx);
        return;
    }
    v = warp_sum(v);
    if (lane == 0) {
        if (kind == HM_RENYI) v = log2(v) / (1.0 - alpha);
        else if (kind == HM_MAX_ENTROPY) {
            // log2 of a COUNT: CUDA's log2 is a 1-ulp function (log2(8.0) comes out as 2.9999999999999996), the host's is exact
            // on powers of two — and so is the reference's log2(rank) (QInfo.jl:336)
            const double cnt = v, l = log2(cnt), r = rint(l);
            v = (exp2(r) == cnt) ? r : l;
        }
        out[item] = v;
    }
}

}  // namespace qm
