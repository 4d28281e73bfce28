// kernels.cuh — hand-written sm_100a kernels of the B200 state-vector engine.
//
//   k_apply_1q        K1/K2: one (controlled) 2x2 gate, streaming, 128-bit ComplexF64 accesses
//                     (replaces the kron + mul! of u2!/controlled!, QSim.jl:52-63, :175-203)
//   k_fused_pass      K3/K4: one HBM pass — a 2^m-amplitude tile staged in shared memory, any
//                     number of gates applied from register blocks of 8 amplitudes, written back
//   k_zpartial/final  K5: all-qubit thresholded marginals in one read (measure_z, QSim.jl:297-318)
//   k_measure_rot     K5: single-qubit marginal after a basis rotation (measure_x/y, QSim.jl:320-330)
//   k_probs           mp = abs2.(s), QSim.jl:293
//   k_dm_*            density-matrix methods on vec(rho): outer product, diagonal, marginals (QSim.jl:39-41, :295, :332-365)
//   k_reduced_*       K6: psi -> reduced density matrix (QSim.jl:39-41 + QInfo.jl:381-405)
//   k_leafsum/k_sample K8: inverse-CDF sampler (defined in oracle/qsim_oracle.c; no reference counterpart)
//
// All of them are HBM-bandwidth kernels: no tensor cores — FP64 has no tcgen05
// kind and the DMMA rate on B200 equals the vector FP64 rate, see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "planner.hpp"

namespace qm {

struct Mat2 { double v[8]; };   // U11 U21 U12 U22 (re, im)

__device__ __forceinline__ uint64_t insert0(uint64_t x, int pos) {
    uint64_t lo = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | lo;
}
__device__ __forceinline__ double2 cmul(double2 u, double2 a) {
    return make_double2(fma(u.x, a.x, -(u.y * a.y)), fma(u.x, a.y, u.y * a.x));
}
// acc + u * a
__device__ __forceinline__ double2 cfma(double2 u, double2 a, double2 acc) {
    acc.x = fma(u.x, a.x, acc.x); acc.x = fma(-u.y, a.y, acc.x);
    acc.y = fma(u.x, a.y, acc.y); acc.y = fma(u.y, a.x, acc.y);
    return acc;
}
// abs2 as two products and one sum, no contraction: bit-identical to the oracle's abs2()
__device__ __forceinline__ double abs2_exact(double2 a) {
    return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y));
}

// ---------------------------------------------------------------------------------------------
// K1/K2  single gate, unfused.  One thread per amplitude pair; pairs failing the control mask are
// neither read nor written (a controlled gate moves 16 * 2^n bytes, an uncontrolled one 32 * 2^n).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_apply_1q(double2* __restrict__ amp, uint64_t npairs, int tbit, uint64_t cmask, Mat2 U) {
    const double2 u11 = make_double2(U.v[0], U.v[1]), u21 = make_double2(U.v[2], U.v[3]);
    const double2 u12 = make_double2(U.v[4], U.v[5]), u22 = make_double2(U.v[6], U.v[7]);
    const uint64_t tm = 1ull << tbit;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < npairs; j += stride) {
        const uint64_t i0 = insert0(j, tbit);
        if ((i0 & cmask) != cmask) continue;
        const double2 a0 = amp[i0], a1 = amp[i0 | tm];
        amp[i0] = cfma(u12, a1, cmul(u11, a0));
        amp[i0 | tm] = cfma(u22, a1, cmul(u21, a0));
    }
}

// amplitudes *= (pr + i pi): the global phase the fused passes left pending (state.hpp gphase)
__global__ void k_scale_phase(double2* __restrict__ amp, uint64_t n, double pr, double pi) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double2 a = amp[i];
        amp[i] = make_double2(a.x * pr - a.y * pi, a.x * pi + a.y * pr);
    }
}

__global__ void k_set_basis(double2* __restrict__ amp, uint64_t reg_stride, int nreg, uint64_t index) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nreg) amp[(uint64_t)r * reg_stride + index] = make_double2(1.0, 0.0);
}

// ---------------------------------------------------------------------------------------------
// K3/K4  fused pass
// Tile-local index i (m bits): bits [0, L) are the L lowest index bits, bit L + j is index bit
// hb[j].  Shared-memory slot of i is i ^ ((i >> 3) & 7) (16-byte units): the XOR-128B pattern,
// conflict-free for 8 consecutive i and for 8 i's that differ in bits 3..5.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <int B>
__device__ __forceinline__ void blk_mat(double2 (&a)[8], double2 u11, double2 u21, double2 u12, double2 u22,
                                        uint32_t cm) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k & (1 << B)) continue;
        if ((k & cm) == cm) {
            const double2 a0 = a[k], a1 = a[k | (1 << B)];
            a[k] = cfma(u12, a1, cmul(u11, a0));
            a[k | (1 << B)] = cfma(u22, a1, cmul(u21, a0));
        }
    }
}

__device__ __forceinline__ void run_subs(double2 (&a)[8], const DevGroup& G, const DevSub* __restrict__ subs,
                                         uint32_t loc, uint64_t gbase, const double* __restrict__ slot_base) {
    for (int s = 0; s < G.nsub; ++s) {
        const DevSub& S = subs[G.sub_off + s];
        if ((loc & S.cmask_loc) != S.cmask_loc) continue;
        if ((gbase & S.cmask_out) != S.cmask_out) continue;
        const double* mp = S.slot >= 0 ? (slot_base + (size_t)S.slot * 8) : S.m;
        const uint32_t cm = S.cmask_blk;
        if (S.kind == SUB_MAT) {
            const double2 u11 = make_double2(mp[0], mp[1]), u21 = make_double2(mp[2], mp[3]);
            const double2 u12 = make_double2(mp[4], mp[5]), u22 = make_double2(mp[6], mp[7]);
            switch (S.b) {
            case 0: blk_mat<0>(a, u11, u21, u12, u22, cm); break;
            case 1: blk_mat<1>(a, u11, u21, u12, u22, cm); break;
            default: blk_mat<2>(a, u11, u21, u12, u22, cm); break;
            }
        } else {
            const double2 d0 = make_double2(mp[0], mp[1]), d1 = make_double2(mp[6], mp[7]);
            if (S.b < 3) {
                const int b = S.b;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((k & cm) == cm) a[k] = cmul(((k >> b) & 1) ? d1 : d0, a[k]);
            } else {
                const bool bit = (S.b == 3) ? ((loc & S.tmask_loc) != 0) : ((gbase & S.tmask_out) != 0);
                const double2 d = bit ? d1 : d0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((k & cm) == cm) a[k] = cmul(d, a[k]);
            }
        }
    }
}

// reg_bits: index bits of one register (n_local); tiles above that belong to the next register
// of the batch.  rank_or: (rank << n_local) for sharded registers, so that predicates on global
// qubits see the full index.  slots: per-register 2x2 matrices [nreg][nslots][8] (batched only).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_fused_pass(double2* __restrict__ amp, const unsigned char* __restrict__ blob, const double* __restrict__ slots,
             int nslots, uint64_t ntiles, int reg_bits, uint64_t rank_or) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevPass* __restrict__ P = reinterpret_cast<const DevPass*>(blob);
    const int m = P->m, L = P->L, nh = P->nhigh, ngroups = P->ngroups;
    const uint32_t tsize = 1u << m;
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    uint64_t* rowoff = reinterpret_cast<uint64_t*>(smem_raw + (size_t)tsize * 16);
    const DevGroup* __restrict__ groups = reinterpret_cast<const DevGroup*>(blob + sizeof(DevPass));
    const DevSub* __restrict__ subs =
        reinterpret_cast<const DevSub*>(blob + sizeof(DevPass) + (size_t)ngroups * sizeof(DevGroup));
    const int tid = threadIdx.x;

    for (uint32_t r = tid; r < (1u << nh); r += THREADS) {
        uint64_t off = 0;
        for (int j = 0; j < nh; ++j) if ((r >> j) & 1u) off |= 1ull << P->hb[j];
        rowoff[r] = off;
    }
    __syncthreads();

    const int tbits = reg_bits - m;              // tiles per register = 2^tbits
    const uint32_t lmask = (1u << L) - 1u;
    for (uint64_t tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const uint64_t breg = tile_id >> tbits;
        uint64_t base = (tile_id & ((1ull << tbits) - 1ull)) << L;
        for (int j = 0; j < nh; ++j) base = insert0(base, P->hb[j]);
        const uint64_t gbase = base | rank_or;
        double2* __restrict__ reg = amp + (breg << reg_bits) + base;
        const double* slot_base = slots ? slots + (size_t)breg * nslots * 8 : nullptr;

        for (uint32_t idx = tid; idx < tsize; idx += THREADS)
            cp_async16(&tile[swz(idx)], reg + rowoff[idx >> L] + (idx & lmask));
        cp_async_commit_wait_all();
        __syncthreads();

        for (int gi = 0; gi < ngroups; ++gi) {
            const DevGroup& G = groups[gi];
            const uint32_t m0 = G.rbmask[0], m1 = G.rbmask[1], m2 = G.rbmask[2];
            const int nfree = G.nfree;
            for (uint32_t item = tid; item < (1u << nfree); item += THREADS) {
                uint32_t loc = 0;
                for (int j = 0; j < nfree; ++j) loc |= ((item >> j) & 1u) << G.fpos[j];
                double2 a[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = ((k & 1) ? m0 : 0u) | ((k & 2) ? m1 : 0u) | ((k & 4) ? m2 : 0u);
                    a[k] = tile[swz(loc | off)];
                }
                run_subs(a, G, subs, loc, gbase, slot_base);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = ((k & 1) ? m0 : 0u) | ((k & 2) ? m1 : 0u) | ((k & 4) ? m2 : 0u);
                    tile[swz(loc | off)] = a[k];
                }
            }
            __syncthreads();
        }

        for (uint32_t idx = tid; idx < tsize; idx += THREADS)
            reg[rowoff[idx >> L] + (idx & lmask)] = tile[swz(idx)];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K5  all-qubit thresholded marginals.  CTA (r, j) sweeps `cpc` consecutive chunks of 2^cb
// amplitudes of register r and emits  [S1_0 .. S1_{cb-1} | H_0 .. H_{hb-1} | total]  where S1_q is
// the kept probability with index bit q set (q inside a chunk), H_b the kept probability of the
// chunks whose chunk-index bit b is set (b inside the CTA's range).  k_zfinal folds the partials
// of a register in a fixed order and resolves the bits above the CTA range from the CTA index.
// ---------------------------------------------------------------------------------------------
static const int kZThreads = 256;
static const int kZMaxChunkBits = 13;
static const int kZMaxHi = 10;
static const int kZVals = kZMaxChunkBits + kZMaxHi + 1;

// FAST: chunk of exactly 2^13 amplitudes (registers of >= 13 qubits).  Element `it*256 + tid` of a chunk
// has index bits 0..7 = tid and bits 8..12 = it, so with the 32-iteration loop unrolled the bit tests
// fold away at compile time: 32 independent 128-bit loads in flight per thread, ~4 adds per amplitude.
template <bool FAST>
__global__ void __launch_bounds__(kZThreads)
k_zpartial(const double2* __restrict__ amp, int reg_bits, int cb, int hbits, int ctas_per_reg_log2, double thr,
           double* __restrict__ partial) {
    const uint32_t cta = blockIdx.x;
    const uint64_t r = cta >> ctas_per_reg_log2, j = cta & ((1u << ctas_per_reg_log2) - 1u);
    const double2* __restrict__ base = amp + (r << reg_bits) + ((j << hbits) << cb);
    double s1[kZMaxChunkBits], hi[kZMaxHi], tot = 0.0;
#pragma unroll
    for (int b = 0; b < kZMaxChunkBits; ++b) s1[b] = 0.0;
#pragma unroll
    for (int b = 0; b < kZMaxHi; ++b) hi[b] = 0.0;
    const uint32_t csize = 1u << cb;
    for (uint32_t c = 0; c < (1u << hbits); ++c) {
        double ctot = 0.0;
        const double2* __restrict__ p = base + ((uint64_t)c << cb);
        if (FAST) {
            double v[32];
#pragma unroll
            for (int it = 0; it < 32; ++it) {
                const double pr = abs2_exact(p[it * kZThreads + threadIdx.x]);
                v[it] = pr > thr ? pr : 0.0;
            }
#pragma unroll
            for (int b = 0; b < 5; ++b) {           // bits 8..12 of the index = bits of `it`
                double acc = 0.0;
#pragma unroll
                for (int it = 0; it < 32; ++it) if ((it >> b) & 1) acc += v[it];
                s1[8 + b] += acc;
            }
#pragma unroll
            for (int w = 16; w > 0; w >>= 1)
#pragma unroll
                for (int it = 0; it < w; ++it) v[it] += v[it + w];
            ctot = v[0];
        } else {
            for (uint32_t idx = threadIdx.x; idx < csize; idx += kZThreads) {
                const double pr = abs2_exact(p[idx]);
                const double v = pr > thr ? pr : 0.0;
                ctot += v;
#pragma unroll
                for (int b = 0; b < kZMaxChunkBits; ++b) s1[b] += ((idx >> b) & 1u) ? v : 0.0;
            }
        }
        tot += ctot;
#pragma unroll
        for (int b = 0; b < kZMaxHi; ++b) hi[b] += ((c >> b) & 1u) ? ctot : 0.0;
    }
    if (FAST) {
#pragma unroll
        for (int b = 0; b < 8; ++b) s1[b] = ((threadIdx.x >> b) & 1u) ? tot : 0.0;   // bits 0..7 = bits of tid
    }
    // block reduction: butterfly inside the warp, then warp 0 adds the warp results in order
    __shared__ double red[kZThreads / 32][kZVals];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto wsum = [&](double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
#pragma unroll
    for (int b = 0; b < kZMaxChunkBits; ++b) { double v = wsum(s1[b]); if (lane == 0) red[warp][b] = v; }
#pragma unroll
    for (int b = 0; b < kZMaxHi; ++b) { double v = wsum(hi[b]); if (lane == 0) red[warp][kZMaxChunkBits + b] = v; }
    { double v = wsum(tot); if (lane == 0) red[warp][kZVals - 1] = v; }
    __syncthreads();
    if (threadIdx.x < kZVals) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kZThreads / 32; ++w) v += red[w][threadIdx.x];
        partial[(size_t)cta * kZVals + threadIdx.x] = v;
    }
}

// one CTA per register; out[r][q] = (p0, p1) for q < reg_bits, plus out_tot[r]
__global__ void __launch_bounds__(256)
k_zfinal(const double* __restrict__ partial, int reg_bits, int cb, int hbits, int ctas_per_reg_log2,
         double* __restrict__ out, double* __restrict__ out_tot) {
    const uint32_t r = blockIdx.x;
    const uint32_t nct = 1u << ctas_per_reg_log2;
    const double* __restrict__ pr = partial + (size_t)r * nct * kZVals;
    __shared__ double sh[256];
    const int nq = reg_bits;
    for (int q = 0; q <= nq; ++q) {          // q == nq: the total
        double acc = 0.0;
        for (uint32_t j = threadIdx.x; j < nct; j += 256) {
            const double* v = pr + (size_t)j * kZVals;
            double x;
            if (q == nq) x = v[kZVals - 1];
            else if (q < cb) x = v[q];
            else if (q < cb + hbits) x = v[kZMaxChunkBits + (q - cb)];
            else x = ((j >> (q - cb - hbits)) & 1u) ? v[kZVals - 1] : 0.0;
            acc += x;
        }
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (q == nq) out_tot[r] = sh[0]; else out[((size_t)r * nq + q) * 2 + 1] = sh[0];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int q = 0; q < nq; ++q) out[((size_t)r * nq + q) * 2] = out_tot[r] - out[((size_t)r * nq + q) * 2 + 1];
}

// K5, folded flavour: the last fused pass of a circuit has left per-CTA sums partial[cta][64] ([b] = kept probability with
// index bit b set, [63] = kept total; V4Fold, fused_tma_kernel.cuh ZFOLD).  One CTA adds the CTAs in order and writes the same
// [reg_bits][2] pairs + total that k_zfinal writes.
__global__ void __launch_bounds__(64)
k_zfold_final(const double* __restrict__ partial, int nctas, int reg_bits, double* __restrict__ out, double* __restrict__ out_tot) {
    __shared__ double sh[64];
    double acc = 0.0;
    for (int c = 0; c < nctas; ++c) acc += partial[(size_t)c * 64 + threadIdx.x];
    sh[threadIdx.x] = acc;
    __syncthreads();
    const double tot = sh[63];
    if ((int)threadIdx.x < reg_bits) { out[2 * threadIdx.x] = tot - sh[threadIdx.x]; out[2 * threadIdx.x + 1] = sh[threadIdx.x]; }
    if (threadIdx.x == 0) out_tot[0] = tot;
}

// K5  all-qubit measure_x / measure_y (QSim.jl:320-330 for every t): a tile of 2^m amplitudes (the L lowest index
// bits + m - L bits starting at hs) is staged in shared memory ONCE and every index bit of the tile selected by
// `tmask` is measured from it — the pair (a[i0], a[i0 | bit]) is rotated by R in registers exactly as k_measure_rot
// does, the threshold applies to the ROTATED probabilities.  The reference makes a copy of the state, multiplies a
// dense 2^n x 2^n operator into it and scans it, per qubit; here a 30-qubit register is read 3 times for all 30.
// CTA c of a register sweeps tiles c, c + ctas_per_reg, ... in order and emits partial[cta][12][2]; k_pauli_final
// adds the CTAs of a register in order and scatters to out[reg][index bit][2].
// ---------------------------------------------------------------------------------------------
static const int kPauliM = 12;
static const int kPauliThreads = 256;
__global__ void __launch_bounds__(kPauliThreads)
k_pauli_tile(const double2* __restrict__ amp, int reg_bits, int m, int L, int hs, uint32_t tmask, Mat2 R, double thr,
             uint32_t ctas_per_reg, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    const double2 u11 = make_double2(R.v[0], R.v[1]), u21 = make_double2(R.v[2], R.v[3]);
    const double2 u12 = make_double2(R.v[4], R.v[5]), u22 = make_double2(R.v[6], R.v[7]);
    const uint32_t reg = blockIdx.x / ctas_per_reg, c = blockIdx.x % ctas_per_reg;
    const double2* __restrict__ base_reg = amp + ((uint64_t)reg << reg_bits);
    const uint32_t tsize = 1u << m, lmask = (1u << L) - 1u;
    const int nh = m - L;
    const uint64_t ntiles = 1ull << (reg_bits - m);
    double p0[kPauliM], p1[kPauliM];
#pragma unroll
    for (int b = 0; b < kPauliM; ++b) { p0[b] = 0.0; p1[b] = 0.0; }
    for (uint64_t t = c; t < ntiles; t += ctas_per_reg) {
        // tile index -> base: the bits below hs above the L low ones, then the bits above the high run
        uint64_t base;
        if (nh == 0) base = t << m;
        else {
            const uint64_t midbits = (uint64_t)(hs - L);
            const uint64_t mid = t & ((1ull << midbits) - 1ull), top = t >> midbits;
            base = (mid << L) | (top << (hs + nh));
        }
        for (uint32_t idx = threadIdx.x; idx < tsize; idx += kPauliThreads)
            tile[idx] = base_reg[base | (idx & lmask) | ((uint64_t)(idx >> L) << hs)];
        __syncthreads();
#pragma unroll
        for (int b = 0; b < kPauliM; ++b) {
            if (!((tmask >> b) & 1u)) continue;
            double a0s = 0.0, a1s = 0.0;
            for (uint32_t j = threadIdx.x; j < (tsize >> 1); j += kPauliThreads) {
                const uint32_t i0 = (uint32_t)insert0(j, b);
                const double2 a0 = tile[i0], a1 = tile[i0 | (1u << b)];
                const double2 b0 = cfma(u12, a1, cmul(u11, a0)), b1 = cfma(u22, a1, cmul(u21, a0));
                const double q0 = abs2_exact(b0), q1 = abs2_exact(b1);
                a0s += q0 > thr ? q0 : 0.0;
                a1s += q1 > thr ? q1 : 0.0;
            }
            p0[b] += a0s; p1[b] += a1s;
        }
        __syncthreads();
    }
    __shared__ double red[kPauliThreads / 32][2 * kPauliM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int b = 0; b < kPauliM; ++b) {
        double v0 = p0[b], v1 = p1[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
        if (lane == 0) { red[warp][2 * b] = v0; red[warp][2 * b + 1] = v1; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kPauliM) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kPauliThreads / 32; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 2 * kPauliM + threadIdx.x] = v;
    }
}
// one thread per (register, tile-local bit, 0/1): out[reg][phys bit][2], only the bits in tmask are written
__global__ void __launch_bounds__(128)
k_pauli_final(const double* __restrict__ partial, int nreg, int nq, int m, int L, int hs, uint32_t tmask, uint32_t ctas_per_reg,
              double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nreg * 2 * kPauliM) return;
    const int reg = e / (2 * kPauliM), r = e % (2 * kPauliM), b = r >> 1;
    if (b >= m || !((tmask >> b) & 1u)) return;
    double v = 0.0;
    for (uint32_t c = 0; c < ctas_per_reg; ++c) v += partial[((size_t)reg * ctas_per_reg + c) * 2 * kPauliM + r];
    const int phys = b < L ? b : hs + (b - L);
    out[((size_t)reg * nq + phys) * 2 + (r & 1)] = v;
}

// K5  (p0, p1) of bit tbit after rotating it by R: measure_x / measure_y (QSim.jl:320-330); with
// R = identity it is measure_z of a single qubit.  Fixed grid -> fixed summation tree.
static const int kMeasBlocks = 1184;    // 148 SMs x 8
__global__ void __launch_bounds__(256)
k_measure_rot(const double2* __restrict__ amp, uint64_t npairs, int tbit, Mat2 R, int identity, double thr,
              double* __restrict__ partial) {
    const double2 u11 = make_double2(R.v[0], R.v[1]), u21 = make_double2(R.v[2], R.v[3]);
    const double2 u12 = make_double2(R.v[4], R.v[5]), u22 = make_double2(R.v[6], R.v[7]);
    const uint64_t tm = 1ull << tbit;
    double p0 = 0.0, p1 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < npairs; j += stride) {
        const uint64_t i0 = insert0(j, tbit);
        double2 a0 = amp[i0], a1 = amp[i0 | tm];
        if (!identity) {
            const double2 b0 = cfma(u12, a1, cmul(u11, a0)), b1 = cfma(u22, a1, cmul(u21, a0));
            a0 = b0; a1 = b1;
        }
        const double q0 = abs2_exact(a0), q1 = abs2_exact(a1);
        p0 += q0 > thr ? q0 : 0.0;
        p1 += q1 > thr ? q1 : 0.0;
    }
    __shared__ double sh0[256], sh1[256];
    sh0[threadIdx.x] = p0; sh1[threadIdx.x] = p1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { sh0[threadIdx.x] += sh0[threadIdx.x + w]; sh1[threadIdx.x] += sh1[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sh0[0]; partial[2 * blockIdx.x + 1] = sh1[0]; }
}
__global__ void __launch_bounds__(256)
k_pair_final(const double* __restrict__ partial, int nblocks, double* __restrict__ out2) {
    __shared__ double sh0[256], sh1[256];
    double p0 = 0.0, p1 = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { p0 += partial[2 * i]; p1 += partial[2 * i + 1]; }
    sh0[threadIdx.x] = p0; sh1[threadIdx.x] = p1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { sh0[threadIdx.x] += sh0[threadIdx.x + w]; sh1[threadIdx.x] += sh1[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = sh0[0]; out2[1] = sh1[0]; }
}

// ---------------------------------------------------------------------------------------------
// Density-matrix methods by vectorisation: rho on q qubits lives in the engine as the
// 2q-qubit vector vec(rho), amplitude r + c 2^q = rho[r, c] (Julia's column-major Matrix memory), so
// rho <- op rho op' (QSim.jl:67-79, :209-232) is `op` on the row bits and conj(op) on the column bits —
// the same fused passes.  Only these three small kernels are specific to it.
//   k_dm_outer   statevector_to_density_matrix, QSim.jl:39-41: rho[r, c] = s[r] conj(s[c])
//   k_dm_diag    mp(rho) = real.(diag(rho)), QSim.jl:295
//   k_dm_measure measure_z / measure_x / measure_y of rho (QSim.jl:332-365): diagonal of R rho R' summed by
//                bit t, entries <= threshold dropped; the 2x2 block of rho in (row bit t, column bit t) is
//                rotated in registers, no copy of rho is made
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dm_outer(const double2* __restrict__ sv, int q, double2* __restrict__ rho) {
    const uint64_t n = 1ull << (2 * q), rm = (1ull << q) - 1ull;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 a = sv[i & rm], b = sv[i >> q];
        rho[i] = make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));       // a * conj(b)
    }
}
__global__ void __launch_bounds__(256)
k_dm_diag(const double2* __restrict__ rho, int q, double* __restrict__ out) {
    const uint64_t d = 1ull << q;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += stride) out[i] = rho[i * (d + 1ull)].x;
}
__global__ void __launch_bounds__(256)
k_dm_measure(const double2* __restrict__ rho, int q, int tbit, Mat2 R, int identity, double thr, double* __restrict__ partial) {
    const double2 u11 = make_double2(R.v[0], R.v[1]), u21 = make_double2(R.v[2], R.v[3]);
    const double2 u12 = make_double2(R.v[4], R.v[5]), u22 = make_double2(R.v[6], R.v[7]);
    const uint64_t d = 1ull << q, tm = 1ull << tbit, npairs = d >> 1;
    double p0 = 0.0, p1 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < npairs; j += stride) {
        const uint64_t i0 = insert0(j, tbit), i1 = i0 | tm;
        double q0, q1;
        if (identity) { q0 = rho[i0 * (d + 1ull)].x; q1 = rho[i1 * (d + 1ull)].x; }
        else {
            const double2 b00 = rho[i0 + i0 * d], b10 = rho[i1 + i0 * d], b01 = rho[i0 + i1 * d], b11 = rho[i1 + i1 * d];
            // (R B R')[s, s] = sum_ab R[s,a] B[a,b] conj(R[s,b]);  w_s,b = sum_a R[s,a] B[a,b]
            const double2 w00 = cfma(u12, b10, cmul(u11, b00)), w01 = cfma(u12, b11, cmul(u11, b01));
            const double2 w10 = cfma(u22, b10, cmul(u21, b00)), w11 = cfma(u22, b11, cmul(u21, b01));
            q0 = (w00.x * u11.x + w00.y * u11.y) + (w01.x * u12.x + w01.y * u12.y);          // Re(w conj(u))
            q1 = (w10.x * u21.x + w10.y * u21.y) + (w11.x * u22.x + w11.y * u22.y);
        }
        p0 += q0 > thr ? q0 : 0.0;
        p1 += q1 > thr ? q1 : 0.0;
    }
    __shared__ double sh0[256], sh1[256];
    sh0[threadIdx.x] = p0; sh1[threadIdx.x] = p1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { sh0[threadIdx.x] += sh0[threadIdx.x + w]; sh1[threadIdx.x] += sh1[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sh0[0]; partial[2 * blockIdx.x + 1] = sh1[0]; }
}

__global__ void __launch_bounds__(256)
k_probs(const double2* __restrict__ amp, uint64_t n, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = abs2_exact(amp[i]);
}

// ---------------------------------------------------------------------------------------------
// K6  reduced density matrix straight from psi.
//   keep_low : rho[i,j] = sum_x psi[x d + i] conj(psi[x d + j])      (partial_trace(rho, dA, dB, 1))
//   keep_high: rho[i,j] = sum_x psi[i R + x] conj(psi[j R + x])      (partial_trace(rho, dA, dB, 2))
// grid.x = x-chunks, grid.y = (d/T)^2 entry tiles of T x T (T = min(d, 16)); every CTA writes one
// partial tile, k_reduced_final adds the x-chunks in order.
// ---------------------------------------------------------------------------------------------
static const int kRedXC = 64;
__global__ void __launch_bounds__(256)
k_reduced_partial(const double2* __restrict__ psi_all, int nbits, int k, int keep_low, uint64_t x_per_cta,
                  double2* __restrict__ partial_all) {
    // blockIdx.z = register of a batch (its amplitudes and its partial tiles follow those of register z - 1)
    const double2* __restrict__ psi = psi_all + ((uint64_t)blockIdx.z << nbits);
    double2* __restrict__ partial = partial_all + (size_t)blockIdx.z * gridDim.x * ((size_t)1 << (2 * k));
    const uint32_t d = 1u << k, T = d < 16 ? d : 16, tiles = d / T;
    const uint64_t R = 1ull << (nbits - k);
    const uint32_t ti0 = (blockIdx.y % tiles) * T, tj0 = (blockIdx.y / tiles) * T;
    __shared__ double2 A[kRedXC][17], Bm[kRedXC][17];
    const uint32_t ti = threadIdx.x % T, tj = threadIdx.x / T;
    const bool active = threadIdx.x < T * T;
    double2 acc = make_double2(0.0, 0.0);
    const uint64_t x0 = (uint64_t)blockIdx.x * x_per_cta, x1 = x0 + x_per_cta;
    for (uint64_t xb = x0; xb < x1; xb += kRedXC) {
        const uint32_t cnt = (uint32_t)((x1 - xb) < (uint64_t)kRedXC ? (x1 - xb) : kRedXC);
        for (uint32_t e = threadIdx.x; e < kRedXC * T; e += 256) {
            uint32_t xx, ii;
            if (keep_low) { xx = e / T; ii = e % T; } else { ii = e / kRedXC; xx = e % kRedXC; }
            double2 va = make_double2(0.0, 0.0), vb = va;
            if (xx < cnt) {
                const uint64_t x = xb + xx;
                va = keep_low ? psi[x * d + ti0 + ii] : psi[(uint64_t)(ti0 + ii) * R + x];
                vb = keep_low ? psi[x * d + tj0 + ii] : psi[(uint64_t)(tj0 + ii) * R + x];
            }
            A[xx][ii] = va; Bm[xx][ii] = vb;
        }
        __syncthreads();
        if (active) {
#pragma unroll 8
            for (uint32_t xx = 0; xx < kRedXC; ++xx) {
                const double2 p = A[xx][ti], q = Bm[xx][tj];
                acc.x = fma(p.x, q.x, acc.x); acc.x = fma(p.y, q.y, acc.x);     // p * conj(q)
                acc.y = fma(p.y, q.x, acc.y); acc.y = fma(-p.x, q.y, acc.y);
            }
        }
        __syncthreads();
    }
    if (active) partial[((size_t)blockIdx.x * d + (tj0 + tj)) * d + (ti0 + ti)] = acc;   // column-major [j][i]
}
__global__ void __launch_bounds__(256)
k_reduced_final(const double2* __restrict__ partial_all, uint32_t nchunks, uint32_t entries, double2* __restrict__ out_all) {
    const double2* __restrict__ partial = partial_all + (size_t)blockIdx.y * nchunks * entries;      // blockIdx.y = register
    double2* __restrict__ out = out_all + (size_t)blockIdx.y * entries;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    double2 acc = make_double2(0.0, 0.0);
    for (uint32_t c = 0; c < nchunks; ++c) { const double2 v = partial[(size_t)c * entries + e]; acc.x += v.x; acc.y += v.y; }
    out[e] = acc;
}

// ---------------------------------------------------------------------------------------------
// K8  sampler (definition: oracle/qsim_oracle.c orc_sample).  Level sums over 4096-ary leaves:
// lane l adds elements l, l+32, ... left to right, then a 5-step xor butterfly.
// ---------------------------------------------------------------------------------------------
static const uint32_t kLeaf = 4096;
template <bool FROM_AMPS>
__global__ void __launch_bounds__(256)
k_leafsum(const void* __restrict__ src, uint64_t n, uint64_t nleaves, double* __restrict__ out) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nleaves) return;
    const uint64_t lo = warp * kLeaf;
    double acc = 0.0;
    for (uint32_t j = 0; j < kLeaf / 32; ++j) {
        const uint64_t i = lo + 32ull * j + lane;
        double v = 0.0;
        if (i < n) v = FROM_AMPS ? abs2_exact(reinterpret_cast<const double2*>(src)[i]) : reinterpret_cast<const double*>(src)[i];
        acc = __dadd_rn(acc, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (lane == 0) out[warp] = acc;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// one thread per shot; s3 has n3 <= 4096 entries
__global__ void __launch_bounds__(128)
k_sample(const double2* __restrict__ amp, uint64_t N, const double* __restrict__ s1, uint64_t n1,
         const double* __restrict__ s2, uint64_t n2, const double* __restrict__ s3, uint64_t n3,
         uint64_t seed, int nshots, uint64_t* __restrict__ out, uint64_t leaf_lo, uint64_t leaf_cnt) {
    const int sh = blockIdx.x * blockDim.x + threadIdx.x;
    if (sh >= nshots) return;
    double total = 0.0;
    for (uint64_t c = 0; c < n3; ++c) total = __dadd_rn(total, s3[c]);
    const double u = __dmul_rn((double)(splitmix64(seed + (uint64_t)sh) >> 11), 0x1.0p-53);
    double r = __dmul_rn(u, total);
    uint64_t i3 = 0; double acc = 0.0;
    for (; i3 + 1 < n3; ++i3) { const double t = __dadd_rn(acc, s3[i3]); if (t > r) break; acc = t; }
    r = __dadd_rn(r, -acc); acc = 0.0;
    uint64_t i2 = i3 * kLeaf, e2 = (i3 + 1) * kLeaf < n2 ? (i3 + 1) * kLeaf : n2;
    for (; i2 + 1 < e2; ++i2) { const double t = __dadd_rn(acc, s2[i2]); if (t > r) break; acc = t; }
    r = __dadd_rn(r, -acc); acc = 0.0;
    uint64_t i1 = i2 * kLeaf, e1 = (i2 + 1) * kLeaf < n1 ? (i2 + 1) * kLeaf : n1;
    for (; i1 + 1 < e1; ++i1) { const double t = __dadd_rn(acc, s1[i1]); if (t > r) break; acc = t; }
    r = __dadd_rn(r, -acc); acc = 0.0;
    // sharded registers: `amp` is this rank's slice (leaves leaf_lo .. leaf_lo + leaf_cnt of the whole register, N
    // amplitudes in all); the rank that owns the leaf finishes the descent, the others report 0 (summed over ranks)
    if (i1 < leaf_lo || i1 >= leaf_lo + leaf_cnt) { out[sh] = 0; return; }
    const double2* a = amp - leaf_lo * kLeaf;
    uint64_t i0 = i1 * kLeaf, e0 = (i1 + 1) * kLeaf < N ? (i1 + 1) * kLeaf : N;
    for (; i0 + 1 < e0; ++i0) { const double t = __dadd_rn(acc, abs2_exact(a[i0])); if (t > r) break; acc = t; }
    out[sh] = i0;
}

}  // namespace qm
