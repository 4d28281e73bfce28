// emulate.hpp — TEST INFRASTRUCTURE, host only: a scalar walk through a V4Program exactly as the fused TMA
// kernel reads it (thread -> tile-local bits from the lane/warp tables, context word, register-block
// offsets recovered from the swizzled byte offsets, records decoded by opcode, every gate in its lifting
// form with the same operation order as tools/gen_v4_dispatch.py emits).
//
// It exists so that the CPU test suite (no GPU in CI) can check planner + encoder against the oracle for
// the programs the GPU will be handed: tests/test_abi_cpu.py calls qm_plan_emulate (tests/native/libqm_emu.so,
// a library of its own — the shipped libqm_b200.so does not contain it) on small registers.
// It is NOT a fallback: nothing in the product path calls it, and state-creating entry points still fail
// with QM_ERR_NO_DEVICE when there is no GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>

#include "../../quromorphic.jl_b200/csrc/encode_v4.hpp"

namespace qm {

struct EmuAmp { double x, y; };

inline void emu_lift(double& x, double& y, double t1, double s1, double t2, int sg) {
    x = fma(t1, y, x);
    y = (sg & 1) ? fma(s1, x, -y) : fma(s1, x, y);
    x = (sg & 2) ? fma(t2, y, -x) : fma(t2, y, x);
}

// Runs one pass program over a host register of nl index bits (rank_or: the sharded bits of this rank).
// slot_lifts: per-state lifting coefficients [nslots][8] of THIS register or nullptr.  Returns false on a
// malformed program (unknown opcode, a control predicate that differs inside a warp, ...).
inline bool emulate_v4(const V4Program& prog, const PlanPass& P, EmuAmp* amp, int nl, uint64_t rank_or, const double* slot_lifts) {
    const int m = P.m, NB = kV4BlockBits, NA = 1 << NB;
    const int nthreads = 1 << (m - NB);
    std::vector<int> tbit(m);                       // tile-local bit -> physical bit
    if (!P.tl.empty()) { for (int b = 0; b < m; ++b) tbit[b] = P.tl[b]; }
    else {
        for (int b = 0; b < P.L; ++b) tbit[b] = b;
        for (size_t j = 0; j < P.hb.size(); ++j) tbit[P.L + j] = P.hb[j];
    }
    uint64_t tilemask = 0; for (int b = 0; b < m; ++b) tilemask |= 1ull << tbit[b];
    const uint64_t ntiles = 1ull << (nl - m);
    std::vector<EmuAmp> tile(1u << m);
    for (uint64_t t = 0; t < ntiles; ++t) {
        // base: spread the tile id over the physical bits that are not tile bits
        uint64_t base = 0, rest = t;
        for (int b = 0; b < nl; ++b) if (!((tilemask >> b) & 1)) { base |= (rest & 1ull) << b; rest >>= 1; }
        const uint64_t gbase = base | rank_or;
        auto phys = [&](uint32_t i) { uint64_t a = base; for (int b = 0; b < m; ++b) if ((i >> b) & 1) a |= 1ull << tbit[b]; return a; };
        for (uint32_t i = 0; i < (1u << m); ++i) tile[i] = amp[phys(i)];
        uint32_t octx = prog.ctx_const | kV4CtxAlways;
        for (int j = 0; j < prog.n_outer; ++j) {
            const int pos = (int)((prog.outer_pack[j / 10] >> (6 * (j % 10))) & 63ull);
            octx |= (uint32_t)((gbase >> pos) & 1ull) << (kV4CtxOuter + j);
        }
        for (int gi = 0; gi < prog.ngroups; ++gi) {
            const V4Group& G = prog.groups[gi];
            uint32_t bm[kV4BlockBits];
            for (int j = 0; j < NB; ++j) { const uint32_t u = G.s[j] >> 4; bm[j] = u ^ ((u >> 3) & 7u); }
            for (int w0 = 0; w0 < nthreads; w0 += 32) {
                // the 32 lanes of a warp walk the record list together (the kernel's branches are .uni)
                const V4Rec* R = reinterpret_cast<const V4Rec*>(reinterpret_cast<const unsigned char*>(&prog) + G.rec_off);
                EmuAmp a[32][16]; uint32_t ctx[32], v[32];
                for (int l = 0; l < 32 && w0 + l < nthreads; ++l) {
                    const int ttid = w0 + l;
                    v[l] = (uint32_t)prog.lane_tab[gi][ttid & 31] | (uint32_t)prog.warp_tab[gi][ttid >> 5];
                    ctx[l] = v[l] | octx;
                    for (int k = 0; k < NA; ++k) { uint32_t off = 0; for (int j = 0; j < NB; ++j) if ((k >> j) & 1) off |= bm[j]; a[l][k] = tile[v[l] | off]; }
                }
                const int nl_ = nthreads - w0 < 32 ? nthreads - w0 : 32;
                double c[32][6]; bool preloaded = false;
                for (int guard = 0; guard < 4 * kV4MaxRecs; ++guard, ++R) {
                    if (R->op == (uint32_t)V4_OP_END) break;
                    if (R->op == (uint32_t)V4_OP_SLOT) {
                        if (!slot_lifts) return false;
                        for (int l = 0; l < nl_; ++l) for (int q = 0; q < 6; ++q) c[l][q] = slot_lifts[(size_t)R->aux * 8 + q];
                        preloaded = true; continue;
                    }
                    // records hold the coefficients pairwise: {c0, c3} {c1, c4} {c2, c5}
                    if (!preloaded) for (int l = 0; l < nl_; ++l) for (int q = 0; q < 3; ++q) { c[l][q] = R->c[2 * q]; c[l][3 + q] = R->c[2 * q + 1]; }
                    preloaded = false;
                    int fam, var, B, CB;
                    uint32_t op = R->op;
                    bool active = true;
                    if (op >= (uint32_t)V4_NUM_GATE_OPS && op < (uint32_t)V4_OP_SLOT) {       // controlled entry
                        op -= V4_NUM_GATE_OPS;
                        active = (ctx[0] & R->cm) == R->cm;
                        for (int l = 0; l < nl_; ++l) if (((ctx[l] & R->cm) == R->cm) != active) return false;   // must be warp-uniform
                    } else if (R->cm) return false;
                    if (!v4_decode_op(op, fam, var, B, CB)) return false;
                    if (!active) continue;
                    for (int l = 0; l < nl_; ++l) {
                        EmuAmp* x = a[l]; double* cc = c[l];
                        auto ok = [&](int k) { return CB < 0 || ((k >> CB) & 1); };
                        if (fam == 4) {
                            if (ctx[l] & R->tm) { cc[0] = cc[3]; cc[1] = cc[4]; cc[2] = cc[5]; }
                            for (int k = 0; k < NA; ++k) {
                                if (!ok(k)) continue;
                                if (var == 2) {
                                    const double tm_ = cc[1] * x[k].x;
                                    x[k].x = cc[0] * x[k].x; x[k].x = fma(-cc[1], x[k].y, x[k].x); x[k].y = fma(cc[0], x[k].y, tm_);
                                } else emu_lift(x[k].x, x[k].y, cc[0], cc[1], cc[2], var == 1 ? 3 : 0);
                            }
                            continue;
                        }
                        for (int k = 0; k < NA; ++k) {
                            if (!ok(k)) continue;
                            if (fam == 2) {
                                static const int sgs[4][2] = { { 0, 0 }, { 0, 3 }, { 3, 0 }, { 3, 3 } };
                                if ((k >> B) & 1) emu_lift(x[k].x, x[k].y, cc[3], cc[4], cc[5], sgs[var][1]);
                                else emu_lift(x[k].x, x[k].y, cc[0], cc[1], cc[2], sgs[var][0]);
                                continue;
                            }
                            if ((k >> B) & 1) continue;
                            const int k1 = k | (1 << B);
                            if (fam == 0) {
                                emu_lift(x[k].x, x[k1].x, cc[0], cc[1], cc[2], var);
                                emu_lift(x[k].y, x[k1].y, cc[0], cc[1], cc[2], var);
                            } else if (fam == 1) {
                                const int sg = var ? 3 : 0;
                                emu_lift(x[k].x, x[k1].y, cc[0], cc[1], cc[2], sg);
                                emu_lift(x[k].y, x[k1].x, -cc[0], -cc[1], -cc[2], sg);
                            } else { const EmuAmp tmp = x[k]; x[k] = x[k1]; x[k1] = tmp; }
                        }
                    }
                }
                // the END record says where the registers go: the load offsets, unless a tail of x! / cnot! gates on
                // block bits was folded into the stores (encode_v4)
                if (R->op != (uint32_t)V4_OP_END) return false;
                uint64_t packed; memcpy(&packed, &R->c[0], 8);
                const uint32_t so[kV4BlockBits] = { R->tm, R->cm, R->aux, (uint32_t)packed };
                uint32_t sb[kV4BlockBits];
                for (int j = 0; j < NB; ++j) { const uint32_t u = so[j] >> 4; sb[j] = u ^ ((u >> 3) & 7u); }
                const uint32_t du = (uint32_t)(packed >> 32) >> 4, dm = du ^ ((du >> 3) & 7u);
                for (int l = 0; l < nl_; ++l)
                    for (int k = 0; k < NA; ++k) { uint32_t off = dm; for (int j = 0; j < NB; ++j) if ((k >> j) & 1) off ^= sb[j]; tile[v[l] ^ off] = a[l][k]; }
            }
        }
        for (uint32_t i = 0; i < (1u << m); ++i) amp[phys(i)] = tile[i];
    }
    return true;
}

}  // namespace qm
