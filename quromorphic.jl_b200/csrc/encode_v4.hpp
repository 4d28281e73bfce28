// encode_v4.hpp — host only: PlanPass -> V4Program, the byte image the fused TMA kernel interprets (layout in
// fused_tma.cuh, opcodes in v4_ops.inc / tools/gen_v4_dispatch.py).  Shared by the engine (qm_api.cu) and by the
// host emulator of the test suite (tests/native/emu_api.cu), so both see the same bytes.
#pragma once
#include <stddef.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "fused_tma.cuh"
#include "lifting.hpp"
#include "planner.hpp"

namespace qm {

// the planner limits of the TMA kernel, in one place (engine, introspection and emulator must agree)
inline void v4_planner_options(PlanOptions& o, double max_cost, int cost_model = 0) {
    o.max_runs = 3; o.max_run_len = 8; o.max_subs = kV4MaxSubs; o.max_groups = kV4MaxGroups;
    o.max_outer = kV4MaxOuter; o.warp_uniform_ctrl = true; o.block_bits = kV4BlockBits;
    o.cost_model = cost_model;
    o.group_cost = cost_model == 1 ? kV4GroupCostCompiled : kV4GroupCost; o.retile = true;
    if (cost_model == 1) { o.fixed_units = 24.0; o.hbm_units = 112.0; }
    if (max_cost > 0 && o.max_cost < o.group_cost + 5.0) o.max_cost = o.group_cost + 5.0;
}

// decode op -> family (0 ROT, 1 ROTX, 2 PHASE, 3 PERMX, 4 THR), variant, B, CB
inline bool v4_decode_op(uint32_t op, int& fam, int& var, int& B, int& CB) {
    auto bc = [&](int idx) {
        int i = 0;
        for (int b = 0; b < kV4BlockBits; ++b) for (int c = -1; c < kV4BlockBits; ++c) { if (c == b) continue; if (i == idx) { B = b; CB = c; return; } ++i; }
    };
    if (op < (uint32_t)V4_OP_ROTX) { fam = 0; var = op / V4_NBC; bc(op % V4_NBC); return true; }
    if (op < (uint32_t)V4_OP_PHASE) { fam = 1; var = (op - V4_OP_ROTX) / V4_NBC; bc((op - V4_OP_ROTX) % V4_NBC); return true; }
    if (op < (uint32_t)V4_OP_PERMX) { fam = 2; var = (op - V4_OP_PHASE) / V4_NBC; bc((op - V4_OP_PHASE) % V4_NBC); return true; }
    if (op < (uint32_t)V4_OP_THR) { fam = 3; var = 0; bc(op - V4_OP_PERMX); return true; }
    if (op < (uint32_t)V4_NUM_GATE_OPS) { fam = 4; var = (op - V4_OP_THR) / (kV4BlockBits + 1); B = -1; CB = (int)((op - V4_OP_THR) % (kV4BlockBits + 1)) - 1; return true; }
    return false;
}

// runs of consecutive bits in hb (ascending), cut after 8 bits: one TMA box dimension holds at most 256 elements
inline int hb_runs(const std::vector<int>& hb, int start[8], int len[8]) {
    int nr = 0;
    for (size_t i = 0; i < hb.size(); ++i) {
        if (nr > 0 && hb[i] == start[nr - 1] + len[nr - 1] && len[nr - 1] < 8) { len[nr - 1]++; continue; }
        if (nr == 8) return 99;
        start[nr] = hb[i]; len[nr] = 1; ++nr;
    }
    return nr;
}

// PlanPass -> V4Program (layout in fused_tma.cuh).  Returns false when the pass breaks a limit of the
// kernel that the planner is supposed to respect (context bits, warp-uniform controls); the caller then
// falls back to the general kernel.
//
// gph (optional, {re, im}, multiplied in place): DIAGONAL GATES IN "D1" FORM.  An uncontrolled diagonal gate diag(d0, d1)
// (rz, z, the phase factors of a general unitary) equals d0 * diag(1, d1 / d0): the scalar d0 is a GLOBAL phase of the
// register — it is multiplied into *gph and never touches an amplitude (the engine applies the accumulated phase when
// somebody reads amplitudes, qm_api.cu materialize_phase; probabilities, expectations, reduced densities and samples do not
// depend on it) — and diag(1, r) only rotates the amplitudes whose target bit is set: HALF the FP64 work of the gate
// (1.5 operations per amplitude instead of 3; rz is a third of the one-qubit gates of config C3 and of every reservoir
// layer).  The form is used when the target is a register-block bit (a PHASE record "controlled" by the target bit itself)
// or a bit that is uniform over a warp (outside the tile or on the warp index: the thread-constant phase entered through
// its controlled entry, so half of the warps skip the record); on a lane bit the gate keeps both coefficient sets.
//
// canon (needs gph; COMPILED pass kernels only — the interpreter ignores a record's aux): CANONICAL SIGNS.  The lifting form of
// a rotation whose cosine is negative uses negated addends (the "sign variants" of the opcode table), so the opcode of a gate
// would depend on its ANGLE and a reservoir step with new angles would never find its compiled kernel again.  In canonical
// form the matrix is negated first (-M has the plain form) and the -1 is put where it is cheapest: into the global phase
// for an uncontrolled rotation, else into aux bit 0 / 1 = "negate the amplitudes that took coefficient set 0 / 1" — one
// integer xor per value in the generated code (jit_codegen.hpp), data of the launch.  What is left in the opcodes is the
// structure of the circuit only: gate family, rotation or reflection, register-block positions, controlled or not.
inline bool encode_v4(const PlanPass& P, V4Program& out, double* gph = nullptr, bool canon = false) {
    if (canon && !gph) return false;
    memset(&out, 0, offsetof(V4Program, recs));
    out.ngroups = (int)P.groups.size();
    int outer_of[64]; for (int i = 0; i < 64; ++i) outer_of[i] = -1;
    int n_outer = 0, nrec = 0;
    bool ok = true;
    // context bit of a physical index bit that is not a register-block bit of the current group
    auto ctx_bit = [&](int phys) -> uint32_t {
        const int tl = tile_local_of(P, phys);
        if (tl >= 0) return 1u << tl;
        if (outer_of[phys] < 0) {
            if (n_outer >= kV4MaxOuter) { ok = false; return 0u; }
            outer_of[phys] = n_outer;
            out.outer_pack[n_outer / 10] |= (uint64_t)phys << (6 * (n_outer % 10));
            ++n_outer;
        }
        return 1u << (kV4CtxOuter + outer_of[phys]);
    };
    const int NB = kV4BlockBits;
    const int nwarp_bits = P.m - NB - 5;
    for (size_t gi = 0; gi < P.groups.size(); ++gi) {
        const PlanGroup& G = P.groups[gi];
        V4Group& g = out.groups[gi];
        int rbl[NB]; for (int i = 0; i < NB; ++i) rbl[i] = -1;
        uint32_t blockmask = 0; int nrb = 0;
        for (size_t i = 0; i < G.rb.size() && nrb < NB; ++i) { rbl[nrb] = tile_local_of(P, G.rb[i]); blockmask |= 1u << rbl[nrb]; ++nrb; }
        // always four block bits: pad with tile bits no gate of this group targets (highest first)
        for (int b = P.m - 1; b >= 0 && nrb < NB; --b) if (!((blockmask >> b) & 1)) { rbl[nrb++] = b; blockmask |= 1u << b; }
        std::sort(rbl, rbl + NB);
        for (int i = 0; i < NB; ++i) { const uint32_t mk = 1u << rbl[i]; g.s[i] = (mk ^ ((mk >> 3) & 7u)) << 4; }
        // Thread-index bit order.  The three lowest thread bits must give distinct 16-byte bank groups under
        // slot = i ^ ((i >> 3) & 7): bit k of the slot's low 3 bits is i_k ^ i_{k+3}, so pick a free bit from
        // {k, k+3} for k = 0, 1, 2.  Bits the planner pinned to the warp index (controls that are not block
        // bits) go last.
        uint32_t wmask = 0;
        for (int b : G.wb) { const int tl = tile_local_of(P, b); if (tl >= 0 && !((blockmask >> tl) & 1)) wmask |= 1u << tl; }
        std::vector<int> order; uint32_t used = blockmask | wmask;
        for (int k = 0; k < 3; ++k) {
            int pick = -1;
            if (!((used >> k) & 1)) pick = k; else if (!((used >> (k + 3)) & 1)) pick = k + 3;
            if (pick >= 0) { order.push_back(pick); used |= 1u << pick; }
        }
        for (int b = 0; b < P.m; ++b) if (!((used >> b) & 1)) { order.push_back(b); used |= 1u << b; }
        for (int b = 0; b < P.m; ++b) if ((wmask >> b) & 1) order.push_back(b);
        if ((int)order.size() != P.m - NB || __builtin_popcount(wmask) > nwarp_bits) { ok = false; break; }
        uint32_t warp_uniform = 0;                    // tile-local bits that sit on the warp index in this group
        for (size_t j = 5; j < order.size(); ++j) warp_uniform |= 1u << order[j];
        for (int lane = 0; lane < 32; ++lane) {
            uint32_t v = 0; for (int j = 0; j < 5; ++j) if ((lane >> j) & 1) v |= 1u << order[j];
            out.lane_tab[gi][lane] = (uint16_t)v;
        }
        for (int w = 0; w < (1 << nwarp_bits); ++w) {
            uint32_t v = 0; for (int j = 0; j < nwarp_bits; ++j) if ((w >> j) & 1) v |= 1u << order[5 + j];
            out.warp_tab[gi][w] = (uint16_t)v;
        }
        g.rec_off = (uint32_t)(offsetof(V4Program, recs) + (size_t)nrec * sizeof(V4Rec));
        // A tail of x! / cnot! gates whose target AND control are register-block bits only permutes the thread's own
        // sixteen amplitudes: register k ends up at block index A k + c over GF(2).  The swizzled offsets are
        // xor-linear, so the tail is folded into the offsets the END record hands to the stores (v4_dispatch.inc)
        // instead of being executed: store offset of block bit j = xor_i A[i][j] s[i], base xor = xor_i c[i] s[i].
        auto block_index = [&](int phys) { const int tl = tile_local_of(P, phys); if (tl >= 0) for (int i = 0; i < NB; ++i) if (rbl[i] == tl) return i; return -1; };
        size_t nexec = G.subs.size();
        while (nexec > 0) {
            const LGate& x = G.subs[nexec - 1];
            LiftCoef Lt;
            if (x.slot >= 0 || x.diag || lift_classify(x.m, Lt) != LK_PERMX || block_index(x.target) < 0) break;
            if (x.ctrl >= 0 && block_index(x.ctrl) < 0) break;
            --nexec;
        }
        uint32_t Arow[NB], cvec = 0;                 // Arow[i] = row i of A as a bit mask over k's bits
        for (int i = 0; i < NB; ++i) Arow[i] = 1u << i;
        for (size_t q = nexec; q < G.subs.size(); ++q) {
            const LGate& x = G.subs[q];
            const int b = block_index(x.target);
            if (x.ctrl < 0) cvec ^= 1u << b;
            else { const int a = block_index(x.ctrl); Arow[b] ^= Arow[a]; cvec ^= ((cvec >> a) & 1u) << b; }
        }
        for (size_t qi = 0; qi < nexec; ++qi) {
            const LGate& x = G.subs[qi];
            if (nrec + 4 > kV4MaxRecs) { ok = false; break; }
            int CB = -1; uint32_t cm = 0;
            if (x.ctrl >= 0) {
                const int cl = tile_local_of(P, x.ctrl);
                int bi = -1; if (cl >= 0) for (int i = 0; i < NB; ++i) if (rbl[i] == cl) bi = i;
                if (bi >= 0) CB = bi;
                else {
                    cm = ctx_bit(x.ctrl);
                    if (cl >= 0 && !((wmask >> cl) & 1)) ok = false;      // the control predicate must be warp-uniform
                }
            }
            if (x.slot >= 0) {                 // per-state coefficients: a SLOT record in front of the gate
                V4Rec& R = out.recs[nrec++]; memset(&R, 0, sizeof(R));
                R.op = V4_OP_SLOT; R.aux = (uint32_t)x.slot;
            }
            V4Rec& S = out.recs[nrec++]; memset(&S, 0, sizeof(S));
            const int tl = tile_local_of(P, x.target);
            int bi = -1; if (tl >= 0) for (int i = 0; i < NB; ++i) if (rbl[i] == tl) bi = i;
            if (!x.diag && bi < 0) ok = false;
            LiftCoef Lc; memset(&Lc, 0, sizeof(Lc));
            // D1 form of an uncontrolled diagonal gate (see the comment above encode_v4)
            bool d1_block = false, d1_ctx = false;
            double md1[8];
            if (gph && x.diag && x.ctrl < 0 && x.slot < 0 && (bi >= 0 || tl < 0 || ((warp_uniform >> tl) & 1u))) {
                const double d0r = x.m[0], d0i = x.m[1], d1r = x.m[6], d1i = x.m[7];
                double rr = d1r * d0r + d1i * d0i, ri = d1i * d0r - d1r * d0i;       // d1 * conj(d0)
                const double h = hypot(rr, ri);
                if (h > 0) { rr /= h; ri /= h; }
                const double md[8] = { rr, ri, 0, 0, 0, 0, rr, ri };
                LiftCoef Lt;
                if (lift_classify(md, Lt) == LK_PHASE) {
                    memcpy(md1, md, sizeof(md1));
                    const double gr = gph[0] * d0r - gph[1] * d0i, gi = gph[0] * d0i + gph[1] * d0r;
                    gph[0] = gr; gph[1] = gi;
                    if (bi >= 0) d1_block = true; else d1_ctx = true;
                }
            }
            const double* xm = (d1_block || d1_ctx) ? md1 : x.m;
            const int kind = x.slot >= 0 ? x.skind : lift_classify(xm, Lc);   // liftable was checked by v4_eligible
            uint32_t negflags = 0;
            if (canon && x.slot < 0 && (kind == LK_ROT || kind == LK_ROTX || kind == LK_PHASE)) {
                double mm[8]; memcpy(mm, xm, sizeof(mm));
                if (kind == LK_PHASE) {
                    if (Lc.mk[0] && Lc.mk[1]) { mm[0] = -mm[0]; mm[1] = -mm[1]; negflags |= 1u; }
                    if (Lc.mk[2] && Lc.mk[3]) { mm[6] = -mm[6]; mm[7] = -mm[7]; negflags |= 2u; }
                } else if (Lc.mk[1]) { for (int i = 0; i < 8; ++i) mm[i] = -mm[i]; negflags = 1u; }
                if (negflags) {
                    LiftCoef L2;
                    if (lift_classify(mm, L2) != kind) { ok = false; break; }
                    Lc = L2;
                    if (kind != LK_PHASE && x.ctrl < 0) {       // -1 on every amplitude: a global phase
                        gph[0] = -gph[0]; gph[1] = -gph[1]; negflags = 0;
                    }
                }
            }
            S.aux = negflags;
            double cc[6]; memcpy(cc, Lc.c, sizeof(cc));
            // static sign variants: sg = (mk[0] ? 1 : 0) | (mk[1] ? 2 : 0); per-state slots are built for sg 0
            const int sgA = (Lc.mk[0] ? 1 : 0) | (Lc.mk[1] ? 2 : 0), sgB = (Lc.mk[2] ? 1 : 0) | (Lc.mk[3] ? 2 : 0);
            if (kind == LK_PHASE && bi < 0) {
                // the target bit is constant for the thread: pick one of two coefficient sets
                S.tm = ctx_bit(x.target);
                int variant;
                if (sgA == sgB && (sgA == 0 || sgA == 3)) variant = sgA == 3 ? 1 : 0;
                else {                          // halves with different signs (z!): plain complex multiply by d0 / d1
                    variant = 2;
                    cc[0] = x.m[0]; cc[1] = x.m[1]; cc[2] = 0; cc[3] = x.m[6]; cc[4] = x.m[7]; cc[5] = 0;
                }
                S.op = (uint32_t)(V4_OP_THR + variant * (NB + 1) + (CB + 1));
                if (d1_ctx) cm |= S.tm;        // D1: only the threads (whole warps) that have the target bit run the record
            } else {
                // D1 on a register-block bit: the PHASE body "controlled" by the target bit itself (both sets hold the same phase)
                const int bc = d1_block ? v4_bc_index((bi + 1) % NB, bi) : v4_bc_index(bi, CB);
                int base;
                if (kind == LK_ROT) base = sgA * V4_NBC;
                else if (kind == LK_ROTX) base = V4_OP_ROTX + (sgA == 3 ? V4_NBC : 0);
                else if (kind == LK_PHASE) base = V4_OP_PHASE + ((sgA == 3 ? 2 : 0) + (sgB == 3 ? 1 : 0)) * V4_NBC;
                else base = V4_OP_PERMX;
                S.op = (uint32_t)(base + bc);
            }
            if (cm) { S.op += V4_NUM_GATE_OPS; S.cm = cm; }      // a control outside the register block: the body's controlled entry
            for (int q = 0; q < 3; ++q) { S.c[2 * q] = cc[q]; S.c[2 * q + 1] = cc[3 + q]; }     // pairwise {set 0, set 1}
        }
        if (!ok) break;
        V4Rec& E = out.recs[nrec++]; memset(&E, 0, sizeof(E));
        E.op = V4_OP_END;
        uint32_t so[NB] = { 0, 0, 0, 0 }, delta = 0;
        for (int i = 0; i < NB; ++i) {
            for (int j = 0; j < NB; ++j) if ((Arow[i] >> j) & 1u) so[j] ^= g.s[i];
            if ((cvec >> i) & 1u) delta ^= g.s[i];
        }
        E.tm = so[0]; E.cm = so[1]; E.aux = so[2];
        const uint64_t packed = (uint64_t)so[3] | ((uint64_t)delta << 32);
        memcpy(&E.c[0], &packed, 8);
    }
    out.nrecs = nrec; out.n_outer = n_outer; out.ctx_const = 0; out.flags = canon ? 1u : 0u;
    return ok;
}


}  // namespace qm
