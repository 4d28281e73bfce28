// planner.hpp — host-side pass planner of the B200 state-vector engine (no CUDA in this file).
//
// Turns a gate list (include/qm_b200.h qm_gate, 0-based textbook bits) into a sequence of
//   * fused HBM passes: every pass names a TILE of m index bits (the L lowest bits, so that
//     every row of a tile is 2^L contiguous amplitudes, plus m-L higher bits) and a list of
//     register-block GROUPS; each group names <= 3 tile bits held in registers and the gates
//     ("subs") applied while they are there, and
//   * (sharded registers only) REMAP steps that exchange the global qubits with the top local
//     ones by an all-to-all.
//
// What the reference does instead: one dense 2^n x 2^n kron operator and a matrix-vector
// product per gate (QSim.jl:52-63, :175-203).  The algebra below only uses facts about that
// operator: a gate touches the amplitudes pairwise along its target bit; a control bit or a
// diagonal target never needs to be resident in the tile (it is a predicate / a phase), and two
// gates commute when every qubit they share is acted on diagonally by both.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../include/qm_b200.h"

namespace qm {

static const int kMaxTileBits = 13;
static const int kBlockBits = 3;        // general (v1) kernel: amplitudes per thread = 2^3
static const int kMaxSubsPerPass = 160;

// ---- device-visible encodings (plain structs, copied byte for byte) -----------------------
enum { SUB_MAT = 0, SUB_DIAG = 1 };

struct DevSub {
    int32_t  kind;       // SUB_MAT | SUB_DIAG
    int32_t  b;          // MAT: block bit 0..2 of the target.  DIAG: 0..2 block bit, 3 tile-local bit, 4 outer bit
    uint32_t cmask_blk;  // control bits among the block bits (bit i = block bit i)
    uint32_t cmask_loc;  // control bits among the other tile-local bits
    uint64_t cmask_out;  // control bits outside the tile (mask on the global amplitude index)
    uint32_t tmask_loc;  // DIAG b==3: tile-local mask of the target bit
    int32_t  slot;       // >= 0: matrix comes from per-state slot (batched), else -1
    uint64_t tmask_out;  // DIAG b==4: global-index mask of the target bit
    double   m[8];       // U11 U21 U12 U22 (re, im) — Julia memory order
    uint64_t pad_;
};
static_assert(sizeof(DevSub) == 112, "DevSub layout");

struct DevGroup {
    int32_t  nsub, sub_off;
    uint32_t rbmask[3];  // tile-local masks of the block bits (0 = unused)
    int32_t  nfree;      // m - (number of block bits)
    uint8_t  fpos[16];   // tile-local bit position of thread-index bit j
    uint32_t pad_[2];
};
static_assert(sizeof(DevGroup) == 48, "DevGroup layout");

struct DevPass {
    int32_t m, L, nhigh, ngroups;
    int32_t nsubs, total_bytes, flags, pad_;
    int32_t hb[16];      // physical bit of tile-local bit L + j
};
static_assert(sizeof(DevPass) == 96, "DevPass layout");

// ---- planner-side structures ---------------------------------------------------------------
struct LGate {           // lowered gate on PHYSICAL index bits
    int target = 0;
    int ctrl = -1;       // -1 = none
    bool diag = false;
    int slot = -1;       // -1 = none, else index into the per-state (sub)slot table
    int skind = -1;      // slot gates: lifting kind of the (sub)slot (lifting.hpp LK_*)
    bool liftable = true;  // has an in-place lifting form (lifting.hpp); false -> its pass uses the v1 kernel
    double m[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
};

// how the caller's per-state slots map onto (sub)slots: slot s -> count[s] consecutive subslots
// starting at first[s], each of lifting kind kind[first[s] + i] (general unitaries take three)
struct SlotMap {
    std::vector<int> first, count, kind;
};

struct PlanSub { LGate g; };
struct PlanGroup {
    std::vector<int> rb;     // register-block bits (physical)
    std::vector<int> wb;     // tile bits that must be warp-uniform for this group (controls that are not block bits)
    std::vector<LGate> subs;
};
struct PlanPass {
    int m = 0, L = 0;
    std::vector<int> hb;     // the tile's index bits above the L low ones, ascending
    // TMA kernel only (empty otherwise): the order of the tile's bits in SHARED MEMORY.  tl[p] = index bit at
    // tile-local position p; positions 0-2 are always bits 0-2 (one 128-byte row), then come the segments
    // segs[] = (first bit, length) in the order chosen by choose_tile_layout().  Without it tile-local position
    // p is bit p for p < L and hb[p - L] above.
    std::vector<int> tl;
    std::vector<std::pair<int, int>> segs;
    std::vector<PlanGroup> groups;
    int ngates = 0;
    double cost = 0.0;   // estimated DFMA-equivalents per amplitude
};
struct PlanStep {
    int kind = 0;        // 0 = pass, 1 = remap (swap global bits with the top local bits)
    PlanPass pass;
};

struct PlanOptions {
    int n_local = 0;     // index bits resident on this rank
    int n_global = 0;    // sharded (rank) bits above them
    int tile_bits = 12;
    int low_bits = 5;
    bool fuse = true;
    int window = 768;    // look-ahead, in gates
    double max_cost = 1e30;
    int max_runs = 99;   // runs of consecutive high tile bits a pass may use (TMA tensor maps: 3)
    int max_run_len = 99;  // ... a run longer than this counts as several (TMA box dimension <= 256: 8)
    int max_subs = kMaxSubsPerPass;
    int max_groups = 1 << 30;
    int max_outer = 64;  // distinct bits outside the tile a pass may use as controls / diagonal targets
    int block_bits = kBlockBits;     // register-block bits of a group (3: general kernel, 4: TMA kernel)
    bool retile = false;             // rebuild the tile around the targets once the gates of a pass are known (TMA kernel)
    double group_cost = 0.0;         // budget units one register-block group costs on top of its gates (TMA kernel: 38)
    double short_mult = 1e9;         // circuits shorter than the window may use up to short_mult x the budget (Planner::budget_for_pending)
    double hbm_units = 114.0;        // fitted pass model in budget units (0.05 ms at 30 qubits): HBM time of a pass,
    double fixed_units = 41.0;       // compute-side time = fixed_units + group_cost per group + gate costs
    int low_bits_short = 4;          // low tile bits of a circuit shorter than the look-ahead window (0: opt.low_bits)
    bool defer_low3 = true;          // ... and let the stretch after a postponement keep 3 low bits when that is what makes the tile fit
    bool defer_bias = true;          // sharded: schedule the passes on the outgoing qubits first (choose_tile, try_defer_last_pass)
    int cost_model = 0;              // gate_cost: 0 the interpreter kernel, 1 the compiled pass kernels
    bool warp_uniform_ctrl = false;  // controls inside the tile must be block bits or one of the (m - block_bits - 5) warp-index bits
};

inline bool mat_is_diag(const double* m) { return m[2] == 0.0 && m[3] == 0.0 && m[4] == 0.0 && m[5] == 0.0; }
inline bool mat_is_identity(const double* m) {
    return mat_is_diag(m) && m[0] == 1.0 && m[1] == 0.0 && m[6] == 1.0 && m[7] == 0.0;
}

// C = A * B for 2x2 complex in Julia order (U11 U21 U12 U22)
inline void mat_mul(const double* A, const double* B, double* C) {
    auto at = [](const double* M, int r, int c, double& re, double& im) { re = M[(c * 2 + r) * 2]; im = M[(c * 2 + r) * 2 + 1]; };
    double out[8];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) {
            double sr = 0, si = 0;
            for (int k = 0; k < 2; ++k) {
                double ar, ai, br, bi; at(A, r, k, ar, ai); at(B, k, c, br, bi);
                sr += ar * br - ai * bi; si += ar * bi + ai * br;
            }
            out[(c * 2 + r) * 2] = sr; out[(c * 2 + r) * 2 + 1] = si;
        }
    memcpy(C, out, sizeof(out));
}

static const double kXmat[8] = { 0, 0, 1, 0, 1, 0, 0, 0 };

// Lower wire gates to LGates on physical bits (perm[logical] = physical), merging runs of
// uncontrolled 1-qubit gates on the same qubit into one matrix when `fuse` is set.
inline int lower_gates(const qm_gate* gates, int ng, int nbits, const int* perm, bool fuse,
                       std::vector<LGate>& out, const SlotMap* slots = nullptr) {
    std::vector<int> pending(nbits, -1);   // index in `out` of a still-mergeable 1q gate per physical bit
    auto touch = [&](int bit) { pending[bit] = -1; };
    for (int i = 0; i < ng; ++i) {
        const qm_gate& w = gates[i];
        switch (w.kind) {
        case QM_GATE_U2: case QM_GATE_U2_SLOT: {
            if (w.q0 < 0 || w.q0 >= nbits) return QM_ERR_ARG;
            int t = perm ? perm[w.q0] : w.q0;
            LGate g; g.target = t; g.ctrl = -1; g.slot = (w.kind == QM_GATE_U2_SLOT) ? w.q1 : -1;
            memcpy(g.m, w.m, sizeof(g.m));
            g.diag = (g.slot < 0) && mat_is_diag(g.m);
            if (g.slot >= 0 && slots) {
                // one LGate per subslot, in application order
                const int s0 = slots->first[g.slot], cnt = slots->count[g.slot];
                for (int i = 0; i < cnt; ++i) {
                    LGate h = g; h.slot = s0 + i; h.skind = slots->kind[s0 + i]; h.diag = (h.skind == 2);
                    out.push_back(h);
                }
                pending[t] = -1;
            } else if (fuse && g.slot < 0 && pending[t] >= 0) {
                LGate& p = out[pending[t]];
                mat_mul(g.m, p.m, p.m);                 // later gate multiplies from the left
                p.diag = mat_is_diag(p.m);
            } else {
                out.push_back(g);
                pending[t] = (g.slot < 0 && fuse) ? (int)out.size() - 1 : -1;
            }
            break;
        }
        case QM_GATE_CU2: {
            if (w.q0 < 0 || w.q0 >= nbits || w.q1 < 0 || w.q1 >= nbits || w.q0 == w.q1) return QM_ERR_ARG;
            LGate g; g.ctrl = perm ? perm[w.q0] : w.q0; g.target = perm ? perm[w.q1] : w.q1; g.slot = -1;
            memcpy(g.m, w.m, sizeof(g.m)); g.diag = mat_is_diag(g.m);
            touch(g.ctrl); touch(g.target);
            out.push_back(g);
            break;
        }
        case QM_GATE_SWAP: {
            if (w.q0 < 0 || w.q0 >= nbits || w.q1 < 0 || w.q1 >= nbits) return QM_ERR_ARG;
            if (w.q0 == w.q1) break;                    // swap!(s,q,q) is a no-op, QSim.jl:266-268
            int a = perm ? perm[w.q0] : w.q0, b = perm ? perm[w.q1] : w.q1;
            touch(a); touch(b);
            for (int k = 0; k < 3; ++k) {               // three CNOTs, QSim.jl:270-272
                LGate g; g.ctrl = (k == 1) ? b : a; g.target = (k == 1) ? a : b; g.slot = -1; g.diag = false;
                memcpy(g.m, kXmat, sizeof(g.m));
                out.push_back(g);
            }
            break;
        }
        default: return QM_ERR_ARG;
        }
    }
    // drop merged identities (e.g. X*X): exact no-ops
    if (fuse) {
        size_t w = 0;
        for (size_t i = 0; i < out.size(); ++i)
            if (!(out[i].ctrl < 0 && out[i].slot < 0 && mat_is_identity(out[i].m))) out[w++] = out[i];
        out.resize(w);
    }
    return QM_OK;
}

// Planner cost model.  One unit = 0.05 ms of a pass at 30 qubits on B200 (it scales with the register size, the
// ratios do not).  Fitted on per-pass CUDA-event times (tools/pass_model.py, ~300 compute-bound passes, rms 0.4 ms):
//     pass ms = 2.04 + 0.89 per register-block group + 0.229 per full-cost record (an uncontrolled rotation or
//     phase: 0.194 ms is the FP64 pipe at peak) + 0.091 per record whose control is a register-block bit
//     + 0.136 per xor exchange (x!, cnot!) + 0.246 per thread-constant phase
// against ~5.6 ms of HBM time.  A matrix that needs phase * rotation * phase is three records.
//
// cost_model 1 — the COMPILED pass kernels (jit.hpp): no dispatch, no record fetch, x!/cnot! on register-block bits are register
// renames.  Same fit on the compiled kernels (profiles/r3b_*: 192 passes at 30 qubits): pass ms = 1.2 + 1.45 per group + 0.037
// per FP64 operation per amplitude — a full-cost record is 0.11 ms (2.3 units), a single group carries 22-24 of them inside the
// 5.5 ms of HBM time, a second group costs as much as 13 records.
inline double gate_cost(const LGate& g, int cost_model = 0) {
    const bool is_x = !g.diag && g.m[0] == 0 && g.m[1] == 0 && g.m[6] == 0 && g.m[7] == 0 && g.m[2] == 1 && g.m[3] == 0 &&
                      g.m[4] == 1 && g.m[5] == 0;
    const bool real_or_rx = (g.m[1] == 0 && g.m[7] == 0) && ((g.m[3] == 0 && g.m[5] == 0) || (g.m[2] == 0 && g.m[4] == 0));
    const double records = (g.diag || real_or_rx || g.slot >= 0) ? 1.0 : 3.0;
    // an uncontrolled diagonal gate runs in D1 form (encode_v4): only the amplitudes that have the target bit are touched
    const bool d1 = g.diag && g.ctrl < 0 && g.slot < 0;
    if (cost_model == 1) return is_x ? 0.4 : d1 ? 1.3 : records == 3.0 ? (g.ctrl >= 0 ? 3.6 : 4.9) : (g.ctrl >= 0 ? 1.2 : 2.3);
    if (is_x) return 2.7;
    if (d1) return 3.0;
    return records == 3.0 ? (g.ctrl >= 0 ? 7.5 : 10.7) : (g.ctrl >= 0 ? 2.5 : 4.6);
}

// Commutation-aware pull: walk the not-yet-done gates in order; take a gate when no earlier
// skipped gate blocks it and allowed(gate) accepts it (the pass level asks for the non-diagonal target
// to be in the tile; the group level grows the <= 3-bit register block and the warp-uniform set).
// `tile_mask` / `max_outer`: a pass may use at most max_outer distinct bits outside tile_mask as controls
// or diagonal targets (the kernel keeps them in a 32-bit per-thread context word).
template <class Allowed>
inline void pull(const std::vector<LGate>& g, const std::vector<char>& done, size_t first, int window,
                 uint64_t all_bits, double max_cost, int max_take, Allowed&& allowed, std::vector<int>& taken,
                 uint64_t tile_mask = ~0ull, int max_outer = 64, int cost_model = 0) {
    uint64_t blockedN = 0, blockedD = 0, outer_used = 0;
    int scanned = 0; double cost = 0;
    for (size_t i = first; i < g.size() && scanned < window; ++i) {
        if (done[i]) continue;
        ++scanned;
        const LGate& x = g[i];
        uint64_t tm = 1ull << x.target, cm = x.ctrl >= 0 ? (1ull << x.ctrl) : 0ull;
        uint64_t actsN = x.diag ? 0ull : tm, actsD = cm | (x.diag ? tm : 0ull);
        bool free_ = !((actsN | actsD) & blockedN) && !(actsN & blockedD);
        bool take = false;
        const uint64_t ob = (cm | (x.diag ? tm : 0ull)) & ~tile_mask;
        if (free_ && (int)taken.size() < max_take && (taken.empty() || cost + gate_cost(x, cost_model) <= max_cost) &&
            __builtin_popcountll(outer_used | ob) <= max_outer)
            take = allowed(x);
        if (take) { taken.push_back((int)i); cost += gate_cost(x, cost_model); outer_used |= ob; }
        else {
            blockedN |= actsN; blockedD |= actsD;
            if ((blockedN & all_bits) == all_bits) break;
        }
    }
}

struct Planner {
    PlanOptions opt;
    std::vector<LGate> gates;
    std::vector<char> done;
    size_t first = 0;
    double budget_now = 0.0;         // budget of the passes being built (budget_for_pending)
    bool budget_fixed = false, mode_known = false, short_circuit = false;
    std::vector<int> last_taken;     // gate indices of the pass next_pass built last
    std::vector<std::pair<PlanPass, std::vector<int>>> replay;   // the chosen plan of a short stretch, handed out by next_pass
    size_t replay_at = 0;
    int saved_low_bits = -1;         // try_defer_last_pass lowered opt.low_bits for one stretch: restored when the stretch ends

    int total_bits() const { return opt.n_local + opt.n_global; }

    // non-diagonal targets still pending on global (sharded) bits?
    bool pending_global_target() const {
        for (size_t i = first; i < gates.size(); ++i)
            if (!done[i] && !gates[i].diag && gates[i].target >= opt.n_local) return true;
        return false;
    }

    std::vector<int> tile_bits_for(const std::vector<int>& hb, int L) const {
        std::vector<int> t; for (int i = 0; i < L; ++i) t.push_back(i);
        t.insert(t.end(), hb.begin(), hb.end()); return t;
    }

    int count_with(uint64_t tilemask) const {
        std::vector<int> taken;
        pull(gates, done, first, opt.window, (1ull << total_bits()) - 1, opt.max_cost, opt.max_subs,
             [&](const LGate& x) { return x.diag || ((tilemask >> x.target) & 1); }, taken, tilemask, opt.max_outer, opt.cost_model);
        return (int)taken.size();
    }

    // choose the high tile bits for the next pass
    // tmask_best: the bits non-diagonal targets may use in the chosen pass (all tile bits, or one 4-bit block)
    void choose_tile(int m, int L, std::vector<int>& hb_best, int& best, uint64_t& tmask_best) const {
        const int nl = opt.n_local, nh = m - L;
        hb_best.clear(); best = -1; tmask_best = ~0ull;
        double best_score = -1.0;
        const bool outgoing_first = opt.n_global > 0 && opt.defer_bias && pending_global_target();
        if (nh <= 0) { best = count_with((1ull << L) - 1); return; }
        uint64_t low = (1ull << L) - 1;
        auto try_set = [&](const std::vector<int>& hb, uint64_t tmask = ~0ull) {
            // runs of consecutive bits, a run being cut after max_run_len bits (one TMA box dimension holds <= 256)
            int runs = 0, len = 0;
            for (size_t i = 0; i < hb.size(); ++i) {
                if (i == 0 || hb[i] != hb[i - 1] + 1 || len == opt.max_run_len) { ++runs; len = 0; }
                ++len;
            }
            if (runs > opt.max_runs) return;
            uint64_t mask = low; for (int b : hb) mask |= 1ull << b;
            int c; double score;
            if (opt.group_cost > 0) {       // what the pass built on this tile would really take
                PlanPass tmp; std::vector<int> tk;
                build_pass(m, L, hb, tmp, tk, tmask);
                c = (int)tk.size(); score = c;
                // Sharded register with an exchange ahead: passes that work on the OUTGOING qubits (the top n_global local
                // bits) go first, so that the last pass before the exchange does not need them and can be postponed to
                // after it (try_defer_last_pass) — worth a few gates of difference between otherwise similar tiles.
                if (outgoing_first && c > 0) {
                    uint64_t out_q = 0;
                    for (int i : tk) if (!gates[i].diag && gates[i].target >= nl - opt.n_global) out_q |= 1ull << gates[i].target;
                    score += 4.0 * __builtin_popcountll(out_q);          // per outgoing qubit the pass finishes
                }
            } else { c = count_with(mask); score = c; }
            if (score > best_score) { best_score = score; best = c; hb_best = hb; tmask_best = tmask; }
        };
        // (1) greedy: the first nh distinct non-diagonal high local targets in dependency order
        {
            std::vector<int> hb; uint64_t have = low;
            std::vector<int> taken;
            pull(gates, done, first, opt.window, (1ull << total_bits()) - 1, 1e30, 1 << 30,
                 [&](const LGate& x) {
                     if (x.diag) return true;
                     const int t = x.target;
                     if (t >= nl) return false;
                     if ((have >> t) & 1) return true;
                     if ((int)hb.size() < nh) { hb.push_back(t); have |= 1ull << t; return true; }
                     return false;
                 }, taken);
            // filler bits: first the ones that lengthen an existing run of high bits (a tile may have at most max_runs runs:
            // a lone filler at bit L next to targets at 12-15, 24-25 and 32 would be a fourth), then the lowest free ones
            for (bool grown = true; grown && (int)hb.size() < nh;) {
                grown = false;
                for (int b = L; b < nl && (int)hb.size() < nh; ++b) {
                    if ((have >> b) & 1) continue;
                    const bool below = b > L && ((have >> (b - 1)) & 1), above = b + 1 < nl && ((have >> (b + 1)) & 1);
                    if ((below || above) && !hb.empty()) { hb.push_back(b); have |= 1ull << b; grown = true; }
                }
            }
            for (int b = L; b < nl && (int)hb.size() < nh; ++b) if (!((have >> b) & 1)) { hb.push_back(b); have |= 1ull << b; }
            std::sort(hb.begin(), hb.end());
            try_set(hb);
        }
        // (2) contiguous windows
        for (int a = L; a + nh <= nl; ++a) {
            std::vector<int> hb; for (int j = 0; j < nh; ++j) hb.push_back(a + j);
            try_set(hb);
        }
        // (3) group-aware budget: a pass is essentially ONE register block, so every window of block_bits
        // consecutive qubits is tried as that block directly (targets restricted to it), the rest of the tile
        // being filler; the greedy grouping inside a wider tile often settles on a worse block
        if (opt.group_cost > 0 && nh >= opt.block_bits) {
            for (int a = 0; a + opt.block_bits <= nl; ++a) {
                uint64_t w = 0; for (int j = 0; j < opt.block_bits; ++j) w |= 1ull << (a + j);
                std::vector<int> hb;
                for (int b = L; b < nl; ++b) if ((w >> b) & 1) hb.push_back(b);
                for (int b = L; b < nl && (int)hb.size() < nh; ++b) if (!((w >> b) & 1)) hb.push_back(b);
                if ((int)hb.size() != nh) continue;
                std::sort(hb.begin(), hb.end());
                try_set(hb, w);
            }
        }
    }

    // Build the next pass.  Returns false when nothing can be pulled (only global targets left).
    bool next_pass(PlanPass& P) {
        while (first < gates.size() && done[first]) ++first;
        if (first >= gates.size()) return false;
        const int nl = opt.n_local;
        int m = std::min(opt.tile_bits, nl); if (m > kMaxTileBits) m = kMaxTileBits;
        int L = std::min(opt.low_bits, m);
        std::vector<int> hb; int best = 0;
        if (!opt.fuse) {
            // debug mode: one gate per pass
            const LGate& x = gates[first];
            if (!x.diag && x.target >= nl) return false;
            if (m > L) {
                for (int b = L; b < nl && (int)hb.size() < m - L; ++b) hb.push_back(b);
                if (!x.diag && x.target >= L && std::find(hb.begin(), hb.end(), x.target) == hb.end()) hb.back() = x.target;
                std::sort(hb.begin(), hb.end());
            }
            P = PlanPass(); P.m = m; P.L = L; P.hb = hb;
            PlanGroup G; if (!x.diag) G.rb.push_back(x.target);
            if (opt.warp_uniform_ctrl && x.ctrl >= 0 && (x.ctrl < L || std::find(hb.begin(), hb.end(), x.ctrl) != hb.end()))
                G.rb.push_back(x.ctrl);
            G.subs.push_back(x); P.groups.push_back(G); P.ngates = 1; P.cost = gate_cost(x, opt.cost_model);
            done[first] = 1;
            return true;
        }
        uint64_t tmask = ~0ull;
        if (!budget_fixed) budget_now = budget_for_pending();
        // a short circuit (a reservoir step) wants as many NEW qubits per pass as it can get — its passes are few and every one
        // of them costs at least the HBM time — so its tiles keep 4 low bits instead of 5 (8 new qubits per pass: a 33-qubit
        // layer is 12 + 8 + 8 + 5 and never leaves a lone gate for a fifth pass); deep circuits keep opt.low_bits
        if (mode_known && short_circuit && opt.low_bits_short > 0 && L > opt.low_bits_short && m - opt.low_bits_short >= 1) L = opt.low_bits_short;
        if (replay_at < replay.size()) {            // the stretch was already planned while choosing its budget
            P = replay[replay_at].first;
            last_taken = replay[replay_at].second;
            for (int i : last_taken) done[i] = 1;
            ++replay_at;
            return true;
        }
        if (!replay.empty()) { replay.clear(); replay_at = 0; budget_fixed = false; return false; }
        choose_tile(m, L, hb, best, tmask);
        if (best <= 0) {                                            // end of a stretch (a remap follows): choose again
            budget_fixed = false;
            if (saved_low_bits >= 0) { opt.low_bits = saved_low_bits; saved_low_bits = -1; }
            return false;
        }
        std::vector<int> taken;
        build_pass(m, L, hb, P, taken, tmask);
        for (int i : taken) done[i] = 1;
        if (opt.retile && !taken.empty()) retile(P);
        last_taken.swap(taken);
        return !last_taken.empty();
    }

    // Once the gates of a pass are known, only the non-diagonal targets have to be tile bits (controls and
    // diagonal targets outside the tile are bits of the thread's context word).  The tile is rebuilt as: as many
    // LOW bits as fit (up to 8: 4 KiB contiguous rows) + the targets above them + the next low bits as filler.
    // A tile whose seven high bits sit at bits 21-27 is 128 rows of 512 bytes 32 MiB apart — 128 different 2 MiB
    // pages per tile — and such passes took 9 ms against 5.7 ms for the same work on contiguous tiles.
    void retile(PlanPass& P) const {
        uint64_t T = 0, outer_users = 0;
        for (const PlanGroup& G : P.groups) {
            for (int b : G.rb) T |= 1ull << b;          // every block bit stays a tile bit (a control may have been made one)
            for (const LGate& x : G.subs) {
                if (!x.diag) T |= 1ull << x.target; else outer_users |= 1ull << x.target;
                if (x.ctrl >= 0) outer_users |= 1ull << x.ctrl;
            }
        }
        const int m = P.m, nl = opt.n_local;
        for (int L2 = std::min(8, m); L2 >= P.L; --L2) {
            const uint64_t high = T & ~((1ull << L2) - 1ull);
            const int nh = m - L2;
            if (__builtin_popcountll(high) > nh) continue;
            std::vector<int> hb2;
            for (int b = L2; b < nl; ++b) if ((high >> b) & 1) hb2.push_back(b);
            for (int b = L2; b < nl && (int)hb2.size() < nh; ++b) if (!((high >> b) & 1)) hb2.push_back(b);
            if ((int)hb2.size() != nh) continue;
            std::sort(hb2.begin(), hb2.end());
            int runs = 0, len = 0;
            for (size_t i = 0; i < hb2.size(); ++i) {
                if (i == 0 || hb2[i] != hb2[i - 1] + 1 || len == opt.max_run_len) { ++runs; len = 0; }
                ++len;
            }
            if (runs > opt.max_runs) continue;
            uint64_t tilemask = (1ull << L2) - 1ull; for (int b : hb2) tilemask |= 1ull << b;
            if (__builtin_popcountll(outer_users & ~tilemask) > opt.max_outer) continue;
            // controls inside the new tile must still be block bits or fit the warp-index bits of their group
            bool ok = true;
            if (opt.warp_uniform_ctrl) {
                const int wcap = m > opt.block_bits + 5 ? m - opt.block_bits - 5 : 0;
                for (PlanGroup& G : P.groups) {
                    uint64_t rbm = 0; for (int b : G.rb) rbm |= 1ull << b;
                    uint64_t w = 0;
                    for (const LGate& x : G.subs) if (x.ctrl >= 0 && ((tilemask >> x.ctrl) & 1) && !((rbm >> x.ctrl) & 1)) w |= 1ull << x.ctrl;
                    if (__builtin_popcountll(w) > wcap) { ok = false; break; }
                }
                if (ok) for (PlanGroup& G : P.groups) {
                    uint64_t rbm = 0; for (int b : G.rb) rbm |= 1ull << b;
                    G.wb.clear();
                    for (const LGate& x : G.subs)
                        if (x.ctrl >= 0 && ((tilemask >> x.ctrl) & 1) && !((rbm >> x.ctrl) & 1) &&
                            std::find(G.wb.begin(), G.wb.end(), x.ctrl) == G.wb.end()) G.wb.push_back(x.ctrl);
                }
            }
            if (!ok) continue;
            P.L = L2; P.hb = hb2;
            return;
        }
    }

    // Build the pass for a given tile: pull gates, form register-block groups, keep the leading groups that fit
    // the budget.  `taken` receives the indices (into `gates`) of the gates of the pass; nothing is marked done.
    //
    // Budget (opt.max_cost) = sum of gate costs + opt.group_cost per group.  Measured on B200 at 30 qubits
    // (tools/pass_model.py, final kernel): a pass costs about 2.04 ms + 0.89 ms per group + 0.05 ms per cost unit
    // against 5.7 ms of HBM time, i.e. one register-block group is worth 18 cost units — a pass with ONE well-filled
    // group (13-14 gates) stays on the HBM roofline, a second group does not.
    //
    // Short circuits: the budget keeps a pass at HBM time, which is the right trade when the circuit is deep enough
    // to fill one register block per pass (C3: 13-14 gates).  A reservoir step is not — 2-4 gates per qubit, then a
    // readout — so every block leaves most of the HBM time idle, yet each block costs a whole pass.  When the whole
    // circuit fits the look-ahead window, the planner therefore plans what is pending at a few multiples of the
    // budget and keeps the one the fitted pass model (PlanOptions::hbm_units etc.) says finishes first: typically two
    // blocks per pass at ~1.3x the pass time instead of twice the passes.  Deep circuits keep the budget as it is.
    double predicted_units(const PlanPass& P) const {
        return std::max(opt.hbm_units, opt.fixed_units + opt.group_cost * (double)P.groups.size() + P.cost);
    }

    double budget_for_pending() {
        if (opt.group_cost <= 0 || opt.short_mult <= 1.0 || opt.max_cost > 1e20) { budget_fixed = true; return opt.max_cost; }
        if (!mode_known) {
            size_t pending = 0;
            for (size_t i = first; i < gates.size(); ++i) pending += !done[i];
            short_circuit = (int)pending < opt.window;
            mode_known = true;
        }
        budget_fixed = true;
        if (!short_circuit) return opt.max_cost;
        static const double mults[] = {1.0, 1.8, 3.0, 1e9};
        double best_t = 1e300, best_b = opt.max_cost;
        std::vector<int> prev_sig;
        for (double mu : mults) {
            if (mu > opt.short_mult) mu = opt.short_mult;
            Planner trial = *this;
            trial.replay.clear(); trial.replay_at = 0;
            trial.budget_now = std::min(mu * opt.max_cost, 1e30);
            double t = 0.0, work = 0.0; PlanPass P;
            std::vector<std::pair<PlanPass, std::vector<int>>> passes;
            while (trial.next_pass(P)) {
                t += predicted_units(P); work += P.cost;
                passes.emplace_back(P, trial.last_taken);
            }
            if (work <= 0.0) break;
            std::vector<int> sig;                // the larger budget changed nothing: neither will the next
            for (const auto& pr : passes) { sig.push_back(pr.first.ngates); sig.push_back((int)pr.first.groups.size()); }
            const bool same = sig == prev_sig;
            prev_sig.swap(sig);
            if (t / work < 0.99 * best_t) { best_t = t / work; best_b = trial.budget_now; replay.swap(passes); }
            if (mu >= opt.short_mult || same) break;
        }
        replay_at = 0;
        return best_b;
    }

    void build_pass(int m, int L, const std::vector<int>& hb, PlanPass& P, std::vector<int>& taken_out,
                    uint64_t target_mask = ~0ull) const {
        build_pass_at(budget_now > 0 ? budget_now : opt.max_cost, m, L, hb, P, taken_out, target_mask);
    }

    void build_pass_at(double max_cost, int m, int L, const std::vector<int>& hb, PlanPass& P, std::vector<int>& taken_out,
                       uint64_t target_mask = ~0ull) const {
        uint64_t tilemask = (1ull << L) - 1; for (int b : hb) tilemask |= 1ull << b;
        const uint64_t tgt = tilemask & target_mask;
        const double gate_budget = max_cost > opt.group_cost + 1.0 ? max_cost - opt.group_cost : 1.0;
        int max_take = opt.max_subs;
        taken_out.clear();
        while (true) {
            std::vector<int> taken;
            pull(gates, done, first, opt.window, (1ull << total_bits()) - 1, gate_budget, max_take,
                 [&](const LGate& x) { return x.diag || ((tgt >> x.target) & 1); }, taken, tilemask, opt.max_outer, opt.cost_model);
            P = PlanPass(); P.m = m; P.L = L; P.hb = hb;
            std::vector<LGate> pg; pg.reserve(taken.size());
            for (int i : taken) pg.push_back(gates[i]);
            std::vector<std::vector<int>> gidx;      // per group: indices into pg
            // register-block groups: same pull with an allowed set that grows up to block_bits bits
            std::vector<char> d2(pg.size(), 0); size_t f2 = 0;
            while (true) {
                while (f2 < pg.size() && d2[f2]) ++f2;
                if (f2 >= pg.size()) break;
                // A group holds <= block_bits register-block bits (every non-diagonal target) and, when the
                // kernel wants warp-uniform control predicates (opt.warp_uniform_ctrl), every control inside
                // the tile is either a block bit or one of the tile bits pinned to the warp index.
                PlanGroup G; uint64_t have = 0, wset = 0;
                const int bb = opt.block_bits;
                const int wcap = m > bb + 5 ? m - bb - 5 : 0;
                std::vector<int> tk;
                pull(pg, d2, f2, 1 << 30, (1ull << total_bits()) - 1, 1e30, 1 << 30,
                     [&](const LGate& x) {
                         uint64_t h2 = have, w2 = wset;
                         if (!x.diag && !((h2 >> x.target) & 1)) {
                             if (__builtin_popcountll(h2) >= bb) return false;
                             h2 |= 1ull << x.target; w2 &= ~(1ull << x.target);
                         }
                         if (opt.warp_uniform_ctrl && x.ctrl >= 0 && ((tilemask >> x.ctrl) & 1) &&
                             !(((h2 | w2) >> x.ctrl) & 1)) {
                             if (__builtin_popcountll(w2) < wcap) w2 |= 1ull << x.ctrl;
                             else if (__builtin_popcountll(h2) < bb) h2 |= 1ull << x.ctrl;
                             else return false;
                         }
                         have = h2; wset = w2;
                         return true;
                     }, tk);
                for (int i : tk) { G.subs.push_back(pg[i]); d2[i] = 1; }
                for (int b = 0; b < 64; ++b) { if ((have >> b) & 1) G.rb.push_back(b); if ((wset >> b) & 1) G.wb.push_back(b); }
                P.groups.push_back(G);
                gidx.push_back(tk);
            }
            if ((int)P.groups.size() > opt.max_groups && taken.size() > 1) {
                // too many register-block groups for the kernel's program buffer: take fewer gates
                max_take = (int)((double)taken.size() * opt.max_groups / (double)P.groups.size());
                if (max_take < 1) max_take = 1;
                continue;
            }
            // keep the leading groups that fit the budget (the first one is cut to fit if it has to be)
            double est = 0.0; size_t keep = 0;
            for (size_t g = 0; g < P.groups.size(); ++g) {
                double gc = opt.group_cost;
                for (const LGate& x : P.groups[g].subs) gc += gate_cost(x, opt.cost_model);
                if (g > 0 && est + gc > max_cost) break;
                bool cut = false;
                if (g == 0 && gc > max_cost) {
                    double c = opt.group_cost; size_t nk = 0;
                    for (const LGate& x : P.groups[0].subs) { if (nk > 0 && c + gate_cost(x, opt.cost_model) > max_cost) break; c += gate_cost(x, opt.cost_model); ++nk; }
                    P.groups[0].subs.resize(nk); gidx[0].resize(nk);
                    gc = c; cut = true;
                }
                est += gc; keep = g + 1;
                if (cut) break;
            }
            P.groups.resize(keep); gidx.resize(keep);
            P.cost = 0.0; P.ngates = 0;
            for (size_t g = 0; g < keep; ++g)
                for (size_t j = 0; j < gidx[g].size(); ++j) { taken_out.push_back(taken[gidx[g][j]]); P.cost += gate_cost(pg[gidx[g][j]], opt.cost_model); ++P.ngates; }
            break;
        }
    }
};

// Sharded registers: should the LAST pass planned before an exchange be postponed to after it?  A pass brings at most
// 12 - low_bits new qubits into its tile, so a reservoir layer over 33 local qubits takes 4 passes and the few gates on the
// ex-global qubits a 5th.  When the last of the 4 does not target an outgoing qubit (the top g local bits become global), it
// and the ex-global gates often fit ONE tile after the exchange: 4 passes per step instead of 5.  Both orders are planned on
// copies of the planner; the postponement is kept only if the stretch after the exchange does not grow.  On true the gates
// of the pass are pending again (the caller drops the pass and takes its global-phase factor back).
inline bool try_defer_last_pass(Planner& pl, const std::vector<int>& taken, int nl, int g) {
    if (g <= 0 || taken.empty()) return false;
    for (int i : taken) { const LGate& x = pl.gates[i]; if (!x.diag && x.target >= nl - g) return false; }
    auto swap_bit = [&](int b) { return b >= nl ? b - g : (b >= nl - g ? b + g : b); };      // dist_bits.hpp dist_swap_bit
    auto passes_after = [&](bool defer, int low_bits) -> int {
        Planner t = pl;
        t.replay.clear(); t.replay_at = 0; t.budget_fixed = false; t.first = 0;
        t.opt.low_bits = low_bits;
        if (defer) for (int i : taken) t.done[i] = 0;
        for (size_t i = 0; i < t.gates.size(); ++i) {
            if (t.done[i]) continue;
            t.gates[i].target = swap_bit(t.gates[i].target);
            if (t.gates[i].ctrl >= 0) t.gates[i].ctrl = swap_bit(t.gates[i].ctrl);
        }
        int np = 0; PlanPass Pt;
        while (np < 64 && t.next_pass(Pt)) ++np;
        return np;
    };
    const int keep = passes_after(false, pl.opt.low_bits);
    int moved = passes_after(true, pl.opt.low_bits);
    // One qubit too many for the tile?  The stretch after the exchange may keep 3 low bits instead of 4 (9 new qubits per
    // pass): its 128-byte rows make that pass ~10 % slower, a whole pass is saved.  pl.stretch_low_bits lasts until the next
    // exchange (Planner::end_stretch).
    int low = pl.opt.low_bits;
    if (moved > keep && pl.opt.low_bits > 3 && pl.opt.defer_low3) {
        const int m3 = passes_after(true, 3);
        if (m3 <= keep) { moved = m3; low = 3; }
    }
    if (moved > keep) return false;
    for (int i : taken) pl.done[i] = 0;
    pl.first = 0; pl.replay.clear(); pl.replay_at = 0; pl.budget_fixed = false;
    if (low != pl.opt.low_bits) { pl.saved_low_bits = pl.opt.low_bits; pl.opt.low_bits = low; }
    return true;
}

// physical bit -> tile-local bit (or -1)
inline int tile_local_of(const PlanPass& P, int phys) {
    if (!P.tl.empty()) {
        for (size_t p = 0; p < P.tl.size(); ++p) if (P.tl[p] == phys) return (int)p;
        return -1;
    }
    if (phys < P.L) return phys;
    for (size_t j = 0; j < P.hb.size(); ++j) if (P.hb[j] == phys) return P.L + (int)j;
    return -1;
}

// TMA kernel: choose the order of the tile's segments in shared memory.  A tile is moved by one 5-d TMA box:
// dim0 = bits 0-2 (one 128-byte row), the other dims are the segment "bits 3..L-1" and the <= 3 runs of high
// bits (a run longer than 8 bits is cut: a box dimension holds <= 256).  Their order is free, and it decides
// which index bits sit at shared-memory address bits 7-9, the ones the 128-byte swizzle xors into the 16-byte
// chunk index.  A quarter-warp is conflict-free when its three lowest lane bits reach three independent chunk
// bits, i.e. for k = 0, 1, 2 one of the tile-local positions {k, k + 3} is neither a register-block bit nor
// pinned to the warp index.  With bits 3 and 4 always at positions 3 and 4, a register block on low qubits
// (say bits 1-4) left two-way conflicts on every LDS/STS.128 (measured: such passes took 9-13 ms against 6).
// All orders are tried (<= 24) and the one with the most conflict-free groups wins.
inline void choose_tile_layout(PlanPass& P) {
    std::vector<std::pair<int, int>> segs;
    if (P.L > 3) segs.push_back({ 3, P.L - 3 });
    for (size_t i = 0; i < P.hb.size(); ++i) {
        const bool cont = i > 0 && P.hb[i] == P.hb[i - 1] + 1 && segs.back().second < 8;
        if (cont) segs.back().second++; else segs.push_back({ P.hb[i], 1 });
    }
    std::vector<int> idx(segs.size()); for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
    std::vector<int> best_idx = idx; int best_score = -1;
    do {
        int pos_bit[6] = { 0, 1, 2, -1, -1, -1 }; int p = 3;
        for (int si : idx) for (int j = 0; j < segs[si].second && p < 6; ++j) pos_bit[p++] = segs[si].first + j;
        int score = 0;
        for (const PlanGroup& G : P.groups) {
            auto busy = [&](int bit) {
                if (bit < 0) return true;
                for (int b : G.rb) if (b == bit) return true;
                for (int b : G.wb) if (b == bit) return true;
                return false;
            };
            int ok = 0;
            for (int k = 0; k < 3; ++k) if (!busy(pos_bit[k]) || !busy(pos_bit[k + 3])) ++ok;
            // block bits not named by the planner are padded from the top positions: they never touch 0-5 unless m is tiny
            score += ok;
        }
        if (score > best_score) { best_score = score; best_idx = idx; }
    } while (std::next_permutation(idx.begin(), idx.end()));
    P.segs.clear(); P.tl.clear();
    for (int b = 0; b < 3 && b < P.m; ++b) P.tl.push_back(b);
    for (int si : best_idx) { P.segs.push_back(segs[si]); for (int j = 0; j < segs[si].second; ++j) P.tl.push_back(segs[si].first + j); }
}

// Encode one pass into the byte blob the kernel reads: DevPass | DevGroup[] | DevSub[]
inline void encode_pass(const PlanPass& P, std::vector<unsigned char>& blob) {
    DevPass hp; memset(&hp, 0, sizeof(hp));
    hp.m = P.m; hp.L = P.L; hp.nhigh = (int)P.hb.size(); hp.ngroups = (int)P.groups.size();
    for (size_t j = 0; j < P.hb.size(); ++j) hp.hb[j] = P.hb[j];
    std::vector<DevGroup> dg; std::vector<DevSub> ds;
    for (const PlanGroup& G : P.groups) {
        DevGroup g; memset(&g, 0, sizeof(g));
        g.sub_off = (int)ds.size(); g.nsub = (int)G.subs.size();
        uint32_t blockmask = 0; int rbl[3] = { -1, -1, -1 };
        for (size_t i = 0; i < G.rb.size(); ++i) { rbl[i] = tile_local_of(P, G.rb[i]); g.rbmask[i] = 1u << rbl[i]; blockmask |= g.rbmask[i]; }
        g.nfree = P.m - (int)G.rb.size();
        // thread-index bit order: the three lowest thread bits must give distinct 16-byte bank
        // groups under the layout slot = i ^ ((i >> 3) & 7): bit k of the slot's low 3 bits is
        // i_k ^ i_{k+3}, so pick a free bit from {k, k+3} for k = 0, 1, 2.
        std::vector<int> order; uint32_t used = blockmask;
        for (int k = 0; k < 3; ++k) {
            int pick = -1;
            if (k < P.m && !((used >> k) & 1)) pick = k;
            else if (k + 3 < P.m && !((used >> (k + 3)) & 1)) pick = k + 3;
            if (pick >= 0) { order.push_back(pick); used |= 1u << pick; }
        }
        for (int b = 0; b < P.m; ++b) if (!((used >> b) & 1)) { order.push_back(b); used |= 1u << b; }
        for (size_t j = 0; j < order.size() && j < 16; ++j) g.fpos[j] = (uint8_t)order[j];
        for (const LGate& x : G.subs) {
            DevSub s; memset(&s, 0, sizeof(s));
            s.slot = x.slot; memcpy(s.m, x.m, sizeof(s.m));
            if (x.ctrl >= 0) {
                int cl = tile_local_of(P, x.ctrl);
                if (cl < 0) s.cmask_out = 1ull << x.ctrl;
                else {
                    int bi = -1; for (int i = 0; i < 3; ++i) if (rbl[i] == cl) bi = i;
                    if (bi >= 0) s.cmask_blk = 1u << bi; else s.cmask_loc = 1u << cl;
                }
            }
            int tl = tile_local_of(P, x.target);
            int bi = -1; if (tl >= 0) for (int i = 0; i < 3; ++i) if (rbl[i] == tl) bi = i;
            if (!x.diag) { s.kind = SUB_MAT; s.b = bi; }
            else {
                s.kind = SUB_DIAG;
                if (bi >= 0) s.b = bi;
                else if (tl >= 0) { s.b = 3; s.tmask_loc = 1u << tl; }
                else { s.b = 4; s.tmask_out = 1ull << x.target; }
            }
            ds.push_back(s);
        }
        dg.push_back(g);
    }
    hp.nsubs = (int)ds.size();
    hp.total_bytes = (int)(sizeof(DevPass) + dg.size() * sizeof(DevGroup) + ds.size() * sizeof(DevSub));
    size_t off = blob.size();
    blob.resize(off + (size_t)hp.total_bytes);
    memcpy(&blob[off], &hp, sizeof(hp)); off += sizeof(hp);
    if (!dg.empty()) memcpy(&blob[off], dg.data(), dg.size() * sizeof(DevGroup));
    off += dg.size() * sizeof(DevGroup);
    if (!ds.empty()) memcpy(&blob[off], ds.data(), ds.size() * sizeof(DevSub));
}

}  // namespace qm
