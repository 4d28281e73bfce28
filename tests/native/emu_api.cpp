// emu_api.cpp — TEST INFRASTRUCTURE, host only, built into tests/native/libqm_emu.so (NOT into libqm_b200.so):
// plans a gate list exactly as the engine does for its TMA kernel, encodes every pass with the engine's own encoder
// (csrc/encode_v4.hpp) and walks the encoded programs over a host vector with the scalar restatement of the
// kernel's interpreter in emulate.hpp.  Lets the CPU suite check planner + encoder against the oracle without a GPU.
#include <stdarg.h>
#include <stdio.h>
#include <vector>

#include "../../include/qm_b200.h"
#include "emulate.hpp"
#include "ptx_interp.hpp"
#include "../../quromorphic.jl_b200/csrc/dist_bits.hpp"

using namespace qm;

// 0: walk the encoded records (what the interpreter kernel does); 1: generate the straight-line PTX text of the compiled pass
// kernels (csrc/jit_codegen.hpp) and execute THAT with the PTX-subset interpreter of ptx_interp.hpp
static int g_defer = 1, g_last_deferred = 0;
extern "C" int32_t qm_emu_set_defer(int32_t on) { g_defer = on ? 1 : 0; return QM_OK; }
extern "C" int32_t qm_emu_last_deferred(void) { return g_last_deferred; }
static int g_jit_mode = 0;
// in mode 1 the planner also prices gates as the engine does when the compiled kernels will run them (planner.hpp cost_model 1)
extern "C" int32_t qm_emu_set_jit(int32_t on) { g_jit_mode = on ? 1 : 0; return QM_OK; }
static void apply_phase(EmuAmp* a, uint64_t n, const double g[2]) {
    if (g[0] == 1.0 && g[1] == 0.0) return;
    double pr = g[0], pi = g[1]; const double h = hypot(pr, pi); if (h > 0) { pr /= h; pi /= h; }
    for (uint64_t i = 0; i < n; ++i) { const double x = a[i].x, y = a[i].y; a[i].x = x * pr - y * pi; a[i].y = x * pi + y * pr; }
}
static bool run_program(const V4Program& prog, const PlanPass& P, EmuAmp* amp, int nl, uint64_t rank_or, std::string& why) {
    why.clear();
    if (!g_jit_mode) { if (!emulate_v4(prog, P, amp, nl, rank_or, nullptr)) { why = "malformed program (opcode / non-uniform control predicate)"; return false; } return true; }
    return emulate_v4_jit(prog, P, amp, nl, rank_or, nullptr, &why);
}

static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
extern "C" int32_t qm_emu_last_error(char* buf, int32_t cap) {
    if (!buf || cap <= 0) return QM_ERR_ARG;
    snprintf(buf, (size_t)cap, "%s", g_err);
    return QM_OK;
}

// TEST INFRASTRUCTURE (host only, see emulate.hpp): plan `gates` exactly as the engine would for the TMA
// kernel, encode every pass and walk the encoded programs over a host state vector.
extern "C" int32_t qm_plan_emulate(int32_t n_qubits, int32_t tile_bits, int32_t low_bits, double max_cost,
                                   const qm_gate* gates, int32_t ngates, double* state, int32_t* n_passes) {
    if (!gates || !state || n_qubits < 11 || n_qubits > 24 || (tile_bits != 11 && tile_bits != 12) || tile_bits > n_qubits ||
        low_bits < 3 || low_bits > 8)
        return set_err(QM_ERR_ARG, "bad argument");
    Planner pl;
    pl.opt.n_local = n_qubits; pl.opt.n_global = 0;
    pl.opt.tile_bits = tile_bits; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost < 9.0 ? 9.0 : max_cost;
    v4_planner_options(pl.opt, max_cost, g_jit_mode);
    int rc = lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates);
    if (rc) return set_err(rc, "bad gate");
    prepare_for_lifting(pl.gates);
    for (const LGate& x : pl.gates) if (!x.liftable) return set_err(QM_ERR_STATE, "a gate has no lifting form");
    pl.done.assign(pl.gates.size(), 0);
    std::vector<V4Program> prog(1);
    PlanPass P; int np = 0;
    double gph[2] = { 1.0, 0.0 };        // the global phase the D1 form of the diagonal gates leaves pending (encode_v4)
    while (pl.next_pass(P)) {
        choose_tile_layout(P);
        if (!encode_v4(P, prog[0], gph, g_jit_mode != 0)) return set_err(QM_ERR_STATE, "pass %d cannot be encoded for the TMA kernel", np);
        std::string why;
        if (!run_program(prog[0], P, reinterpret_cast<EmuAmp*>(state), n_qubits, 0, why))
            return set_err(QM_ERR_STATE, "pass %d: %s", np, why.c_str());
        ++np;
    }
    for (size_t i = 0; i < pl.gates.size(); ++i) if (!pl.done[i]) return set_err(QM_ERR_STATE, "planner left a gate behind");
    apply_phase(reinterpret_cast<EmuAmp*>(state), 1ull << n_qubits, gph);       // what the engine's materialize_phase does
    if (n_passes) *n_passes = np;
    return QM_OK;
}

// TEST INFRASTRUCTURE (host only): qm_plan_emulate for a sharded register, every rank emulated in this process.
extern "C" int32_t qm_plan_emulate_sharded(int32_t n_qubits, int32_t global_bits, int32_t tile_bits, int32_t low_bits, double max_cost,
                                           const qm_gate* gates, int32_t ngates, double* state, int32_t* n_passes, int32_t* n_remaps) {
    const int g = global_bits, nl = n_qubits - global_bits;
    if (!gates || !state || g < 1 || g > 4 || nl < 11 || nl > 22 || 2 * g > nl || (tile_bits != 11 && tile_bits != 12) || tile_bits > nl ||
        low_bits < 3 || low_bits > 8)
        return set_err(QM_ERR_ARG, "bad argument");
    Planner pl;
    pl.opt.n_local = nl; pl.opt.n_global = g;
    pl.opt.tile_bits = tile_bits; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost < 9.0 ? 9.0 : max_cost;
    v4_planner_options(pl.opt, max_cost, g_jit_mode);
    int rc = lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates);
    if (rc) return set_err(rc, "bad gate");
    prepare_for_lifting(pl.gates);
    for (const LGate& x : pl.gates) if (!x.liftable) return set_err(QM_ERR_STATE, "a gate has no lifting form");
    pl.done.assign(pl.gates.size(), 0);
    EmuAmp* amp = reinterpret_cast<EmuAmp*>(state);
    const int W = 1 << g;
    const uint64_t shard = 1ull << nl, chunk = 1ull << (nl - g);
    auto exchange = [&]() {                          // what k_remap_swap does, on the host
        for (int r = 0; r < W; ++r)
            for (int j = r + 1; j < W; ++j)
                for (uint64_t i = 0; i < chunk; ++i)
                    std::swap(amp[(uint64_t)r * shard + (uint64_t)j * chunk + i], amp[(uint64_t)j * shard + (uint64_t)r * chunk + i]);
    };
    std::vector<V4Program> prog(1);
    PlanPass P; int np = 0, nr = 0; bool swapped = false;
    double gph[2] = { 1.0, 0.0 };
    // like the engine, ONE pass behind the planner: the last pass planned before an exchange may be postponed to after it
    // (planner.hpp try_defer_last_pass), which un-does its gates
    struct Held { bool valid = false; PlanPass P; std::vector<int> taken; };
    Held held;
    auto run_held = [&]() -> int {
        if (!held.valid) return QM_OK;
        held.valid = false;
        if (!encode_v4(held.P, prog[0], gph, g_jit_mode != 0)) return set_err(QM_ERR_STATE, "pass %d cannot be encoded for the TMA kernel", np);
        for (int r = 0; r < W; ++r) {
            std::string why;
            if (!run_program(prog[0], held.P, amp + (uint64_t)r * shard, nl, (uint64_t)r << nl, why))
                return set_err(QM_ERR_STATE, "pass %d, rank %d: %s", np, r, why.c_str());
        }
        ++np;
        return QM_OK;
    };
    int ndefer = 0;
    for (int guard = 0; guard < 4 * ngates + 8; ++guard) {
        while (pl.next_pass(P)) {
            choose_tile_layout(P);
            int rc2 = run_held(); if (rc2) return rc2;
            held.valid = true; held.P = P; held.taken = pl.last_taken;
        }
        bool left = false;
        for (size_t i = 0; i < pl.gates.size(); ++i) if (!pl.done[i]) { left = true; break; }
        if (!left) { int rc2 = run_held(); if (rc2) return rc2; break; }
        if (!pl.pending_global_target()) return set_err(QM_ERR_STATE, "planner stalled");
        if (held.valid && g_defer && try_defer_last_pass(pl, held.taken, nl, g)) { held.valid = false; ++ndefer; }
        { int rc2 = run_held(); if (rc2) return rc2; }
        exchange(); swapped = !swapped; ++nr;
        for (size_t i = 0; i < pl.gates.size(); ++i) {
            if (pl.done[i]) continue;
            pl.gates[i].target = dist_swap_bit(pl.gates[i].target, nl, g);
            if (pl.gates[i].ctrl >= 0) pl.gates[i].ctrl = dist_swap_bit(pl.gates[i].ctrl, nl, g);
        }
    }
    g_last_deferred = ndefer;
    for (size_t i = 0; i < pl.gates.size(); ++i) if (!pl.done[i]) return set_err(QM_ERR_STATE, "planner left a gate behind");
    if (swapped) exchange();                         // back to the canonical layout (dist_canonicalize)
    apply_phase(amp, 1ull << n_qubits, gph);
    if (n_passes) *n_passes = np;
    if (n_remaps) *n_remaps = nr;
    return QM_OK;
}


// TEST INFRASTRUCTURE (host only): the pass / exchange / postponement counts of a sharded circuit at ANY size — the planning
// loop of qm_plan_emulate_sharded without walking the programs (a 36-qubit register does not fit a host vector).
extern "C" int32_t qm_emu_plan_count_sharded(int32_t n_qubits, int32_t global_bits, int32_t low_bits, double max_cost, int32_t cost_model,
                                             const qm_gate* gates, int32_t ngates, int32_t* n_passes, int32_t* n_remaps, int32_t* n_deferred,
                                             int32_t* pass_gates, int32_t cap) {
    const int g = global_bits, nl = n_qubits - global_bits;
    if (!gates || g < 0 || nl < 12) return set_err(QM_ERR_ARG, "bad argument");
    Planner pl;
    pl.opt.n_local = nl; pl.opt.n_global = g;
    pl.opt.tile_bits = 12; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost < 9.0 ? 9.0 : max_cost;
    v4_planner_options(pl.opt, max_cost, cost_model);
    int rc = lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates);
    if (rc) return set_err(rc, "bad gate");
    prepare_for_lifting(pl.gates);
    pl.done.assign(pl.gates.size(), 0);
    int np = 0, nr = 0, nd = 0;
    bool held = false; std::vector<int> taken; PlanPass P;
    auto emit = [&](int ng_) { if (pass_gates && np < cap) pass_gates[np] = ng_; ++np; };
    for (int guard = 0; guard < 4 * ngates + 8; ++guard) {
        while (pl.next_pass(P)) { if (held) emit((int)taken.size()); held = true; taken = pl.last_taken; }
        bool left = false;
        for (size_t i = 0; i < pl.gates.size(); ++i) if (!pl.done[i]) { left = true; break; }
        if (!left) { if (held) emit((int)taken.size()); break; }
        if (g == 0 || !pl.pending_global_target()) return set_err(QM_ERR_STATE, "planner stalled");
        if (held && g_defer && try_defer_last_pass(pl, taken, nl, g)) { held = false; ++nd; }
        if (held) { emit((int)taken.size()); held = false; }
        if (pass_gates && np < cap) pass_gates[np] = -1;      // marks the exchange
        ++np; ++nr;
        for (size_t i = 0; i < pl.gates.size(); ++i) {
            if (pl.done[i]) continue;
            pl.gates[i].target = dist_swap_bit(pl.gates[i].target, nl, g);
            if (pl.gates[i].ctrl >= 0) pl.gates[i].ctrl = dist_swap_bit(pl.gates[i].ctrl, nl, g);
        }
    }
    if (n_passes) *n_passes = np - nr;
    if (n_remaps) *n_remaps = nr;
    if (n_deferred) *n_deferred = nd;
    return QM_OK;
}
