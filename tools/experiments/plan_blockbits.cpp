// tools/experiments/plan_blockbits.cpp — PLANNING STUDY, host only (never linked into the engine): how many gates does one
// fused pass absorb when a register-block group holds 4 (today) or 5 qubits?  Runs the engine's own planner (planner.hpp, the
// compiled kernels' cost model) over a gate list with PlanOptions::block_bits overridden and reports, per pass, gates / groups /
// planner cost.  Driven by tools/experiments/plan_blockbits.py.  Nothing here can run a pass: a 5-bit block has no kernel yet.
#include <stdint.h>
#include <vector>

#include "../../include/qm_b200.h"
#include "../../quromorphic.jl_b200/csrc/fused_tma.cuh"
#include "../../quromorphic.jl_b200/csrc/lifting.hpp"
#include "../../quromorphic.jl_b200/csrc/planner.hpp"
#include "../../quromorphic.jl_b200/csrc/encode_v4.hpp"

using namespace qm;

// out_rows: per pass {gates, groups, cost * 16, records at full cost}; returns the number of passes, or -1
extern "C" int32_t plan_blockbits(int32_t n_qubits, int32_t low_bits, double max_cost, int32_t block_bits, double group_cost,
                                  const qm_gate* gates, int32_t ngates, int32_t* out_rows, int32_t cap) {
    Planner pl;
    pl.opt.n_local = n_qubits; pl.opt.n_global = 0;
    pl.opt.tile_bits = 12; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost;
    v4_planner_options(pl.opt, max_cost, 1);
    pl.opt.block_bits = block_bits;
    pl.opt.max_subs = 4096;                       // the record budget of today's program layout is not the question here
    if (group_cost > 0) pl.opt.group_cost = group_cost;
    if (lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates)) return -1;
    prepare_for_lifting(pl.gates);
    pl.done.assign(pl.gates.size(), 0);
    int np = 0;
    PlanPass P;
    while (pl.next_pass(P)) {
        if (np < cap) {
            out_rows[4 * np] = P.ngates; out_rows[4 * np + 1] = (int)P.groups.size(); out_rows[4 * np + 2] = (int)(P.cost * 16.0);
            int full = 0;
            for (const PlanGroup& G : P.groups) for (const LGate& x : G.subs) full += gate_cost(x, 1) >= 2.0;
            out_rows[4 * np + 3] = full;
        }
        ++np;
    }
    for (size_t i = 0; i < pl.gates.size(); ++i) if (!pl.done[i]) return -2;
    return np;
}
