// jit_codegen.hpp — host only: V4Program -> straight-line PTX for the register-block groups of ONE pass program.
//
// Why: the fused TMA pass interprets its program (v4_dispatch.inc): per gate a record fetch from shared memory, an indexed
// jump through a constant-bank table (LDC -> BRX, ~110 cycles that two warps per scheduler cannot hide — ncu r2m: 12 % of
// all warp samples sit in the five instructions of the dispatcher) and coefficient loads.  The FP64 work of a gate is only
// 48 DFMA per thread.  A pass program whose STRUCTURE repeats (a reservoir step, every step of a benchmark circuit) is
// instead compiled once: the same kernel (fused_tma_kernel.cuh, JIT = true) with the interpreter replaced by the text this
// file emits — the 16 loads, every gate as plain in-place FMAs on named registers, the 16 stores.  No dispatch, no record
// fetch; coefficients are kernel-parameter constants (ptxas turns them into uniform-register operands: one LDCU.64 per 16
// DFMA), x!/cnot! inside the register block are register renames, and ptxas schedules across gates and across the
// load/store phases.  What is baked into the code is only the opcode sequence of the program (gate family, sign variant,
// block positions, controlled / plain entry, SLOT prefixes); everything else — coefficients, control masks, store offsets,
// index tables — stays data of the launch, so one compiled kernel serves every pass with the same opcode sequence.
//
// The emitted text refers to the operands of the marker block by the register names found in the template
// ("// QMJIT_BEGIN r0 .. r9": rp (unused), ctx, slot_s, tile_s, b0, s0..s3, gi) and to the program by the name of the kernel
// parameter.  tests/native interprets the same text on the host (a PTX-subset interpreter) against the oracle.
#pragma once
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "fused_tma.cuh"
#include "encode_v4.hpp"

namespace qm {

struct JitOperands {
    std::string ctx, slot_s, tile_s, b0, s[4], gi;   // register names (or symbolic names in the host interpreter)
    std::string prog;                                 // name of the V4Program kernel parameter
};

// what the code depends on: the opcode sequence, group by group
inline void jit_structure_bytes(const V4Program& p, int m, std::vector<uint32_t>& out) {
    out.clear();
    out.push_back(0x4a495431u); out.push_back((uint32_t)m); out.push_back((uint32_t)p.ngroups); out.push_back(p.flags);
    for (int g = 0; g < p.ngroups; ++g) {
        const V4Rec* R = reinterpret_cast<const V4Rec*>(reinterpret_cast<const unsigned char*>(&p) + p.groups[g].rec_off);
        for (int guard = 0; guard < kV4MaxRecs + 1; ++guard, ++R) { out.push_back(R->op); if (R->op == (uint32_t)V4_OP_END) break; }
    }
}

inline void jit_hash(const std::vector<uint32_t>& v, uint64_t key[2]) {
    uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x9e3779b97f4a7c15ull;
    for (uint32_t w : v) {
        h1 = (h1 ^ w) * 0x100000001b3ull;
        h2 ^= (uint64_t)w + 0x9e3779b97f4a7c15ull + (h2 << 6) + (h2 >> 2); h2 *= 0xff51afd7ed558ccdull; h2 ^= h2 >> 33;
    }
    key[0] = h1; key[1] = h2 ^ ((uint64_t)v.size() << 48);
}

class JitEmitter {
public:
    JitEmitter(const V4Program& p, const JitOperands& o) : P(p), O(o) {}

    // returns false on a program the emitter does not understand (the caller keeps the interpreter)
    bool emit(std::string& out) {
        body.clear(); ncoef = 0; nlabel = 0;
        const int ng = P.ngroups;
        for (int g = 0; g < ng; ++g) {
            if (ng > 1) line("QJ_G%d:", g);
            if (!group(g)) return false;
            if (ng > 1) line("bra.uni QJ_DONE;");
        }
        std::string head = "{\n";
        char buf[256];
        snprintf(buf, sizeof(buf), "\t.reg .f64 x<16>, y<16>, ng, tq<2>, sw, c<%d>;\n", ncoef + 1); head += buf;
        head += "\t.reg .b32 t, ad, o<4>, e<4>, hcm, ng0, ng1, ngs, wlo, whi;\n\t.reg .pred pact, phi, pg;\n\t.reg .b64 sa, sb;\n";
        if (ng > 1) {       // which group: a warp-uniform branch per group (the C++ loop around the block passes gi)
            for (int g = 1; g < ng; ++g) {
                snprintf(buf, sizeof(buf), "\tsetp.eq.u32 pg, %s, %d;\n\t@pg bra.uni QJ_G%d;\n", O.gi.c_str(), g, g); head += buf;
            }
        }
        out = head + body;
        if (ng > 1) out += "QJ_DONE:\n";
        out += "}\n";
        return true;
    }

private:
    const V4Program& P;
    const JitOperands& O;
    std::string body;
    int ncoef = 0, nlabel = 0;
    int reg_of[16];

    void line(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
        const size_t n = strlen(buf);
        if (n && buf[n - 1] == ':') { body += buf; body += "\n"; } else { body += "\t"; body += buf; body += "\n"; }
    }
    std::string X(int k) const { return "x" + std::to_string(reg_of[k]); }
    std::string Y(int k) const { return "y" + std::to_string(reg_of[k]); }
    static bool ctrl_ok(int k, int CB) { return CB < 0 || ((k >> CB) & 1); }

    // three shears of one real pair; coefficient names may carry a leading '-' (negated copy requested by the caller)
    struct Chain { std::string x, y, t1, s1, t2; int sg; };
    void lift_all(const std::vector<Chain>& ch) {
        for (const Chain& c : ch) line("fma.rn.f64 %s, %s, %s, %s;", c.x.c_str(), c.t1.c_str(), c.y.c_str(), c.x.c_str());
        for (const Chain& c : ch) {
            if (c.sg & 1) { line("neg.f64 ng, %s;", c.y.c_str()); line("fma.rn.f64 %s, %s, %s, ng;", c.y.c_str(), c.s1.c_str(), c.x.c_str()); }
            else line("fma.rn.f64 %s, %s, %s, %s;", c.y.c_str(), c.s1.c_str(), c.x.c_str(), c.y.c_str());
        }
        for (const Chain& c : ch) {
            if (c.sg & 2) { line("neg.f64 ng, %s;", c.x.c_str()); line("fma.rn.f64 %s, %s, %s, ng;", c.x.c_str(), c.t2.c_str(), c.y.c_str()); }
            else line("fma.rn.f64 %s, %s, %s, %s;", c.x.c_str(), c.t2.c_str(), c.y.c_str(), c.x.c_str());
        }
    }

    // canonical signs (V4Program::flags bit 0): the record's aux bit 0 / 1 = negate the values that took coefficient set 0 / 1.
    // ng0 / ng1 become 0 or 0x80000000 and are xor-ed into the high word of every value of the set: an integer operation per
    // value beside the FP64 pipe, data of the launch — so the code does not depend on the signs of the angles.
    void neg_masks(size_t rec_off) {
        line("ld.param.u32 t, [%s+%zu];", O.prog.c_str(), rec_off + 12);
        line("shl.b32 ng0, t, 31;");
        line("and.b32 ng1, t, 2;");
        line("shl.b32 ng1, ng1, 30;");
    }
    void neg_value(const std::string& v, const char* mask) {
        line("mov.b64 {wlo, whi}, %s;", v.c_str());
        line("xor.b32 whi, whi, %s;", mask);
        line("mov.b64 %s, {wlo, whi};", v.c_str());
    }

    bool group(int g) {
        for (int k = 0; k < 16; ++k) reg_of[k] = k;
        const size_t goff = offsetof(V4Program, groups) + sizeof(V4Group) * (size_t)g;
        (void)goff;
        // addresses: off_j (j < 4) = b0 xor the swizzled offsets of the low two block bits set in j
        line("mov.b32 o0, %s;", O.b0.c_str());
        line("xor.b32 o1, o0, %s;", O.s[0].c_str());
        line("xor.b32 o2, o0, %s;", O.s[1].c_str());
        line("xor.b32 o3, o1, %s;", O.s[1].c_str());
        for (int k = 0; k < 16; ++k) {
            std::string src = "o" + std::to_string(k & 3);
            if (k & 4) { line("xor.b32 ad, %s, %s;", src.c_str(), O.s[2].c_str()); src = "ad"; }
            if (k & 8) { line("xor.b32 ad, %s, %s;", src.c_str(), O.s[3].c_str()); src = "ad"; }
            line("add.u32 ad, %s, %s;", src.c_str(), O.tile_s.c_str());
            line("ld.shared.v2.f64 {x%d, y%d}, [ad];", k, k);
        }
        const size_t rec0 = P.groups[g].rec_off;
        const V4Rec* R = reinterpret_cast<const V4Rec*>(reinterpret_cast<const unsigned char*>(&P) + rec0);
        size_t off = rec0;
        bool from_slot = false; int slot_c0 = 0;
        for (int guard = 0; guard < kV4MaxRecs + 1; ++guard, ++R, off += sizeof(V4Rec)) {
            if (off + sizeof(V4Rec) > sizeof(V4Program)) return false;
            uint32_t op = R->op;
            if (op == (uint32_t)V4_OP_END) { end_record(off); return true; }
            if (op == (uint32_t)V4_OP_SLOT) {
                // the six coefficients of the NEXT record come from the per-state table of the tile's register
                slot_c0 = ncoef; ncoef += 6; from_slot = true;
                line("ld.param.u32 t, [%s+%zu];", O.prog.c_str(), off + 12);
                line("ld.shared.u64 sa, [%s];", O.slot_s.c_str());
                line("mul.wide.u32 sb, t, 64;");
                line("add.u64 sa, sa, sb;");
                for (int q = 0; q < 3; ++q) line("ld.global.nc.v2.f64 {c%d, c%d}, [sa+%d];", slot_c0 + 2 * q, slot_c0 + 2 * q + 1, 16 * q);
                continue;
            }
            bool controlled = false;
            if (op >= (uint32_t)V4_NUM_GATE_OPS) { op -= V4_NUM_GATE_OPS; controlled = true; }
            int fam, var, B, CB;
            if (!v4_decode_op(op, fam, var, B, CB)) return false;
            // coefficient registers of this record: c[base + q], q = 0..5 (set 0: 0..2, set 1: 3..5)
            int cb;
            if (from_slot) { cb = slot_c0; from_slot = false; }
            else {
                cb = ncoef; ncoef += 6;
                const bool set1 = fam == 2 || fam == 4;
                if (fam != 3) for (int q = 0; q < 3; ++q) {
                    line("ld.param.f64 c%d, [%s+%zu];", cb + q, O.prog.c_str(), off + 16 + 16 * q);
                    if (set1) line("ld.param.f64 c%d, [%s+%zu];", cb + 3 + q, O.prog.c_str(), off + 24 + 16 * q);
                }
            }
            auto C = [&](int q) { return "c" + std::to_string(cb + q); };
            int skip = -1;
            if (controlled) {
                skip = nlabel++;
                line("ld.param.u32 hcm, [%s+%zu];", O.prog.c_str(), off + 8);
                line("and.b32 t, %s, hcm;", O.ctx.c_str());
                line("setp.ne.u32 pact, t, hcm;");
                line("@pact bra.uni QJ_S%d;", skip);
            }
            const bool canon = (P.flags & 1u) != 0;
            if (canon && ((fam == 0 && var > 1) || (fam == 1 && var != 0) || (fam == 2 && var != 0) || (fam == 4 && var != 0))) return false;
            std::vector<Chain> ch;
            if (fam == 0) {
                for (int k = 0; k < 16; ++k) {
                    if (!ctrl_ok(k, CB) || ((k >> B) & 1)) continue;
                    const int k1 = k | (1 << B);
                    ch.push_back({ X(k), X(k1), C(0), C(1), C(2), var });
                    ch.push_back({ Y(k), Y(k1), C(0), C(1), C(2), var });
                }
                lift_all(ch);
                if (canon && (CB >= 0 || controlled)) {
                    neg_masks(off);
                    for (int k = 0; k < 16; ++k) if (ctrl_ok(k, CB)) { neg_value(X(k), "ng0"); neg_value(Y(k), "ng0"); }
                }
            } else if (fam == 1) {
                const int sg = var ? 3 : 0;
                line("neg.f64 tq0, %s;", C(0).c_str());      // mirrored pair: coefficients negated
                line("neg.f64 tq1, %s;", C(1).c_str());
                line("neg.f64 sw, %s;", C(2).c_str());
                for (int k = 0; k < 16; ++k) {
                    if (!ctrl_ok(k, CB) || ((k >> B) & 1)) continue;
                    const int k1 = k | (1 << B);
                    ch.push_back({ X(k), Y(k1), C(0), C(1), C(2), sg });
                    ch.push_back({ Y(k), X(k1), "tq0", "tq1", "sw", sg });
                }
                lift_all(ch);
                if (canon && (CB >= 0 || controlled)) {
                    neg_masks(off);
                    for (int k = 0; k < 16; ++k) if (ctrl_ok(k, CB)) { neg_value(X(k), "ng0"); neg_value(Y(k), "ng0"); }
                }
            } else if (fam == 2) {
                static const int sgs[4][2] = { { 0, 0 }, { 0, 3 }, { 3, 0 }, { 3, 3 } };
                for (int k = 0; k < 16; ++k) {
                    if (!ctrl_ok(k, CB)) continue;
                    if ((k >> B) & 1) ch.push_back({ X(k), Y(k), C(3), C(4), C(5), sgs[var][1] });
                    else ch.push_back({ X(k), Y(k), C(0), C(1), C(2), sgs[var][0] });
                }
                lift_all(ch);
                if (canon) {
                    neg_masks(off);
                    for (int k = 0; k < 16; ++k) {
                        if (!ctrl_ok(k, CB)) continue;
                        const char* mk = ((k >> B) & 1) ? "ng1" : "ng0";
                        neg_value(X(k), mk); neg_value(Y(k), mk);
                    }
                }
            } else if (fam == 3) {
                for (int k = 0; k < 16; ++k) {
                    if (!ctrl_ok(k, CB) || ((k >> B) & 1)) continue;
                    const int k1 = k | (1 << B);
                    if (!controlled) { const int t = reg_of[k]; reg_of[k] = reg_of[k1]; reg_of[k1] = t; }     // a rename
                    else {      // under a predicate the exchange has to happen
                        line("mov.f64 sw, %s;", X(k).c_str()); line("mov.f64 %s, %s;", X(k).c_str(), X(k1).c_str()); line("mov.f64 %s, sw;", X(k1).c_str());
                        line("mov.f64 sw, %s;", Y(k).c_str()); line("mov.f64 %s, %s;", Y(k).c_str(), Y(k1).c_str()); line("mov.f64 %s, sw;", Y(k1).c_str());
                    }
                }
            } else {
                // THR: the target bit is a constant of the thread, which picks coefficient set 1 when its context has it
                line("ld.param.u32 t, [%s+%zu];", O.prog.c_str(), off + 4);
                line("and.b32 t, %s, t;", O.ctx.c_str());
                line("setp.ne.u32 phi, t, 0;");
                const int nsel = var == 2 ? 2 : 3;
                for (int q = 0; q < nsel; ++q) line("selp.f64 %s, %s, %s, phi;", C(q).c_str(), C(3 + q).c_str(), C(q).c_str());
                if (var == 2) {
                    // (x, y) <- (cr x - ci y, ci x + cr y), cr = c0, ci = c1
                    line("neg.f64 sw, %s;", C(1).c_str());
                    int h = 0;
                    for (int k = 0; k < 16; ++k) {
                        if (!ctrl_ok(k, CB)) continue;
                        const char* tq = (h++ & 1) ? "tq1" : "tq0";
                        line("mul.f64 %s, %s, %s;", tq, C(1).c_str(), X(k).c_str());
                        line("mul.f64 %s, %s, %s;", X(k).c_str(), C(0).c_str(), X(k).c_str());
                        line("fma.rn.f64 %s, sw, %s, %s;", X(k).c_str(), Y(k).c_str(), X(k).c_str());
                        line("fma.rn.f64 %s, %s, %s, %s;", Y(k).c_str(), C(0).c_str(), Y(k).c_str(), tq);
                    }
                } else {
                    for (int k = 0; k < 16; ++k) if (ctrl_ok(k, CB)) ch.push_back({ X(k), Y(k), C(0), C(1), C(2), var == 1 ? 3 : 0 });
                    lift_all(ch);
                    if (canon) {
                        neg_masks(off);
                        line("selp.b32 ngs, ng1, ng0, phi;");
                        for (int k = 0; k < 16; ++k) if (ctrl_ok(k, CB)) { neg_value(X(k), "ngs"); neg_value(Y(k), "ngs"); }
                    }
                }
            }
            if (controlled) line("QJ_S%d:", skip);
        }
        return false;       // no END record
    }

    // stores: the END record carries the store offsets of the four block bits and an xor for the thread's base (they
    // equal the load offsets unless a tail of x! / cnot! gates was folded into them, encode_v4); amplitude k sits in
    // register reg_of[k]
    void end_record(size_t off) {
        line("ld.param.u32 e0, [%s+%zu];", O.prog.c_str(), off + 4);
        line("ld.param.u32 e1, [%s+%zu];", O.prog.c_str(), off + 8);
        line("ld.param.u32 e2, [%s+%zu];", O.prog.c_str(), off + 12);
        line("ld.param.u32 e3, [%s+%zu];", O.prog.c_str(), off + 16);
        line("ld.param.u32 t, [%s+%zu];", O.prog.c_str(), off + 20);
        line("xor.b32 o0, %s, t;", O.b0.c_str());
        line("xor.b32 o1, o0, e0;");
        line("xor.b32 o2, o0, e1;");
        line("xor.b32 o3, o1, e1;");
        for (int k = 0; k < 16; ++k) {
            std::string src = "o" + std::to_string(k & 3);
            if (k & 4) { line("xor.b32 ad, %s, e2;", src.c_str()); src = "ad"; }
            if (k & 8) { line("xor.b32 ad, %s, e3;", src.c_str()); src = "ad"; }
            line("add.u32 ad, %s, %s;", src.c_str(), O.tile_s.c_str());
            line("st.shared.v2.f64 [ad], {%s, %s};", X(k).c_str(), Y(k).c_str());
        }
    }
};

// Splices the generated block into the template text (between "// QMJIT_BEGIN ..." and "// QMJIT_END").
// Returns false when the template has no marker or the program cannot be emitted.
inline bool jit_build_ptx(const std::string& tmpl, const V4Program& prog, std::string& out) {
    const size_t b = tmpl.find("// QMJIT_BEGIN");
    if (b == std::string::npos) return false;
    const size_t eol = tmpl.find('\n', b);
    const size_t e = tmpl.find("// QMJIT_END", b);
    if (eol == std::string::npos || e == std::string::npos) return false;
    // operand register names
    std::vector<std::string> ops;
    {
        std::string ln = tmpl.substr(b + strlen("// QMJIT_BEGIN"), eol - b - strlen("// QMJIT_BEGIN"));
        size_t i = 0;
        while (i < ln.size()) {
            while (i < ln.size() && (ln[i] == ' ' || ln[i] == '\t' || ln[i] == '\r')) ++i;
            size_t j = i; while (j < ln.size() && ln[j] != ' ' && ln[j] != '\t' && ln[j] != '\r') ++j;
            if (j > i) ops.push_back(ln.substr(i, j - i));
            i = j;
        }
    }
    if (ops.size() != 10) return false;
    // the V4Program parameter: "<entry>_param_1"
    const size_t en = tmpl.find(".entry ");
    if (en == std::string::npos) return false;
    const size_t en2 = tmpl.find('(', en);
    std::string entry = tmpl.substr(en + 7, en2 - en - 7);
    while (!entry.empty() && (entry.back() == ' ' || entry.back() == '\n')) entry.pop_back();
    JitOperands O;
    O.ctx = ops[1]; O.slot_s = ops[2]; O.tile_s = ops[3]; O.b0 = ops[4];
    for (int i = 0; i < 4; ++i) O.s[i] = ops[5 + i];
    O.gi = ops[9];
    O.prog = entry + "_param_1";
    std::string code;
    JitEmitter em(prog, O);
    if (!em.emit(code)) return false;
    out.reserve(tmpl.size() + code.size() + 16);
    out.assign(tmpl, 0, eol + 1);
    out += code;
    out.append(tmpl, e, std::string::npos);
    return true;
}

inline std::string jit_entry_name(const std::string& tmpl) {
    const size_t en = tmpl.find(".entry ");
    if (en == std::string::npos) return std::string();
    const size_t en2 = tmpl.find('(', en);
    std::string entry = tmpl.substr(en + 7, en2 - en - 7);
    while (!entry.empty() && (entry.back() == ' ' || entry.back() == '\n')) entry.pop_back();
    return entry;
}

}  // namespace qm
