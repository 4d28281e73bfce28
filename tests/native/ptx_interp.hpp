// ptx_interp.hpp — TEST INFRASTRUCTURE, host only: an interpreter for the small PTX subset that jit_codegen.hpp emits, and a
// walk of a pass program through the GENERATED TEXT (thread by thread, tile in the swizzled shared-memory layout, the program
// bytes as "parameter space").  It is to the compiled pass kernels what emulate.hpp is to the interpreter kernel: the CPU
// suite checks the text the GPU's PTX compiler will be handed against the oracle, without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "emulate.hpp"
#include "../../quromorphic.jl_b200/csrc/jit_codegen.hpp"

namespace qm {

struct PtxIns {
    enum Op { MOV32, MOVF, XOR, AND, ADD32, ADD64, MULWIDE, SETP_EQ, SETP_NE, SELP, NEG, MUL, FMA, LDP32, LDPF, LDS_V2, LDS_U64, LDG_V2,
              STS_V2, BRA, LABEL, SHL, UNPACK, PACK } op;
    int pred = -1; bool pneg = false;      // guard predicate register
    int d = -1, d2 = -1, a = -1, b = -1, c = -1;     // register indices (-1: unused)
    bool a_imm = false, b_imm = false; int64_t imm_a = 0, imm_b = 0, off = 0;
    int target = -1;                       // BRA: instruction index
    std::string label;
};

struct PtxProgram {
    std::vector<PtxIns> ins;
    std::map<std::string, int> reg;        // name -> slot in the 64-bit register file
    std::string err;
    int R(const std::string& n) { auto it = reg.find(n); if (it != reg.end()) return it->second; int i = (int)reg.size(); reg[n] = i; return i; }
};

inline std::vector<std::string> ptx_split(const std::string& s) {
    // operands separated by commas; braces and brackets kept as part of the token
    std::vector<std::string> out; std::string cur; int depth = 0;
    for (char ch : s) {
        if (ch == '{' || ch == '[') ++depth;
        if (ch == '}' || ch == ']') --depth;
        if (ch == ',' && depth == 0) { out.push_back(cur); cur.clear(); continue; }
        if (ch == ' ' || ch == '\t') { if (depth == 0) continue; }
        cur += ch;
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}

inline bool ptx_is_imm(const std::string& s) { return !s.empty() && (isdigit((unsigned char)s[0]) || s[0] == '-'); }

inline bool ptx_parse(const std::string& text, PtxProgram& P) {
    std::map<std::string, int> labels;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t e = text.find('\n', pos); if (e == std::string::npos) e = text.size();
        std::string ln = text.substr(pos, e - pos); pos = e + 1;
        size_t a = ln.find_first_not_of(" \t"); if (a == std::string::npos) continue;
        ln = ln.substr(a);
        while (!ln.empty() && (ln.back() == ' ' || ln.back() == '\r' || ln.back() == '\t')) ln.pop_back();
        if (ln.empty() || ln == "{" || ln == "}" || ln[0] == '.' || ln.compare(0, 2, "//") == 0) continue;
        if (ln.back() == ':') { PtxIns I; I.op = PtxIns::LABEL; I.label = ln.substr(0, ln.size() - 1); labels[I.label] = (int)P.ins.size(); P.ins.push_back(I); continue; }
        if (ln.back() != ';') { P.err = "no semicolon: " + ln; return false; }
        ln.pop_back();
        PtxIns I;
        if (ln[0] == '@') {
            size_t sp = ln.find(' ');
            std::string p = ln.substr(1, sp - 1);
            if (p[0] == '!') { I.pneg = true; p = p.substr(1); }
            I.pred = P.R(p); ln = ln.substr(sp + 1);
        }
        size_t sp = ln.find(' ');
        const std::string opc = ln.substr(0, sp);
        std::vector<std::string> o = ptx_split(sp == std::string::npos ? std::string() : ln.substr(sp + 1));
        auto mem = [&](const std::string& m, int& base, int64_t& off) {       // "[name+123]" / "[name]"
            std::string in = m.substr(1, m.size() - 2);
            size_t plus = in.find('+');
            off = 0;
            if (plus != std::string::npos) { off = atoll(in.substr(plus + 1).c_str()); in = in.substr(0, plus); }
            base = P.R(in);
        };
        auto pair = [&](const std::string& m, int& r0, int& r1) {             // "{a, b}"
            std::string in = m.substr(1, m.size() - 2);
            size_t comma = in.find(',');
            std::string s0 = in.substr(0, comma), s1 = in.substr(comma + 1);
            auto trim = [](std::string& s) { while (!s.empty() && s[0] == ' ') s.erase(0, 1); while (!s.empty() && s.back() == ' ') s.pop_back(); };
            trim(s0); trim(s1); r0 = P.R(s0); r1 = P.R(s1);
        };
        auto src = [&](const std::string& s, int& r, bool& is_imm, int64_t& imm) {
            if (ptx_is_imm(s)) { is_imm = true; imm = atoll(s.c_str()); r = -1; } else r = P.R(s);
        };
        bool dummy = false; int64_t dimm = 0;
        if (opc == "mov.b32" || opc == "mov.f64") { I.op = opc == "mov.b32" ? PtxIns::MOV32 : PtxIns::MOVF; I.d = P.R(o[0]); src(o[1], I.a, I.a_imm, I.imm_a); }
        else if (opc == "xor.b32" || opc == "and.b32" || opc == "add.u32" || opc == "add.u64") {
            I.op = opc == "xor.b32" ? PtxIns::XOR : opc == "and.b32" ? PtxIns::AND : opc == "add.u32" ? PtxIns::ADD32 : PtxIns::ADD64;
            I.d = P.R(o[0]); src(o[1], I.a, dummy, dimm); src(o[2], I.b, I.b_imm, I.imm_b);
        }
        else if (opc == "shl.b32") { I.op = PtxIns::SHL; I.d = P.R(o[0]); I.a = P.R(o[1]); src(o[2], I.b, I.b_imm, I.imm_b); }
        else if (opc == "mov.b64" && o[0][0] == '{') { I.op = PtxIns::UNPACK; pair(o[0], I.d, I.d2); I.a = P.R(o[1]); }
        else if (opc == "mov.b64" && o[1][0] == '{') { I.op = PtxIns::PACK; I.d = P.R(o[0]); pair(o[1], I.a, I.b); }
        else if (opc == "mul.wide.u32") { I.op = PtxIns::MULWIDE; I.d = P.R(o[0]); I.a = P.R(o[1]); src(o[2], I.b, I.b_imm, I.imm_b); }
        else if (opc == "setp.eq.u32" || opc == "setp.ne.u32") { I.op = opc == "setp.eq.u32" ? PtxIns::SETP_EQ : PtxIns::SETP_NE; I.d = P.R(o[0]); I.a = P.R(o[1]); src(o[2], I.b, I.b_imm, I.imm_b); }
        else if (opc == "selp.f64" || opc == "selp.b32") { I.op = PtxIns::SELP; I.d = P.R(o[0]); I.a = P.R(o[1]); I.b = P.R(o[2]); I.c = P.R(o[3]); }
        else if (opc == "neg.f64") { I.op = PtxIns::NEG; I.d = P.R(o[0]); I.a = P.R(o[1]); }
        else if (opc == "mul.f64") { I.op = PtxIns::MUL; I.d = P.R(o[0]); I.a = P.R(o[1]); I.b = P.R(o[2]); }
        else if (opc == "fma.rn.f64") { I.op = PtxIns::FMA; I.d = P.R(o[0]); I.a = P.R(o[1]); I.b = P.R(o[2]); I.c = P.R(o[3]); }
        else if (opc == "ld.param.u32" || opc == "ld.param.f64") { I.op = opc == "ld.param.u32" ? PtxIns::LDP32 : PtxIns::LDPF; I.d = P.R(o[0]); mem(o[1], I.a, I.off); }
        else if (opc == "ld.shared.v2.f64") { I.op = PtxIns::LDS_V2; pair(o[0], I.d, I.d2); mem(o[1], I.a, I.off); }
        else if (opc == "ld.shared.u64") { I.op = PtxIns::LDS_U64; I.d = P.R(o[0]); mem(o[1], I.a, I.off); }
        else if (opc == "ld.global.nc.v2.f64") { I.op = PtxIns::LDG_V2; pair(o[0], I.d, I.d2); mem(o[1], I.a, I.off); }
        else if (opc == "st.shared.v2.f64") { I.op = PtxIns::STS_V2; mem(o[0], I.a, I.off); pair(o[1], I.d, I.d2); }
        else if (opc == "bra.uni" || opc == "bra") { I.op = PtxIns::BRA; I.label = o[0]; }
        else { P.err = "unknown instruction: " + opc; return false; }
        P.ins.push_back(I);
    }
    for (PtxIns& I : P.ins) if (I.op == PtxIns::BRA) {
        auto it = labels.find(I.label);
        if (it == labels.end()) { P.err = "unknown label " + I.label; return false; }
        I.target = it->second;
    }
    return true;
}

// one thread through the block.  Memory: `param` = the program bytes; shared addresses [kTileS, kTileS + tile bytes) = the
// tile, kSlotS = an 8-byte cell holding the per-state coefficient base (a host pointer); global = host memory.
static const uint32_t kPtxTileS = 0x100000u, kPtxSlotS = 0x80000u;

inline bool ptx_run(const PtxProgram& P, std::vector<uint64_t>& regs, const unsigned char* param, size_t param_bytes, unsigned char* tile,
                    size_t tile_bytes, uint64_t slot_base, std::vector<int>* trace_branches) {
    auto F = [&](int r) { double v; memcpy(&v, &regs[r], 8); return v; };
    auto SF = [&](int r, double v) { memcpy(&regs[r], &v, 8); };
    for (size_t pc = 0, steps = 0; pc < P.ins.size(); ++pc) {
        if (++steps > 4000000) return false;
        const PtxIns& I = P.ins[pc];
        if (I.op == PtxIns::LABEL) continue;
        if (I.pred >= 0) { const bool p = regs[I.pred] != 0; if (p == I.pneg) continue; }
        const uint64_t bv = I.b_imm ? (uint64_t)I.imm_b : (I.b >= 0 ? regs[I.b] : 0);
        switch (I.op) {
        case PtxIns::MOV32: regs[I.d] = (uint32_t)(I.a_imm ? (uint64_t)I.imm_a : regs[I.a]); break;
        case PtxIns::MOVF: regs[I.d] = regs[I.a]; break;
        case PtxIns::XOR: regs[I.d] = (uint32_t)(regs[I.a] ^ bv); break;
        case PtxIns::AND: regs[I.d] = (uint32_t)(regs[I.a] & bv); break;
        case PtxIns::ADD32: regs[I.d] = (uint32_t)(regs[I.a] + bv); break;
        case PtxIns::ADD64: regs[I.d] = regs[I.a] + bv; break;
        case PtxIns::SHL: regs[I.d] = (uint32_t)((uint32_t)regs[I.a] << (bv & 31)); break;
        case PtxIns::UNPACK: { const uint64_t v = regs[I.a]; regs[I.d] = (uint32_t)v; regs[I.d2] = (uint32_t)(v >> 32); break; }
        case PtxIns::PACK: regs[I.d] = (uint64_t)(uint32_t)regs[I.a] | ((uint64_t)(uint32_t)regs[I.b] << 32); break;
        case PtxIns::MULWIDE: regs[I.d] = (uint64_t)(uint32_t)regs[I.a] * (uint64_t)(uint32_t)bv; break;
        case PtxIns::SETP_EQ: regs[I.d] = (uint32_t)regs[I.a] == (uint32_t)bv; break;
        case PtxIns::SETP_NE: regs[I.d] = (uint32_t)regs[I.a] != (uint32_t)bv; break;
        case PtxIns::SELP: regs[I.d] = regs[I.c] ? regs[I.a] : regs[I.b]; break;
        case PtxIns::NEG: SF(I.d, -F(I.a)); break;
        case PtxIns::MUL: SF(I.d, F(I.a) * F(I.b)); break;
        case PtxIns::FMA: SF(I.d, fma(F(I.a), F(I.b), F(I.c))); break;
        case PtxIns::LDP32: case PtxIns::LDPF: {
            const size_t n = I.op == PtxIns::LDP32 ? 4 : 8;
            if ((size_t)I.off + n > param_bytes) return false;
            uint64_t v = 0; memcpy(&v, param + I.off, n); regs[I.d] = v; break;
        }
        case PtxIns::LDS_V2: case PtxIns::STS_V2: {
            const uint64_t ad = (uint32_t)regs[I.a] + (uint64_t)I.off;
            if (ad < kPtxTileS || ad + 16 > kPtxTileS + tile_bytes || (ad & 15)) return false;
            if (I.op == PtxIns::LDS_V2) { memcpy(&regs[I.d], tile + (ad - kPtxTileS), 8); memcpy(&regs[I.d2], tile + (ad - kPtxTileS) + 8, 8); }
            else { memcpy(tile + (ad - kPtxTileS), &regs[I.d], 8); memcpy(tile + (ad - kPtxTileS) + 8, &regs[I.d2], 8); }
            break;
        }
        case PtxIns::LDS_U64: if ((uint32_t)regs[I.a] + I.off != kPtxSlotS) return false; regs[I.d] = slot_base; break;
        case PtxIns::LDG_V2: {
            if (!slot_base) return false;
            const double* p = reinterpret_cast<const double*>(regs[I.a] + (uint64_t)I.off);
            memcpy(&regs[I.d], p, 8); memcpy(&regs[I.d2], p + 1, 8); break;
        }
        case PtxIns::BRA: if (trace_branches) trace_branches->push_back((int)pc); pc = (size_t)I.target; break;
        default: return false;
        }
    }
    return true;
}

// emulate_v4 with the register-block groups executed from the generated PTX text
// (the caller encodes with canonical signs when it wants the text the engine compiles: encode_v4(..., canon = true))
inline bool emulate_v4_jit(const V4Program& prog, const PlanPass& Pp, EmuAmp* amp, int nl, uint64_t rank_or, const double* slot_lifts,
                           std::string* err = nullptr) {
    JitOperands O;
    O.ctx = "%ctx"; O.slot_s = "%slot"; O.tile_s = "%tile"; O.b0 = "%b0"; O.gi = "%gi"; O.prog = "PROG";
    for (int i = 0; i < 4; ++i) O.s[i] = "%s" + std::to_string(i);
    std::string text;
    JitEmitter em(prog, O);
    if (!em.emit(text)) { if (err) *err = "emitter refused the program"; return false; }
    PtxProgram PP;
    if (!ptx_parse(text, PP)) { if (err) *err = PP.err; return false; }
    const int r_ctx = PP.R("%ctx"), r_slot = PP.R("%slot"), r_tile = PP.R("%tile"), r_b0 = PP.R("%b0"), r_gi = PP.R("%gi");
    int r_s[4]; for (int i = 0; i < 4; ++i) r_s[i] = PP.R("%s" + std::to_string(i));
    PP.R("PROG");
    const int m = Pp.m, NB = kV4BlockBits;
    const int nthreads = 1 << (m - NB);
    std::vector<int> tbit(m);
    if (!Pp.tl.empty()) { for (int b = 0; b < m; ++b) tbit[b] = Pp.tl[b]; }
    else { for (int b = 0; b < Pp.L; ++b) tbit[b] = b; for (size_t j = 0; j < Pp.hb.size(); ++j) tbit[Pp.L + j] = Pp.hb[j]; }
    uint64_t tilemask = 0; for (int b = 0; b < m; ++b) tilemask |= 1ull << tbit[b];
    const uint64_t ntiles = 1ull << (nl - m);
    const size_t tile_bytes = (size_t)16 << m;
    std::vector<unsigned char> tile(tile_bytes);
    std::vector<uint64_t> regs(PP.reg.size() + 8, 0);
    auto swz = [](uint32_t i) { return i ^ ((i >> 3) & 7u); };
    for (uint64_t t = 0; t < ntiles; ++t) {
        uint64_t base = 0, rest = t;
        for (int b = 0; b < nl; ++b) if (!((tilemask >> b) & 1)) { base |= (rest & 1ull) << b; rest >>= 1; }
        const uint64_t gbase = base | rank_or;
        auto phys = [&](uint32_t i) { uint64_t a = base; for (int b = 0; b < m; ++b) if ((i >> b) & 1) a |= 1ull << tbit[b]; return a; };
        for (uint32_t i = 0; i < (1u << m); ++i) memcpy(&tile[(size_t)16 * swz(i)], &amp[phys(i)], 16);
        uint32_t octx = prog.ctx_const | kV4CtxAlways;
        for (int j = 0; j < prog.n_outer; ++j) {
            const int pos = (int)((prog.outer_pack[j / 10] >> (6 * (j % 10))) & 63ull);
            octx |= (uint32_t)((gbase >> pos) & 1ull) << (kV4CtxOuter + j);
        }
        for (int gi = 0; gi < prog.ngroups; ++gi) {
            const V4Group& G = prog.groups[gi];
            for (int w0 = 0; w0 < nthreads; w0 += 32) {
                std::vector<int> br0;
                for (int l = 0; l < 32 && w0 + l < nthreads; ++l) {
                    const int ttid = w0 + l;
                    const uint32_t v = (uint32_t)prog.lane_tab[gi][ttid & 31] | (uint32_t)prog.warp_tab[gi][ttid >> 5];
                    std::fill(regs.begin(), regs.end(), 0);
                    regs[r_ctx] = v | octx; regs[r_slot] = kPtxSlotS; regs[r_tile] = kPtxTileS; regs[r_b0] = swz(v) << 4; regs[r_gi] = (uint32_t)gi;
                    for (int i = 0; i < 4; ++i) regs[r_s[i]] = G.s[i];
                    std::vector<int> br;
                    if (!ptx_run(PP, regs, reinterpret_cast<const unsigned char*>(&prog), sizeof(V4Program), tile.data(), tile_bytes,
                                 (uint64_t)(uintptr_t)slot_lifts, &br)) { if (err) *err = "execution fault"; return false; }
                    // every branch of the generated code is .uni: the lanes of a warp must take the same path
                    if (l == 0) br0 = br; else if (br != br0) { if (err) *err = "divergent .uni branch"; return false; }
                }
            }
        }
        for (uint32_t i = 0; i < (1u << m); ++i) memcpy(&amp[phys(i)], &tile[(size_t)16 * swz(i)], 16);
    }
    return true;
}

}  // namespace qm
