#!/usr/bin/env python
"""Generates quromorphic.jl_b200/csrc/v4_dispatch.inc: one register-block GROUP of the fused TMA pass as ONE
inline-PTX block — address arithmetic, the 16 amplitude loads, the gate interpreter ("threaded code") and
the 16 stores.

Design notes (all measured on B200, see profiles/ and DESIGN.md):
  * 16 amplitudes (4 index bits) per thread.  With 8 the per-gate overhead of an interpreter — fetching the
    gate record, the indexed jump, the wait for the coefficients — cost more warp time than the FP64 work
    it dispatched (ncu: 46 % of stall samples in the gate bodies outside DFMA against 21 % on DFMA, FP64
    pipe 26 % busy).  Doubling the block doubles the FMAs behind every dispatch (48 per full-cost gate in
    16 independent chains), and a 12-bit tile is covered by 3 groups instead of 4.
  * Every gate body ends with its own dispatch of the next record (`brx.idx.uni`: the opcode is
    warp-uniform), so there is no loop counter (the list ends with an END record), no flag decoding and no
    branch back.  The next record's header is fetched at the top of a body, its coefficient pairs in between
    the three shear steps (see LOAD_PAIR).
  * Records are read from shared memory (16-byte broadcasts).  Reading them with LDC from the kernel-parameter
    constant bank takes the load off the LSU (tools/microbench/ldc_bench.cu: a warp-uniform 64-byte record
    costs 8 LSU wavefronts) but the constant-cache latency (~80 cycles hit, more on a miss with 16 warps at
    different places of a 15 KB program) made every variant tried slower: 27.8 ms against 21.7 ms per
    compute-bound pass at 30 qubits.
  * Every body has a second, CONTROLLED entry for a control that is not a register-block bit: it tests the record's
    control mask against the thread's context word and either falls into the body or jumps to one shared skip
    block.  The planner keeps such controls on warp-index bits or outside the tile, so the predicate is
    warp-uniform.  (A separate CTRL record in front of the gate cost a whole extra dispatch — and a pass costs
    ~0.2 ms per dispatched record at 30 qubits whatever the record does, tools/pass_model.py.)

A record is 64 bytes:  u32 op, tm, cm, aux;  f64 c0, c3, c1, c4, c2, c5.
  END record: tm, cm, aux, low word of c0 = store offsets of block bits 0..3, high word of c0 = xor for the thread's base.
  cm   controlled entry: context bits that must all be set          tm   THR: context bit selecting the coefficient set
  aux  SLOT: per-state coefficient slot
Operands:  %0 shared address of the current record ("+r")   %1 context word   %2 shared address of the per-state coefficient base
           %3 shared address of the tile buffer      %4 the thread's swizzled byte offset in the tile
           %5..%8 swizzled byte offsets of the four block bits
Opcodes (must match encode_v4 in qm_api.cu), NBC = 16 legal (B, CB) pairs:
   ROT   sign variant 0..3 x NBC      ROTX  sign variant 0, 3 x NBC      PHASE (sg set 0, sg set 1) in 00 03 30 33 x NBC
   PERMX x NBC       THR (lift sg 0, lift sg 3, complex multiply) x 5 (CB = -1, 0..3);   + V4_NUM_GATE_OPS: controlled entry
   SLOT   END
"""
import os

NB = 4                      # block bits
NA = 1 << NB                # amplitudes per thread
BCS = [(B, CB) for B in range(NB) for CB in range(-1, NB) if CB != B]
NBC = len(BCS)              # 16


def X(k): return "x%d" % k
def Y(k): return "y%d" % k


# coefficient registers: set 0 = (c0, c1, c2) = (t1, s1, t2); set 1 = (c3, c4, c5).  In a record they are stored
# pairwise, {c0, c3} {c1, c4} {c2, c5}: a lift uses t1 only in its first shear, s1 in the second, t2 in the third,
# so the NEXT record's {c0, c3} can be fetched as soon as the first shears of this body have issued, and so on —
# the coefficient latency hides behind the remaining FMAs without a second register set.
C = ["c%d" % i for i in range(6)]
LOAD_PAIR = ["ld.shared.v2.f64 {c0, c3}, [%0+16];", "ld.shared.v2.f64 {c1, c4}, [%0+32];", "ld.shared.v2.f64 {c2, c5}, [%0+48];"]


def lift_steps(x, y, t1, s1, t2, sg):
    """the three shears of one real pair, as three lists of instructions (dependent in this order)"""
    a = ["fma.rn.f64 %s, %s, %s, %s;" % (x, t1, y, x)]
    if sg & 1:
        b = ["neg.f64 ng, %s;" % y, "fma.rn.f64 %s, %s, %s, ng;" % (y, s1, x)]
    else:
        b = ["fma.rn.f64 %s, %s, %s, %s;" % (y, s1, x, y)]
    if sg & 2:
        c = ["neg.f64 ng, %s;" % x, "fma.rn.f64 %s, %s, %s, ng;" % (x, t2, y)]
    else:
        c = ["fma.rn.f64 %s, %s, %s, %s;" % (x, t2, y, x)]
    return a, b, c


def interleave(chains, pre=(None, None, None)):
    """step-major order: step 1 of every chain, then step 2, then step 3; after step j the next record's
    j-th coefficient pair is loaded; pre[j] = instructions that must precede step j (coefficient selects)"""
    out = []
    for step in range(3):
        if pre[step]:
            out += pre[step]
        for ch in chains:
            out += ch[step]
        out.append(LOAD_PAIR[step])
    return out


def ctrl_ok(k, CB):
    return CB < 0 or (k >> CB) & 1


def body_rot(B, CB, sg):
    ch = []
    for k in range(NA):
        if not ctrl_ok(k, CB) or (k >> B) & 1:
            continue
        k1 = k | (1 << B)
        ch.append(lift_steps(X(k), X(k1), C[0], C[1], C[2], sg))
        ch.append(lift_steps(Y(k), Y(k1), C[0], C[1], C[2], sg))
    return interleave(ch)


def body_rotx(B, CB, sg):
    ch = []
    for k in range(NA):
        if not ctrl_ok(k, CB) or (k >> B) & 1:
            continue
        k1 = k | (1 << B)
        ch.append(lift_steps(X(k), Y(k1), C[0], C[1], C[2], sg))
        ch.append(lift_steps(Y(k), X(k1), "n0", "n1", "n2", sg))
    return interleave(ch, (["neg.f64 n0, c0;"], ["neg.f64 n1, c1;"], ["neg.f64 n2, c2;"]))


def body_phase(B, CB, sg0, sg1):
    ch = []
    for k in range(NA):
        if not ctrl_ok(k, CB):
            continue
        if (k >> B) & 1:
            ch.append(lift_steps(X(k), Y(k), C[3], C[4], C[5], sg1))
        else:
            ch.append(lift_steps(X(k), Y(k), C[0], C[1], C[2], sg0))
    return interleave(ch)


def body_permx(B, CB):
    o = list(LOAD_PAIR)
    for k in range(NA):
        if not ctrl_ok(k, CB) or (k >> B) & 1:
            continue
        k1 = k | (1 << B)
        for a, b in ((X(k), X(k1)), (Y(k), Y(k1))):
            o += ["xor.b64 %s, %s, %s;" % (a, a, b), "xor.b64 %s, %s, %s;" % (b, b, a), "xor.b64 %s, %s, %s;" % (a, a, b)]
    return o


# THR: the thread picks coefficient set 1 when its context has the target bit (htm is read before the header
# prefetch overwrites it; the selects sit right before the step that uses them)
THR_PHI = ["and.b32 t, %1, htm;", "setp.ne.u32 phi, t, 0;"]


def body_thr_lift(CB, sg):
    ch = [lift_steps(X(k), Y(k), C[0], C[1], C[2], sg) for k in range(NA) if ctrl_ok(k, CB)]
    return interleave(ch, (["selp.f64 c0, c3, c0, phi;"], ["selp.f64 c1, c4, c1, phi;"], ["selp.f64 c2, c5, c2, phi;"]))


def body_thr_mul(CB):
    """(x, y) <- (cr x - ci y, ci x + cr y) with cr = c0, ci = c1: any unit phase, no sign variants"""
    o = ["selp.f64 c0, c3, c0, phi;", "selp.f64 c1, c4, c1, phi;", "neg.f64 n1, c1;", LOAD_PAIR[2]]
    ks = [k for k in range(NA) if ctrl_ok(k, CB)]
    for h in range(0, len(ks), 2):
        kk = ks[h:h + 2]
        o += ["mul.f64 tm%d, c1, %s;" % (k & 1, X(k)) for k in kk]
        o += ["mul.f64 %s, c0, %s;" % (X(k), X(k)) for k in kk]
        o += ["fma.rn.f64 %s, n1, %s, %s;" % (X(k), Y(k), X(k)) for k in kk]
        o += ["fma.rn.f64 %s, c0, %s, tm%d;" % (Y(k), Y(k), k & 1) for k in kk]
    return o + LOAD_PAIR[:2]


# every body: advance to the next record, fetch its header, run, (the coefficient pairs are fetched inside the
# body), jump.  The pointer is advanced BEFORE the loads that use it: an add after them has to wait until the
# loads have read their address register.
ADVANCE_HDR = ["add.u32 %0, %0, 64;", "ld.shared.v2.u32 {hop, htm}, [%0];"]
# One shared dispatcher: with a `brx.idx` at the end of every body ptxas emitted a private 384-entry table of BRA
# instructions per site (58k of 70k SASS instructions, ~1 MB of code) and instruction-fetch stalls tripled.
SHARED_DISPATCH = os.environ.get("QM_V4_SHARED_DISPATCH", "1") == "1"
DISPATCH = ["bra.uni V4DISP;"] if SHARED_DISPATCH else ["brx.idx.uni hop, V4T;"]


def addresses(tag):
    """off_j (j < 4) = b0 xor the swizzled offsets of the low two block bits set in j; the address of amplitude
    k is tile + (off_(k & 3) xor the offsets of block bits 2 and 3 as set in k).  Only four offsets stay live
    across the gate loop (sixteen addresses did not fit the register budget and were spilled)."""
    if tag == "setup":
        return ["mov.b32 o0, %4;", "xor.b32 o1, o0, %5;", "xor.b32 o2, o0, %6;", "xor.b32 o3, o1, %6;"]
    # stores: the END record carries the store offsets of the four block bits and an xor for the thread's base.  They
    # equal the load offsets unless the group ended in x! / cnot! gates on register-block bits: such a tail only
    # permutes the thread's own sixteen amplitudes, an affine map over GF(2) of the block bits, and the swizzle is
    # xor-linear, so the whole tail is folded into WHERE the registers are stored (encode_v4) and costs nothing.
    o = []
    b2, b3 = ("%7", "%8") if tag == "load" else ("e2", "e3")
    if tag == "store":
        o += ["ld.shared.u32 e0, [%0+4];", "ld.shared.v2.u32 {e1, e2}, [%0+8];", "ld.shared.v2.u32 {e3, t}, [%0+16];",
              "xor.b32 o0, %4, t;", "xor.b32 o1, o0, e0;", "xor.b32 o2, o0, e1;", "xor.b32 o3, o1, e1;"]
    for k in range(NA):
        src = "o%d" % (k & 3)
        if k & 4:
            o.append("xor.b32 ad, %s, %s;" % (src, b2))
            src = "ad"
        if k & 8:
            o.append("xor.b32 ad, %s, %s;" % (src, b3))
            src = "ad"
        o.append("add.u32 ad, %s, %%3;" % src)
        if tag == "load":
            o.append("ld.shared.v2.f64 {x%d, y%d}, [ad];" % (k, k))
        else:
            o.append("st.shared.v2.f64 [ad], {x%d, y%d};" % (k, k))
    return o


def main():
    bodies = []          # (instructions before the header prefetch, main instructions)
    for sg in range(4):
        bodies += [([], body_rot(B, CB, sg)) for B, CB in BCS]
    op_rotx = len(bodies)
    for sg in (0, 3):
        bodies += [([], body_rotx(B, CB, sg)) for B, CB in BCS]
    op_phase = len(bodies)
    for s0, s1 in ((0, 0), (0, 3), (3, 0), (3, 3)):
        bodies += [([], body_phase(B, CB, s0, s1)) for B, CB in BCS]
    op_permx = len(bodies)
    bodies += [([], body_permx(B, CB)) for B, CB in BCS]
    op_thr = len(bodies)
    for kind in (0, 3, "mul"):
        for CB in range(-1, NB):
            bodies.append((THR_PHI, body_thr_mul(CB) if kind == "mul" else body_thr_lift(CB, kind)))
    n_gate_ops = len(bodies)
    # opcode k < n_gate_ops: plain entry of body k; n_gate_ops + k: its CONTROLLED entry (the record's cm = context
    # bits that must all be set; the planner keeps such controls warp-uniform, so the warp runs the body or skips it)
    OP_SLOT, OP_END = 2 * n_gate_ops, 2 * n_gate_ops + 1
    nops = 2 * n_gate_ops + 2
    targets = ["V4L%d" % i for i in range(n_gate_ops)] + ["V4C%d" % i for i in range(n_gate_ops)] + ["V4L%d" % OP_SLOT, "V4L%d" % OP_END]
    lines = ["{", ".reg .f64 x<%d>, y<%d>, c<6>, ng, n0, n1, n2, tm<2>;" % (NA, NA), ".reg .b32 hop, hcm, htm, hax, t, ad, o<4>, e<4>;",
             ".reg .pred pact, phi;", ".reg .b64 sa, sb;",
             "V4T: .branchtargets " + ", ".join(targets) + ";"]
    lines += addresses("setup") + addresses("load")
    lines += ["ld.shared.v2.u32 {hop, htm}, [%0];"] + LOAD_PAIR
    lines += ["V4DISP:", "brx.idx.uni hop, V4T;"] if SHARED_DISPATCH else DISPATCH
    for i, (pre, b) in enumerate(bodies):
        lines.append("V4L%d:" % i)
        lines += pre + ADVANCE_HDR + b + DISPATCH
    # controlled entries: small stubs kept apart from the bodies (a stub falling through into its body made ptxas
    # clone the bodies: 70k instead of 16k SASS instructions, and the instruction cache felt it)
    for i in range(n_gate_ops):
        lines.append("V4C%d:" % i)
        lines += ["ld.shared.u32 hcm, [%0+8];", "and.b32 t, %1, hcm;", "setp.eq.u32 pact, t, hcm;", "@!pact bra.uni V4SKIP;",
                  "bra.uni V4L%d;" % i]
    # a controlled gate whose control is off: on to the next record (header, all coefficient pairs, jump)
    lines.append("V4SKIP:")
    lines += ADVANCE_HDR + LOAD_PAIR + DISPATCH
    # SLOT (aux = slot index): all six coefficients of the record that follows come from the per-state table,
    # whose base address for this tile's register sits in shared memory at %2
    lines.append("V4L%d:" % OP_SLOT)
    lines += ["ld.shared.u32 t, [%0+12];", "ld.shared.u64 sa, [%2];", "mul.wide.u32 sb, t, 64;", "add.u64 sa, sa, sb;"] + ADVANCE_HDR
    lines += ["ld.global.nc.v2.f64 {c0, c1}, [sa];", "ld.global.nc.v2.f64 {c2, c3}, [sa+16];",
              "ld.global.nc.v2.f64 {c4, c5}, [sa+32];"] + DISPATCH
    lines.append("V4L%d:" % OP_END)
    lines += addresses("store")
    lines.append("}")
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "..", "quromorphic.jl_b200", "csrc", "v4_dispatch.inc")
    defs = ("#define V4_NUM_OPS %d\n#define V4_NBC %d\n#define V4_OP_ROTX %d\n#define V4_OP_PHASE %d\n#define V4_OP_PERMX %d\n"
            "#define V4_OP_THR %d\n#define V4_NUM_GATE_OPS %d\n#define V4_OP_SLOT %d\n#define V4_OP_END %d\n"
            % (nops, NBC, op_rotx, op_phase, op_permx, op_thr, n_gate_ops, OP_SLOT, OP_END))
    with open(os.path.join(os.path.dirname(out), "v4_ops.inc"), "w") as f:
        f.write("// GENERATED by tools/gen_v4_dispatch.py — do not edit.  Opcode numbering of the gate interpreter.\n" + defs)
    with open(out, "w") as f:
        f.write("// GENERATED by tools/gen_v4_dispatch.py — do not edit.  %d gate bodies (plain + controlled entry) + SLOT + END, threaded dispatch.\n" % n_gate_ops)
        f.write("#define V4_GROUP_PTX \\\n")
        for ln in lines:
            f.write('    "%s\\n\\t" \\\n' % ln)
        f.write('    ""\n')
    print("wrote", out, "ops", nops, "ptx lines", len(lines))


if __name__ == "__main__":
    main()
