// fused_tma.cuh — K3/K4: the persistent, warp-specialised, TMA-pipelined fused pass (sm_100a).
//
// One CTA per SM, resident for the whole pass: a producer warpgroup (one elected lane drives TMA) and
// TEAMS compute teams of 2^(M-4) threads.  A ring of STAGES tile buffers lives in shared memory; every
// tile (2^M amplitudes: the 2^L lowest index bits contiguous in memory, plus up to three runs of higher
// bits) is moved by ONE cp.async.bulk.tensor (TMA) load and ONE TMA store through a 5-d tensor map with
// SWIZZLE_128B, i.e. the tile lands in shared memory already in the bank-conflict-free layout
// slot(i) = i ^ ((i >> 3) & 7).  Tile i of the CTA goes to stage i % STAGES and to team i % TEAMS.
//
//   producer lane : wait computed[s] -> TMA store(stage s) -> wait until the store has read the
//                   buffer -> TMA load of the tile STAGES ahead into the same stage -> ...
//   compute team  : wait full[s] -> for every register-block group: 16 amplitudes per thread out of
//                   shared memory, all gates of the group from registers, back to shared memory,
//                   team barrier -> fence.proxy.async, arrive computed[s]
//
// So HBM traffic per pass is exactly one read and one write of the state (32 * 2^n bytes), the copy
// engine works while the FP64 pipe does, no thread ever spends an instruction on moving a tile, and the
// teams run out of phase: one team's shared-memory and barrier phases overlap the other's FP64 phases.
// Registers are rebalanced with setmaxnreg: the kernel starts at 96 per thread (640 threads), the producer
// warpgroup drops to 24 and releases 128 x 72 registers to the CTA pool, the 512 compute threads take 16 more each (112).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "planner.hpp"

namespace qm {

// ---- gate kinds: every body is IN PLACE -----------------------------------------------------------
// A 2x2 update written as  a0' = u11 a0 + u12 a1, a1' = u21 a0 + u22 a1  needs both old values while
// it produces both new ones; inside an interpreter the compiler then parks results in fresh registers
// and copies the whole register block at every edge (measured with ncu on the first version: 46 %
// of all executed instructions were IMAD.MOV/MOV, FP64 13 %).  So the engine never does that.  Every
// unitary 2x2 is applied as real plane rotations written as three shears ("lifting"):
//        x += t1*y ;   y = s1*x + (+-y) ;   x = t2*y + (+-x)
// each FMA overwrites an operand that is dead after it, the negations are operand modifiers, and a
// rotation costs 3 FP64 ops per real pair instead of 4.  The host (qm_api.cu, lifting.hpp) derives
// the coefficients and checks the reconstruction to a few ulp:
//   ROT   real orthogonal 2x2 (ry, H ...): lift on (a0.x, a1.x) and on (a0.y, a1.y)
//   ROTX  real diagonal / imaginary off-diagonal (rx, Y ...): lift on (a0.x, a1.y), mirrored on (a0.y, a1.x)
//   PHASE unit-modulus diagonal (rz, z, crz ...): lift on (a.x, a.y) of each amplitude
//   PERMX exactly X (x!, cnot!, swap!): xor exchange of the pair, bit-exact, no arithmetic
//   THR   unit-modulus diagonal whose target bit is constant for the thread (a thread-index bit or a bit
//         outside the tile): the thread picks one of two coefficient sets; lift when both halves share a
//         sign variant, a plain complex multiply (4 FP64 ops) otherwise
// A general unitary is sent as PHASE * ROT * PHASE (9 FP64 ops per amplitude against 8 for the
// direct form, but in place).  A non-unitary matrix sends the whole circuit to the general (v1) kernel.
//
// The interpreter is threaded code (tools/gen_v4_dispatch.py -> v4_dispatch.inc, which documents the
// opcode table, the record format and the measurements behind the design).
#include "v4_ops.inc"

static const int kV4BlockBits = 4;     // amplitudes per thread = 16
static const int kV4CtxOuter = 13;     // context word: bits [0, 13) tile-local index bits of the thread, [13, 31) outer bits, bit 31 = 1
static const int kV4MaxOuter = 31 - kV4CtxOuter;
static const uint32_t kV4CtxAlways = 0x80000000u;
static const int kV4MaxRecs = 240;     // gate records + SLOT prefixes + one END per group
static const int kV4MaxGroups = 48;
static const int kV4MaxSubs = 96;
// planner budget units (0.05 ms at 30 qubits) one register-block group costs on top of its gates, and the fixed cost
// of a pass; see planner.hpp gate_cost for the fit
static const double kV4GroupCost = 18.0;
static const double kV4GroupCostCompiled = 29.0;     // the compiled pass kernels (jit.hpp): 1.45 ms per group against 0.11 ms per record

struct V4Rec {              // 64 bytes, read by the PTX interpreter with shared-memory broadcasts
    uint32_t op;            // opcode
    uint32_t tm;            // THR: context bit selecting the coefficient set
    uint32_t cm;            // controlled entry: context bits that must all be set (a control outside the register block)
    uint32_t aux;           // SLOT: per-state slot index
    double   c[6];          // pairwise {t1, t1'} {s1, s1'} {t2, t2'} (set 0, set 1)  |  THR multiply: {cr, cr'} {ci, ci'} -
};
static_assert(sizeof(V4Rec) == 64, "V4Rec layout");

struct V4Group {            // 32 bytes
    uint32_t s[4];          // swizzled byte offsets of the four register-block bits inside a tile buffer
    uint32_t rec_off;       // byte offset of the group's first record from the start of the program
    uint32_t pad_[3];
};

// The whole pass program travels as a KERNEL PARAMETER (__grid_constant__, ~22 KB of the 32 KB the
// launch ABI allows): no upload and no staging buffer on the host side.  The CTA copies the index tables
// and the gate records to shared memory once; group records are read from the parameter bank (LDC).
struct V4Program {
    int32_t  ngroups, nrecs, n_outer;
    uint32_t ctx_const;
    uint64_t outer_pack[2]; // 6-bit fields: global index bit behind context bit kV4CtxOuter + j (10 per word)
    uint32_t flags;         // bit 0: canonical signs (compiled pass kernels only, encode_v4 `canon`): no sign-variant opcodes, a record's aux bits 0 / 1 = negate the amplitudes that took coefficient set 0 / 1
    uint32_t pad_[3];
    V4Group  groups[kV4MaxGroups];
    uint16_t lane_tab[kV4MaxGroups][32];    // tile-local index bits contributed by the lane
    uint16_t warp_tab[kV4MaxGroups][8];     // ... by the warp within its team
    V4Rec    recs[kV4MaxRecs];
};
static_assert(offsetof(V4Program, groups) % 16 == 0 && offsetof(V4Program, recs) % 16 == 0 && offsetof(V4Program, lane_tab) % 16 == 0, "V4Program alignment");
static_assert(sizeof(V4Program) + 512 < 32764, "kernel parameter space");
static const int kV4TabBytes = (int)(sizeof(uint16_t) * kV4MaxGroups * 40);

__host__ __device__ inline int v4_bc_index(int B, int CB) {     // position of (B, CB) among the 16 legal pairs
    int i = 0;
    for (int b = 0; b < kV4BlockBits; ++b) for (int c = -1; c < kV4BlockBits; ++c) { if (c == b) continue; if (b == B && c == CB) return i; ++i; }
    return 0;
}

static const int kV4Gaps = 6;          // <= 3 runs of high tile bits + <= 3 pinned slab bits, merged in ascending order
static const int kV4TraceStride = 8;   // phase trace: 8 clock stamps per tile of CTA 0 (diagnostics)

struct V4Launch {
    uint64_t ntiles;        // tiles this launch covers (the whole batch, or one slab of it)
    int32_t  reg_bits;      // index bits of one register
    int32_t  L;             // low bits
    // The tile index enumerates the index bits that are NOT tile bits and not pinned; gaps are opened for the runs
    // of high tile bits (they stay zero: the TMA box spans them) and for pinned bits (fixed_or supplies their value).
    int32_t  gap_start[kV4Gaps];
    int32_t  gap_len[kV4Gaps];
    uint64_t fixed_or;      // slab launches (sharded registers, exchange overlapped with passes): index bits pinned to a value
    uint64_t rank_or;
    int32_t  nslots;
    int32_t  rowdim;        // which TMA dimension (1..4) takes the tile's base row as its coordinate
    int32_t  sm_cap;        // 0 = all SMs; otherwise the persistent grid leaves the other SMs to a concurrent exchange kernel
    int32_t  nfixed;        // number of pinned bits (they are not enumerated by the tile index)
    union {                      // same size and offset either way: the parameter layout of every kernel is unchanged
        unsigned long long* trace;   // nullptr, or device buffer of kV4TraceStride stamps per tile of CTA 0
        const struct V4Fold* fold;   // ZFOLD instantiations (a pass that also sums the all-qubit <Z> statistics): never null there
    };
};

// Device-side parameters of a ZFOLD pass.  The pass is the LAST one of a circuit whose caller wants measure_z of every qubit
// (QSim.jl:297-318) right after it: each team sums the thresholded |a|^2 of the tile it has just finished from shared memory,
// per index bit, so the state is not read from HBM again by the readout kernels.
static const int kV4FoldVals = 64;     // [b] = kept probability with physical index bit b set (b < 63), [63] = kept total
struct V4Fold {
    double   thr;                      // probabilities not above thr are dropped (strict >, QSim.jl:308); negative: keep all
    double*  part;                     // [gridDim.x][kV4FoldVals], written by every CTA of the launch
    uint32_t tl[16];                   // physical index bit of tile-local position k (PlanPass::tl)
};

static const int kV4ProducerThreads = 128;     // a whole warpgroup, so that setmaxnreg can hand its registers over

// One translation unit per kernel shape (fused_m12.cu, fused_m11.cu): the interpreter is ~16k SASS instructions
// per instantiation and ptxas needs minutes for it, so the two are compiled in parallel.
//   M = 12: 2 teams x 256 threads, 3 stages of 64 KiB        M = 11: 4 teams x 128 threads, 6 stages of 32 KiB
int launch_v4_m12(cudaStream_t stream, int sms, const CUtensorMap& tm, const V4Program& prog, const double* d_slots, const V4Launch& lp);
int launch_v4_m11(cudaStream_t stream, int sms, const CUtensorMap& tm, const V4Program& prog, const double* d_slots, const V4Launch& lp);


}  // namespace qm
