// fused_tma.cuh — K3/K4 v2: the persistent, warp-specialised, TMA-pipelined fused pass (sm_100a).
//
// One CTA = one producer warp + 2^(M-3) compute threads, resident for the whole pass.  A ring of
// STAGES tile buffers lives in shared memory; every tile (2^M amplitudes: the 2^L lowest index bits
// contiguous in memory, plus up to three runs of higher bits) is moved by ONE cp.async.bulk.tensor
// (TMA) load and ONE TMA store through a 5-d tensor map with SWIZZLE_128B, i.e. the tile lands in
// shared memory already in the bank-conflict-free layout  slot(i) = i ^ ((i >> 3) & 7).
//
//   producer lane : wait computed[s] -> TMA store(stage s) -> wait until the store has read the
//                   buffer -> TMA load of the tile after next into the same stage -> ...
//   compute warps : wait full[s] -> for every register-block group: 8 amplitudes per thread out of
//                   shared memory, all gates of the group from registers, back to shared memory,
//                   named barrier -> fence.proxy.async, arrive computed[s]
//
// So HBM traffic per pass is exactly one read and one write of the state (32 * 2^n bytes), the
// copy engine works while the FP64 pipe does, and no thread ever spends an instruction on moving
// a tile.  The pass program (groups + gates) is copied to shared memory once per CTA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "planner.hpp"

namespace qm {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// producer-side wait: back off between probes so the spinning lane does not eat issue slots of the
// compute warps that share its scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(1024);
    }
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(tm), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- gate kinds: every body is IN PLACE -----------------------------------------------------------
// A 2x2 update written as  a0' = u11 a0 + u12 a1, a1' = u21 a0 + u22 a1  needs both old values while
// it produces both new ones; inside an interpreter loop the compiler then parks results in fresh
// registers and copies the whole 32-register block at every loop edge (measured with ncu: 46 % of
// all executed instructions were IMAD.MOV/MOV, FP64 13 %).  So the engine never does that.  Every
// unitary 2x2 is applied as real plane rotations written as three shears ("lifting"):
//        x += t1*y ;   y = s1*x + (+-y) ;   x = t2*y + (+-x)
// each FMA overwrites an operand that is dead after it, the two signs are XOR masks on the high
// word (ALU pipe), and a rotation costs 3 FP64 ops per real pair instead of 4.  The host
// (qm_api.cu) derives the coefficients and checks the reconstruction to a few ulp:
//   V2_ROT   real orthogonal 2x2 (ry, H ...): lift on (a0.x, a1.x) and on (a0.y, a1.y)
//   V2_ROTX  real diagonal / imaginary off-diagonal (rx, Y ...): lift on (a0.x, a1.y), mirrored on (a0.y, a1.x)
//   V2_PHASE unit-modulus diagonal (rz, z, crz ...): lift on (a.x, a.y) of each amplitude
//   V2_PERMX exactly X (x!, cnot!, swap!): xor exchange of the pair, bit-exact, no arithmetic
// A general unitary is sent as PHASE * ROT * PHASE (9 FP64 ops per amplitude against 8 for the
// direct form, but in place).  A non-unitary matrix makes its pass take the v1 kernel.
enum { V2_ROT = 0, V2_ROTX = 1, V2_PHASE = 2, V2_PERMX = 3 };

struct V2Sub {              // 96 bytes
    uint8_t  op;            // opcode (see v2_run_subs); for a thread-level PHASE: the opcode when the target bit is 0
    uint8_t  op_hi;         // thread-level PHASE: the opcode when the target bit is 1
    uint8_t  cm_blk;        // (informational) control bit among the block bits
    uint8_t  flags;         // bit 0: per-state slot coefficients, bit 1: thread-level/outer control, bit 2: thread-level PHASE
    uint32_t cm_loc;        // control bits among the tile-local (thread) bits
    uint32_t t_loc;         // fam 1: tile-local mask of the target
    int32_t  slot;
    uint64_t cm_out;        // control bits outside the tile
    uint64_t t_out;         // fam 1: outer mask of the target
    double   c[8];          // t1 s1 t2 (set 0), t1 s1 t2 (set 1: PHASE, target bit = 1)
};
static_assert(sizeof(V2Sub) == 96, "V2Sub layout");

struct V2Group {            // 16 bytes
    uint16_t nsub, sub_off;
    uint16_t rbmask[3];
    uint16_t pad_[3];
};
static const int kV2MaxSubs = 96;
static const int kV2MaxGroups = 48;

struct V2Program {
    int32_t ngroups, nsubs;
    int32_t pad_[2];
    V2Group groups[kV2MaxGroups];
    uint16_t lane_tab[kV2MaxGroups][32];   // tile-local index bits contributed by the lane
    uint16_t warp_tab[kV2MaxGroups][16];   // ... by the warp
    V2Sub subs[kV2MaxSubs];
};
static_assert(sizeof(V2Program) % 16 == 0, "V2Program must be copyable in 16-byte units");

// The gate bodies and their dispatch live in v2_dispatch.inc (generated by tools/gen_v2_dispatch.py):
// one inline-PTX block, `brx.idx.uni` on the warp-uniform opcode, every body a straight run of in-place
// DFMAs (or xor exchanges for X) on the 16 amplitude registers.  Opcodes:
//   [0, 36)   ROT   sign variant 0..3   x 9 (B, CB)        [36, 54)  ROTX  sign variant 0, 3  x 9
//   [54, 90)  PHASE (sg bit0, sg bit1) in 00 03 30 33  x 9  [90, 99)  PERMX x 9
//   [99, 107) PHASE on a thread-level / outer bit (coefficient set picked per thread): sign 0, 3  x 4 (CB)
// sign variant: bit 0 -> the middle shear takes -y, bit 1 -> the last shear takes -x (operand negations).
#include "v2_dispatch.inc"
static const int kV2OpThr = 99;
__host__ __device__ inline int v2_bc_index(int B, int CB) {     // position of (B, CB) among the 9 legal pairs
    int i = 0;
    for (int b = 0; b < 3; ++b) for (int c = -1; c < 3; ++c) { if (c == b) continue; if (b == B && c == CB) return i; ++i; }
    return 0;
}

__device__ __forceinline__ void v2_run_subs(double2 (&a)[8], const V2Program* __restrict__ P, const V2Group& G, uint32_t loc,
                                            uint64_t gbase, const double* __restrict__ slot_base) {
    const int s0 = G.sub_off, s1 = s0 + G.nsub;
    const V2Sub* S = &P->subs[s0];
    uint4 hdr = *reinterpret_cast<const uint4*>(S);         // op | op_hi << 8 | cm_blk << 16 | flags << 24, cm_loc, t_loc, slot
    for (int s = s0; s < s1; ++s) {
        const V2Sub* cur = S;
        const uint4 h = hdr;
        // coefficients of this sub (they arrive while the header is decoded) and the header of the next one
        const double2 q0 = *reinterpret_cast<const double2*>(&cur->c[0]);
        const double2 q1 = *reinterpret_cast<const double2*>(&cur->c[2]);
        const double2 q2 = *reinterpret_cast<const double2*>(&cur->c[4]);
        if (s + 1 < s1) { S = cur + 1; hdr = *reinterpret_cast<const uint4*>(S); }
        double c0 = q0.x, c1 = q0.y, c2 = q1.x, c3 = q1.y, c4 = q2.x, c5 = q2.y;
        uint32_t op = h.x & 0xffu;
        const uint32_t flags = h.x >> 24;
        if (flags) {
            if (flags & 2u) {                                            // thread-level or outer control
                if ((loc & h.y) != h.y) continue;
                if ((gbase & cur->cm_out) != cur->cm_out) continue;
            }
            if (flags & 1u) {                                            // per-state coefficients
                const double* g = slot_base + (size_t)h.w * 8;
                c0 = g[0]; c1 = g[1]; c2 = g[2]; c3 = g[3]; c4 = g[4]; c5 = g[5];
            }
            if (flags & 4u) {                                            // PHASE on a bit this thread sees as a constant
                const bool bit = h.z ? ((loc & h.z) != 0) : ((gbase & cur->t_out) != 0);
                if (bit) { c0 = c3; c1 = c4; c2 = c5; op = (h.x >> 8) & 0xffu; }
            }
        }
        asm volatile(V2_DISPATCH_PTX
                     : "+d"(a[0].x), "+d"(a[0].y), "+d"(a[1].x), "+d"(a[1].y), "+d"(a[2].x), "+d"(a[2].y), "+d"(a[3].x), "+d"(a[3].y),
                       "+d"(a[4].x), "+d"(a[4].y), "+d"(a[5].x), "+d"(a[5].y), "+d"(a[6].x), "+d"(a[6].y), "+d"(a[7].x), "+d"(a[7].y)
                     : "d"(c0), "d"(c1), "d"(c2), "d"(c3), "d"(c4), "d"(c5), "r"(op));
    }
}

struct V2Launch {
    uint64_t ntiles;        // over the whole batch
    int32_t  reg_bits;      // index bits of one register
    int32_t  L;             // low bits
    int32_t  nh;            // high tile bits
    int32_t  hb[16];
    uint64_t rank_or;
    int32_t  nslots;
    int32_t  pad_;
};

template <int M, int STAGES>
__global__ void __launch_bounds__((1 << (M - 3)) + 32, (M >= 12 ? 1 : 2))
k_fused_tma(const __grid_constant__ CUtensorMap tmap, const V2Program* __restrict__ gprog, const double* __restrict__ slots,
            const V2Launch lp) {
    constexpr int NCOMP = 1 << (M - 3);
    constexpr uint32_t TILE_BYTES = 16u << M;
    extern __shared__ unsigned char smem_dyn[];
    // [stages][TILE_BYTES] | V2Program | barriers ; SWIZZLE_128B needs the stages 1024-byte aligned
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;
    V2Program* P = reinterpret_cast<V2Program*>(smem + (size_t)STAGES * TILE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(P) + sizeof(V2Program));
    uint64_t* computed = full + STAGES;

    const int tid = threadIdx.x;
    // program -> shared memory (16-byte copies by every thread)
    {
        const int4* src = reinterpret_cast<const int4*>(gprog);
        int4* dst = reinterpret_cast<int4*>(P);
        for (int i = tid; i < (int)(sizeof(V2Program) / 16); i += NCOMP + 32) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&computed[s], NCOMP / 32); }
        fence_mbar_init();
    }
    __syncthreads();

    const int tbits = lp.reg_bits - M;
    const uint64_t first = blockIdx.x, step = gridDim.x;
    const uint64_t ntiles = lp.ntiles;
    const uint64_t my_tiles = first < ntiles ? (ntiles - first + step - 1) / step : 0;

    auto tile_base = [&](uint64_t tile_id, uint64_t& gbase_out, uint64_t& breg_out) -> uint64_t {
        const uint64_t breg = tile_id >> tbits;
        uint64_t base = (tile_id & ((1ull << tbits) - 1ull)) << lp.L;
        for (int j = 0; j < lp.nh; ++j) base = insert0(base, lp.hb[j]);
        gbase_out = base | lp.rank_or; breg_out = breg;
        return (breg << lp.reg_bits) + base;      // amplitude index of the tile's first element in the whole array
    };

    if (tid >= NCOMP) {
        // ===================== producer warp: one elected lane drives TMA =====================
        if (tid == NCOMP) {
            uint64_t gb, br;
            const uint64_t pro = my_tiles < (uint64_t)STAGES ? my_tiles : (uint64_t)STAGES;
            for (uint64_t i = 0; i < pro; ++i) {
                const uint64_t base = tile_base(first + i * step, gb, br);
                mbar_arrive_expect_tx(&full[i], TILE_BYTES);
                tma_load_5d(stage_base + i * TILE_BYTES, &tmap, &full[i], 0, (int)(base >> 3), 0, 0, 0);
            }
            for (uint64_t i = 0; i < my_tiles; ++i) {
                const int s = (int)(i % STAGES);
                const uint32_t ph = (uint32_t)((i / STAGES) & 1);
                mbar_wait_sleep(&computed[s], ph);
                const uint64_t base = tile_base(first + i * step, gb, br);
                tma_store_5d(&tmap, stage_base + (size_t)s * TILE_BYTES, 0, (int)(base >> 3), 0, 0, 0);
                tma_commit();
                const uint64_t nxt = i + STAGES;
                if (nxt < my_tiles) {
                    tma_wait_read0();            // the store has read the buffer: safe to overwrite
                    const uint64_t nbase = tile_base(first + nxt * step, gb, br);
                    mbar_arrive_expect_tx(&full[s], TILE_BYTES);
                    tma_load_5d(stage_base + (size_t)s * TILE_BYTES, &tmap, &full[s], 0, (int)(nbase >> 3), 0, 0, 0);
                }
            }
            tma_wait_all0();
        }
        return;
    }

    // ================================ compute warps ================================
    const int lane = tid & 31, warp = tid >> 5;
    const int ngroups = P->ngroups;
    for (uint64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % STAGES);
        const uint32_t ph = (uint32_t)((i / STAGES) & 1);
        uint64_t gbase, breg;
        tile_base(first + i * step, gbase, breg);
        const double* slot_base = slots ? slots + (size_t)breg * lp.nslots * 8 : nullptr;
        double2* tile = reinterpret_cast<double2*>(stage_base + (size_t)s * TILE_BYTES);
        mbar_wait(&full[s], ph);
        for (int gi = 0; gi < ngroups; ++gi) {
            const V2Group& G = P->groups[gi];
            const uint32_t loc = (uint32_t)P->lane_tab[gi][lane] | (uint32_t)P->warp_tab[gi][warp];
            const uint32_t m0 = G.rbmask[0], m1 = G.rbmask[1], m2 = G.rbmask[2];
            double2 a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t off = ((k & 1) ? m0 : 0u) | ((k & 2) ? m1 : 0u) | ((k & 4) ? m2 : 0u);
                a[k] = tile[swz(loc | off)];
            }
            v2_run_subs(a, P, G, loc, gbase, slot_base);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t off = ((k & 1) ? m0 : 0u) | ((k & 2) ? m1 : 0u) | ((k & 4) ? m2 : 0u);
                tile[swz(loc | off)] = a[k];
            }
            if (gi + 1 < ngroups) bar_sync_named(1, NCOMP);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&computed[s]);
    }
}

}  // namespace qm
