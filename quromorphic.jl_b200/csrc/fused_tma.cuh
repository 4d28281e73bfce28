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

// ---- specialised gate kinds (chosen on the host from the matrix bits) -------------------------
enum {
    V2_MAT = 0,    // general complex 2x2                         16 FMA-class ops per pair
    V2_MATR = 1,   // all four entries real (ry, H, Z*ry ...)        8
    V2_MATX = 2,   // real diagonal, imaginary off-diagonal (rx, Y)  8
    V2_PERMX = 3,  // exactly X: exchange the pair, no arithmetic    0   (x!, cnot!, swap!: bit-exact)
    V2_DIAG = 4    // diagonal (rz, z, crz ...)                      4 per amplitude
};

struct V2Sub {              // 96 bytes
    uint8_t  kind;          // V2_*
    uint8_t  b;             // block bit 0..2 of the target; V2_DIAG: 3 = tile-local bit, 4 = outer bit
    uint8_t  cm_blk;        // control bits among the block bits
    uint8_t  has_slot;
    uint32_t cm_loc;        // control bits among the tile-local (thread) bits
    uint32_t t_loc;         // V2_DIAG b == 3: tile-local mask of the target
    int32_t  slot;
    uint64_t cm_out;        // control bits outside the tile
    uint64_t t_out;         // V2_DIAG b == 4: outer mask of the target
    double   c[8];          // MAT: U11 U21 U12 U22 (re, im) ; MATR: u11 u21 u12 u22 ; MATX: p s r q ; DIAG: d0 (re,im), d1 (re,im)
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

template <int B>
__device__ __forceinline__ void v2_mat(double2 (&a)[8], const double* __restrict__ m, uint32_t cm) {
    const double2 u11 = make_double2(m[0], m[1]), u21 = make_double2(m[2], m[3]);
    const double2 u12 = make_double2(m[4], m[5]), u22 = make_double2(m[6], m[7]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k & (1 << B)) continue;
        if ((k & cm) == cm) {
            const double2 a0 = a[k], a1 = a[k | (1 << B)];
            a[k] = cfma(u12, a1, cmul(u11, a0));
            a[k | (1 << B)] = cfma(u22, a1, cmul(u21, a0));
        }
    }
}
template <int B>
__device__ __forceinline__ void v2_matr(double2 (&a)[8], const double* __restrict__ c, uint32_t cm) {
    const double u11 = c[0], u21 = c[1], u12 = c[2], u22 = c[3];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k & (1 << B)) continue;
        if ((k & cm) == cm) {
            const double2 a0 = a[k], a1 = a[k | (1 << B)];
            a[k] = make_double2(fma(u12, a1.x, u11 * a0.x), fma(u12, a1.y, u11 * a0.y));
            a[k | (1 << B)] = make_double2(fma(u22, a1.x, u21 * a0.x), fma(u22, a1.y, u21 * a0.y));
        }
    }
}
// u11 = p, u21 = i s, u12 = i r, u22 = q
template <int B>
__device__ __forceinline__ void v2_matx(double2 (&a)[8], const double* __restrict__ c, uint32_t cm) {
    const double p = c[0], s = c[1], r = c[2], q = c[3];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k & (1 << B)) continue;
        if ((k & cm) == cm) {
            const double2 a0 = a[k], a1 = a[k | (1 << B)];
            a[k] = make_double2(fma(-r, a1.y, p * a0.x), fma(r, a1.x, p * a0.y));
            a[k | (1 << B)] = make_double2(fma(-s, a0.y, q * a1.x), fma(s, a0.x, q * a1.y));
        }
    }
}
template <int B>
__device__ __forceinline__ void v2_permx(double2 (&a)[8], uint32_t cm) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k & (1 << B)) continue;
        if ((k & cm) == cm) { const double2 t = a[k]; a[k] = a[k | (1 << B)]; a[k | (1 << B)] = t; }
    }
}

__device__ __forceinline__ void v2_run_subs(double2 (&a)[8], const V2Program* __restrict__ P, const V2Group& G, uint32_t loc,
                                            uint64_t gbase, const double* __restrict__ slot_base) {
    const int s0 = G.sub_off, s1 = s0 + G.nsub;
    for (int s = s0; s < s1; ++s) {
        const V2Sub& S = P->subs[s];
        if ((loc & S.cm_loc) != S.cm_loc) continue;
        if ((gbase & S.cm_out) != S.cm_out) continue;
        const uint32_t cm = S.cm_blk;
        switch (S.kind) {
        case V2_MAT: {
            const double* m = S.has_slot ? (slot_base + (size_t)S.slot * 8) : S.c;
            switch (S.b) { case 0: v2_mat<0>(a, m, cm); break; case 1: v2_mat<1>(a, m, cm); break; default: v2_mat<2>(a, m, cm); }
            break;
        }
        case V2_MATR:
            switch (S.b) { case 0: v2_matr<0>(a, S.c, cm); break; case 1: v2_matr<1>(a, S.c, cm); break; default: v2_matr<2>(a, S.c, cm); }
            break;
        case V2_MATX:
            switch (S.b) { case 0: v2_matx<0>(a, S.c, cm); break; case 1: v2_matx<1>(a, S.c, cm); break; default: v2_matx<2>(a, S.c, cm); }
            break;
        case V2_PERMX:
            switch (S.b) { case 0: v2_permx<0>(a, cm); break; case 1: v2_permx<1>(a, cm); break; default: v2_permx<2>(a, cm); }
            break;
        default: {   // V2_DIAG
            const double2 d0 = make_double2(S.c[0], S.c[1]), d1 = make_double2(S.c[2], S.c[3]);
            if (S.b < 3) {
                const int b = S.b;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((k & cm) == cm) a[k] = cmul(((k >> b) & 1) ? d1 : d0, a[k]);
            } else {
                const bool bit = (S.b == 3) ? ((loc & S.t_loc) != 0) : ((gbase & S.t_out) != 0);
                const double2 d = bit ? d1 : d0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((k & cm) == cm) a[k] = cmul(d, a[k]);
            }
        }
        }
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
                mbar_wait(&computed[s], ph);
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
