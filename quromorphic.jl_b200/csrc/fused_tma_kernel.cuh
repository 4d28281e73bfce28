// fused_tma_kernel.cuh — the kernel template of the fused TMA pass (see fused_tma.cuh for the design and the
// program layout).  Included only by the per-shape translation units fused_m12.cu / fused_m11.cu.
#pragma once
#include "fused_tma.cuh"
#include "state.hpp"
#include "v4_dispatch.inc"

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
// producer-side wait: back off a little between probes so the spinning lane does not eat issue slots of the
// compute warps that share its scheduler (try_wait itself suspends the lane for a hardware-defined time; the first
// version slept 1 us here and the phase trace showed the producer noticing a finished tile ~1,100 cycles late)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(64);
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

__device__ __forceinline__ uint64_t ldp64(uint64_t pa) {
    uint64_t v;
    asm volatile("ld.param.u64 %0, [%1];" : "=l"(v) : "l"(pa));
    return v;
}
__device__ __forceinline__ uint32_t ldp32(uint64_t pa) {
    uint32_t v;
    asm volatile("ld.param.u32 %0, [%1];" : "=r"(v) : "l"(pa));
    return v;
}


// the tile's base row goes into TMA coordinate lp.rowdim, the other coordinates are 0
#define TMA_COORDS(b) const int tr_ = (int)((b) >> 3), tc1 = lp.rowdim == 1 ? tr_ : 0, tc2 = lp.rowdim == 2 ? tr_ : 0, \
                      tc3 = lp.rowdim == 3 ? tr_ : 0, tc4 = lp.rowdim == 4 ? tr_ : 0

// JIT = true: the same kernel with the interpreter cut out.  The register-block group is an EMPTY inline-PTX block between two
// marker comments; the build turns this instantiation into PTX text (fused_jit.cu -> jit_template_m*.inc) and at run time the
// engine splices straight-line code generated for one pass program between the markers (jit.cpp) and hands the text to the
// driver's PTX compiler.  The records never go to shared memory: the generated code reads its coefficients from the
// kernel-parameter bank at fixed offsets.
//
// ZFOLD = true: the pass also sums the all-qubit <Z> statistics of what it stores (V4Fold, fused_tma.cuh).  After the last
// group of a tile the team reads the finished tile back from shared memory — element j * TEAM + ttid, conflict-free under the
// swizzle — so tile-local bits 0..M-5 are bits of the thread index and bits M-4..M-1 bits of j; the kept probability of
// the whole tile goes to the index bits outside the tile that are set in the tile's base (lane l of every warp keeps bits l
// and l + 32).  Everything is accumulated in registers over the CTA's tiles and folded ONCE per CTA in a fixed order
// (shuffles inside a warp, then warp by warp), so the result does not depend on timing.  A separate instantiation: the
// kernels without the fold are byte for byte what they were.
template <int M, int TEAMS, int STAGES, bool JIT = false, bool ZFOLD = false>
__global__ void __launch_bounds__(TEAMS * (1 << (M - 4)) + kV4ProducerThreads, 1)
k_fused_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ V4Program P, const double* __restrict__ slots,
            const V4Launch lp) {
    constexpr int TEAM = 1 << (M - kV4BlockBits);          // threads of one team
    constexpr int NCOMP = TEAMS * TEAM;
    constexpr uint32_t TILE_BYTES = 16u << M;
    static_assert(TEAMS <= STAGES && NCOMP % 128 == 0 && TEAM % 32 == 0, "team shape");
    static_assert(!ZFOLD || (3 * STAGES <= 16 && JIT && (NCOMP / 32) * kV4FoldVals * 8 <= (int)(sizeof(V4Rec) * kV4MaxRecs) && M - 4 >= 5 && M <= 16),
                  "ZFOLD: a third descriptor word per stage, and the record area (unused by compiled kernels) as reduction scratch");
    extern __shared__ unsigned char smem_dyn[];
    // [stages][TILE_BYTES] | lane_tab, warp_tab | gate records | per-stage tile descriptors | barriers ; SWIZZLE_128B needs the stages 1024-byte aligned
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;
    unsigned char* tabs = smem + (size_t)STAGES * TILE_BYTES;
    unsigned char* recs = tabs + kV4TabBytes;
    // One 16-byte descriptor per stage, written by the producer lane together with the TMA load of the tile that goes
    // into the stage: {per-state coefficient base of the tile's register, outer context word}.  Everything a compute
    // thread needs to know about WHICH tile it holds — so the 512 compute threads never decode a tile index (the
    // first version did: 145 of ~1270 instructions per thread per tile, ncu r1m).
    uint64_t* tdesc = reinterpret_cast<uint64_t*>(recs + sizeof(V4Rec) * kV4MaxRecs);          // [STAGES][2], room for 8 stages
    uint64_t* full = tdesc + 16;
    uint64_t* computed = full + 2 * STAGES;
    // full[] holds TWO barriers per stage, used alternately (use k = i / STAGES of a stage takes barrier
    // s + STAGES * (k & 1), parity (k >> 1) & 1).  With one barrier a parity wait cannot tell "use k has
    // landed" from "use k-1 has not landed yet", and a team CAN get there: tile i and tile i + STAGES of a
    // stage belong to different teams, TMA loads complete out of order, and a team that finishes a light
    // tile before an earlier tile's load has landed would sail through the wait for the following use of
    // that stage (seen on B200 as rare wrong amplitudes or a hang with cheap passes).

    const int tid = threadIdx.x;
    uint64_t pa;                                  // param-space address of the program
    asm("cvta.to.param.u64 %0, %1;" : "=l"(pa) : "l"(&P));
    // index tables -> shared memory (the rows of the groups in use)
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(tabs);
        const int nl = P.ngroups * 16, nw = P.ngroups * 4;       // 32-bit words of lane_tab / warp_tab in use
        for (int i = tid; i < nl; i += NCOMP + kV4ProducerThreads) dst[i] = ldp32(pa + offsetof(V4Program, lane_tab) + 4u * i);
        for (int i = tid; i < nw; i += NCOMP + kV4ProducerThreads) dst[kV4MaxGroups * 16 + i] = ldp32(pa + offsetof(V4Program, warp_tab) + 4u * i);
    }
    if (!JIT) {   // gate records -> shared memory (the interpreter reads them with 16-byte broadcasts)
        uint64_t* dst = reinterpret_cast<uint64_t*>(recs);
        const int n8 = P.nrecs * 8;
        for (int i = tid; i < n8; i += NCOMP + kV4ProducerThreads) dst[i] = ldp64(pa + offsetof(V4Program, recs) + 8u * i);
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&full[STAGES + s], 1); mbar_init(&computed[s], TEAM / 32); }
        fence_mbar_init();
    }
    __syncthreads();

    const int tbits = lp.reg_bits - M - lp.nfixed;      // tile-index bits per register
    const uint64_t first = blockIdx.x, step = gridDim.x;
    const uint64_t ntiles = lp.ntiles;
    const uint64_t my_tiles = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    // optional phase trace (diagnostics, tools/tile_timeline.py): CTA 0 writes SM-clock stamps per tile
    unsigned long long* const trace = (!ZFOLD && blockIdx.x == 0) ? lp.trace : nullptr;

    if (tid >= NCOMP) {
        // ===================== producer warpgroup: one elected lane drives TMA =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == NCOMP) {
            const int n_outer = P.n_outer;
            const uint32_t ctx_const = P.ctx_const | kV4CtxAlways;
            const uint64_t opk0 = P.outer_pack[0], opk1 = P.outer_pack[1];
            // tile index -> amplitude index of its first element in the whole array (TMA row coordinate), plus the
            // stage descriptor {slot base, outer context word}
            auto tile_base = [&](uint64_t tile_id, uint64_t& d0, uint64_t& d1) -> uint64_t {
                const uint64_t breg = tile_id >> tbits;
                uint64_t base = (tile_id & ((1ull << tbits) - 1ull)) << lp.L;
#pragma unroll
                for (int j = 0; j < kV4Gaps; ++j) {     // open a gap of gap_len zero bits at gap_start (ascending)
                    const uint64_t lo = base & ((1ull << lp.gap_start[j]) - 1ull);
                    base = ((base >> lp.gap_start[j]) << (lp.gap_start[j] + lp.gap_len[j])) | lo;
                }
                base |= lp.fixed_or;                    // slab launches: some index bits pinned to a value
                const uint64_t gbase = base | lp.rank_or;
                // outer part of the context word: the index bits outside the tile that some gate of this pass
                // uses as a control or as the target of a diagonal gate
                uint32_t octx = ctx_const;
                uint64_t pk = opk0;
                for (int j = 0; j < n_outer; ++j) {
                    octx |= (uint32_t)((gbase >> (pk & 63ull)) & 1ull) << (kV4CtxOuter + j);
                    pk >>= 6;
                    if (j == 9) pk = opk1;
                }
                d0 = slots ? (uint64_t)(slots + (size_t)breg * lp.nslots * 8) : 0ull;
                d1 = octx;
                return (breg << lp.reg_bits) + base;
            };
            uint64_t d0, d1;
            const uint64_t pro = my_tiles < (uint64_t)STAGES ? my_tiles : (uint64_t)STAGES;
            for (uint64_t i = 0; i < pro; ++i) {
                const uint64_t base = tile_base(first + i * step, d0, d1);
                tdesc[2 * i] = d0; tdesc[2 * i + 1] = d1;
                if (ZFOLD) tdesc[2 * STAGES + i] = base;            // the tile's index bits outside the tile (one register: base == index)
                mbar_arrive_expect_tx(&full[i], TILE_BYTES);        // release: the descriptor is visible to whoever sees the phase
                TMA_COORDS(base);
                tma_load_5d(stage_base + i * TILE_BYTES, &tmap, &full[i], 0, tc1, tc2, tc3, tc4);
            }
            for (uint64_t i = 0; i < my_tiles; ++i) {
                const int s = (int)(i % STAGES);
                const uint32_t ph = (uint32_t)((i / STAGES) & 1);
                const uint64_t nxt = i + STAGES;
                // everything that does not depend on the wait is done before it
                uint64_t n0 = 0, n1 = 0, nbase = 0, dd0, dd1;
                const uint64_t base = tile_base(first + i * step, dd0, dd1);
                if (nxt < my_tiles) nbase = tile_base(first + nxt * step, n0, n1);
                mbar_wait_sleep(&computed[s], ph);
                if (trace) trace[kV4TraceStride * i + 6] = clock64();
                { TMA_COORDS(base); tma_store_5d(&tmap, stage_base + (size_t)s * TILE_BYTES, 0, tc1, tc2, tc3, tc4); }
                tma_commit();
                if (nxt < my_tiles) {
                    tma_wait_read0();            // the store has read the buffer: safe to overwrite
                    uint64_t* fb = &full[s + STAGES * (int)((nxt / STAGES) & 1)];
                    tdesc[2 * s] = n0; tdesc[2 * s + 1] = n1;
                    if (ZFOLD) tdesc[2 * STAGES + s] = nbase;
                    mbar_arrive_expect_tx(fb, TILE_BYTES);
                    TMA_COORDS(nbase);
                    tma_load_5d(stage_base + (size_t)s * TILE_BYTES, &tmap, fb, 0, tc1, tc2, tc3, tc4);
                    if (trace) trace[kV4TraceStride * i + 7] = clock64();
                }
            }
            tma_wait_all0();
        }
        return;
    }

    // ================================ compute teams ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int team = tid / TEAM, ttid = tid % TEAM;
    const int lane = ttid & 31, warp = ttid >> 5;
    const int ngroups = P.ngroups;
    const uint32_t stage_s = smem_u32(stage_base);
    const uint32_t desc_s = smem_u32(tdesc);
    const uint32_t lane_s = smem_u32(tabs) + 2u * lane, warp_s = smem_u32(tabs) + 2u * (kV4MaxGroups * 32 + warp);
    // ZFOLD accumulators of this thread over the CTA's tiles: kept probability of its 16 elements per tile, the same split by
    // the four bits of j, and — per lane — the tile totals of the tiles whose base has index bit lane / lane + 32 set
    double zT = 0.0, zB0 = 0.0, zB1 = 0.0, zB2 = 0.0, zB3 = 0.0, zO = 0.0, zO2 = 0.0;
    double zthr = 0.0;
    if (ZFOLD) zthr = lp.fold->thr;
    int s = team % STAGES; uint32_t k = 0;        // stage and use count of the stage (tile i: s = i % STAGES, k = i / STAGES)
#pragma unroll 1
    for (uint32_t i = team; i < (uint32_t)my_tiles; i += TEAMS) {
        const uint32_t tile_s = stage_s + (uint32_t)s * TILE_BYTES;
        const uint32_t slot_s = desc_s + 16u * (uint32_t)s;        // the SLOT record reads the coefficient base from here
        const bool tr = trace && ttid == 0;
        if (tr) trace[kV4TraceStride * i + 0] = clock64();
        mbar_wait(&full[s + STAGES * (int)(k & 1u)], (k >> 1) & 1u);
        if (tr) trace[kV4TraceStride * i + 1] = clock64();
        uint32_t octx;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(octx) : "r"(slot_s + 8u));
#pragma unroll 1
        for (int gi = 0; gi < ngroups; ++gi) {
            const uint64_t ga = pa + offsetof(V4Program, groups) + 32u * gi;
            const uint64_t g01 = ldp64(ga), g23 = ldp64(ga + 8);
            const uint32_t gw = ldp32(ga + 16);
            uint32_t v, vw;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(lane_s + 64u * gi));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(vw) : "r"(warp_s + 16u * gi));
            v |= vw;
            // the swizzle slot(i) = i ^ ((i >> 3) & 7) is linear over xor, so the sixteen addresses of the
            // register block are the thread's base xor the (swizzled) block-bit offsets
            const uint32_t b0 = (v ^ ((v >> 3) & 7u)) << 4;
            const uint32_t ctx = v | octx;
            uint32_t rp = smem_u32(recs) + (gw - (uint32_t)offsetof(V4Program, recs));
            if (JIT)
                asm volatile("// QMJIT_BEGIN %0 %1 %2 %3 %4 %5 %6 %7 %8 %9\n\t// QMJIT_END\n\t"
                             : "+r"(rp)
                             : "r"(ctx), "r"(slot_s), "r"(tile_s), "r"(b0), "r"((uint32_t)g01), "r"((uint32_t)(g01 >> 32)),
                               "r"((uint32_t)g23), "r"((uint32_t)(g23 >> 32)), "r"(gi)
                             : "memory");
            else
            asm volatile(V4_GROUP_PTX
                         : "+r"(rp)
                         : "r"(ctx), "r"(slot_s), "r"(tile_s), "r"(b0), "r"((uint32_t)g01), "r"((uint32_t)(g01 >> 32)),
                           "r"((uint32_t)g23), "r"((uint32_t)(g23 >> 32))
                         : "memory");
            if (tr && gi < 3) trace[kV4TraceStride * i + 2 + gi] = clock64();
            if (gi + 1 < ngroups) bar_sync_named(1 + team, TEAM);
        }
        if (ZFOLD) {
            bar_sync_named(1 + team, TEAM);          // the last group's stores of the whole team are in the tile
            uint64_t tbase;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(tbase) : "r"(desc_s + 8u * (uint32_t)(2 * STAGES + s)));
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t e = (uint32_t)j * TEAM + (uint32_t)ttid;
                const uint32_t slot = e ^ ((e >> 3) & 7u);
                double re, im;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(re), "=d"(im) : "r"(tile_s + 16u * slot));
                const double pr = __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));      // abs2_exact (kernels.cuh): the oracle's abs2, no contraction
                v[j] = pr > zthr ? pr : 0.0;
            }
            {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (j & 1) a0 += v[j];
                    if (j & 2) a1 += v[j];
                    if (j & 4) a2 += v[j];
                    if (j & 8) a3 += v[j];
                }
                zB0 += a0; zB1 += a1; zB2 += a2; zB3 += a3;
            }
#pragma unroll
            for (int w = 8; w > 0; w >>= 1)
#pragma unroll
                for (int j = 0; j < w; ++j) v[j] += v[j + w];
            double tw = v[0];
            zT += tw;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tw += __shfl_xor_sync(0xffffffffu, tw, o);      // the warp's share of the tile
            zO += ((tbase >> lane) & 1ull) ? tw : 0.0;
            zO2 += ((tbase >> (lane + 32)) & 1ull) ? tw : 0.0;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&computed[s]);
        if (tr) trace[kV4TraceStride * i + 5] = clock64();
        s += TEAMS; if (s >= STAGES) { s -= STAGES; ++k; }
    }
    if (ZFOLD) {
        // one fold per CTA, fixed order.  Row w of zred (the record area: compiled kernels keep no records in shared memory)
        // belongs to compute warp w: the lanes write their outer-bit sums — exactly zero at tile-bit positions and at 63,
        // because a tile's base has no such bit — then lane 0 writes the tile-bit sums and the total over them.
        const V4Fold* const zf = lp.fold;
        double* const zred = reinterpret_cast<double*>(recs);
        const int cw = tid >> 5;                                    // compute warp of the CTA
        double* const row = zred + cw * kV4FoldVals;
        auto wsum = [](double x) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            return x;
        };
        row[lane] = zO; row[lane + 32] = zO2;
        __syncwarp();
        const double wt = wsum(zT);
        double tb[M];                                               // kept probability with tile-local bit k set, this warp
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) tb[kk] = wsum(((lane >> kk) & 1) ? zT : 0.0);
#pragma unroll
        for (int kk = 5; kk < M - 4; ++kk) tb[kk] = ((warp >> (kk - 5)) & 1) ? wt : 0.0;
        tb[M - 4] = wsum(zB0); tb[M - 3] = wsum(zB1); tb[M - 2] = wsum(zB2); tb[M - 1] = wsum(zB3);
        if (lane == 0) {
#pragma unroll
            for (int kk = 0; kk < M; ++kk) row[zf->tl[kk] & 63u] = tb[kk];
            row[kV4FoldVals - 1] = wt;
        }
        bar_sync_named(8, NCOMP);                                   // the compute warps only: the producer warpgroup is not here
        if (tid < kV4FoldVals) {
            double acc = 0.0;
#pragma unroll 1
            for (int w = 0; w < NCOMP / 32; ++w) acc += zred[w * kV4FoldVals + tid];
            zf->part[(size_t)blockIdx.x * kV4FoldVals + tid] = acc;
        }
    }
}

template <int M, int TEAMS, int STAGES>
static int launch_v4_t(cudaStream_t stream, int sms, const CUtensorMap& tm, const V4Program& prog, const double* d_slots,
                       const V4Launch& lp) {
    constexpr int threads = TEAMS * (1 << (M - kV4BlockBits)) + kV4ProducerThreads;
    constexpr size_t smem = (size_t)STAGES * (16u << M) + kV4TabBytes + sizeof(V4Rec) * kV4MaxRecs + 128 + 3 * STAGES * 8 + 1024;
    // the attribute belongs to the device, not to the process: one flag per device (qm_create takes `device`)
    static bool attr[64] = { false };
    int dev = 0; CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        CU(cudaFuncSetAttribute(k_fused_tma<M, TEAMS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    const uint64_t cap = (uint64_t)(lp.sm_cap > 0 && lp.sm_cap < sms ? lp.sm_cap : sms);
    const int grid = (int)(lp.ntiles < cap ? lp.ntiles : cap);
    k_fused_tma<M, TEAMS, STAGES><<<grid, threads, smem, stream>>>(tm, prog, d_slots, lp);
    CU(cudaGetLastError());
    return QM_OK;
}

}  // namespace qm
