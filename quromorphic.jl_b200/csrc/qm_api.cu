// qm_api.cu — the C ABI of include/qm_b200.h: state handle, gate queue, pass launches, readouts.
// There is NO CPU fallback here: without a CUDA device every state-creating call returns
// QM_ERR_NO_DEVICE.  The only GPU-free entry points are the planner introspection and version.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <math.h>
#include <stdio.h>
#include <chrono>
#include <deque>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qm_b200.h"
#include "dist.hpp"
#include "encode_v4.hpp"
#include "jit.hpp"
#include "jit_codegen.hpp"
#include "dist_bits.hpp"
#include "fused_tma.cuh"
#include "kernels.cuh"
#include "lifting.hpp"
#include "planner.hpp"
#include "qinfo_kernels.cuh"
#include "state.hpp"

using namespace qm;

static thread_local char g_err[512] = "";
int set_err(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}

// Scratch buffers only grow.  The outgrown buffer is NOT freed here: cudaFree / cudaFreeHost wait for the whole device,
// and when several ranks of a sharded register share a device (same-process worlds) another rank's stream may be parked
// on a flag this very call is about to post.  Outgrown buffers are released by qm_destroy.
int ensure_dev(void** p, size_t* cap, size_t need, std::vector<void*>* retired) {
    if (*cap >= need) return QM_OK;
    if (*p) retired->push_back(*p);
    *p = nullptr; *cap = 0;
    size_t c = need < 4096 ? 4096 : need;
    CU(cudaMalloc(p, c)); *cap = c;
    return QM_OK;
}
int ensure_pinned(void** p, size_t* cap, size_t need, std::vector<void*>* retired) {
    if (*cap >= need) return QM_OK;
    if (*p) retired->push_back(*p);
    *p = nullptr; *cap = 0;
    size_t c = need < 4096 ? 4096 : need;
    CU(cudaMallocHost(p, c)); *cap = c;
    return QM_OK;
}

static const size_t kTraceBytes = 8u << 20;      // QM_OPT_TRACE buffer: 8 stamps x 8 bytes x up to 131072 tiles of CTA 0

static int grid_for(uint64_t work_items, int threads, int sms, int per_sm) {
    uint64_t blocks = (work_items + threads - 1) / threads;
    uint64_t cap = (uint64_t)sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
int qm_create_common(int n, uint64_t basis, int batch, int device, int rank, int world, qm_state** out) {
    if (!out) return set_err(QM_ERR_ARG, "out is null");
    *out = nullptr;
    if (n < 1 || n > 62) return set_err(QM_ERR_ARG, "n_qubits %d out of range", n);
    if (batch < 1) return set_err(QM_ERR_ARG, "batch must be >= 1");
    int g = 0; while ((1 << g) < world) ++g;
    if ((1 << g) != world || rank < 0 || rank >= world) return set_err(QM_ERR_ARG, "world must be a power of two, 0 <= rank < world");
    if (g >= n) return set_err(QM_ERR_ARG, "more ranks than amplitudes");
    if (world > 1 && batch != 1) return set_err(QM_ERR_ARG, "sharded registers are not batched");
    if (basis >> n) return set_err(QM_ERR_ARG, "basis index out of range");   // statevector(n,m): BoundsError, QSim.jl:35
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return set_err(QM_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return set_err(QM_ERR_ARG, "device %d of %d", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
    qm_state* st = new qm_state();
    st->n = n; st->g = g; st->nl = n - g; st->rank = rank; st->world = world; st->batch = batch; st->device = device;
    // sharded registers keep 4 low bits per tile by default (8 new qubits per pass instead of 7): with the last pass before an
    // exchange postponed across it (try_defer_last_pass) a reservoir step then takes 4.0-4.5 passes instead of 5
    // (profiles/r3l_*: 34 qubits on 2 GPUs 330 -> 295 ms per step; 3 low bits: 303 ms, the 128-byte rows cost more than they save)
    if (g > 0) st->low_bits = 4;
    st->sms = prop.multiProcessorCount;
    st->perm.resize(n); for (int i = 0; i < n; ++i) st->perm[i] = i;
    // a sharded register carries its peer window (flags + mailboxes, dist.cu) right behind the shard, in the same
    // allocation: one IPC handle maps both into the other ranks
    const size_t bytes = (size_t)total_amps(st) * 16 + (world > 1 ? dist_window_bytes() : 0);
    cudaError_t e = cudaMalloc((void**)&st->amp, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); delete st; return set_err(QM_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    int rc = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&st->ev0)); CU(cudaEventCreate(&st->ev1));
        CU(cudaEventCreateWithFlags(&st->ev_stage, cudaEventDisableTiming));
        return qm_reset(st, basis);
    }();
    if (rc != QM_OK) { qm_destroy(st); return rc; }
    *out = st;
    return QM_OK;
}

extern "C" int32_t qm_create(int32_t n, uint64_t basis, int32_t batch, int32_t device, qm_state** out) {
    return qm_create_common(n, basis, batch, device, 0, 1, out);
}

extern "C" int32_t qm_destroy(qm_state* st) {
    if (!st) return QM_OK;
    cudaSetDevice(st->device);
    if (st->stream) cudaStreamSynchronize(st->stream);
    if (dist_owns_stream(st)) st->stream = nullptr;
    if (st->dist) { dist_destroy(st->dist); st->dist = nullptr; }
    for (auto& pr : st->remap_events) { if (pr.first) cudaEventDestroy(pr.first); if (pr.second) cudaEventDestroy(pr.second); }
    for (cudaEvent_t ev : st->pass_events) cudaEventDestroy(ev);
    st->remap_events.clear(); st->pass_events.clear();
    cudaFree(st->trace_buf);
    cudaFree(st->amp); cudaFree(st->d_blob); cudaFree(st->d_scratch); cudaFree(st->d_slots); cudaFree(st->d_zfold);
    cudaFreeHost(st->h_blob); cudaFreeHost(st->h_scratch); cudaFreeHost(st->h_slots); cudaFreeHost(st->h_zfold);
    for (int i = 0; i < 2; ++i) if (st->ev_slots[i]) cudaEventDestroy(st->ev_slots[i]);
    for (int i = 0; i < 2; ++i) { if (st->ev_zq[i]) cudaEventDestroy(st->ev_zq[i]); cudaFreeHost(st->h_zq[i]); }
    for (void* q : st->retired_dev) cudaFree(q);
    for (void* q : st->retired_pinned) cudaFreeHost(q);
    if (st->ev0) cudaEventDestroy(st->ev0);
    if (st->ev1) cudaEventDestroy(st->ev1);
    if (st->ev_stage) cudaEventDestroy(st->ev_stage);
    if (st->stream) cudaStreamDestroy(st->stream);
    delete st;
    return QM_OK;
}

extern "C" int32_t qm_reset(qm_state* st, uint64_t basis) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    if (basis >> st->n) return set_err(QM_ERR_ARG, "basis index out of range");
    CU(cudaSetDevice(st->device));
    st->queue.clear();
    st->gphase[0] = 1.0; st->gphase[1] = 0.0;
    for (int i = 0; i < st->n; ++i) st->perm[i] = i;
    dist_reset(st);
    CU(cudaMemsetAsync(st->amp, 0, (size_t)total_amps(st) * 16, st->stream));
    uint64_t owner = st->g ? (basis >> st->nl) : 0;
    if ((int)owner == st->rank) {
        uint64_t local = basis & (local_amps(st) - 1);
        k_set_basis<<<(st->batch + 255) / 256, 256, 0, st->stream>>>(st->amp, local_amps(st), st->batch, local);
        CU(cudaGetLastError());
        st->ctr.launches++;
    }
    return QM_OK;
}

extern "C" int32_t qm_set_option(qm_state* st, int32_t key, int64_t value) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    switch (key) {
    case QM_OPT_EAGER: st->eager = value != 0; break;
    case QM_OPT_TILE_BITS:
        if (value != 0 && (value < 4 || value > kMaxTileBits)) return set_err(QM_ERR_ARG, "tile bits must be 0 (automatic) or in [4, %d]", kMaxTileBits);
        st->tile_bits = (int)value; break;
    case QM_OPT_LOW_BITS:
        if (value < 0 || value > 8) return set_err(QM_ERR_ARG, "low bits must be in [0, 8]");
        st->low_bits = (int)value; break;
    case QM_OPT_FUSE: st->fuse = value != 0; break;
    case QM_OPT_TRACE:
        CU(cudaSetDevice(st->device));
        if (value && !st->trace_buf) { CU(cudaMalloc((void**)&st->trace_buf, kTraceBytes)); CU(cudaMemset(st->trace_buf, 0, kTraceBytes)); }
        if (!value && st->trace_buf) { CU(cudaStreamSynchronize(st->stream)); cudaFree(st->trace_buf); st->trace_buf = nullptr; }
        break;
    case QM_OPT_PIPELINE: st->pipeline = value != 0 ? 1 : 0; break;
    case QM_OPT_MAX_COST: st->max_cost = (double)value; break;
    case QM_OPT_TIME_PASSES: st->time_passes = value != 0; break;
    case QM_OPT_OVERLAP: st->overlap = value != 0; break;
    case QM_OPT_XSM:
        if (value < 1 || value > 64) return set_err(QM_ERR_ARG, "exchange SMs must be in [1, 64]");
        st->xsm = (int)value; break;
    case QM_OPT_XUN:
        if (value != 4 && value != 8) return set_err(QM_ERR_ARG, "pairs in flight per exchange thread: 4 or 8");
        st->xun = (int)value; break;
    case QM_OPT_OVERLAP_DEPTH:
        if (value < 1 || value > 4) return set_err(QM_ERR_ARG, "overlap depth must be in [1, 4]");
        st->overlap_depth = (int)value; break;
    case QM_OPT_SLAB_BITS:
        if (value < 1 || value > 3) return set_err(QM_ERR_ARG, "slab bits must be in [1, 3]");
        st->slab_bits = (int)value; break;
    case QM_OPT_SHORT_PCT:
        if (value != 0 && value < 100) return set_err(QM_ERR_ARG, "short-circuit per cent must be 0 (no limit) or >= 100");
        st->short_mult = value == 0 ? 1e9 : (double)value / 100.0; break;
    case QM_OPT_DEFER_PASS:
        st->defer_pass = value != 0; break;
    case QM_OPT_ZFOLD:
        st->zfold = value != 0; break;
    case QM_OPT_JIT:
        if (value < 0 || value > 1000000) return set_err(QM_ERR_ARG, "jit threshold must be >= 0");
        st->jit = (int)value; break;
    default: return set_err(QM_ERR_ARG, "unknown option %d", key);
    }
    return QM_OK;
}

extern "C" int32_t qm_jit_cache_dir(const char* path) { jit_set_cache_dir(path); return QM_OK; }
extern "C" int32_t qm_jit_cache_stats(uint64_t out[2]) {
    if (!out) return set_err(QM_ERR_ARG, "null argument");
    jit_disk_stats(out);
    return QM_OK;
}
extern "C" int32_t qm_jit_stats(uint64_t out[4], double* compile_ms) {
    if (!out) return set_err(QM_ERR_ARG, "null argument");
    const JitStats s = jit_stats();
    out[0] = s.compiled; out[1] = s.failed; out[2] = s.cached; out[3] = (uint64_t)s.backend;
    if (compile_ms) *compile_ms = s.compile_ms;
    return QM_OK;
}

extern "C" int32_t qm_num_qubits(const qm_state* st, int32_t* n, int32_t* batch) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    if (n) *n = st->n;
    if (batch) *batch = st->batch;
    return QM_OK;
}

// ---------------------------------------------------------------------------------------------
// gate execution
// ---------------------------------------------------------------------------------------------
int launch_pass_v4(qm_state* st, const V4Program& prog, const PlanPass& P, const double* d_slots, int nslots,
                   const PassSlab* slab, cudaStream_t stream, JitKernel* jk = nullptr, const V4Fold* d_fold = nullptr);

static int launch_1q(qm_state* st, int tbit, int cbit, const double* m) {
    // physical bits; control on a sharded bit is a rank predicate, target must be local
    uint64_t cmask = 0;
    if (cbit >= st->nl) { if (!((st->rank >> (cbit - st->nl)) & 1)) return QM_OK; }
    else if (cbit >= 0) cmask = 1ull << cbit;
    Mat2 U; memcpy(U.v, m, sizeof(U.v));
    uint64_t npairs = total_amps(st) >> 1;
    int grid = grid_for(npairs, 256, st->sms, 16);
    k_apply_1q<<<grid, 256, 0, st->stream>>>(st->amp, npairs, tbit, cmask, U);
    CU(cudaGetLastError());
    st->ctr.launches++;
    st->ctr.bytes_hbm += (cbit >= 0 && cbit < st->nl ? 16ull : 32ull) * total_amps(st);
    return QM_OK;
}

static int launch_pass(qm_state* st, const unsigned char* d_pass, const PlanPass& P, const double* d_slots, int nslots) {
    const int m = P.m;
    uint64_t ntiles = total_amps(st) >> m;
    size_t smem = ((size_t)16 << m) + ((size_t)8 << P.hb.size());
    const int threads = 256;
    // the attribute belongs to the device, not to the process (qm_create takes `device`)
    static bool attr[64] = { false };
    const int dev = st->device;
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        CU(cudaFuncSetAttribute(k_fused_pass<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 8) per_sm = 8;
    uint64_t cap = (uint64_t)st->sms * per_sm;
    int grid = (int)(ntiles < cap ? ntiles : cap);
    uint64_t rank_or = st->g ? ((uint64_t)st->rank << st->nl) : 0;
    k_fused_pass<256><<<grid, threads, smem, st->stream>>>(st->amp, d_pass, d_slots, nslots, ntiles, st->nl, rank_or);
    CU(cudaGetLastError());
    st->ctr.launches++; st->ctr.passes++;
    st->ctr.bytes_hbm += 32ull * total_amps(st);
    return QM_OK;
}

// ---- v2: persistent TMA-pipelined pass -------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr; static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Compiled pass kernels for this register?  QM_OPT_JIT 0: never; 1: every pass program, at first sight (tests, tools);
// >= 2: an opcode sequence seen that many times — but only on registers whose pass is long enough for a compilation
// (~0.25 s of one host core, 16 at a time) to pay for itself within a few hundred launches: 2^26 amplitudes and more
// (a 1 GiB pass takes 0.36 ms; a 20-qubit reservoir step takes microseconds and would never earn it back).
static int jit_threshold(const qm_state* st) {
    if (st->dry_run) return 1;                 // qm_precompile_circuit: plan for the compiled kernels and compile at first sight
    if (st->jit <= 1) return st->jit;
    return total_amps(st) >= (1ull << 26) ? st->jit : 0;
}

// L2 promotion of the tile loads: 256 B unless QM_TMA_L2PROMO says otherwise (0 none, 1 64 B, 2 128 B, 3 256 B; a tuning knob of
// the measurement tools, read once)
static CUtensorMapL2promotion tma_l2_promotion() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("QM_TMA_L2PROMO"); v = (e && *e >= '0' && *e <= '3') ? (*e - '0') : 3; }
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// can this pass go through the TMA kernel?
static bool v4_eligible(const qm_state* st, const PlanPass& P) {
    if (!st->pipeline || !get_encode()) return false;
    if (P.m != 11 && P.m != 12) return false;
    if (P.L < 3 || P.L > 8) return false;
    if (P.tl.size() != (size_t)P.m || P.segs.size() > 4) return false;
    for (const auto& sg : P.segs) if (sg.second > 8) return false;
    if ((int)P.groups.size() > kV4MaxGroups) return false;
    size_t nrec = P.groups.size();
    for (const PlanGroup& G : P.groups) {
        if ((int)G.rb.size() > kV4BlockBits) return false;
        for (const LGate& x : G.subs) { if (!x.liftable) return false; nrec += 1 + (x.slot >= 0); }
    }
    if (nrec > (size_t)kV4MaxRecs) return false;
    if ((total_amps(st) >> 3) > 0x7fffffffull) return false;      // TMA coordinates are int32
    return true;
}

// slab: nullptr for a pass over the whole register; otherwise the pass covers only the amplitudes whose index bits
// slab->bit[] have the values in slab->value_or (sharded registers: the exchange of one slab runs beside the passes
// of its neighbours, dist.cu), on a grid of slab->sm_cap CTAs.
int launch_pass_v4(qm_state* st, const V4Program& prog, const PlanPass& P, const double* d_slots, int nslots,
                   const PassSlab* slab, cudaStream_t stream, JitKernel* jk, const V4Fold* d_fold) {
    // tensor map: dim0 = one 128-byte row (8 amplitudes = 16 doubles); dim1 = row index over the whole
    // array (the box takes 2^(L-3) consecutive rows at the tile's base); dims 2..4 = the runs of high
    // tile bits (size 2^len, stride 16 * 2^start bytes), unused dims have size 1.
    // dim0 = one 128-byte row; the other dims are P.segs in their shared-memory order.  The segment "bits 3..L-1"
    // is the dim that spans the whole array in rows (its coordinate selects the tile: base >> 3); the runs of high
    // bits are dims of size 2^len and stride 16 * 2^start bytes at coordinate 0.
    const uint64_t amps = total_amps(st);
    cuuint64_t gdim[5] = { 16, 1, 1, 1, 1 };
    cuuint64_t gstr[4] = { 128, 128, 128, 128 };
    cuuint32_t box[5] = { 16, 1, 1, 1, 1 };
    cuuint32_t estr[5] = { 1, 1, 1, 1, 1 };
    int rowdim = -1, nd = 1;
    for (const auto& sg : P.segs) {
        if (sg.first == 3 && sg.first < P.L) { rowdim = nd; gdim[nd] = amps >> 3; gstr[nd - 1] = 128; box[nd] = 1u << sg.second; }
        else { gdim[nd] = 1ull << sg.second; gstr[nd - 1] = 16ull << sg.first; box[nd] = 1u << sg.second; }
        ++nd;
    }
    if (rowdim < 0) {           // L == 3: no row segment; a dim of box 1 carries the tile's base row
        if (nd >= 5) return set_err(QM_ERR_STATE, "tile layout needs more than 5 TMA dimensions (internal error)");
        rowdim = nd; gdim[nd] = amps >> 3; gstr[nd - 1] = 128; box[nd] = 1; ++nd;
    }
    CUtensorMap tm;
    CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, st->amp, gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tma_l2_promotion(),
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(QM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    int rs[8], rl[8];
    const int nr = hb_runs(P.hb, rs, rl);
    V4Launch lp; memset(&lp, 0, sizeof(lp));
    lp.ntiles = amps >> P.m; lp.reg_bits = st->nl; lp.L = P.L; lp.rowdim = rowdim;
    // gaps of the tile index, ascending: the runs of high tile bits and the pinned slab bits
    std::vector<std::pair<int, int>> gaps;
    for (int i = 0; i < nr; ++i) gaps.push_back({ rs[i], rl[i] });
    if (slab) {
        if (st->batch != 1 || slab->nbits > 3) return set_err(QM_ERR_STATE, "slab launch: bad shape (internal error)");
        for (int i = 0; i < slab->nbits; ++i) {
            if (tile_local_of(P, slab->bit[i]) >= 0 || slab->bit[i] >= st->nl) return set_err(QM_ERR_STATE, "slab bit %d is a tile bit (internal error)", slab->bit[i]);
            gaps.push_back({ slab->bit[i], 1 });
        }
        lp.nfixed = slab->nbits; lp.fixed_or = slab->value_or; lp.sm_cap = slab->sm_cap;
        lp.ntiles >>= slab->nbits;
    }
    std::sort(gaps.begin(), gaps.end());
    if ((int)gaps.size() > kV4Gaps) return set_err(QM_ERR_STATE, "tile layout needs more than %d index gaps (internal error)", kV4Gaps);
    for (int i = 0; i < kV4Gaps; ++i) { lp.gap_start[i] = i < (int)gaps.size() ? gaps[i].first : 0; lp.gap_len[i] = i < (int)gaps.size() ? gaps[i].second : 0; }
    lp.rank_or = st->g ? ((uint64_t)st->rank << st->nl) : 0; lp.nslots = nslots;
    lp.trace = st->trace_buf;
    if (d_fold) {           // jk is a ZFOLD kernel (fused_tma_kernel.cuh): it reads its parameters through this pointer and never traces
        if (!jk || slab || st->batch != 1) return set_err(QM_ERR_STATE, "folded readout: bad launch (internal error)");
        lp.fold = d_fold;
    }
    // a pass program whose opcode sequence has been compiled (jit.hpp) runs as straight-line code; otherwise the interpreter
    if (jk) { QM_TRY(jit_launch(jk, stream, st->sms, P.m, tm, prog, d_slots, lp)); st->ctr.compiled_passes += (!slab || slab->index == 0); }
    else if (P.m == 11) QM_TRY(launch_v4_m11(stream, st->sms, tm, prog, d_slots, lp));
    else QM_TRY(launch_v4_m12(stream, st->sms, tm, prog, d_slots, lp));
    // the program crosses the host link with the launch: counted as host->device bytes
    st->ctr.bytes_h2d += offsetof(V4Program, recs) + (size_t)prog.nrecs * sizeof(V4Rec);
    st->ctr.launches++;
    if (!slab || slab->index == 0) st->ctr.passes++;
    st->ctr.bytes_hbm += (32ull * amps) >> (slab ? slab->nbits : 0);
    return QM_OK;
}

// run `gates` (logical qubits) now: plan, upload the pass programs, launch
static int run_circuit(qm_state* st, const qm_gate* gates, int ngates, const double* d_slots, int nslots,
                       const SlotMap* smap = nullptr, const double* d_slots_lift = nullptr) {
    if (ngates <= 0) return QM_OK;
    CU(cudaSetDevice(st->device));
    auto t0 = std::chrono::steady_clock::now();
    if (st->eager) {
        std::vector<LGate> lg;
        int rc = lower_gates(gates, ngates, st->n, st->perm.data(), false, lg);
        if (rc) return set_err(rc, "bad gate in circuit");
        for (const LGate& x : lg) {
            if (x.slot >= 0) return set_err(QM_ERR_STATE, "per-state slot gates need the fused path");
            if (x.target >= st->nl) {
                if (!x.diag) return set_err(QM_ERR_STATE, "eager mode cannot apply a non-diagonal gate to a sharded qubit");
                // diagonal on a global bit: a rank-dependent scalar on the controlled half
                return set_err(QM_ERR_STATE, "eager mode does not handle diagonal gates on sharded qubits");
            }
            QM_TRY(launch_1q(st, x.target, x.ctrl, x.m));
        }
        st->ctr.gates += ngates;
        return QM_OK;
    }
    // tile size: 2^12 amplitudes unless the caller asks for 2^11 (four teams of 128 threads, six stages): the two are
    // within 2 % of each other for single-group passes, 2^12 holds more qubits per pass for deep fusion
    const int tile_bits = st->tile_bits ? st->tile_bits : 12;
    Planner pl;
    pl.opt.n_local = st->nl; pl.opt.n_global = st->g;
    pl.opt.tile_bits = tile_bits; pl.opt.low_bits = st->low_bits; pl.opt.fuse = st->fuse;
    if (st->max_cost > 0) pl.opt.max_cost = st->max_cost < 9.0 ? 9.0 : st->max_cost;
    pl.opt.short_mult = st->short_mult;
    bool first_event = true;
    // lower once under the current logical->physical map; a remap later relabels what is left
    int rc = lower_gates(gates, ngates, st->n, st->perm.data(), st->fuse, pl.gates, smap);
    if (rc) return set_err(rc, "bad gate in circuit");
    // The TMA kernel takes the circuit when tiles of 2^11 / 2^12 amplitudes fit a register and every gate has
    // an in-place lifting form; otherwise the whole circuit is planned for the general kernel (3-bit blocks).
    bool tma = st->pipeline && get_encode() && (tile_bits == 11 || tile_bits == 12) && st->nl >= tile_bits &&
               st->low_bits >= 3 && (total_amps(st) >> 3) <= 0x7fffffffull;
    if (tma) {
        std::vector<LGate> lifted = pl.gates;
        prepare_for_lifting(lifted);
        for (const LGate& x : lifted) if (!x.liftable) { tma = false; break; }
        if (tma) pl.gates.swap(lifted);
    }
    if (tma) {
        v4_planner_options(pl.opt, st->max_cost, jit_threshold(st) > 0 ? 1 : 0);     // the planner prices gates for the kernel that will run them
        // every per-state gate takes two records (SLOT + gate); two records of slack for the encoder's END records
        if (smap) pl.opt.max_subs = (kV4MaxRecs - kV4MaxGroups) / 2 - 2;
    }
    pl.done.assign(pl.gates.size(), 0);
    // Plan cache (unsharded, fused, TMA kernel): same STRUCTURE as a recent circuit -> reuse its passes with this
    // circuit's matrices.  The key is everything the planner reads — the options and, per lowered gate, target, control,
    // diagonal / per-state-slot flag and cost class — folded into two independent 64-bit hashes (a deep circuit has
    // thousands of gates; the gate count is compared too).
    const bool cacheable = tma && st->g == 0 && st->fuse;
    uint64_t key[3] = { 0, 0, 0 };
    qm_state::CachedPlan* hit = nullptr;
    std::vector<std::pair<PlanPass, std::vector<int>>> fresh;
    if (cacheable) {
        uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x9e3779b97f4a7c15ull;
        auto mix = [&](uint64_t v) {
            h1 = (h1 ^ v) * 0x100000001b3ull;
            h2 ^= v + 0x9e3779b97f4a7c15ull + (h2 << 6) + (h2 >> 2); h2 *= 0xff51afd7ed558ccdull; h2 ^= h2 >> 33;
        };
        mix((uint64_t)st->nl); mix((uint64_t)tile_bits); mix((uint64_t)st->low_bits); mix((uint64_t)(pl.opt.max_cost * 16));
        mix((uint64_t)(pl.opt.short_mult > 1e8 ? -1 : pl.opt.short_mult * 100)); mix((uint64_t)pl.opt.max_subs); mix((uint64_t)pl.opt.cost_model);
        for (const LGate& x : pl.gates)
            mix((uint64_t)x.target | ((uint64_t)(x.ctrl + 1) << 8) | ((uint64_t)x.diag << 16) | ((uint64_t)(x.slot >= 0) << 17) |
                ((uint64_t)(gate_cost(x, pl.opt.cost_model) * 16.0) << 20));
        key[0] = h1; key[1] = h2; key[2] = pl.gates.size();
        for (auto& c : st->plan_cache) if (c.key[0] == key[0] && c.key[1] == key[1] && c.key[2] == key[2]) { hit = &c; break; }
    }
    size_t replay_i = 0;
    // the next pass: planned now, or replayed from the cache with this circuit's matrices
    auto plan_one = [&](PlanPass& out) -> bool {
        if (!hit) {
            if (!pl.next_pass(out)) return false;
            if (tma) choose_tile_layout(out);
            if (cacheable) fresh.emplace_back(out, pl.last_taken);
            return true;
        }
        if (replay_i >= hit->passes.size()) return false;
        out = hit->passes[replay_i].first;
        const std::vector<int>& idx = hit->passes[replay_i].second;
        size_t j = 0;
        for (PlanGroup& G : out.groups) for (LGate& x : G.subs) x = pl.gates[idx[j++]];
        for (int i : idx) pl.done[i] = 1;
        ++replay_i;
        return true;
    };
    // A planned pass ready to launch: the TMA kernel's program (a launch parameter), the general kernel's blob, or —
    // when the encoder refuses a pass the planner built for the TMA kernel — its gates one by one (kind 2)
    struct Item {
        PlanPass P; int kind = 0; V4Program prog; std::vector<unsigned char> blob; JitKernel* jk = nullptr;
        std::vector<int> taken;                  // indices (into pl.gates) of the pass's gates, when it was planned here (not replayed)
        double gf[2] = { 1.0, 0.0 };             // the global-phase factor the pass contributed (D1 form), so that a cancelled pass can take it back
        bool fold = false;                       // jk is the ZFOLD flavour of the pass's kernel: the launch also takes the all-qubit <Z> sums
    };
    // Compiled pass kernels (jit.hpp).  A replayed plan knows all its passes in advance: they are encoded once up front and the
    // opcode sequences that have been seen often enough are compiled together on several host threads (a deep circuit has
    // hundreds of passes, ~0.25 s of PTX compilation each); otherwise a pass asks for its kernel when it is prepared.
    std::vector<JitKernel*> pre_jit;
    // qm_precompile_circuit, first walk (dry_run 1): only the plan is wanted — it goes into the plan cache and the second walk
    // (dry_run 2) finds it there and compiles all its kernels together on several host threads, not one pass at a time
    const int jit_thr = st->dry_run == 1 ? 0 : jit_threshold(st);
    if (hit && tma && jit_thr > 0) {
        std::vector<V4Program> progs(hit->passes.size());
        std::vector<const V4Program*> ptrs; std::vector<size_t> which;
        for (size_t r = 0; r < hit->passes.size(); ++r) {
            PlanPass P = hit->passes[r].first;
            const std::vector<int>& idx = hit->passes[r].second;
            size_t j = 0;
            for (PlanGroup& G : P.groups) for (LGate& x : G.subs) x = pl.gates[idx[j++]];
            double scratch_phase[2] = { 1.0, 0.0 };        // this encoding only names the kernels; the phase is counted in prepare()
            if (v4_eligible(st, P) && P.m == 12 && encode_v4(P, progs[r], scratch_phase, true)) { ptrs.push_back(&progs[r]); which.push_back(r); }
        }
        std::vector<JitKernel*> got;
        jit_get_many(st->device, 12, ptrs, jit_thr, got);
        pre_jit.assign(hit->passes.size(), nullptr);
        for (size_t i = 0; i < which.size(); ++i) pre_jit[which[i]] = got[i];
    }
    auto prepare = [&](Item& it) -> int {
        it.kind = 1; it.blob.clear(); it.jk = nullptr;
        if (tma) {
            const bool eligible = v4_eligible(st, it.P);
            // compiled kernel first: the program in canonical-sign form (what the kernel's opcode sequence is keyed by);
            // only when there is no kernel (yet) is the pass encoded for the interpreter
            auto commit_phase = [&](const double f[2]) {          // global phase of the pass's diagonal gates (D1 form)
                it.gf[0] = f[0]; it.gf[1] = f[1];
                const double r = st->gphase[0] * f[0] - st->gphase[1] * f[1], i = st->gphase[0] * f[1] + st->gphase[1] * f[0];
                st->gphase[0] = r; st->gphase[1] = i;
            };
            it.gf[0] = 1.0; it.gf[1] = 0.0;
            if (eligible && jit_thr > 0 && it.P.m == 12) {
                double f[2] = { 1.0, 0.0 };
                if (encode_v4(it.P, it.prog, f, true)) {
                    if (!pre_jit.empty()) it.jk = replay_i >= 1 && replay_i <= pre_jit.size() ? pre_jit[replay_i - 1] : nullptr;
                    else it.jk = jit_get(st->device, it.P.m, it.prog, jit_thr);
                    if (it.jk) { it.kind = 0; commit_phase(f); return QM_OK; }
                }
            }
            double f[2] = { 1.0, 0.0 };
            if (eligible && encode_v4(it.P, it.prog, f)) {
                it.kind = 0;
                commit_phase(f);
                return QM_OK;
            }
            // The planner is supposed to respect the kernel's limits; if a pass slips through, it is still applied
            // (one unfused launch per gate) rather than dropped with the queue already swapped out.
            for (const PlanGroup& G : it.P.groups)
                for (const LGate& x : G.subs)
                    if (x.slot >= 0 || x.target >= st->nl || (x.ctrl >= st->nl && x.diag))
                        return set_err(QM_ERR_STATE, "planner produced a pass the TMA kernel cannot run (internal error)");
            it.kind = 2;
            return QM_OK;
        }
        encode_pass(it.P, it.blob);
        return QM_OK;
    };
    auto launch_full = [&](Item& it) -> int {
        if (st->dry_run) return QM_OK;            // qm_precompile_circuit: plan, encode, compile — launch nothing
        if (first_event) { CU(cudaEventRecord(st->ev0, st->stream)); first_event = false; }
        cudaEvent_t ea = nullptr, eb = nullptr;
        if (st->time_passes) { CU(cudaEventCreate(&ea)); CU(cudaEventCreate(&eb)); CU(cudaEventRecord(ea, st->stream)); }
        if (it.kind == 0 && it.fold) {
            // V4Fold | per-CTA partials | pairs [nl][2] | total — the grid is the one jit_launch will use
            const uint64_t ntiles = total_amps(st) >> it.P.m;
            const int grid = (int)(ntiles < (uint64_t)st->sms ? ntiles : (uint64_t)st->sms);
            const size_t head = 256, part = (size_t)grid * kV4FoldVals * 8, outb = ((size_t)st->nl * 2 + 1) * 8;
            QM_TRY(ENSURE_DEV(st, d_zfold, head + part + outb));
            QM_TRY(ENSURE_PIN(st, h_zfold, head));
            V4Fold hf; memset(&hf, 0, sizeof(hf));
            hf.thr = st->zfold_thr;
            hf.part = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(st->d_zfold) + head);
            for (int k = 0; k < it.P.m && k < 16; ++k) hf.tl[k] = (uint32_t)it.P.tl[k];
            static_assert(sizeof(V4Fold) <= 256, "V4Fold block");
            memcpy(st->h_zfold, &hf, sizeof(hf));
            CU(cudaMemcpyAsync(st->d_zfold, st->h_zfold, sizeof(hf), cudaMemcpyHostToDevice, st->stream));
            st->ctr.bytes_h2d += sizeof(hf);
            QM_TRY(launch_pass_v4(st, it.prog, it.P, d_slots_lift, nslots, nullptr, st->stream, it.jk,
                                  reinterpret_cast<const V4Fold*>(st->d_zfold)));
            double* d_out = hf.part + (size_t)grid * kV4FoldVals;
            k_zfold_final<<<1, 64, 0, st->stream>>>(hf.part, grid, st->nl, d_out, d_out + (size_t)st->nl * 2);
            CU(cudaGetLastError());
            st->ctr.launches++;
            st->zfold_done = true;
        }
        else if (it.kind == 0) QM_TRY(launch_pass_v4(st, it.prog, it.P, d_slots_lift, nslots, nullptr, st->stream, it.jk));
        else if (it.kind == 2) {
            for (const PlanGroup& G : it.P.groups) for (const LGate& x : G.subs) QM_TRY(launch_1q(st, x.target, x.ctrl, x.m));
            st->ctr.passes++;
        } else {
            QM_TRY(ENSURE_DEV(st, d_blob, it.blob.size()));
            CU(cudaEventSynchronize(st->ev_stage));      // pinned staging buffer free again?
            QM_TRY(ENSURE_PIN(st, h_blob, it.blob.size()));
            memcpy(st->h_blob, it.blob.data(), it.blob.size());
            CU(cudaMemcpyAsync(st->d_blob, st->h_blob, it.blob.size(), cudaMemcpyHostToDevice, st->stream));
            st->ctr.bytes_h2d += it.blob.size();
            CU(cudaEventRecord(st->ev_stage, st->stream));
            QM_TRY(launch_pass(st, st->d_blob, it.P, d_slots, nslots));
        }
        if (st->time_passes) {
            CU(cudaEventRecord(eb, st->stream));
            st->pass_events.push_back(ea); st->pass_events.push_back(eb);
            st->pass_desc.push_back((float)it.P.groups.size());
            st->pass_desc.push_back((float)it.P.ngates);
            st->pass_desc.push_back((float)it.P.cost);
            // records by kind: full-cost FP64 (uncontrolled or context-controlled), half-cost (control in the
            // register block), xor exchanges, thread-constant phases
            float kinds[4] = { 0, 0, 0, 0 };
            if (it.kind == 0) {
                const V4Program& pr = it.prog;
                for (int r = 0; r < pr.nrecs; ++r) {
                    uint32_t op = pr.recs[r].op;
                    if (op >= (uint32_t)V4_NUM_GATE_OPS) { if (op >= (uint32_t)V4_OP_SLOT) continue; op -= V4_NUM_GATE_OPS; }
                    int fam, var, B, CB;
                    if (!v4_decode_op(op, fam, var, B, CB)) continue;
                    if (fam == 3) kinds[2] += 1; else if (fam == 4) kinds[3] += 1; else if (CB >= 0) kinds[1] += 1; else kinds[0] += 1;
                }
            }
            for (int q = 0; q < 4; ++q) st->pass_desc.push_back(kinds[q]);
        }
        return QM_OK;
    };
    // Sharded registers: index bits that pin a slab — below the half-chunk bit of the exchange (so that the kernel's
    // lower-rank / higher-rank halves and the exchanged bits stay whole), at or above bit 12 (whole 64 KiB segments),
    // and outside the tiles of the pass before and the pass after the exchange.  Highest first: long contiguous runs.
    auto tile_has = [](const PlanPass& P, int b) { return b < P.L || std::find(P.hb.begin(), P.hb.end(), b) != P.hb.end(); };
    auto choose_slab_bits = [&](const std::vector<const PlanPass*>& tiles, std::vector<int>& bits) -> bool {
        bits.clear();
        const int want = st->slab_bits < 1 ? 1 : (st->slab_bits > 3 ? 3 : st->slab_bits);
        for (int b = st->nl - st->g - 2; b >= 12 && (int)bits.size() < want; --b) {
            bool free_ = true;
            for (const PlanPass* T : tiles) free_ = free_ && !tile_has(*T, b);
            if (free_) bits.push_back(b);
        }
        std::sort(bits.begin(), bits.end());
        // a pass over a slab must still find enough tiles for its grid; one slab bit less than asked for is accepted
        return (int)bits.size() >= (want > 1 ? want - 1 : 1);
    };

    // Passes are launched as they are planned, ONE pass behind the planner: while the GPU runs pass i the host plans
    // pass i + 1 (a 30-qubit depth-200 circuit takes ~70 ms to plan against ~3.6 s of kernels — but all of it used to
    // sit in front of the first launch).  The look-ahead also tells a sharded register which pass is the last before
    // an exchange: that pass and the first one after it are cut into slabs and run beside the exchange of the
    // neighbouring slabs (dist.cu).
    // `held`: the planned passes not launched yet, oldest first.  One for an ordinary register; a sharded register that can
    // overlap keeps the last st->overlap_depth passes back, because ALL of them — not only the very last — can run slab by
    // slab in front of the exchange (pass p + 1 on a slab only needs pass p on the same slab when the slab bits lie outside
    // both tiles), which lets the exchange of slab 0 start one or two passes earlier.
    const size_t hold_max = (st->g > 0 && tma && dist_can_overlap(st)) ? (size_t)(st->overlap_depth < 1 ? 1 : st->overlap_depth) : 1;
    std::deque<Item> held;
    Item nxt;
    bool seg_had_pass = false;
    auto launch_all_held = [&]() -> int {
        while (!held.empty()) { QM_TRY(launch_full(held.front())); held.pop_front(); }
        return QM_OK;
    };
    bool finished = false;
    for (int guard = 0; guard < 8 * ngates + 64; ++guard) {
        PlanPass P;
        if (plan_one(P)) {
            nxt.P = P;
            nxt.taken.clear(); if (!hit) nxt.taken = pl.last_taken;
            QM_TRY(prepare(nxt));
            held.push_back(nxt); seg_had_pass = true;
            if (held.size() > hold_max) { QM_TRY(launch_full(held.front())); held.pop_front(); }
            continue;
        }
        bool leftover = false;
        for (size_t i = pl.first; i < pl.gates.size(); ++i) if (!pl.done[i]) { leftover = true; break; }
        if (!leftover) {
            // qm_apply_circuit_expect_z: the very last pass also takes the <Z> sums when it is a compiled kernel over one
            // whole register and the ZFOLD flavour of that kernel is there (same threshold rule as every compiled kernel)
            // (qm_precompile_circuit compiles that flavour of the last pass as well: the caller may be about to use either call)
            if ((st->zfold_want || st->dry_run == 2) && st->zfold && tma && st->batch == 1 && st->nl < 63 && !held.empty()) {
                Item& last = held.back();
                if (last.kind == 0 && last.jk && last.P.m == 12 && last.P.tl.size() == 12) {
                    JitKernel* jz = jit_get(st->device, 12, last.prog, jit_thr, true);
                    if (jz && !st->dry_run) { last.jk = jz; last.fold = true; }
                }
            }
            QM_TRY(launch_all_held()); finished = true; break;
        }
        if (st->g == 0) return set_err(QM_ERR_STATE, "planner could not place a gate (internal error)");
        // sharded register: what is left needs a non-diagonal target on a global qubit.  Exchange the
        // global qubits with the top local ones (all-to-all over NVLink) and relabel the pending gates.
        if (!seg_had_pass && !pl.pending_global_target()) return set_err(QM_ERR_STATE, "planner stalled (internal error)");
        seg_had_pass = false;
        // the last pass planned before the exchange is postponed to after it when that saves a pass (planner.hpp try_defer_last_pass)
        if (st->g > 0 && st->defer_pass && !held.empty() && try_defer_last_pass(pl, held.back().taken, st->nl, st->g)) {
            // the pass had already put its global phase in: take it back (unit modulus: divide = multiply by the conjugate)
            const Item& last = held.back();
            const double r = st->gphase[0] * last.gf[0] + st->gphase[1] * last.gf[1], im = st->gphase[1] * last.gf[0] - st->gphase[0] * last.gf[1];
            st->gphase[0] = r; st->gphase[1] = im;
            held.pop_back();
            st->ctr.deferred_passes++;
        }
        bool ov = !held.empty() && tma && dist_can_overlap(st);
        for (const Item& it : held) ov = ov && it.kind == 0;
        if (!ov) {
            QM_TRY(launch_all_held());
            QM_TRY(dist_remap(st, &pl.gates, &pl.done));
            continue;
        }
        // overlapped exchange: relabel first (host only), plan the first pass of the new layout, pick the slab bits —
        // outside the tiles of every held pass and of the new one; if there are too few such bits, the oldest held pass
        // is launched whole and the choice is made again with one pass less
        dist_relabel(st, &pl.gates, &pl.done);
        PlanPass P2;
        const bool got2 = plan_one(P2);
        if (got2) { nxt.P = P2; nxt.taken = pl.last_taken; QM_TRY(prepare(nxt)); seg_had_pass = true; }
        std::vector<int> sb;
        bool ok = got2 && nxt.kind == 0;
        while (ok) {
            std::vector<const PlanPass*> tiles; for (const Item& it : held) tiles.push_back(&it.P);
            tiles.push_back(&nxt.P);
            if (choose_slab_bits(tiles, sb)) break;
            if (held.size() <= 1) { ok = false; break; }
            QM_TRY(launch_full(held.front())); held.pop_front();
        }
        if (!ok) {
            QM_TRY(launch_all_held());
            QM_TRY(dist_exchange_now(st));
            if (got2) held.push_back(nxt);
            continue;
        }
        if (first_event) { CU(cudaEventRecord(st->ev0, st->stream)); first_event = false; }
        uint32_t ep_pass = 0, ep_exch = 0;
        dist_next_epochs(st, &ep_pass, &ep_exch);
        const int ns = 1 << sb.size();
        std::vector<PassSlab> slabs(ns);
        for (int h = 0; h < ns; ++h) {
            PassSlab& S = slabs[h];
            S.nbits = (int)sb.size(); S.index = h; S.value_or = 0;
            S.sm_cap = st->sms - st->xsm > 8 ? st->sms - st->xsm : st->sms;
            for (size_t i = 0; i < sb.size(); ++i) { S.bit[i] = sb[i]; if ((h >> i) & 1) S.value_or |= 1ull << sb[i]; }
        }
        for (int h = 0; h < ns; ++h) {
            for (Item& it : held) QM_TRY(launch_pass_v4(st, it.prog, it.P, d_slots_lift, nslots, &slabs[h], st->stream, it.jk));
            QM_TRY(dist_slab_pass_done(st, slabs[h], ep_pass));
        }
        for (int h = 0; h < ns; ++h) QM_TRY(dist_slab_exchange(st, slabs[h], ep_pass, ep_exch, h == 0, h == ns - 1));
        for (int h = 0; h < ns; ++h) {
            QM_TRY(dist_slab_wait_exchanged(st, slabs[h], ep_exch));
            QM_TRY(launch_pass_v4(st, nxt.prog, nxt.P, d_slots_lift, nslots, &slabs[h], st->stream, nxt.jk));
        }
        st->ctr.overlapped_remaps++;
        st->ctr.overlapped_passes += held.size() + 1;
        held.clear();
    }
    if (!finished) return set_err(QM_ERR_STATE, "planner did not terminate (internal error)");
    if (cacheable && !hit) {
        qm_state::CachedPlan c; c.key[0] = key[0]; c.key[1] = key[1]; c.key[2] = key[2]; c.passes.swap(fresh);
        st->plan_cache.insert(st->plan_cache.begin(), std::move(c));
        if (st->plan_cache.size() > 8) st->plan_cache.pop_back();
    }
    if (!first_event) { CU(cudaEventRecord(st->ev1, st->stream)); st->ev_pending = true; }
    if (!st->dry_run) st->ctr.gates += ngates;
    st->ctr.ms_plan = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return QM_OK;
}

static int flush(qm_state* st) {
    if (st->queue.empty()) return QM_OK;
    std::vector<qm_gate> q; q.swap(st->queue);
    return run_circuit(st, q.data(), (int)q.size(), nullptr, 0);
}

static int enqueue(qm_state* st, const qm_gate& g) {
    st->queue.push_back(g);
    if (st->eager || st->queue.size() >= 65536) return flush(st);
    return QM_OK;
}

static int check_bit(const qm_state* st, int b, const char* what) {
    if (b < 0 || b >= st->n) return set_err(QM_ERR_ARG, "Qubit index out of range (%s bit %d, n = %d)", what, b, st->n);
    return QM_OK;
}

extern "C" int32_t qm_apply_u2(qm_state* st, int32_t t0, const double U[8]) {
    if (!st || !U) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(check_bit(st, t0, "target"));
    qm_gate g; g.kind = QM_GATE_U2; g.q0 = t0; g.q1 = 0; g.flags = 0; memcpy(g.m, U, sizeof(g.m));
    return enqueue(st, g);
}

extern "C" int32_t qm_apply_controlled(qm_state* st, int32_t c0, int32_t t0, const double V[8]) {
    if (!st || !V) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(check_bit(st, c0, "control")); QM_TRY(check_bit(st, t0, "target"));
    if (c0 == t0) return set_err(QM_ERR_ARG, "Control and target cannot be the same qubit");
    qm_gate g; g.kind = QM_GATE_CU2; g.q0 = c0; g.q1 = t0; g.flags = 0; memcpy(g.m, V, sizeof(g.m));
    return enqueue(st, g);
}

extern "C" int32_t qm_apply_swap(qm_state* st, int32_t a0, int32_t b0) {
    if (!st) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(check_bit(st, a0, "first")); QM_TRY(check_bit(st, b0, "second"));
    if (a0 == b0) return QM_OK;
    qm_gate g; memset(&g, 0, sizeof(g)); g.kind = QM_GATE_SWAP; g.q0 = a0; g.q1 = b0;
    return enqueue(st, g);
}

static int validate_gates(const qm_state* st, const qm_gate* gates, int ng, int nslots) {
    for (int i = 0; i < ng; ++i) {
        const qm_gate& g = gates[i];
        switch (g.kind) {
        case QM_GATE_U2: QM_TRY(check_bit(st, g.q0, "target")); break;
        case QM_GATE_U2_SLOT:
            QM_TRY(check_bit(st, g.q0, "target"));
            if (g.q1 < 0 || g.q1 >= nslots) return set_err(QM_ERR_ARG, "gate %d: slot %d of %d", i, g.q1, nslots);
            break;
        case QM_GATE_CU2:
            QM_TRY(check_bit(st, g.q0, "control")); QM_TRY(check_bit(st, g.q1, "target"));
            if (g.q0 == g.q1) return set_err(QM_ERR_ARG, "Control and target cannot be the same qubit (gate %d)", i);
            break;
        case QM_GATE_SWAP: QM_TRY(check_bit(st, g.q0, "first")); QM_TRY(check_bit(st, g.q1, "second")); break;
        default: return set_err(QM_ERR_ARG, "gate %d: unknown kind %d", i, g.kind);
        }
    }
    return QM_OK;
}

// Plans `gates` exactly as qm_apply_circuit would, compiles the pass kernels it will need and keeps the plan — without touching
// the register.  A caller that knows its circuit ahead of time (a reservoir loop, a benchmark) calls this once, e.g. while data
// is still loading, and the FIRST application already runs compiled instead of going through the interpreter kernel.
// Does nothing when compiled kernels are not in use for this register (QM_OPT_JIT 0, or the default on a small register).
extern "C" int32_t qm_precompile_circuit(qm_state* st, const qm_gate* gates, int32_t ngates, int32_t* n_compiled) {
    if (!st || (!gates && ngates > 0) || ngates < 0) return set_err(QM_ERR_ARG, "null argument");
    if (n_compiled) *n_compiled = 0;
    if (st->dist || st->eager) return set_err(QM_ERR_STATE, "precompilation is for unsharded registers on the fused path");
    QM_TRY(validate_gates(st, gates, ngates, 0));
    QM_TRY(flush(st));
    if (jit_threshold(st) == 0) return QM_OK;
    const uint64_t before = jit_stats().compiled;
    const double g0 = st->gphase[0], g1 = st->gphase[1];
    st->dry_run = 1;
    int rc = run_circuit(st, gates, ngates, nullptr, 0);                 // plan -> plan cache
    st->dry_run = 2;
    if (rc == QM_OK) rc = run_circuit(st, gates, ngates, nullptr, 0);    // replay the cached plan: compile what is missing
    st->dry_run = 0;
    st->gphase[0] = g0; st->gphase[1] = g1;          // nothing was applied: the encodings' global-phase factors are not owed
    if (n_compiled) *n_compiled = (int32_t)(jit_stats().compiled - before);
    return rc;
}

static int zfold_result(qm_state* st, double** h_pairs, double** h_tot) {
    // the sums the last pass left in d_zfold -> pinned host memory (same layout as qm_zstats: [nl][2] pairs, then the total)
    const size_t out_doubles = (size_t)st->nl * 2 + 1;
    const uint64_t ntiles = total_amps(st) >> 12;
    const int grid = (int)(ntiles < (uint64_t)st->sms ? ntiles : (uint64_t)st->sms);
    const double* d_out = reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(st->d_zfold) + 256) + (size_t)grid * kV4FoldVals;
    QM_TRY(ENSURE_PIN(st, h_scratch, out_doubles * 8));
    CU(cudaMemcpyAsync(st->h_scratch, d_out, out_doubles * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += out_doubles * 8;
    QM_TRY(qm_sync(st));
    *h_pairs = st->h_scratch; *h_tot = st->h_scratch + (size_t)st->nl * 2;
    return QM_OK;
}

// qm_apply_circuit + qm_expect_z_all in one call; the readout is folded into the last pass when it can be (QM_OPT_ZFOLD)
extern "C" int32_t qm_apply_circuit_expect_z(qm_state* st, const qm_gate* gates, int32_t ngates, double threshold, double* out_host) {
    if (!st || !out_host || (!gates && ngates > 0) || ngates < 0) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(validate_gates(st, gates, ngates, 0));
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    st->zfold_want = true; st->zfold_done = false; st->zfold_thr = threshold;
    const int rc = run_circuit(st, gates, ngates, nullptr, 0);
    st->zfold_want = false;
    if (rc != QM_OK) { st->zfold_done = false; return rc; }
    double *pairs, *tot;
    if (st->zfold_done) { st->zfold_done = false; st->ctr.folded_readouts++; QM_TRY(zfold_result(st, &pairs, &tot)); }
    else QM_TRY(qm_zstats(st, threshold, &pairs, &tot));
    if (st->dist) return dist_finish_z(st, pairs, tot, out_host);
    memcpy(out_host, pairs, (size_t)st->batch * st->nl * 2 * sizeof(double));
    return QM_OK;
}

extern "C" int32_t qm_apply_circuit(qm_state* st, const qm_gate* gates, int32_t ngates) {
    if (!st || (!gates && ngates > 0) || ngates < 0) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(validate_gates(st, gates, ngates, 0));
    QM_TRY(flush(st));
    return run_circuit(st, gates, ngates, nullptr, 0);
}

extern "C" int32_t qm_apply_batched(qm_state* st, const qm_gate* gates, int32_t ngates, const double* per_state_mats,
                                    int32_t nslots) {
    if (!st || (!gates && ngates > 0) || ngates < 0 || nslots < 0) return set_err(QM_ERR_ARG, "null argument");
    if (nslots > 0 && !per_state_mats) return set_err(QM_ERR_ARG, "per_state_mats is null");
    if (st->dist) return set_err(QM_ERR_STATE, "sharded registers are not batched");
    QM_TRY(validate_gates(st, gates, ngates, nslots));
    QM_TRY(flush(st));
    if (st->eager) return set_err(QM_ERR_STATE, "batched application needs the fused path");
    CU(cudaSetDevice(st->device));
    if (nslots == 0) return run_circuit(st, gates, ngates, nullptr, 0);
    // Per-state matrices -> (sub)slots.  The v2 kernel's sign variants are static, so a slot stays ONE
    // subslot only when every state's matrix has the same lifting kind with plain signs (ry(theta),
    // 0 <= theta < pi, is the reservoir-encoding case).  Otherwise every state's matrix is sent as
    //   sqrt(R) sqrt(R) * Y * sqrt(L) sqrt(L)      (R, L unit-modulus diagonals, Y = [c -s; s c], c, s >= 0)
    // five subslots whose lifts all have plain signs for any angle.  Both forms are uploaded:
    // matrices for the v1 kernel, lifting coefficients for v2.
    const int B = st->batch;
    SlotMap sm; sm.first.resize(nslots); sm.count.resize(nslots);
    std::vector<int> kinds(nslots);
    auto plain_kind = [](const double* m) {
        LiftCoef L; int k = lift_classify(m, L);
        if (k == LK_PERMX) return (int)LK_NONE;
        if (k != LK_NONE && (L.mk[0] | L.mk[1] | L.mk[2] | L.mk[3])) return (int)LK_NONE;
        return k;
    };
    // The host side of a batched step is B x nslots small classifications (65,536 at 4096 x 16): spread over host
    // threads, states in contiguous blocks.  It used to be a single loop and, with the synchronous pageable upload
    // after it, 46 % of a C2 step.
    int nthreads = (int)std::thread::hardware_concurrency(); if (nthreads < 1) nthreads = 1; if (nthreads > 16) nthreads = 16;
    if ((long long)B * nslots < 4096) nthreads = 1;
    auto parallel_states = [&](const std::function<void(int, int, int)>& fn) {        // fn(thread, b0, b1)
        if (nthreads == 1) { fn(0, 0, B); return; }
        std::vector<std::thread> th;
        const int per = (B + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            const int b0 = t * per, b1 = b0 + per < B ? b0 + per : B;
            if (b0 >= b1) break;
            th.emplace_back(fn, t, b0, b1);
        }
        for (auto& x : th) x.join();
    };
    {   // slot kinds: one lifting kind with plain signs for every state, or the general five-piece form
        std::vector<std::vector<int>> part(nthreads, std::vector<int>(nslots, -2));       // -2 = no state seen
        parallel_states([&](int t, int b0, int b1) {
            std::vector<int>& k = part[t];
            for (int b = b0; b < b1; ++b)
                for (int s = 0; s < nslots; ++s) {
                    if (k[s] == LK_NONE) continue;
                    const int kk = plain_kind(per_state_mats + ((size_t)b * nslots + s) * 8);
                    k[s] = (k[s] == -2 || k[s] == kk) ? kk : (int)LK_NONE;
                }
        });
        for (int s = 0; s < nslots; ++s) {
            int k0 = -2;
            for (int t = 0; t < nthreads; ++t) { const int k = part[t][s]; if (k == -2) continue; k0 = (k0 == -2 || k0 == k) ? k : (int)LK_NONE; }
            if (k0 == -2) k0 = LK_NONE;
            kinds[s] = k0;
            sm.first[s] = (int)sm.kind.size(); sm.count[s] = (k0 == LK_NONE) ? 5 : 1;
            if (k0 == LK_NONE) { const int ks[5] = { LK_PHASE, LK_PHASE, LK_ROT, LK_PHASE, LK_PHASE }; sm.kind.insert(sm.kind.end(), ks, ks + 5); }
            else sm.kind.push_back(k0);
        }
    }
    const int nsub = (int)sm.kind.size();
    const size_t ndoubles = (size_t)B * nsub * 8, bytes = ndoubles * sizeof(double);
    // pinned staging, two halves used alternately (a half is reused once its upload has finished): [mats | lifts] each
    // (an outgrown staging buffer is only retired, never freed under a copy in flight: ensure_pinned).
    // No stream synchronisation: the call returns as soon as the step is enqueued, so the caller's preparation of the
    // next step overlaps this step's kernels.
    const int half = st->slots_half; st->slots_half ^= 1;
    if (!st->ev_slots[0]) { CU(cudaEventCreateWithFlags(&st->ev_slots[0], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&st->ev_slots[1], cudaEventDisableTiming)); }
    QM_TRY(ENSURE_PIN(st, h_slots, 4 * bytes));
    CU(cudaEventSynchronize(st->ev_slots[half]));
    double* mats = st->h_slots + (size_t)half * 2 * ndoubles;
    double* lifts = mats + ndoubles;
    memset(lifts, 0, bytes);
    static const double I8[8] = { 1, 0, 0, 0, 0, 0, 1, 0 };
    std::vector<char> thread_ok(nthreads, 1);
    parallel_states([&](int t, int b0, int b1) {
        bool ok = true;
        for (int b = b0; b < b1; ++b)
            for (int s = 0; s < nslots; ++s) {
                const double* m = per_state_mats + ((size_t)b * nslots + s) * 8;
                double* mo = &mats[((size_t)b * nsub + sm.first[s]) * 8];
                double* lo = &lifts[((size_t)b * nsub + sm.first[s]) * 8];
                if (kinds[s] != LK_NONE) {
                    memcpy(mo, m, 64);
                    LiftCoef L; lift_classify(m, L); memcpy(lo, &L, 64);
                    continue;
                }
                double f[3][8];
                if (!decompose_unitary(m, f)) {
                    ok = false;      // v1 only: the first subslot carries the matrix, the rest are identities
                    memcpy(mo, m, 64); for (int i = 1; i < 5; ++i) memcpy(mo + 8 * i, I8, 64);
                    continue;
                }
                double piece[5][8];
                for (int side = 0; side < 2; ++side) {          // f[0] = right phase, f[2] = left phase
                    const double* d = f[side == 0 ? 0 : 2];
                    double h[8] = { 1, 0, 0, 0, 0, 0, 1, 0 };
                    unit_sqrt(d[0], d[1], h[0], h[1]); unit_sqrt(d[6], d[7], h[6], h[7]);
                    memcpy(piece[side == 0 ? 0 : 3], h, 64); memcpy(piece[side == 0 ? 1 : 4], h, 64);
                }
                memcpy(piece[2], f[1], 64);
                for (int i = 0; i < 5; ++i) {
                    memcpy(mo + 8 * i, piece[i], 64);
                    LiftCoef L; const int k = lift_classify(piece[i], L);
                    // an identity piece classifies as PHASE with zero shears; a rotation by exactly 0 is diagonal -> PHASE too
                    if (k != sm.kind[sm.first[s] + i] || (L.mk[0] | L.mk[1] | L.mk[2] | L.mk[3])) {
                        if (i == 2 && k == LK_PHASE && !(L.mk[0] | L.mk[1] | L.mk[2] | L.mk[3])) { memset(&L, 0, sizeof(L)); }  // identity rotation: zero shears
                        else ok = false;
                    }
                    memcpy(lo + 8 * i, &L, 64);
                }
            }
        thread_ok[t] = ok ? 1 : 0;
    });
    bool all_liftable = true;
    for (char c : thread_ok) all_liftable = all_liftable && c;
    if (!all_liftable) for (size_t i = 0; i < sm.kind.size(); ++i) sm.kind[i] = LK_NONE;
    QM_TRY(ENSURE_DEV(st, d_slots, 2 * bytes));
    // the TMA kernel reads only the lifting coefficients, the general kernel only the matrices: upload what will be read
    const int tb = st->tile_bits ? st->tile_bits : 12;
    const bool tma_path = all_liftable && st->pipeline && get_encode() && (tb == 11 || tb == 12) && st->nl >= tb && st->low_bits >= 3 &&
                          st->fuse && (total_amps(st) >> 3) <= 0x7fffffffull;
    if (!tma_path) {
        CU(cudaMemcpyAsync(st->d_slots, mats, bytes, cudaMemcpyHostToDevice, st->stream));
        st->ctr.bytes_h2d += bytes;
    }
    if (all_liftable) {
        CU(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(st->d_slots) + bytes, lifts, bytes, cudaMemcpyHostToDevice, st->stream));
        st->ctr.bytes_h2d += bytes;
    }
    CU(cudaEventRecord(st->ev_slots[half], st->stream));
    const int saved = st->pipeline;
    if (!all_liftable) st->pipeline = 0;
    int rc = run_circuit(st, gates, ngates, st->d_slots, nsub, &sm, st->d_slots + ndoubles);
    st->pipeline = saved;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// state transfer
// ---------------------------------------------------------------------------------------------
int materialize_phase(qm_state* st) {
    if (st->gphase[0] == 1.0 && st->gphase[1] == 0.0) return QM_OK;
    CU(cudaSetDevice(st->device));
    double pr = st->gphase[0], pi = st->gphase[1];
    const double h = hypot(pr, pi);            // a product of unit-modulus factors: keep it on the unit circle
    if (h > 0) { pr /= h; pi /= h; }
    const uint64_t N = total_amps(st);
    k_scale_phase<<<grid_for(N, 256, st->sms, 16), 256, 0, st->stream>>>(st->amp, N, pr, pi);
    CU(cudaGetLastError());
    st->ctr.launches++; st->ctr.bytes_hbm += 32ull * N;
    st->gphase[0] = 1.0; st->gphase[1] = 0.0;
    return QM_OK;
}

extern "C" int32_t qm_sync(qm_state* st) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    CU(cudaStreamSynchronize(st->stream));
    if (st->ev_pending) {
        float ms = 0.f; CU(cudaEventElapsedTime(&ms, st->ev0, st->ev1));
        st->ctr.ms_device = ms; st->ev_pending = false;
    }
    for (auto& pr : st->remap_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) st->ctr.ms_remap += ms; else cudaGetLastError();
        cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    st->remap_events.clear();
    return QM_OK;
}

extern "C" int32_t qm_upload(qm_state* st, const double* host, uint64_t n_amps, int32_t batch_idx) {
    if (!st || !host) return set_err(QM_ERR_ARG, "null argument");
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    if (n_amps != local_amps(st)) return set_err(QM_ERR_ARG, "expected %llu amplitudes (this rank's shard)", (unsigned long long)local_amps(st));
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    if (st->dist) QM_TRY(dist_canonicalize(st));
    QM_TRY(materialize_phase(st));          // the other registers of the batch keep their phase; this one is overwritten
    CU(cudaMemcpyAsync(st->amp + ((uint64_t)batch_idx << st->nl), host, (size_t)n_amps * 16, cudaMemcpyHostToDevice, st->stream));
    st->ctr.bytes_h2d += n_amps * 16;
    CU(cudaStreamSynchronize(st->stream));
    return QM_OK;
}

extern "C" int32_t qm_download(qm_state* st, double* host, uint64_t n_amps, int32_t batch_idx) {
    if (!st || !host) return set_err(QM_ERR_ARG, "null argument");
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    if (n_amps != local_amps(st)) return set_err(QM_ERR_ARG, "expected %llu amplitudes (this rank's shard)", (unsigned long long)local_amps(st));
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    if (st->dist) QM_TRY(dist_canonicalize(st));
    QM_TRY(materialize_phase(st));
    CU(cudaMemcpyAsync(host, st->amp + ((uint64_t)batch_idx << st->nl), (size_t)n_amps * 16, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += n_amps * 16;
    QM_TRY(qm_sync(st));
    return QM_OK;
}

// ---------------------------------------------------------------------------------------------
// readouts
// ---------------------------------------------------------------------------------------------
extern "C" int32_t qm_probs(qm_state* st, double* out_host, int32_t batch_idx) {
    if (!st || !out_host) return set_err(QM_ERR_ARG, "null argument");
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    if (st->dist) QM_TRY(dist_canonicalize(st));
    uint64_t N = local_amps(st);
    QM_TRY(ENSURE_DEV(st, d_scratch, (size_t)N * 8));
    k_probs<<<grid_for(N, 256, st->sms, 16), 256, 0, st->stream>>>(st->amp + ((uint64_t)batch_idx << st->nl), N, st->d_scratch);
    CU(cudaGetLastError());
    st->ctr.launches++; st->ctr.bytes_hbm += 24ull * N;
    CU(cudaMemcpyAsync(out_host, st->d_scratch, (size_t)N * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += N * 8;
    return qm_sync(st);
}

// all-qubit z marginals of every register into st->h_scratch: [batch][nl][2] then [batch] totals
// enqueue the all-qubit Z statistics and their copy into h_dst (pinned, out_doubles doubles); no synchronisation
static int zstats_enqueue(qm_state* st, double thr, double* h_dst, size_t* out_doubles_p);

int qm_zstats(qm_state* st, double thr, double** h_pairs, double** h_tot) {
    size_t out_doubles = (size_t)st->batch * st->nl * 2 + st->batch;
    QM_TRY(ENSURE_PIN(st, h_scratch, out_doubles * 8));
    QM_TRY(zstats_enqueue(st, thr, st->h_scratch, nullptr));
    QM_TRY(qm_sync(st));
    *h_pairs = st->h_scratch; *h_tot = st->h_scratch + (size_t)st->batch * st->nl * 2;
    return QM_OK;
}

static int zstats_enqueue(qm_state* st, double thr, double* h_dst, size_t* out_doubles_p) {
    const int nl = st->nl;
    int cb = nl < kZMaxChunkBits ? nl : kZMaxChunkBits;
    int rest = nl - cb;                                    // chunk-index bits per register
    int ctas_log2 = rest < 10 ? rest : 10; if (st->batch > 64 && ctas_log2 > 3) ctas_log2 = rest < 3 ? rest : 3;
    int hbits = rest - ctas_log2;
    if (hbits > kZMaxHi) { ctas_log2 = rest - kZMaxHi; hbits = kZMaxHi; }
    uint64_t nctas = (uint64_t)st->batch << ctas_log2;
    if (nctas > 0x7fffffffull) return set_err(QM_ERR_ARG, "register too large for the readout grid");
    size_t part_bytes = (size_t)nctas * kZVals * 8;
    size_t out_doubles = (size_t)st->batch * nl * 2 + st->batch;
    QM_TRY(ENSURE_DEV(st, d_scratch, part_bytes + out_doubles * 8));
    if (out_doubles_p) *out_doubles_p = out_doubles;
    double* d_part = st->d_scratch;
    double* d_out = st->d_scratch + (size_t)nctas * kZVals;
    double* d_tot = d_out + (size_t)st->batch * nl * 2;
    if (cb == kZMaxChunkBits) k_zpartial<true><<<(unsigned)nctas, kZThreads, 0, st->stream>>>(st->amp, nl, cb, hbits, ctas_log2, thr, d_part);
    else k_zpartial<false><<<(unsigned)nctas, kZThreads, 0, st->stream>>>(st->amp, nl, cb, hbits, ctas_log2, thr, d_part);
    CU(cudaGetLastError());
    k_zfinal<<<st->batch, 256, 0, st->stream>>>(d_part, nl, cb, hbits, ctas_log2, d_out, d_tot);
    CU(cudaGetLastError());
    st->ctr.launches += 2; st->ctr.bytes_hbm += 16ull * total_amps(st);
    CU(cudaMemcpyAsync(h_dst, d_out, out_doubles * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += out_doubles * 8;
    return QM_OK;
}

// Split readout for pipelines (a reservoir loop whose host side prepares step t + 1 while the device still runs step t):
// _begin enqueues the all-qubit <Z> of the state AS IT IS AT THIS POINT OF THE STREAM and returns at once, later gate calls
// may follow immediately; _end hands out the oldest pending result (at most two may be pending).  Unsharded registers.
extern "C" int32_t qm_expect_z_all_begin(qm_state* st, double threshold) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    if (st->dist) return set_err(QM_ERR_STATE, "split readout is for unsharded registers");
    if (st->zq_count >= 2) return set_err(QM_ERR_STATE, "two readouts are already pending: call qm_expect_z_all_end first");
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    const int slot = (st->zq_head + st->zq_count) & 1;
    const size_t out_doubles = (size_t)st->batch * st->nl * 2 + st->batch;
    QM_TRY(ensure_pinned((void**)&st->h_zq[slot], &st->h_zq_cap[slot], out_doubles * 8, &st->retired_pinned));
    if (!st->ev_zq[slot]) CU(cudaEventCreateWithFlags(&st->ev_zq[slot], cudaEventDisableTiming));
    QM_TRY(zstats_enqueue(st, threshold, st->h_zq[slot], nullptr));
    CU(cudaEventRecord(st->ev_zq[slot], st->stream));
    st->zq_count++;
    return QM_OK;
}

extern "C" int32_t qm_expect_z_all_end(qm_state* st, double* out_host) {
    if (!st || !out_host) return set_err(QM_ERR_ARG, "null argument");
    if (st->zq_count <= 0) return set_err(QM_ERR_STATE, "no readout is pending");
    CU(cudaSetDevice(st->device));
    const int slot = st->zq_head;
    CU(cudaEventSynchronize(st->ev_zq[slot]));
    memcpy(out_host, st->h_zq[slot], (size_t)st->batch * st->nl * 2 * sizeof(double));
    st->zq_head ^= 1; st->zq_count--;
    return QM_OK;
}

extern "C" int32_t qm_expect_z_all(qm_state* st, double threshold, double* out_host) {
    if (!st || !out_host) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    double *pairs, *tot;
    QM_TRY(qm_zstats(st, threshold, &pairs, &tot));
    if (st->dist) return dist_finish_z(st, pairs, tot, out_host);
    memcpy(out_host, pairs, (size_t)st->batch * st->nl * 2 * sizeof(double));
    return QM_OK;
}

extern "C" int32_t qm_norm2(qm_state* st, int32_t batch_idx, double* out) {
    if (!st || !out) return set_err(QM_ERR_ARG, "null argument");
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    double *pairs, *tot;
    QM_TRY(qm_zstats(st, -1.0, &pairs, &tot));
    if (st->dist) return dist_finish_norm(st, tot[0], out);
    *out = tot[batch_idx];
    return QM_OK;
}

static const double kHmat[8] = { 0.7071067811865475, 0, 0.7071067811865475, 0, 0.7071067811865475, 0, -0.7071067811865475, 0 };

int qm_measure_phys(qm_state* st, int tb, double thr, int batch_idx, const double* R, double out[2]) {
    if (tb >= st->nl) return set_err(QM_ERR_STATE, "internal: measurement target is not local");
    Mat2 U; int identity = (R == nullptr);
    if (R) memcpy(U.v, R, sizeof(U.v)); else memset(&U, 0, sizeof(U));
    uint64_t npairs = local_amps(st) >> 1;
    int blocks = grid_for(npairs, 256, st->sms, 8); if (blocks > kMeasBlocks) blocks = kMeasBlocks;
    QM_TRY(ENSURE_DEV(st, d_scratch, (size_t)(2 * kMeasBlocks + 2) * 8));
    QM_TRY(ENSURE_PIN(st, h_scratch, 64));
    const double2* base = st->amp + ((uint64_t)batch_idx << st->nl);
    k_measure_rot<<<blocks, 256, 0, st->stream>>>(base, npairs, tb, U, identity, thr, st->d_scratch);
    CU(cudaGetLastError());
    k_pair_final<<<1, 256, 0, st->stream>>>(st->d_scratch, blocks, st->d_scratch + 2 * kMeasBlocks);
    CU(cudaGetLastError());
    st->ctr.launches += 2; st->ctr.bytes_hbm += 16ull * local_amps(st);
    CU(cudaMemcpyAsync(st->h_scratch, st->d_scratch + 2 * kMeasBlocks, 16, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += 16;
    QM_TRY(qm_sync(st));
    out[0] = st->h_scratch[0]; out[1] = st->h_scratch[1];
    return QM_OK;
}

// rx_gate(pi/2) exactly as QSim.jl:119-121 builds it: c = cos(pi/4), s = sin(pi/4); [c -is; -is c].
// The host shim may pass its own matrix through qm_measure_with (bit-identical to Julia's); this
// default is what the C library uses for QM_BASIS_Y.
static void y_rotation(double R[8]) {
    const double c = cos(0.78539816339744830962), s = sin(0.78539816339744830962);
    R[0] = c; R[1] = 0; R[2] = 0; R[3] = -s; R[4] = 0; R[5] = -s; R[6] = c; R[7] = 0;
}

extern "C" int32_t qm_measure(qm_state* st, int32_t basis, int32_t t0, double threshold, int32_t batch_idx, double out[2]) {
    if (!st || !out) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(check_bit(st, t0, "target"));
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    if (basis < QM_BASIS_Z || basis > QM_BASIS_Y) return set_err(QM_ERR_ARG, "unknown basis %d", basis);
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    double R[8];
    const double* Rp = nullptr;
    if (basis == QM_BASIS_X) Rp = kHmat;
    else if (basis == QM_BASIS_Y) { y_rotation(R); Rp = R; }
    if (st->dist) return dist_measure(st, t0, threshold, Rp, out);
    return qm_measure_phys(st, st->perm[t0], threshold, batch_idx, Rp, out);
}

extern "C" int32_t qm_measure_with(qm_state* st, int32_t t0, const double* R, double threshold, int32_t batch_idx, double out[2]) {
    if (!st || !out) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(check_bit(st, t0, "target"));
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    if (st->dist) return dist_measure(st, t0, threshold, R, out);
    return qm_measure_phys(st, st->perm[t0], threshold, batch_idx, R, out);
}

// all-qubit marginals in the X or Y basis (or any caller-supplied rotation R): ceil((n - 12) / 9) + 1 reads of the state
static int expect_all_rot(qm_state* st, const double* Rm, double thr, double* out_host) {
    const int nl = st->nl, B = st->batch;
    Mat2 R; memcpy(R.v, Rm, sizeof(R.v));
    const int m = nl < kPauliM ? nl : kPauliM;
    const uint64_t tiles_per_reg = 1ull << (nl - m);
    uint64_t cpr = (uint64_t)st->sms * 4 / (uint64_t)B; if (cpr < 1) cpr = 1; if (cpr > tiles_per_reg) cpr = tiles_per_reg;
    const uint64_t nctas = cpr * (uint64_t)B;
    if (nctas > 0x7fffffffull) return set_err(QM_ERR_ARG, "batch too large for the readout grid");
    const size_t part_doubles = (size_t)nctas * 2 * kPauliM, out_doubles = (size_t)B * nl * 2;
    QM_TRY(ENSURE_DEV(st, d_scratch, (part_doubles + out_doubles) * 8));
    QM_TRY(ENSURE_PIN(st, h_scratch, out_doubles * 8));
    double* d_part = st->d_scratch; double* d_out = d_part + part_doubles;
    static bool attr[64] = { false };
    if (st->device < 0 || st->device >= 64 || !attr[st->device]) {
        CU(cudaFuncSetAttribute(k_pauli_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << kPauliM));
        if (st->device >= 0 && st->device < 64) attr[st->device] = true;
    }
    // pass 0: the m lowest bits (contiguous tiles); then windows of m - 3 high bits on 128-byte rows, the last one
    // pulled back so that it ends at the top bit (bits measured twice are masked out)
    int done = 0;
    for (int pass = 0; done < nl; ++pass) {
        int L, hs; uint32_t tmask;
        if (pass == 0) { L = m; hs = m; tmask = (1u << m) - 1u; done = m; }
        else {
            L = 3; const int nh = m - L;
            hs = done; if (hs + nh > nl) hs = nl - nh;
            tmask = 0; for (int j = 0; j < nh; ++j) if (hs + j >= done) tmask |= 1u << (L + j);
            done = hs + nh;
        }
        k_pauli_tile<<<(unsigned)nctas, kPauliThreads, (size_t)16 << m, st->stream>>>(st->amp, nl, m, L, hs, tmask, R, thr, (uint32_t)cpr, d_part);
        CU(cudaGetLastError());
        k_pauli_final<<<(unsigned)((B * 2 * kPauliM + 127) / 128), 128, 0, st->stream>>>(d_part, B, nl, m, L, hs, tmask, (uint32_t)cpr, d_out);
        CU(cudaGetLastError());
        st->ctr.launches += 2; st->ctr.bytes_hbm += 16ull * total_amps(st);
    }
    CU(cudaMemcpyAsync(st->h_scratch, d_out, out_doubles * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += out_doubles * 8;
    QM_TRY(qm_sync(st));
    memcpy(out_host, st->h_scratch, out_doubles * 8);
    return QM_OK;
}

extern "C" int32_t qm_expect_all_with(qm_state* st, const double* R, double threshold, double* out_host) {
    if (!st || !out_host) return set_err(QM_ERR_ARG, "null argument");
    if (!R) return qm_expect_z_all(st, threshold, out_host);
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    if (st->dist) {         // sharded: qubit by qubit (a rotated basis on a global qubit brings it home first, dist_measure)
        for (int q = 0; q < st->n; ++q) QM_TRY(dist_measure(st, q, threshold, R, out_host + 2 * (size_t)q));
        return QM_OK;
    }
    return expect_all_rot(st, R, threshold, out_host);
}

extern "C" int32_t qm_expect_all(qm_state* st, int32_t basis, double threshold, double* out_host) {
    if (!st || !out_host) return set_err(QM_ERR_ARG, "null argument");
    if (basis < QM_BASIS_Z || basis > QM_BASIS_Y) return set_err(QM_ERR_ARG, "unknown basis %d", basis);
    if (basis == QM_BASIS_Z) return qm_expect_z_all(st, threshold, out_host);
    double R[8];
    if (basis == QM_BASIS_X) memcpy(R, kHmat, sizeof(R)); else y_rotation(R);
    return qm_expect_all_with(st, R, threshold, out_host);
}

int qm_reduced_local(qm_state* st, int keep_low, int k, int batch_idx, double2** d_out_p) {
    const uint32_t d = 1u << k, T = d < 16 ? d : 16, tiles = d / T;
    const uint64_t R = 1ull << (st->nl - k);
    uint64_t xchunks = R / kRedXC; if (xchunks < 1) xchunks = 1; if (xchunks > 592) { xchunks = 512; }
    while (R % xchunks) --xchunks;
    uint64_t x_per = R / xchunks;
    size_t entries = (size_t)d * d;
    QM_TRY(ENSURE_DEV(st, d_scratch, (xchunks + 1) * entries * 16));
    QM_TRY(ENSURE_PIN(st, h_scratch, entries * 16));
    double2* d_part = reinterpret_cast<double2*>(st->d_scratch);
    double2* d_out = d_part + xchunks * entries;
    dim3 grid((unsigned)xchunks, tiles * tiles);
    k_reduced_partial<<<grid, 256, 0, st->stream>>>(st->amp + ((uint64_t)batch_idx << st->nl), st->nl, k, keep_low ? 1 : 0, x_per, d_part);
    CU(cudaGetLastError());
    k_reduced_final<<<(unsigned)((entries + 255) / 256), 256, 0, st->stream>>>(d_part, (uint32_t)xchunks, (uint32_t)entries, d_out);
    CU(cudaGetLastError());
    st->ctr.launches += 2; st->ctr.bytes_hbm += 16ull * local_amps(st);
    *d_out_p = d_out;
    return QM_OK;
}

extern "C" int32_t qm_reduced_density(qm_state* st, int32_t keep_low, int32_t k, int32_t batch_idx, double* out) {
    if (!st || !out) return set_err(QM_ERR_ARG, "null argument");
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    if (k < 1 || k >= st->n || k > 10) return set_err(QM_ERR_ARG, "Dimension mismatch: k = %d of n = %d (1 <= k < n, k <= 10)", k, st->n);
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    if (st->dist) return dist_reduced_density(st, keep_low, k, out);
    double2* d_out = nullptr;
    QM_TRY(qm_reduced_local(st, keep_low, k, batch_idx, &d_out));
    const size_t entries = (size_t)1 << (2 * k);
    CU(cudaMemcpyAsync(st->h_scratch, d_out, entries * 16, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += entries * 16;
    QM_TRY(qm_sync(st));
    memcpy(out, st->h_scratch, entries * 16);
    return QM_OK;
}

// ---------------------------------------------------------------------------------------------
// reduced matrices and QInfo measures for a whole batch of registers, on the device
// ---------------------------------------------------------------------------------------------
// reduced density of EVERY register of the batch: d_out [batch][4^k] double2 on the device
static int reduced_batch_local(qm_state* st, int keep_low, int k, double2** d_out_p) {
    const uint32_t d = 1u << k, T = d < 16 ? d : 16, tiles = d / T;
    const uint64_t R = 1ull << (st->nl - k);
    const int B = st->batch;
    if (B > 65535) return set_err(QM_ERR_ARG, "batch too large for the readout grid");
    uint64_t xchunks = R / kRedXC; if (xchunks < 1) xchunks = 1;
    uint64_t cap = (uint64_t)st->sms * 8 / (uint64_t)B; if (cap < 1) cap = 1; if (cap > 512) cap = 512;
    if (xchunks > cap) xchunks = cap;
    while (R % xchunks) --xchunks;
    const uint64_t x_per = R / xchunks;
    const size_t entries = (size_t)d * d;
    QM_TRY(ENSURE_DEV(st, d_scratch, ((size_t)B * (xchunks + 1) * entries) * 16 + (size_t)B * 8 * (kHermMaxD + 1)));
    double2* d_part = reinterpret_cast<double2*>(st->d_scratch);
    double2* d_out = d_part + (size_t)B * xchunks * entries;
    dim3 grid((unsigned)xchunks, tiles * tiles, (unsigned)B);
    k_reduced_partial<<<grid, 256, 0, st->stream>>>(st->amp, st->nl, k, keep_low ? 1 : 0, x_per, d_part);
    CU(cudaGetLastError());
    dim3 g2((unsigned)((entries + 255) / 256), (unsigned)B);
    k_reduced_final<<<g2, 256, 0, st->stream>>>(d_part, (uint32_t)xchunks, (uint32_t)entries, d_out);
    CU(cudaGetLastError());
    st->ctr.launches += 2; st->ctr.bytes_hbm += 16ull * total_amps(st);
    *d_out_p = d_out;
    return QM_OK;
}

extern "C" int32_t qm_reduced_density_batch(qm_state* st, int32_t keep_low, int32_t k, double* out) {
    if (!st || !out) return set_err(QM_ERR_ARG, "null argument");
    if (st->dist) return set_err(QM_ERR_STATE, "sharded registers are not batched");
    if (k < 1 || k >= st->n || k > 10) return set_err(QM_ERR_ARG, "Dimension mismatch: k = %d of n = %d (1 <= k < n, k <= 10)", k, st->n);
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    double2* d_out = nullptr;
    QM_TRY(reduced_batch_local(st, keep_low, k, &d_out));
    const size_t bytes = ((size_t)st->batch << (2 * k)) * 16;
    CU(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += bytes;
    return qm_sync(st);
}

// von Neumann / Renyi / min / max entropy of the reduced matrix of every register of the batch: one read of the states,
// one warp-per-matrix Jacobi launch, `batch` doubles back (QInfo.jl:178-182 entanglement_entropy for a whole batch)
extern "C" int32_t qm_entropy_batch(qm_state* st, int32_t keep_low, int32_t k, int32_t kind, double alpha, double* out_host) {
    if (!st || !out_host) return set_err(QM_ERR_ARG, "null argument");
    if (st->dist) return set_err(QM_ERR_STATE, "sharded registers are not batched");
    if (k < 1 || k >= st->n || k > 4) return set_err(QM_ERR_ARG, "Dimension mismatch: k = %d of n = %d (1 <= k < n, k <= 4 on the device)", k, st->n);
    if (kind < HM_VON_NEUMANN || kind > HM_MAX_ENTROPY) return set_err(QM_ERR_ARG, "unknown entropy kind %d", kind);
    if (kind == HM_RENYI && alpha == 1.0) return set_err(QM_ERR_ARG, "Renyi entropy undefined for alpha = 1; use von Neumann entropy");
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    double2* d_rho = nullptr;
    QM_TRY(reduced_batch_local(st, keep_low, k, &d_rho));
    double* d_ent = reinterpret_cast<double*>(d_rho + ((size_t)st->batch << (2 * k)));
    QM_TRY(ENSURE_PIN(st, h_scratch, (size_t)st->batch * 8));
    k_herm_measure<<<st->batch, 32, 0, st->stream>>>(d_rho, nullptr, 1 << k, st->batch, kind, alpha, d_ent);
    CU(cudaGetLastError());
    st->ctr.launches++;
    CU(cudaMemcpyAsync(st->h_scratch, d_ent, (size_t)st->batch * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += (size_t)st->batch * 8;
    QM_TRY(qm_sync(st));
    memcpy(out_host, st->h_scratch, (size_t)st->batch * 8);
    return QM_OK;
}

// QInfo measures of `count` host matrices (d x d complex, d <= 16, column-major), computed on `device`.
// kind: 0 von Neumann, 1 Renyi(alpha), 2 min-entropy, 3 max-entropy, 4 trace distance of (A - B), 5 fidelity(A, B),
// 6 eigenvalues (out: [count][d] ascending).  B may be NULL for the one-matrix kinds.
extern "C" int32_t qm_herm_measure_batch(int32_t device, int32_t kind, double alpha, const double* A, const double* B,
                                         int32_t d, int32_t count, double* out_host) {
    if (!A || !out_host || count < 0) return set_err(QM_ERR_ARG, "null argument");
    if (d < 1 || d > kHermMaxD) return set_err(QM_ERR_ARG, "matrix dimension %d (1 <= d <= %d)", d, kHermMaxD);
    if (kind < HM_VON_NEUMANN || kind > HM_EIGVALS) return set_err(QM_ERR_ARG, "unknown measure %d", kind);
    if ((kind == HM_FIDELITY || kind == HM_TRACE_DISTANCE) && !B) return set_err(QM_ERR_ARG, "this measure takes two matrices");
    if (kind == HM_RENYI && alpha == 1.0) return set_err(QM_ERR_ARG, "Renyi entropy undefined for alpha = 1");
    if (count == 0) return QM_OK;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return set_err(QM_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return set_err(QM_ERR_ARG, "device %d of %d", device, ndev);
    CU(cudaSetDevice(device));
    const size_t mbytes = (size_t)count * d * d * 16, obytes = (size_t)count * (kind == HM_EIGVALS ? d : 1) * 8;
    double2 *dA = nullptr, *dB = nullptr; double* dO = nullptr;
    int rc = [&]() -> int {
        CU(cudaMalloc((void**)&dA, mbytes)); CU(cudaMalloc((void**)&dO, obytes));
        if (kind == HM_TRACE_DISTANCE) {            // eigenvalues of A - B: subtract on the host
            std::vector<double> D((size_t)count * d * d * 2);
            for (size_t i = 0; i < D.size(); ++i) D[i] = A[i] - B[i];
            CU(cudaMemcpy(dA, D.data(), mbytes, cudaMemcpyHostToDevice));
        } else CU(cudaMemcpy(dA, A, mbytes, cudaMemcpyHostToDevice));
        if (kind == HM_FIDELITY) { CU(cudaMalloc((void**)&dB, mbytes)); CU(cudaMemcpy(dB, B, mbytes, cudaMemcpyHostToDevice)); }
        k_herm_measure<<<count, 32>>>(dA, dB, d, count, kind, alpha, dO);
        CU(cudaGetLastError());
        CU(cudaMemcpy(out_host, dO, obytes, cudaMemcpyDeviceToHost));
        return QM_OK;
    }();
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// density matrices as 2q-qubit registers (vec(rho), column-major): see kernels.cuh k_dm_*
// ---------------------------------------------------------------------------------------------
static int dm_check(qm_state* st, int* q) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    if (st->dist || st->batch != 1) return set_err(QM_ERR_STATE, "density-matrix calls need a single unsharded register");
    if (st->n & 1) return set_err(QM_ERR_ARG, "a density matrix on q qubits is a register of 2q qubits (got %d)", st->n);
    *q = st->n / 2;
    return QM_OK;
}

extern "C" int32_t qm_dm_from_state(qm_state* sv, qm_state* rho) {
    int q = 0;
    QM_TRY(dm_check(rho, &q));
    if (!sv || sv->dist || sv->batch != 1 || sv->n != q || sv->device != rho->device)
        return set_err(QM_ERR_ARG, "source must be an unsharded %d-qubit register on the same device", q);
    QM_TRY(flush(sv)); QM_TRY(flush(rho));
    CU(cudaSetDevice(rho->device));
    CU(cudaStreamSynchronize(sv->stream));
    for (int i = 0; i < rho->n; ++i) rho->perm[i] = i;
    rho->gphase[0] = 1.0; rho->gphase[1] = 0.0;          // overwritten whole; psi psi' does not see psi's pending phase
    k_dm_outer<<<grid_for(total_amps(rho), 256, rho->sms, 16), 256, 0, rho->stream>>>(sv->amp, q, rho->amp);
    CU(cudaGetLastError());
    rho->ctr.launches++; rho->ctr.bytes_hbm += 16ull * total_amps(rho);
    return QM_OK;
}

extern "C" int32_t qm_dm_diag(qm_state* rho, double* out_host) {
    int q = 0;
    QM_TRY(dm_check(rho, &q));
    if (!out_host) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(flush(rho));
    CU(cudaSetDevice(rho->device));
    QM_TRY(materialize_phase(rho));          // the diagonal of rho is linear in the stored amplitudes
    const uint64_t d = 1ull << q;
    QM_TRY(ENSURE_DEV(rho, d_scratch, (size_t)d * 8));
    k_dm_diag<<<grid_for(d, 256, rho->sms, 16), 256, 0, rho->stream>>>(rho->amp, q, rho->d_scratch);
    CU(cudaGetLastError());
    rho->ctr.launches++; rho->ctr.bytes_hbm += 16ull * d;
    CU(cudaMemcpyAsync(out_host, rho->d_scratch, (size_t)d * 8, cudaMemcpyDeviceToHost, rho->stream));
    rho->ctr.bytes_d2h += d * 8;
    return qm_sync(rho);
}

extern "C" int32_t qm_dm_measure(qm_state* rho, int32_t t0, const double* R, double threshold, double out[2]) {
    int q = 0;
    QM_TRY(dm_check(rho, &q));
    if (!out) return set_err(QM_ERR_ARG, "null argument");
    if (t0 < 0 || t0 >= q) return set_err(QM_ERR_ARG, "Qubit index out of range (target bit %d, q = %d)", t0, q);
    QM_TRY(flush(rho));
    CU(cudaSetDevice(rho->device));
    QM_TRY(materialize_phase(rho));
    Mat2 U; const int identity = (R == nullptr);
    if (R) memcpy(U.v, R, sizeof(U.v)); else memset(&U, 0, sizeof(U));
    const uint64_t npairs = (1ull << q) >> 1;
    int blocks = grid_for(npairs, 256, rho->sms, 8); if (blocks > kMeasBlocks) blocks = kMeasBlocks;
    QM_TRY(ENSURE_DEV(rho, d_scratch, (size_t)(2 * kMeasBlocks + 2) * 8));
    QM_TRY(ENSURE_PIN(rho, h_scratch, 64));
    k_dm_measure<<<blocks, 256, 0, rho->stream>>>(rho->amp, q, t0, U, identity, threshold, rho->d_scratch);
    CU(cudaGetLastError());
    k_pair_final<<<1, 256, 0, rho->stream>>>(rho->d_scratch, blocks, rho->d_scratch + 2 * kMeasBlocks);
    CU(cudaGetLastError());
    rho->ctr.launches += 2; rho->ctr.bytes_hbm += 64ull << q;
    CU(cudaMemcpyAsync(rho->h_scratch, rho->d_scratch + 2 * kMeasBlocks, 16, cudaMemcpyDeviceToHost, rho->stream));
    rho->ctr.bytes_d2h += 16;
    QM_TRY(qm_sync(rho));
    out[0] = rho->h_scratch[0]; out[1] = rho->h_scratch[1];
    return QM_OK;
}

int qm_flush(qm_state* st) { return flush(st); }

int qm_sample_leaves(qm_state* st, int batch_idx, double* d_s1_local) {
    const uint64_t N = local_amps(st), n1 = (N + kLeaf - 1) / kLeaf;
    k_leafsum<true><<<(unsigned)((n1 * 32 + 255) / 256), 256, 0, st->stream>>>(st->amp + ((uint64_t)batch_idx << st->nl), N, n1, d_s1_local);
    CU(cudaGetLastError());
    st->ctr.launches += 1; st->ctr.bytes_hbm += 16ull * N;
    return QM_OK;
}

// d_levels: room for n2 + n3 doubles (the upper levels, rebuilt from the full level-1 array on every rank alike)
int qm_sample_descend(qm_state* st, int batch_idx, const double* d_s1, uint64_t n1, uint64_t leaf_lo, uint64_t seed, int nshots,
                      double* d_levels, uint64_t* d_out) {
    const uint64_t n2 = (n1 + kLeaf - 1) / kLeaf, n3 = (n2 + kLeaf - 1) / kLeaf;
    if (n3 > kLeaf) return set_err(QM_ERR_ARG, "register too large for the sampler");
    double* s2 = d_levels; double* s3 = s2 + n2;
    k_leafsum<false><<<(unsigned)((n2 * 32 + 255) / 256), 256, 0, st->stream>>>(d_s1, n1, n2, s2);
    k_leafsum<false><<<(unsigned)((n3 * 32 + 255) / 256), 256, 0, st->stream>>>(s2, n2, n3, s3);
    const uint64_t Nloc = local_amps(st), leaf_cnt = (Nloc + kLeaf - 1) / kLeaf;
    const uint64_t Ntot = st->dist ? (1ull << st->n) : Nloc;
    k_sample<<<(nshots + 127) / 128, 128, 0, st->stream>>>(st->amp + ((uint64_t)batch_idx << st->nl), Ntot, d_s1, n1, s2, n2, s3, n3,
                                                          seed, nshots, d_out, leaf_lo, leaf_cnt);
    CU(cudaGetLastError());
    st->ctr.launches += 3;
    return QM_OK;
}

extern "C" int32_t qm_sample(qm_state* st, uint64_t seed, int32_t nshots, int32_t batch_idx, uint64_t* out) {
    if (!st || (!out && nshots > 0) || nshots < 0) return set_err(QM_ERR_ARG, "null argument");
    if (batch_idx < 0 || batch_idx >= st->batch) return set_err(QM_ERR_ARG, "batch index out of range");
    if (nshots == 0) return QM_OK;
    QM_TRY(flush(st));
    CU(cudaSetDevice(st->device));
    if (st->dist) return dist_sample(st, seed, nshots, out);
    const uint64_t N = local_amps(st);
    const uint64_t n1 = (N + kLeaf - 1) / kLeaf, n2 = (n1 + kLeaf - 1) / kLeaf, n3 = (n2 + kLeaf - 1) / kLeaf;
    size_t need = (n1 + n2 + n3) * 8 + (size_t)nshots * 8 + 64;
    QM_TRY(ENSURE_DEV(st, d_scratch, need));
    QM_TRY(ENSURE_PIN(st, h_scratch, (size_t)nshots * 8));
    double* s1 = st->d_scratch;
    uint64_t* d_out = reinterpret_cast<uint64_t*>(s1 + n1 + n2 + n3);
    QM_TRY(qm_sample_leaves(st, batch_idx, s1));
    QM_TRY(qm_sample_descend(st, batch_idx, s1, n1, 0, seed, nshots, s1 + n1, d_out));
    CU(cudaMemcpyAsync(st->h_scratch, d_out, (size_t)nshots * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += (size_t)nshots * 8;
    QM_TRY(qm_sync(st));
    memcpy(out, st->h_scratch, (size_t)nshots * 8);
    return QM_OK;
}

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
extern "C" int32_t qm_last_error(char* buf, int32_t cap) {
    if (!buf || cap <= 0) return QM_ERR_ARG;
    snprintf(buf, (size_t)cap, "%s", g_err);
    return QM_OK;
}
extern "C" int32_t qm_counters(qm_state* st, qm_counters_t* out) {
    if (!st || !out) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(qm_sync(st));
    *out = st->ctr;
    return QM_OK;
}
extern "C" int32_t qm_counters_reset(qm_state* st) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    memset(&st->ctr, 0, sizeof(st->ctr));
    return QM_OK;
}
extern "C" int32_t qm_abi_version(void) { return QM_ABI_VERSION; }
extern "C" int32_t qm_device_ptr(qm_state* st, void** amps, void** stream) {
    if (!st) return set_err(QM_ERR_ARG, "null state");
    if (amps) { QM_TRY(flush(st)); QM_TRY(materialize_phase(st)); *amps = st->amp; }     // the caller will read amplitudes
    if (stream) *stream = st->stream;
    return QM_OK;
}

// Planner introspection (host only).  Flat int32 layout, per step:
//   [kind(0 pass | 1 remap), ngates, m, L, nhigh, hb[0..nhigh), ngroups, {nrb, rb[0..nrb), nsub}*ngroups]
extern "C" int32_t qm_plan_describe_ex(int32_t n_qubits, int32_t global_bits, int32_t tile_bits, int32_t low_bits,
                                       double max_cost, int32_t pipeline, const qm_gate* gates, int32_t ngates,
                                       int32_t* out, int64_t cap, int64_t* out_len) {
    if (!gates || !out_len || n_qubits < 1 || global_bits < 0 || global_bits >= n_qubits) return set_err(QM_ERR_ARG, "bad argument");
    Planner pl;
    pl.opt.n_local = n_qubits - global_bits; pl.opt.n_global = global_bits;
    pl.opt.tile_bits = tile_bits; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost < 9.0 ? 9.0 : max_cost;
    if (pipeline) {
        v4_planner_options(pl.opt, max_cost, pipeline == 2 ? 1 : 0);      // pipeline 2: priced for the compiled pass kernels
    }
    int rc = lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates);
    if (rc) return set_err(rc, "bad gate");
    if (pipeline) prepare_for_lifting(pl.gates);
    pl.done.assign(pl.gates.size(), 0);
    std::vector<int32_t> v;
    PlanPass P;
    while (pl.next_pass(P)) {
        v.push_back(0); v.push_back(P.ngates); v.push_back(P.m); v.push_back(P.L); v.push_back((int)P.hb.size());
        for (int b : P.hb) v.push_back(b);
        v.push_back((int)P.groups.size());
        for (const PlanGroup& G : P.groups) {
            v.push_back((int)G.rb.size()); for (int b : G.rb) v.push_back(b);
            v.push_back((int)G.subs.size());
        }
    }
    int left = 0; for (size_t i = 0; i < pl.gates.size(); ++i) if (!pl.done[i]) ++left;
    if (left) { v.push_back(1); v.push_back(left); }
    *out_len = (int64_t)v.size();
    if ((int64_t)v.size() > cap || !out) return set_err(QM_ERR_ARG, "output too small: need %lld", (long long)v.size());
    memcpy(out, v.data(), v.size() * sizeof(int32_t));
    return QM_OK;
}

extern "C" int32_t qm_plan_gates(int32_t n_qubits, int32_t global_bits, int32_t tile_bits, int32_t low_bits,
                                 double max_cost, const qm_gate* gates, int32_t ngates, qm_gate* out,
                                 int32_t* pass_of, int32_t cap, int32_t* out_len) {
    if (!gates || !out_len || n_qubits < 1 || global_bits < 0 || global_bits >= n_qubits) return set_err(QM_ERR_ARG, "bad argument");
    Planner pl;
    pl.opt.n_local = n_qubits - global_bits; pl.opt.n_global = global_bits;
    pl.opt.tile_bits = tile_bits; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost < 9.0 ? 9.0 : max_cost;
    int rc = lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates);
    if (rc) return set_err(rc, "bad gate");
    pl.done.assign(pl.gates.size(), 0);
    std::vector<qm_gate> v; std::vector<int32_t> po;
    PlanPass P; int pi = 0;
    const int nl_ = n_qubits - global_bits;
    for (int guard = 0; guard < 4 * ngates + 8; ++guard) {
    while (pl.next_pass(P)) {
        for (const PlanGroup& G : P.groups)
            for (const LGate& x : G.subs) {
                qm_gate w; memset(&w, 0, sizeof(w));
                if (x.ctrl >= 0) { w.kind = QM_GATE_CU2; w.q0 = x.ctrl; w.q1 = x.target; }
                else if (x.slot >= 0) { w.kind = QM_GATE_U2_SLOT; w.q0 = x.target; w.q1 = x.slot; }
                else { w.kind = QM_GATE_U2; w.q0 = x.target; }
                memcpy(w.m, x.m, sizeof(w.m));
                // every non-diagonal target must be resident in the tile
                if (!x.diag && tile_local_of(P, x.target) < 0) return set_err(QM_ERR_STATE, "planner placed a non-diagonal target outside its tile");
                v.push_back(w); po.push_back(pi);
            }
        ++pi;
    }
    bool left = false;
    for (size_t i = 0; i < pl.gates.size(); ++i) if (!pl.done[i]) { left = true; break; }
    if (!left) break;
    if (global_bits == 0) return set_err(QM_ERR_STATE, "planner stalled");
    // what the sharded engine does here: exchange global and top local bits, relabel the pending gates.
    // The remap is reported as a pseudo-gate of kind -1 (q0 = local bit base, q1 = number of bits).
    qm_gate mk; memset(&mk, 0, sizeof(mk)); mk.kind = -1; mk.q0 = nl_ - global_bits; mk.q1 = global_bits;
    v.push_back(mk); po.push_back(pi);
    for (size_t i = 0; i < pl.gates.size(); ++i) {
        if (pl.done[i]) continue;
        pl.gates[i].target = dist_swap_bit(pl.gates[i].target, nl_, global_bits);
        if (pl.gates[i].ctrl >= 0) pl.gates[i].ctrl = dist_swap_bit(pl.gates[i].ctrl, nl_, global_bits);
    }
    }
    *out_len = (int32_t)v.size();
    if ((int32_t)v.size() > cap || !out) return set_err(QM_ERR_ARG, "output too small: need %d", (int)v.size());
    memcpy(out, v.data(), v.size() * sizeof(qm_gate));
    if (pass_of) memcpy(pass_of, po.data(), po.size() * sizeof(int32_t));
    return QM_OK;
}

// Diagnostics (QM_OPT_TIME_PASSES): per fused pass since the last call {ms, groups, gates, planner cost, records: full, half, xor, thr}.
extern "C" int32_t qm_pass_times(qm_state* st, float* out, int32_t cap_passes, int32_t* n_passes) {
    if (!st || !n_passes) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(qm_sync(st));
    const int n = (int)(st->pass_events.size() / 2);
    *n_passes = n;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, st->pass_events[2 * i], st->pass_events[2 * i + 1]));
        if (out && i < cap_passes) { out[8 * i] = ms; for (int q = 0; q < 7; ++q) out[8 * i + 1 + q] = st->pass_desc[7 * i + q]; }
        cudaEventDestroy(st->pass_events[2 * i]); cudaEventDestroy(st->pass_events[2 * i + 1]);
    }
    st->pass_events.clear(); st->pass_desc.clear();
    return QM_OK;
}

// Diagnostics (QM_OPT_TRACE): SM-clock stamps of CTA 0 for the LAST fused pass launched, 8 per tile:
// {team: wait start, tile landed, after group 0, 1, 2, arrived | producer: computed seen, next load issued}
extern "C" int32_t qm_trace_read(qm_state* st, uint64_t* out, int64_t cap_u64) {
    if (!st || !out || cap_u64 <= 0) return set_err(QM_ERR_ARG, "null argument");
    if (!st->trace_buf) return set_err(QM_ERR_STATE, "QM_OPT_TRACE is off");
    QM_TRY(qm_sync(st));
    size_t bytes = (size_t)cap_u64 * 8; if (bytes > kTraceBytes) bytes = kTraceBytes;
    CU(cudaMemcpy(out, st->trace_buf, bytes, cudaMemcpyDeviceToHost));
    return QM_OK;
}

// Host-only introspection: plans and encodes `gates` as for the TMA kernel and counts the records by family:
// hist[0..7] = ROT, ROTX, PHASE, PERMX, THR, SLOT, gates entered through the controlled entry, groups; hist[8] = passes.
extern "C" int32_t qm_plan_ophist(int32_t n_qubits, int32_t tile_bits, int32_t low_bits, double max_cost,
                                  const qm_gate* gates, int32_t ngates, int64_t hist[9]) {
    if (!gates || !hist || n_qubits < 11 || (tile_bits != 11 && tile_bits != 12) || tile_bits > n_qubits) return set_err(QM_ERR_ARG, "bad argument");
    Planner pl;
    pl.opt.n_local = n_qubits; pl.opt.n_global = 0;
    pl.opt.tile_bits = tile_bits; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost;
    v4_planner_options(pl.opt, max_cost);
    int rc = lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates);
    if (rc) return set_err(rc, "bad gate");
    prepare_for_lifting(pl.gates);
    pl.done.assign(pl.gates.size(), 0);
    for (int i = 0; i < 9; ++i) hist[i] = 0;
    std::vector<V4Program> prog(1);
    PlanPass P;
    while (pl.next_pass(P)) {
        choose_tile_layout(P);
        double gph[2] = { 1.0, 0.0 };
        if (!encode_v4(P, prog[0], gph)) return set_err(QM_ERR_STATE, "pass cannot be encoded");
        hist[8]++; hist[7] += prog[0].ngroups;
        for (int i = 0; i < prog[0].nrecs; ++i) {
            uint32_t op = prog[0].recs[i].op;
            if (op >= (uint32_t)V4_NUM_GATE_OPS && op < (uint32_t)V4_OP_SLOT) { op -= V4_NUM_GATE_OPS; hist[6]++; }
            if (op < (uint32_t)V4_OP_ROTX) hist[0]++;
            else if (op < (uint32_t)V4_OP_PHASE) hist[1]++;
            else if (op < (uint32_t)V4_OP_PERMX) hist[2]++;
            else if (op < (uint32_t)V4_OP_THR) hist[3]++;
            else if (op < (uint32_t)V4_NUM_GATE_OPS) hist[4]++;
            else if (op == (uint32_t)V4_OP_SLOT) hist[5]++;
        }
    }
    return QM_OK;
}

// Host-only introspection of the compiled pass kernels (no GPU needed: the PTX text is generated and, when libnvJitLink is
// present, assembled into a cubin for sm_100a): plans and encodes `gates` as for the TMA kernel and puts every pass of tile
// size 2^12 through jit_ptx_for / jit_compile_cubin.  out = {passes, passes with PTX text, passes compiled to a cubin,
// distinct opcode sequences, bytes of the largest cubin, a hash over the opcode sequences of all passes in order}; ptx_out
// (optional) receives the text of the first pass.  Programs are encoded as the engine encodes them for a compiled kernel
// (canonical signs): two circuits of the same structure with different angles give the same hash.
extern "C" int32_t qm_jit_check(int32_t n_qubits, int32_t low_bits, double max_cost, const qm_gate* gates, int32_t ngates,
                                int32_t compile, int64_t out[6], char* ptx_out, int64_t ptx_cap) {
    if (!gates || !out || n_qubits < 12) return set_err(QM_ERR_ARG, "bad argument");
    Planner pl;
    pl.opt.n_local = n_qubits; pl.opt.n_global = 0;
    pl.opt.tile_bits = 12; pl.opt.low_bits = low_bits; pl.opt.fuse = true;
    if (max_cost > 0) pl.opt.max_cost = max_cost;
    v4_planner_options(pl.opt, max_cost, 1);
    int rc = lower_gates(gates, ngates, n_qubits, nullptr, true, pl.gates);
    if (rc) return set_err(rc, "bad gate");
    prepare_for_lifting(pl.gates);
    pl.done.assign(pl.gates.size(), 0);
    for (int i = 0; i < 6; ++i) out[i] = 0;
    std::vector<V4Program> prog(1);
    std::vector<std::vector<uint32_t>> seen;
    PlanPass P;
    while (pl.next_pass(P)) {
        choose_tile_layout(P);
        double gph[2] = { 1.0, 0.0 };
        if (!encode_v4(P, prog[0], gph, true)) return set_err(QM_ERR_STATE, "pass cannot be encoded");
        out[0]++;
        std::string ptx;
        if (!jit_ptx_for(P.m, prog[0], ptx, (compile & 2) != 0)) continue;
        if (out[1] == 0 && ptx_out && ptx_cap > 0) snprintf(ptx_out, (size_t)ptx_cap, "%s", ptx.c_str());
        out[1]++;
        std::vector<uint32_t> sig; jit_structure_bytes(prog[0], P.m, sig);
        if (std::find(seen.begin(), seen.end(), sig) == seen.end()) { seen.push_back(sig); out[3]++; }
        { uint64_t k[2]; jit_hash(sig, k); out[5] = (int64_t)(((uint64_t)out[5] * 0x100000001b3ull) ^ k[0] ^ (k[1] << 1)); }
        if (compile & 1) {
            std::vector<char> cubin; std::string log;
            if (!jit_cubin_for(ptx, cubin, log, nullptr)) return set_err(QM_ERR_STATE, "pass %lld does not assemble: %s", (long long)out[0] - 1, log.c_str());
            out[2]++;
            if ((int64_t)cubin.size() > out[4]) out[4] = (int64_t)cubin.size();
        }
    }
    return QM_OK;
}
