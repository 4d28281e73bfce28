// dist.cu — sharded registers.  Rank r of G = 2^g holds the 2^(n-g) amplitudes whose top g index bits
// equal r.  Everything that does not need a non-diagonal target on a global qubit runs locally (a
// control or a diagonal target on a global qubit is a rank predicate / rank-dependent phase inside
// the ordinary kernels).  When the planner runs out of local work, the g global qubits are exchanged
// with the top g local ones: viewed as [2^g chunks][2^(n-2g)], chunk j of rank r trades places with
// chunk r of rank j — an all-to-all done IN PLACE, because a 2^33-amplitude shard (128 GiB) leaves no
// room for an out-of-place copy on a 180 GB part.
//
// Two transports:
//  * PEER MEMORY (preferred; CUDA IPC between processes, or plain pointers between the ranks of one process):
//    every rank sees the others' shards and a small WINDOW (flags + mailboxes) that sits right behind each shard
//    in the same allocation.  The exchange is one kernel of loads/stores over NVLink (k_remap_swap); ranks are
//    ordered by FLAGS — a one-warp kernel stores an epoch number into every rank's window after
//    __threadfence_system(), the consumer's STREAM waits on its own window with cuStreamWaitValue32 — so no
//    fence costs an SM or a NCCL kernel, and no kernel ever spins on another GPU.  With that the exchange is cut
//    into SLABS (a few index bits pinned) and runs on a second stream BESIDE the fused passes of the neighbouring
//    slabs (run_circuit in qm_api.cu): the pass before the exchange and the pass after it hide most of it.
//    The small collectives of the readouts (sums of a few doubles, the sampler's level-1 all-gather) go through
//    the mailboxes and are folded in rank order, so every rank computes bit-identical results.
//  * NCCL (fallback when peer mapping is refused): grouped ncclSend/ncclRecv through a staging ring and
//    ncclAllReduce / ncclAllGather for the readouts.  NCCL is dlopen'ed on first use.
#include "dist.hpp"

#include <cuda.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

using namespace qm;

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSum = 0, kNcclUint64 = 5, kNcclFloat64 = 8 };

struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.h) return QM_OK;
    const char* names[] = { getenv("QM_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* nm : names) { if (nm && *nm) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; } }
    if (!h) return set_err(QM_ERR_COMM, "cannot dlopen libnccl.so.2 (set QM_NCCL_LIB): %s", dlerror());
#define LOADSYM(field, sym) do { *(void**)(&g_nccl.field) = dlsym(h, sym); if (!g_nccl.field) return set_err(QM_ERR_COMM, "libnccl: missing %s", sym); } while (0)
    LOADSYM(GetUniqueId, "ncclGetUniqueId"); LOADSYM(CommInitRank, "ncclCommInitRank"); LOADSYM(CommDestroy, "ncclCommDestroy");
    LOADSYM(Send, "ncclSend"); LOADSYM(Recv, "ncclRecv"); LOADSYM(AllReduce, "ncclAllReduce"); LOADSYM(AllGather, "ncclAllGather");
    LOADSYM(GroupStart, "ncclGroupStart"); LOADSYM(GroupEnd, "ncclGroupEnd"); LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
    g_nccl.h = h;
    return QM_OK;
}
#define NC(x) do { int r_ = (x); if (r_ != 0) return set_err(QM_ERR_COMM, "%s failed: %s", #x, g_nccl.GetErrorString(r_)); } while (0)

// ---- stream memory operations (driver API, resolved at run time) ---------------------------------------------
typedef CUresult (*PFN_waitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static PFN_waitValue32 get_wait_value() {
    static PFN_waitValue32 fn = nullptr; static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_waitValue32)p;
    }
    return fn;
}

// ---------------------------------------------------------------------------------------------
// The window behind every shard:  flags[1024] (uint32)  |  mailboxes [2 parities][16 sources][kMboxBytes]
// flag index = kind * 256 + source rank * 16 + slab
// ---------------------------------------------------------------------------------------------
enum { FK_PASS = 0, FK_EXCH = 1, FK_MBOX = 2, FK_FENCE = 3 };
static const size_t kFlagBytes = 4096;
static const size_t kMboxBytes = 128u << 10;
static const size_t kWinBytes = kFlagBytes + 2 * 16 * kMboxBytes;
size_t dist_window_bytes() { return kWinBytes; }

struct PeerWins { unsigned char* p[16]; };
struct PeerPtrs { double2* p[16]; };

// after everything earlier in the stream: make it visible system-wide, then post `epoch` in every rank's window
__global__ void k_signal(const __grid_constant__ PeerWins pw, int world, int flag_index, uint32_t epoch) {
    const int j = threadIdx.x;
    __threadfence_system();
    if (j < world) {
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(pw.p[j]) + flag_index;
        *f = epoch;
    }
}

// block j copies n16 16-byte words from src into mailbox [parity][me] of rank j, then posts the epoch there
__global__ void __launch_bounds__(256)
k_mbox_put(const __grid_constant__ PeerWins pw, int me, int parity, const uint4* __restrict__ src, uint32_t n16, uint32_t epoch) {
    const int j = blockIdx.x;
    uint4* dst = reinterpret_cast<uint4*>(pw.p[j] + kFlagBytes + ((size_t)parity * 16 + me) * kMboxBytes);
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(pw.p[j]) + (FK_MBOX * 256 + me * 16);
        *f = epoch;
    }
}
// out[i] = sum over ranks in rank order of mailbox[parity][r][i]: the same bits on every rank
template <typename T>
__global__ void k_mbox_sum(const unsigned char* win, int world, int parity, uint32_t count, T* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    T acc = 0;
    for (int r = 0; r < world; ++r)
        acc += reinterpret_cast<const volatile T*>(win + kFlagBytes + ((size_t)parity * 16 + r) * kMboxBytes)[i];
    out[i] = acc;
}
// out[r * stride + i] = mailbox[parity][r][i]
__global__ void k_mbox_gather(const unsigned char* win, int world, int parity, uint32_t count, size_t stride, double* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    for (int r = 0; r < world; ++r)
        out[(size_t)r * stride + i] = reinterpret_cast<const volatile double*>(win + kFlagBytes + ((size_t)parity * 16 + r) * kMboxBytes)[i];
}

// ---------------------------------------------------------------------------------------------
// K7  global<->local qubit exchange over peer memory.  Chunk j of rank r trades places with chunk r of rank j.
// For each peer a rank owns half of the chunk pair (the lower rank the first half, the higher rank the second),
// loads its own 16 bytes and the peer's 16 bytes (a load over NVLink) and stores them crosswise (a store over
// NVLink).  In place, no staging, both link directions busy; each amplitude pair is touched by exactly one thread
// of one rank.  Work is cut into 64 KiB segments dealt round-robin over the peers, and every rank starts its round at
// a different peer ((rank + 1 + k) mod world), so every GPU's NVLink ingress is fed by all the others evenly.
// A SLAB launch covers only the offsets whose pinned bits (positions >= 12 inside the half-chunk offset, so whole
// segments) have the given values; UN = loads in flight per thread.
// ---------------------------------------------------------------------------------------------
struct SwapSlab { int nbits; int bit[3]; unsigned long long value_or; };

template <int UN, int THREADS>
__global__ void __launch_bounds__(THREADS, (UN == 4 && THREADS == 512) ? 2 : 1)
k_remap_swap(double2* __restrict__ mine, const __grid_constant__ PeerPtrs peers, int rank, int world, unsigned long long chunk,
             const __grid_constant__ SwapSlab slab) {
    const unsigned long long half = chunk >> 1;
    const int seg_log2 = half < 4096ull ? (63 - __clzll((long long)half)) : 12;    // amplitudes per segment: 2^12 (64 KiB) or the whole half
    const unsigned long long seg_mask = (1ull << seg_log2) - 1ull;
    const unsigned long long half_c = half >> slab.nbits;                  // compressed offsets this launch covers
    const unsigned long long total = half_c * (unsigned long long)(world - 1);
    const unsigned long long stride = (unsigned long long)gridDim.x * THREADS;
    const uint32_t wm1 = (uint32_t)(world - 1);
    for (unsigned long long i = (unsigned long long)blockIdx.x * THREADS + threadIdx.x; i < total; i += UN * stride) {
        double2 a[UN], b[UN]; double2 *mp[UN], *pp[UN]; bool ok[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const unsigned long long ii = i + (unsigned long long)u * stride;
            ok[u] = ii < total;
            const unsigned long long q = ok[u] ? ii : 0;
            const uint32_t sg = (uint32_t)(q >> seg_log2);                          // < 2^32: a shard has < 2^44 amplitudes
            const uint32_t sgq = sg / wm1, k = sg - sgq * wm1;                      // which peer, round-robin
            unsigned long long o = ((unsigned long long)sgq << seg_log2) | (q & seg_mask);
            for (int s = 0; s < slab.nbits; ++s) {                                  // open the pinned bits (ascending)
                const unsigned long long lo = o & ((1ull << slab.bit[s]) - 1ull);
                o = ((o >> slab.bit[s]) << (slab.bit[s] + 1)) | lo;
            }
            o |= slab.value_or;
            int j = rank + 1 + (int)k; if (j >= world) j -= world;                  // rank-rotated peer order
            const unsigned long long off = (rank < j ? 0ull : half) + o;
            mp[u] = mine + ((unsigned long long)j * chunk + off); pp[u] = peers.p[j] + ((unsigned long long)rank * chunk + off);
            if (ok[u]) { a[u] = *mp[u]; b[u] = *pp[u]; }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) if (ok[u]) { *mp[u] = b[u]; *pp[u] = a[u]; }
    }
}

// A same-process world whose ranks all sit on ONE device runs every rank on one shared, in-order stream.  Flags and
// stream waits between streams of the same device are ordering the CUDA scheduler cannot see (it may park the kernel that
// posts a flag behind the stream that waits for it), so there the ordering is made of what it does see: stream order,
// with a host-side handshake (a rank enqueues work that depends on another rank's only after that rank has ENQUEUED its
// part).  The kernels, the exchange, the mailboxes and the slab bookkeeping are the ones every other transport runs.
struct LocalWorld {
    std::vector<qm_state*> ranks;
    cudaStream_t shared = nullptr;
    int refs = 0;
    std::atomic<uint32_t> issued[4][16][16];      // [kind][rank][slab]: last epoch whose signal has been enqueued
    LocalWorld() { for (auto& a : issued) for (auto& b : a) for (auto& c : b) c.store(0); }
};

struct DistCtx {
    std::shared_ptr<LocalWorld> lw;                     // same-process worlds only
    bool shared_stream = false;                         // ... with every rank on one device
    cudaEvent_t ev_sig[4][16] = { { nullptr } };        // this rank's own signals as CUDA events: dependencies between its own
                                                        // two streams must be visible to the scheduler
    double2* peer[16] = { nullptr };                    // the other ranks' shards (IPC-mapped or same-process pointers); null: NCCL path
    unsigned char* win[16] = { nullptr };               // every rank's window, own included
    bool have_peers = false, ipc_mapped = false;
    ncclComm_t comm = nullptr;
    double2* stage = nullptr; size_t stage_amps = 0;     // NCCL path: (world - 1) x sub-chunk
    double* d_small = nullptr;                           // collective scratch (4096 doubles)
    bool swapped = false;                                // layout B: global qubits currently sit on the top local bits
    cudaStream_t xstream = nullptr;                      // the exchange runs here when it is overlapped with passes
    uint32_t ep[4] = { 0, 0, 0, 0 };                     // epochs per flag kind; every rank advances them alike (SPMD calls)
    int mbox_parity = 0;
    unsigned long long local_key = 0; bool local_world = false;
};

// ---- same-process worlds (all ranks in one process: the 1-GPU test of the sharded path) ------------------------
static std::mutex g_local_mu;
static std::map<unsigned long long, std::shared_ptr<LocalWorld>> g_local_worlds;

int dist_destroy(DistCtx* d) {
    if (!d) return QM_OK;
    if (d->ipc_mapped) for (int j = 0; j < 16; ++j) if (d->peer[j]) cudaIpcCloseMemHandle(d->peer[j]);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    if (d->xstream && !d->shared_stream) cudaStreamDestroy(d->xstream);
    for (auto& a : d->ev_sig) for (cudaEvent_t e : a) if (e) cudaEventDestroy(e);
    cudaFree(d->stage); cudaFree(d->d_small);
    if (d->local_world) {
        std::lock_guard<std::mutex> lk(g_local_mu);
        if (d->lw && d->shared_stream && --d->lw->refs == 0 && d->lw->shared) { cudaStreamDestroy(d->lw->shared); d->lw->shared = nullptr; }
        g_local_worlds.erase(d->local_key);
    }
    delete d;
    return QM_OK;
}
// the state's stream belongs to the world, not to the state
bool dist_owns_stream(const qm_state* st) { return st->dist && st->dist->shared_stream; }

extern "C" int32_t qm_dist_unique_id(uint8_t id_out[128]) {
    if (!id_out) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(nccl_load());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return QM_OK;
}

static unsigned char* own_window(const qm_state* st) {
    return reinterpret_cast<unsigned char*>(st->amp) + (size_t)local_amps(st) * 16;
}

static int dist_common_init(qm_state* st, DistCtx* d) {
    CU(cudaMalloc((void**)&d->d_small, 4096 * sizeof(double)));
    CU(cudaMemset(d->d_small, 0, 4096 * sizeof(double)));
    CU(cudaMemset(own_window(st), 0, kWinBytes));
    CU(cudaStreamCreateWithFlags(&d->xstream, cudaStreamNonBlocking));
    // Readout scratch is allocated NOW, large enough for every collective readout of a moderate register (all-qubit
    // marginals, single-qubit measurements, reduced densities up to k = 6, the sampler up to 2^30 amplitudes):
    // cudaMalloc / cudaMallocHost synchronise with the other streams of the device, and when several ranks share a
    // device (same-process worlds) another rank's stream may be parked on a flag this rank has yet to post.
    QM_TRY(ENSURE_DEV(st, d_scratch, (size_t)48 << 20));
    QM_TRY(ENSURE_PIN(st, h_scratch, (size_t)1 << 20));
    CU(cudaDeviceSynchronize());
    d->win[st->rank] = own_window(st);
    return QM_OK;
}

extern "C" int32_t qm_create_dist(int32_t n, uint64_t basis, int32_t device, int32_t rank, int32_t world,
                                  const uint8_t nccl_id[128], qm_state** out) {
    if (!nccl_id || !out) return set_err(QM_ERR_ARG, "null argument");
    *out = nullptr;
    if (world < 2 || world > 16) return set_err(QM_ERR_ARG, "world must be in [2, 16] (use qm_create for one GPU)");
    QM_TRY(nccl_load());
    qm_state* st = nullptr;
    QM_TRY(qm_create_common(n, basis, 1, device, rank, world, &st));
    if (2 * st->g > st->n) { qm_destroy(st); return set_err(QM_ERR_ARG, "register too small for %d ranks", world); }
    DistCtx* d = new DistCtx();
    st->dist = d;                                       // from here on qm_destroy releases everything
    ncclUniqueId id; memcpy(id.internal, nccl_id, 128);
    int r = g_nccl.CommInitRank(&d->comm, world, id, rank);
    if (r != 0) { d->comm = nullptr; qm_destroy(st); return set_err(QM_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    // NCCL path staging ring: (world - 1) sub-chunks of at most 2^24 amplitudes (256 MiB) each, allocated on first use
    const uint64_t chunk = 1ull << (st->nl - st->g);
    d->stage_amps = chunk < (1ull << 24) ? chunk : (1ull << 24);
    int rc = dist_common_init(st, d);
    if (rc != QM_OK) { qm_destroy(st); return rc; }
    *out = st;
    return QM_OK;
}

// All ranks of a register inside ONE process (one host thread per rank; one device, or several that can reach each
// other): no NCCL, peers are plain pointers.  `key` names the world; every rank calls qm_create_dist_local with the
// same key, then — after all of them have returned (the caller's own barrier) — qm_dist_local_connect.
extern "C" int32_t qm_create_dist_local(int32_t n, uint64_t basis, int32_t device, int32_t rank, int32_t world,
                                        uint64_t key, qm_state** out) {
    if (!out) return set_err(QM_ERR_ARG, "null argument");
    *out = nullptr;
    if (world < 2 || world > 16) return set_err(QM_ERR_ARG, "world must be in [2, 16]");
    qm_state* st = nullptr;
    QM_TRY(qm_create_common(n, basis, 1, device, rank, world, &st));
    if (!get_wait_value()) { qm_destroy(st); return set_err(QM_ERR_COMM, "cuStreamWaitValue32 is not available: a same-process world needs stream memory operations"); }
    if (2 * st->g > st->n) { qm_destroy(st); return set_err(QM_ERR_ARG, "register too small for %d ranks", world); }
    DistCtx* d = new DistCtx();
    st->dist = d;
    d->local_world = true; d->local_key = key;
    int rc = dist_common_init(st, d);
    if (rc != QM_OK) { qm_destroy(st); return rc; }
    {
        std::lock_guard<std::mutex> lk(g_local_mu);
        std::shared_ptr<LocalWorld>& w = g_local_worlds[key];
        if (!w) w = std::make_shared<LocalWorld>();
        if (w->ranks.size() < (size_t)world) w->ranks.resize(world, nullptr);
        w->ranks[rank] = st;
        d->lw = w;
    }
    *out = st;
    return QM_OK;
}

extern "C" int32_t qm_dist_local_connect(qm_state* st) {
    if (!st || !st->dist || !st->dist->local_world) return set_err(QM_ERR_ARG, "not a rank of a same-process world");
    DistCtx* d = st->dist;
    CU(cudaSetDevice(st->device));
    std::lock_guard<std::mutex> lk(g_local_mu);
    if (!d->lw) return set_err(QM_ERR_STATE, "world not registered");
    const std::vector<qm_state*>& v = d->lw->ranks;
    bool one_device = true;
    for (int j = 0; j < st->world; ++j) one_device = one_device && (size_t)j < v.size() && v[j] && v[j]->device == st->device;
    for (int j = 0; j < st->world; ++j) {
        if ((size_t)j >= v.size() || !v[j]) return set_err(QM_ERR_STATE, "rank %d of the world has not been created yet", j);
        if (v[j]->n != st->n || v[j]->world != st->world) return set_err(QM_ERR_ARG, "rank %d has a different shape", j);
        if (j == st->rank) continue;
        if (v[j]->device != st->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(v[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return set_err(QM_ERR_COMM, "cannot enable peer access to device %d", v[j]->device); }
            cudaGetLastError();
        }
        d->peer[j] = v[j]->amp;
        d->win[j] = own_window(v[j]);
    }
    if (one_device) {           // every rank on one device: one shared in-order stream for all of them (see LocalWorld)
        if (!d->lw->shared) CU(cudaStreamCreateWithFlags(&d->lw->shared, cudaStreamNonBlocking));
        CU(cudaStreamSynchronize(st->stream)); CU(cudaStreamSynchronize(d->xstream));
        cudaStreamDestroy(st->stream); cudaStreamDestroy(d->xstream);
        st->stream = d->lw->shared; d->xstream = d->lw->shared;
        d->shared_stream = true; d->lw->refs++;
    }
    d->have_peers = true;
    return QM_OK;
}

// ---- flag fences and mailbox collectives (peer-memory transport) ----------------------------------------------
static PeerWins wins_of(const DistCtx* d) { PeerWins pw; for (int j = 0; j < 16; ++j) pw.p[j] = d->win[j]; return pw; }

static int signal_all(qm_state* st, cudaStream_t stream, int kind, int slab, uint32_t epoch) {
    DistCtx* d = st->dist;
    if (d->shared_stream) {         // stream order is the ordering: publish that this rank's part has been enqueued
        d->lw->issued[kind][st->rank][slab].store(epoch, std::memory_order_release);
        return QM_OK;
    }
    k_signal<<<1, 32, 0, stream>>>(wins_of(d), st->world, kind * 256 + st->rank * 16 + slab, epoch);
    CU(cudaGetLastError());
    st->ctr.launches++;
    // this rank's own streams see the signal as an event (a dependency the scheduler knows about)
    if (!d->ev_sig[kind][slab]) CU(cudaEventCreateWithFlags(&d->ev_sig[kind][slab], cudaEventDisableTiming));
    CU(cudaEventRecord(d->ev_sig[kind][slab], stream));
    return QM_OK;
}
// the stream waits until every rank (this one included) has posted `epoch` for (kind, slab)
static int wait_all(qm_state* st, cudaStream_t stream, int kind, int slab, uint32_t epoch) {
    DistCtx* d = st->dist;
    if (d->shared_stream) {
        const auto t0 = std::chrono::steady_clock::now();
        for (int j = 0; j < st->world; ++j) {
            int spins = 0;
            while ((int32_t)(d->lw->issued[kind][j][slab].load(std::memory_order_acquire) - epoch) < 0) {
                if (++spins < 200) std::this_thread::yield(); else std::this_thread::sleep_for(std::chrono::microseconds(50));
                if ((spins & 1023) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
                    return set_err(QM_ERR_COMM, "rank %d of the same-process world did not reach this collective (kind %d, slab %d)", j, kind, slab);
            }
        }
        return QM_OK;
    }
    PFN_waitValue32 wv = get_wait_value();
    if (!wv) return set_err(QM_ERR_COMM, "cuStreamWaitValue32 is not available");
    unsigned char* w = d->win[st->rank];
    for (int j = 0; j < st->world; ++j) {
        if (j == st->rank) {
            if (d->ev_sig[kind][slab]) CU(cudaStreamWaitEvent(stream, d->ev_sig[kind][slab], 0));
            continue;
        }
        CUresult r = wv((CUstream)stream, (CUdeviceptr)(w + 4u * (kind * 256 + j * 16 + slab)), epoch, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) return set_err(QM_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)r);
    }
    return QM_OK;
}
static int fence_all(qm_state* st, cudaStream_t stream) {
    const uint32_t e = ++st->dist->ep[FK_FENCE];
    QM_TRY(signal_all(st, stream, FK_FENCE, 0, e));
    return wait_all(st, stream, FK_FENCE, 0, e);
}
// every rank's `bytes` (<= kMboxBytes, multiple of 16, device memory) land in every rank's mailboxes [parity][rank].
// Two parities suffice: a rank can only start exchange c + 2 after it has seen every rank's post for c + 1, which every
// rank issues (in stream order) after its kernel that read the mailboxes of c.
static int mbox_exchange(qm_state* st, const void* d_src, size_t bytes, int* parity_out) {
    DistCtx* d = st->dist;
    if (bytes > kMboxBytes || (bytes & 15)) return set_err(QM_ERR_ARG, "mailbox payload %zu", bytes);
    const uint32_t e = ++d->ep[FK_MBOX];
    const int parity = d->mbox_parity; d->mbox_parity ^= 1;
    k_mbox_put<<<st->world, 256, 0, st->stream>>>(wins_of(d), st->rank, parity, reinterpret_cast<const uint4*>(d_src), (uint32_t)(bytes / 16), e);
    CU(cudaGetLastError());
    st->ctr.launches++;
    if (d->shared_stream) d->lw->issued[FK_MBOX][st->rank][0].store(e, std::memory_order_release);
    QM_TRY(wait_all(st, st->stream, FK_MBOX, 0, e));
    *parity_out = parity;
    return QM_OK;
}
// in place over d_buf (room for the count rounded up to 16 bytes)
template <typename T>
static int win_allreduce(qm_state* st, T* d_buf, size_t count) {
    const size_t per = kMboxBytes / sizeof(T);
    for (size_t off = 0; off < count; off += per) {
        const size_t c = count - off < per ? count - off : per;
        int parity = 0;
        QM_TRY(mbox_exchange(st, d_buf + off, (c * sizeof(T) + 15) & ~(size_t)15, &parity));
        k_mbox_sum<T><<<(unsigned)((c + 255) / 256), 256, 0, st->stream>>>(st->dist->win[st->rank], st->world, parity, (uint32_t)c, d_buf + off);
        CU(cudaGetLastError());
        st->ctr.launches++;
    }
    return QM_OK;
}
// d_src: this rank's `count` doubles (readable up to the next multiple of 2); d_out: [world][count]
static int win_allgather(qm_state* st, const double* d_src, double* d_out, size_t count) {
    const size_t per = kMboxBytes / sizeof(double);
    for (size_t off = 0; off < count; off += per) {
        const size_t c = count - off < per ? count - off : per;
        int parity = 0;
        QM_TRY(mbox_exchange(st, d_src + off, (c * 8 + 15) & ~(size_t)15, &parity));
        k_mbox_gather<<<(unsigned)((c + 255) / 256), 256, 0, st->stream>>>(st->dist->win[st->rank], st->world, parity, (uint32_t)c, count, d_out + off);
        CU(cudaGetLastError());
        st->ctr.launches++;
    }
    return QM_OK;
}

static int allreduce_dev_f64(qm_state* st, double* d_buf, size_t count) {
    DistCtx* d = st->dist;
    if (d->have_peers) return win_allreduce<double>(st, d_buf, count);
    NC(g_nccl.AllReduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, d->comm, st->stream));
    return QM_OK;
}
static int allreduce_dev_u64(qm_state* st, unsigned long long* d_buf, size_t count) {
    DistCtx* d = st->dist;
    if (d->have_peers) return win_allreduce<unsigned long long>(st, d_buf, count);
    NC(g_nccl.AllReduce(d_buf, d_buf, count, kNcclUint64, kNcclSum, d->comm, st->stream));
    return QM_OK;
}

// ---- the exchange -------------------------------------------------------------------------------------------------
static const int kSwapSmem = 120 * 1024;     // an overlapped exchange CTA claims an SM for itself: it can share one
                                             // neither with a fused-pass CTA (> 210 KB) nor with a second exchange CTA

template <int UN, int THREADS>
static int launch_swap(qm_state* st, cudaStream_t stream, int ctas, bool exclusive, const SwapSlab& slab) {
    DistCtx* d = st->dist;
    PeerPtrs pp; for (int j = 0; j < 16; ++j) pp.p[j] = d->peer[j];
    const uint64_t chunk = 1ull << (st->nl - st->g);
    static bool attr[64] = { false };
    if (st->device < 0 || st->device >= 64 || !attr[st->device]) {
        CU(cudaFuncSetAttribute(k_remap_swap<UN, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwapSmem));
        if (st->device >= 0 && st->device < 64) attr[st->device] = true;
    }
    k_remap_swap<UN, THREADS><<<ctas, THREADS, exclusive ? kSwapSmem : 0, stream>>>(st->amp, pp, st->rank, st->world, chunk, slab);
    CU(cudaGetLastError());
    st->ctr.launches++;
    return QM_OK;
}

// relabel after an exchange: perm, the pending lowered gates, the layout flag, the traffic counters
void dist_relabel(qm_state* st, std::vector<LGate>* gates, std::vector<char>* done) {
    DistCtx* d = st->dist;
    const int g = st->g, nl = st->nl, W = st->world;
    const uint64_t chunk = 1ull << (nl - g);
    st->ctr.bytes_link += (uint64_t)(W - 1) * chunk * 16;
    st->ctr.bytes_hbm += 2ull * (uint64_t)(W - 1) * chunk * 16 * 2;
    d->swapped = !d->swapped;
    for (int q = 0; q < st->n; ++q) st->perm[q] = dist_swap_bit(st->perm[q], nl, g);
    if (gates)
        for (size_t i = 0; i < gates->size(); ++i) {
            if (done && (*done)[i]) continue;
            LGate& x = (*gates)[i];
            x.target = dist_swap_bit(x.target, nl, g);
            if (x.ctrl >= 0) x.ctrl = dist_swap_bit(x.ctrl, nl, g);
        }
}

// can the exchange run beside the passes?  (peer memory, stream waits, and enough offset bits for whole-segment slabs)
bool dist_can_overlap(const qm_state* st) {
    const DistCtx* d = st->dist;
    return d && d->have_peers && get_wait_value() && st->overlap && (st->nl - st->g - 1) >= 14;
}
void dist_next_epochs(qm_state* st, uint32_t* ep_pass, uint32_t* ep_exch) {
    *ep_pass = ++st->dist->ep[FK_PASS]; *ep_exch = ++st->dist->ep[FK_EXCH];
}
// Overlapped exchange of one slab, three calls.  (1) after the slab's last pass before the exchange has been enqueued on
// the pass stream: post "pass done" to every rank.
int dist_slab_pass_done(qm_state* st, const PassSlab& slab, uint32_t ep_pass) {
    return signal_all(st, st->stream, FK_PASS, slab.index, ep_pass);
}
// (2) on the exchange stream: wait for every rank's post, exchange the slab (a grid of st->xsm CTAs that each claim an
// SM), post "exchange done" to every rank.  slab.bit[] are positions in the shard's index; they lie below the half-chunk bit.
int dist_slab_exchange(qm_state* st, const PassSlab& slab, uint32_t ep_pass, uint32_t ep_exch, bool first, bool last) {
    DistCtx* d = st->dist;
    QM_TRY(wait_all(st, d->xstream, FK_PASS, slab.index, ep_pass));
    if (first) {
        cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        st->remap_events.push_back(std::make_pair(e0, e1));
        CU(cudaEventRecord(e0, d->xstream));
    }
    SwapSlab ss; ss.nbits = slab.nbits; ss.value_or = slab.value_or;
    for (int i = 0; i < 3; ++i) ss.bit[i] = i < slab.nbits ? slab.bit[i] : 0;
    const int ctas = st->xsm > 0 ? st->xsm : 40;
    // one CTA per claimed SM: 512 threads with 8 (or 1024 with 4) amplitude pairs in flight each.  (A variant that first
    // issues 16 remote loads per thread and fetches its own amplitudes only when they are back — 96 KB instead of 64 KB of
    // remote loads in flight per SM — was measured and dropped: 372 against 478 GB/s at 40 SMs, profiles/r2k_*.)
    if (st->xun == 4) QM_TRY((launch_swap<4, 1024>(st, d->xstream, ctas, true, ss)));
    else QM_TRY((launch_swap<8, 512>(st, d->xstream, ctas, true, ss)));
    if (last) CU(cudaEventRecord(st->remap_events.back().second, d->xstream));
    return signal_all(st, d->xstream, FK_EXCH, slab.index, ep_exch);
}
// (3) before the slab's first pass after the exchange is enqueued on the pass stream: wait for every rank's "exchange done"
int dist_slab_wait_exchanged(qm_state* st, const PassSlab& slab, uint32_t ep_exch) {
    return wait_all(st, st->stream, FK_EXCH, slab.index, ep_exch);
}

// the exchange itself, whole shard at once on the pass stream (no relabelling)
int dist_exchange_now(qm_state* st) {
    DistCtx* d = st->dist;
    if (!d) return set_err(QM_ERR_STATE, "not a sharded register");
    CU(cudaSetDevice(st->device));
    const int g = st->g, nl = st->nl, W = st->world, r = st->rank;
    const uint64_t chunk = 1ull << (nl - g);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (d->have_peers && chunk >= 2) {
        // fence (flags), one exchange kernel over peer memory on the whole grid, fence
        QM_TRY(fence_all(st, st->stream));
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        st->remap_events.push_back(std::make_pair(e0, e1));      // owned by the state from here on (qm_sync / qm_destroy)
        CU(cudaEventRecord(e0, st->stream));
        const uint64_t total = (chunk >> 1) * (uint64_t)(W - 1);
        uint64_t blocks = (total + 512ull * 4 - 1) / (512ull * 4);
        const uint64_t cap = (uint64_t)st->sms * 4; if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
        SwapSlab none; memset(&none, 0, sizeof(none));
        QM_TRY((launch_swap<4, 512>(st, st->stream, (int)blocks, false, none)));
        CU(cudaEventRecord(e1, st->stream));
        QM_TRY(fence_all(st, st->stream));
    } else {
        if (!d->comm) return set_err(QM_ERR_COMM, "no transport: peers are not mapped and there is no NCCL communicator");
        const uint64_t sub = d->stage_amps;
        if (!d->stage) CU(cudaMalloc((void**)&d->stage, (size_t)(W - 1) * sub * 16));
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        st->remap_events.push_back(std::make_pair(e0, e1));
        CU(cudaEventRecord(e0, st->stream));
        for (uint64_t off = 0; off < chunk; off += sub) {
            NC(g_nccl.GroupStart());
            int slot = 0;
            for (int j = 0; j < W; ++j) {
                if (j == r) continue;
                NC(g_nccl.Send(st->amp + (uint64_t)j * chunk + off, (size_t)sub * 2, kNcclFloat64, j, d->comm, st->stream));
                NC(g_nccl.Recv(d->stage + (uint64_t)slot * sub, (size_t)sub * 2, kNcclFloat64, j, d->comm, st->stream));
                ++slot;
            }
            NC(g_nccl.GroupEnd());
            slot = 0;
            for (int j = 0; j < W; ++j) {
                if (j == r) continue;
                CU(cudaMemcpyAsync(st->amp + (uint64_t)j * chunk + off, d->stage + (uint64_t)slot * sub, (size_t)sub * 16,
                                   cudaMemcpyDeviceToDevice, st->stream));
                ++slot;
            }
        }
        CU(cudaEventRecord(e1, st->stream));
        st->ctr.launches += (uint64_t)((chunk + sub - 1) / sub);
    }
    return QM_OK;
}

int dist_remap(qm_state* st, std::vector<LGate>* gates, std::vector<char>* done) {
    QM_TRY(dist_exchange_now(st));
    dist_relabel(st, gates, done);
    return QM_OK;
}

// qm_reset rewrites the basis state in the canonical layout and sets perm to the identity: the layout flag follows
void dist_reset(qm_state* st) { if (st->dist) st->dist->swapped = false; }

int dist_canonicalize(qm_state* st) {
    if (!st->dist || !st->dist->swapped) return QM_OK;
    return dist_remap(st, nullptr, nullptr);
}

// sum `count` doubles over the ranks (host in/out); goes through a small device buffer
static int allreduce_host(qm_state* st, double* v, int count) {
    DistCtx* d = st->dist;
    if (count > 4096) return set_err(QM_ERR_ARG, "allreduce too large");
    CU(cudaMemcpyAsync(d->d_small, v, count * sizeof(double), cudaMemcpyHostToDevice, st->stream));
    QM_TRY(allreduce_dev_f64(st, d->d_small, (size_t)count));
    CU(cudaMemcpyAsync(v, d->d_small, count * sizeof(double), cudaMemcpyDeviceToHost, st->stream));
    CU(cudaStreamSynchronize(st->stream));
    return QM_OK;
}

// pairs: local [nl][2] (p0, p1) per physical local bit, tot: local kept mass.  out: [n][2] per LOGICAL qubit.
int dist_finish_z(qm_state* st, const double* pairs, const double* tot, double* out_host) {
    const int n = st->n, nl = st->nl, g = st->g;
    std::vector<double> v(n + 1, 0.0);
    for (int b = 0; b < nl; ++b) v[b] = pairs[2 * b + 1];                       // kept mass with physical bit b set
    for (int j = 0; j < g; ++j) v[nl + j] = ((st->rank >> j) & 1) ? tot[0] : 0.0;
    v[n] = tot[0];
    QM_TRY(allreduce_host(st, v.data(), n + 1));
    for (int q = 0; q < n; ++q) {
        const double p1 = v[st->perm[q]];
        out_host[2 * q] = v[n] - p1; out_host[2 * q + 1] = p1;
    }
    return QM_OK;
}

int dist_finish_norm(qm_state* st, double local_tot, double* out) {
    double v = local_tot;
    QM_TRY(allreduce_host(st, &v, 1));
    *out = v;
    return QM_OK;
}

int dist_measure(qm_state* st, int t0, double thr, const double* R, double out[2]) {
    int tb = st->perm[t0];
    if (tb >= st->nl && R) {            // a rotated basis needs the pair on one rank: bring the qubit home
        QM_TRY(dist_remap(st, nullptr, nullptr));
        tb = st->perm[t0];
    }
    double v[2];
    if (tb < st->nl) QM_TRY(qm_measure_phys(st, tb, thr, 0, R, v));
    else {
        double *pairs, *tot;
        QM_TRY(qm_zstats(st, thr, &pairs, &tot));
        const bool set = (st->rank >> (tb - st->nl)) & 1;
        v[0] = set ? 0.0 : tot[0]; v[1] = set ? tot[0] : 0.0;
    }
    QM_TRY(allreduce_host(st, v, 2));
    out[0] = v[0]; out[1] = v[1];
    return QM_OK;
}

// Reduced density of a sharded register.  Keeping the LOW k qubits: once they sit on local bits 0..k-1 every rank
// holds whole rows of Psi (A = the other qubits, rank bits included), so rho_B is the sum over ranks of the local
// Psi^T conj(Psi) — the single-device kernel on the shard, then one all-reduce of 4^k entries (QInfo.jl:385-391
// with rho = psi psi').  Keeping the HIGH k qubits would need products of amplitudes that live on different ranks;
// the kept qubits are first swapped with the low ones (bit-exact exchanges through the ordinary circuit path,
// remaps included), which turns the request into the case above with the same index order, and swapped back.
int dist_reduced_density(qm_state* st, int keep_low, int k, double* out) {
    if (k > st->nl) return set_err(QM_ERR_ARG, "Dimension mismatch: k = %d exceeds the %d local bits of a sharded register", k, st->nl);
    if (!keep_low) {
        if (2 * k > st->n) return set_err(QM_ERR_ARG, "sharded register: keeping the high k qubits needs k <= n/2 (k = %d, n = %d)", k, st->n);
        for (int j = 0; j < k; ++j) QM_TRY(qm_apply_swap(st, st->n - k + j, j));
        const int rc = dist_reduced_density(st, 1, k, out);
        for (int j = 0; j < k; ++j) QM_TRY(qm_apply_swap(st, st->n - k + j, j));     // queued; runs with the next circuit
        return rc;
    }
    QM_TRY(qm_flush(st));
    CU(cudaSetDevice(st->device));
    bool home = true; for (int j = 0; j < k; ++j) home = home && st->perm[j] == j;
    if (!home) {
        QM_TRY(dist_canonicalize(st));
        for (int j = 0; j < k; ++j) if (st->perm[j] != j) return set_err(QM_ERR_STATE, "kept qubit %d is not on its home bit", j);
    }
    double2* d_out = nullptr;
    QM_TRY(qm_reduced_local(st, 1, k, 0, &d_out));
    const size_t entries = (size_t)1 << (2 * k);
    QM_TRY(allreduce_dev_f64(st, reinterpret_cast<double*>(d_out), entries * 2));
    CU(cudaMemcpyAsync(st->h_scratch, d_out, entries * 16, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += entries * 16;
    QM_TRY(qm_sync(st));
    memcpy(out, st->h_scratch, entries * 16);
    return QM_OK;
}

// Sampler of a sharded register, bit for bit the single-device one (oracle/qsim_oracle.c orc_sample): in canonical
// layout the leaves (4096 consecutive amplitudes) of rank r are leaves r * 2^(nl-12) ... of the whole register.  Every
// rank sums its own leaves, the level-1 array is all-gathered (<= 2^(n-12) doubles: 128 MiB at 36 qubits), every rank
// rebuilds the upper levels and walks every shot down to its leaf identically; the rank that owns the leaf finishes
// inside it, the others contribute 0 to a sum over ranks.
int dist_sample(qm_state* st, uint64_t seed, int nshots, uint64_t* out) {
    DistCtx* d = st->dist;
    if (st->nl < 12) return set_err(QM_ERR_STATE, "sampling a sharded register needs >= 12 local bits");
    QM_TRY(qm_flush(st));
    CU(cudaSetDevice(st->device));
    QM_TRY(dist_canonicalize(st));
    const uint64_t n1l = local_amps(st) >> 12, n1 = n1l * (uint64_t)st->world;
    const uint64_t n2 = (n1 + 4095) >> 12, n3 = (n2 + 4095) >> 12;
    // [level 1, all ranks][levels 2, 3 (+ pad to an even count)][this rank's level 1: the mailbox source][shots]
    const uint64_t lev = (n1 + n2 + n3 + 1) & ~1ull, minepad = (n1l + 1) & ~1ull;
    QM_TRY(ENSURE_DEV(st, d_scratch, (lev + minepad) * 8 + (((size_t)nshots + 1) & ~(size_t)1) * 8 + 64));
    QM_TRY(ENSURE_PIN(st, h_scratch, (size_t)nshots * 8));
    double* s1 = st->d_scratch;
    double* mine = s1 + lev;
    uint64_t* d_out = reinterpret_cast<uint64_t*>(mine + minepad);
    if (d->have_peers) {
        QM_TRY(qm_sample_leaves(st, 0, mine));
        QM_TRY(win_allgather(st, mine, s1, (size_t)n1l));
    } else {
        QM_TRY(qm_sample_leaves(st, 0, s1 + (uint64_t)st->rank * n1l));
        NC(g_nccl.AllGather(s1 + (uint64_t)st->rank * n1l, s1, (size_t)n1l, kNcclFloat64, d->comm, st->stream));
    }
    QM_TRY(qm_sample_descend(st, 0, s1, n1, (uint64_t)st->rank * n1l, seed, nshots, s1 + n1, d_out));
    QM_TRY(allreduce_dev_u64(st, reinterpret_cast<unsigned long long*>(d_out), (size_t)nshots));
    CU(cudaMemcpyAsync(st->h_scratch, d_out, (size_t)nshots * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += (size_t)nshots * 8;
    QM_TRY(qm_sync(st));
    memcpy(out, st->h_scratch, (size_t)nshots * 8);
    return QM_OK;
}

// ---- peer-memory mapping (CUDA IPC): every rank exports its allocation (shard + window), the host all-gathers the
// 64-byte handles, every rank imports the others'.  Optional: without it everything goes through NCCL.
extern "C" int32_t qm_dist_ipc_export(qm_state* st, uint8_t handle_out[64]) {
    if (!st || !st->dist || !handle_out) return set_err(QM_ERR_ARG, "not a sharded register");
    CU(cudaSetDevice(st->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, st->amp));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, 64);
    return QM_OK;
}

extern "C" int32_t qm_dist_ipc_import(qm_state* st, const uint8_t* handles /* [world][64] */) {
    if (!st || !st->dist || !handles) return set_err(QM_ERR_ARG, "not a sharded register");
    if (st->world > 16) return set_err(QM_ERR_ARG, "at most 16 ranks");
    CU(cudaSetDevice(st->device));
    DistCtx* d = st->dist;
    if (!get_wait_value()) return set_err(QM_ERR_COMM, "cuStreamWaitValue32 is not available — exchanges stay on NCCL");
    for (int j = 0; j < st->world; ++j) {
        if (j == st->rank) continue;
        cudaIpcMemHandle_t h; memcpy(&h, handles + 64 * j, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k < 16; ++k) {
                if (d->peer[k]) { cudaIpcCloseMemHandle(d->peer[k]); d->peer[k] = nullptr; }
                if (k != st->rank) d->win[k] = nullptr;
            }
            d->have_peers = false; d->ipc_mapped = false;
            return set_err(QM_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s — exchanges stay on NCCL", j, cudaGetErrorString(e));
        }
        d->peer[j] = reinterpret_cast<double2*>(p);
        d->win[j] = reinterpret_cast<unsigned char*>(p) + (size_t)local_amps(st) * 16;
        d->ipc_mapped = true;
    }
    d->have_peers = true;
    return QM_OK;
}
