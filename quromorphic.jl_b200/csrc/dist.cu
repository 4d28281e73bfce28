// dist.cu — sharded registers.  Rank r of G = 2^g holds the 2^(n-g) amplitudes whose top g index bits
// equal r.  Everything that does not need a non-diagonal target on a global qubit runs locally (a
// control or a diagonal target on a global qubit is a rank predicate / rank-dependent phase inside
// the ordinary kernels).  When the planner runs out of local work, the g global qubits are exchanged
// with the top g local ones: viewed as [2^g chunks][2^(n-2g)], chunk j of rank r trades places with
// chunk r of rank j — an all-to-all done IN PLACE through a small staging ring, because a 2^33
// amplitude shard (128 GiB) leaves no room for an out-of-place copy on a 180 GB part.
// NCCL (grouped ncclSend/ncclRecv over NVLink 5 / NVSwitch) is dlopen'ed on first use.
#include "dist.hpp"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

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

// ---------------------------------------------------------------------------------------------
// K7  global<->local qubit exchange over peer memory.  Chunk j of rank r trades places with chunk r
// of rank j.  Every rank runs ONE kernel: for each peer it owns half of the chunk pair (the lower
// rank the first half, the higher rank the second), loads its own 16 bytes and the peer's 16 bytes
// (a load over NVLink), and stores them crosswise (a store over NVLink).  In place, no staging, both
// link directions busy; each amplitude pair is touched by exactly one thread of one rank.
// Ranks are fenced before and after by a one-element NCCL all-reduce on the same stream (the kernel
// itself never waits on another GPU).
// ---------------------------------------------------------------------------------------------
struct PeerPtrs { double2* p[16]; };

// Work is cut into 64 KiB segments that are dealt round-robin over the peers, and every rank starts
// its round at a different peer ((rank + 1 + k) mod world), so at any moment each GPU's NVLink ingress
// is fed by all the others evenly instead of everyone hammering the same peer.
__global__ void __launch_bounds__(512)
k_remap_swap(double2* __restrict__ mine, PeerPtrs peers, int rank, int world, unsigned long long chunk) {
    const unsigned long long half = chunk >> 1;
    const unsigned long long SEG = half < 4096ull ? half : 4096ull;       // amplitudes per segment (64 KiB)
    const unsigned long long segs_per_peer = half / SEG;
    const unsigned long long total = half * (unsigned long long)(world - 1);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 4 * stride) {
        double2 a[4], b[4]; unsigned long long mo[4], po[4]; double2* pp[4]; bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned long long ii = i + (unsigned long long)u * stride;
            ok[u] = ii < total;
            const unsigned long long q = ok[u] ? ii : 0;
            const unsigned long long sg = q / SEG, within = q % SEG;
            const int k = (int)(sg % (unsigned long long)(world - 1));             // which peer, round-robin
            const unsigned long long o = (sg / (unsigned long long)(world - 1)) * SEG + within;
            int j = rank + 1 + k; if (j >= world) j -= world;                      // rank-rotated peer order
            const unsigned long long off = (rank < j ? 0ull : half) + o;
            mo[u] = (unsigned long long)j * chunk + off; po[u] = (unsigned long long)rank * chunk + off; pp[u] = peers.p[j];
            if (ok[u]) { a[u] = mine[mo[u]]; b[u] = pp[u][po[u]]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (ok[u]) { mine[mo[u]] = b[u]; pp[u][po[u]] = a[u]; }
    }
    (void)segs_per_peer;
}

struct DistCtx {
    double2* peer[16] = { nullptr };                    // IPC-mapped shards of the other ranks (null: NCCL path)
    bool have_peers = false;
    ncclComm_t comm = nullptr;
    double2* stage = nullptr; size_t stage_amps = 0;     // (world - 1) x sub-chunk
    double* d_small = nullptr;                           // all-reduce scratch
    bool swapped = false;                                // layout B: global qubits currently sit on the top local bits
};

int dist_destroy(DistCtx* d) {
    if (!d) return QM_OK;
    for (int j = 0; j < 16; ++j) if (d->peer[j]) cudaIpcCloseMemHandle(d->peer[j]);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    cudaFree(d->stage); cudaFree(d->d_small);
    delete d;
    return QM_OK;
}

extern "C" int32_t qm_dist_unique_id(uint8_t id_out[128]) {
    if (!id_out) return set_err(QM_ERR_ARG, "null argument");
    QM_TRY(nccl_load());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return QM_OK;
}

extern "C" int32_t qm_create_dist(int32_t n, uint64_t basis, int32_t device, int32_t rank, int32_t world,
                                  const uint8_t nccl_id[128], qm_state** out) {
    if (!nccl_id || !out) return set_err(QM_ERR_ARG, "null argument");
    if (world < 2) return set_err(QM_ERR_ARG, "world must be >= 2 (use qm_create for one GPU)");
    QM_TRY(nccl_load());
    QM_TRY(qm_create_common(n, basis, 1, device, rank, world, out));
    qm_state* st = *out;
    if (2 * st->g > st->n) { qm_destroy(st); *out = nullptr; return set_err(QM_ERR_ARG, "register too small for %d ranks", world); }
    DistCtx* d = new DistCtx();
    ncclUniqueId id; memcpy(id.internal, nccl_id, 128);
    int r = g_nccl.CommInitRank(&d->comm, world, id, rank);
    if (r != 0) { delete d; qm_destroy(st); *out = nullptr; return set_err(QM_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    // staging ring: (world - 1) sub-chunks of at most 2^24 amplitudes (256 MiB) each
    const uint64_t chunk = 1ull << (st->nl - st->g);
    uint64_t sub = chunk < (1ull << 24) ? chunk : (1ull << 24);
    d->stage_amps = sub;
    CU(cudaMalloc((void**)&d->stage, (size_t)(world - 1) * sub * 16));
    CU(cudaMalloc((void**)&d->d_small, 4096 * sizeof(double)));
    st->dist = d;
    return QM_OK;
}

int dist_remap(qm_state* st, std::vector<LGate>* gates, std::vector<char>* done) {
    DistCtx* d = st->dist;
    if (!d) return set_err(QM_ERR_STATE, "not a sharded register");
    CU(cudaSetDevice(st->device));
    const int g = st->g, nl = st->nl, W = st->world, r = st->rank;
    const uint64_t chunk = 1ull << (nl - g), sub = d->stage_amps;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    if (d->have_peers && chunk >= 2) {
        // fence, one exchange kernel over peer memory, fence
        NC(g_nccl.AllReduce(d->d_small, d->d_small, 1, kNcclFloat64, kNcclSum, d->comm, st->stream));
        CU(cudaEventRecord(e0, st->stream));
        PeerPtrs pp; for (int j = 0; j < 16; ++j) pp.p[j] = d->peer[j];
        const uint64_t total = (chunk >> 1) * (uint64_t)(W - 1);
        uint64_t blocks = (total + 512ull * 4 - 1) / (512ull * 4);
        const uint64_t cap = (uint64_t)st->sms * 4; if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
        k_remap_swap<<<(unsigned)blocks, 512, 0, st->stream>>>(st->amp, pp, r, W, chunk);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e1, st->stream));
        NC(g_nccl.AllReduce(d->d_small, d->d_small, 1, kNcclFloat64, kNcclSum, d->comm, st->stream));
        st->ctr.launches += 1;
    } else {
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
    st->remap_events.push_back(std::make_pair(e0, e1));
    st->ctr.bytes_link += (uint64_t)(W - 1) * chunk * 16;
    st->ctr.bytes_hbm += 2ull * (uint64_t)(W - 1) * chunk * 16 * 2;      // send read + stage write, stage read + place write
    d->swapped = !d->swapped;
    for (int q = 0; q < st->n; ++q) st->perm[q] = dist_swap_bit(st->perm[q], nl, g);
    if (gates)
        for (size_t i = 0; i < gates->size(); ++i) {
            if (done && (*done)[i]) continue;
            LGate& x = (*gates)[i];
            x.target = dist_swap_bit(x.target, nl, g);
            if (x.ctrl >= 0) x.ctrl = dist_swap_bit(x.ctrl, nl, g);
        }
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
    NC(g_nccl.AllReduce(d->d_small, d->d_small, (size_t)count, kNcclFloat64, kNcclSum, d->comm, st->stream));
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
    DistCtx* d = st->dist;
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
    NC(g_nccl.AllReduce(d_out, d_out, entries * 2, kNcclFloat64, kNcclSum, d->comm, st->stream));
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
    QM_TRY(ENSURE_DEV(st, d_scratch, (n1 + n2 + n3) * 8 + (size_t)nshots * 8 + 64));
    QM_TRY(ENSURE_PIN(st, h_scratch, (size_t)nshots * 8));
    double* s1 = st->d_scratch;
    uint64_t* d_out = reinterpret_cast<uint64_t*>(s1 + n1 + n2 + n3);
    QM_TRY(qm_sample_leaves(st, 0, s1 + (uint64_t)st->rank * n1l));
    NC(g_nccl.AllGather(s1 + (uint64_t)st->rank * n1l, s1, (size_t)n1l, kNcclFloat64, d->comm, st->stream));
    QM_TRY(qm_sample_descend(st, 0, s1, n1, (uint64_t)st->rank * n1l, seed, nshots, s1 + n1, d_out));
    NC(g_nccl.AllReduce(d_out, d_out, (size_t)nshots, kNcclUint64, kNcclSum, d->comm, st->stream));
    CU(cudaMemcpyAsync(st->h_scratch, d_out, (size_t)nshots * 8, cudaMemcpyDeviceToHost, st->stream));
    st->ctr.bytes_d2h += (size_t)nshots * 8;
    QM_TRY(qm_sync(st));
    memcpy(out, st->h_scratch, (size_t)nshots * 8);
    return QM_OK;
}

// ---- peer-memory mapping (CUDA IPC): every rank exports its shard, the host all-gathers the 64-byte
// handles, every rank imports the others'.  Optional: without it the remap goes through NCCL.
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
    for (int j = 0; j < st->world; ++j) {
        if (j == st->rank) continue;
        cudaIpcMemHandle_t h; memcpy(&h, handles + 64 * j, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k < 16; ++k) if (d->peer[k]) { cudaIpcCloseMemHandle(d->peer[k]); d->peer[k] = nullptr; }
            d->have_peers = false;
            return set_err(QM_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s — remaps stay on NCCL", j, cudaGetErrorString(e));
        }
        d->peer[j] = reinterpret_cast<double2*>(p);
    }
    d->have_peers = true;
    return QM_OK;
}
