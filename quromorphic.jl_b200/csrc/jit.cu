// jit.cu — compile / cache / launch of the straight-line pass kernels (jit.hpp, jit_codegen.hpp).
#include "jit.hpp"

#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>

#include "jit_codegen.hpp"
#include "state.hpp"

namespace qm {

// the JIT templates: PTX text of k_fused_tma<M, TEAMS, STAGES, true>, produced by the build (Makefile: fused_jit.cu -> .ptx -> .inc)
static const char kTemplateM12[] =
#include "build/jit_template_m12.inc"
    ;

static const char kTemplateM12Z[] =              // ... and its ZFOLD instantiation (the pass also takes the all-qubit <Z> sums)
#include "build/jit_template_m12z.inc"
    ;

static const std::string& template_for(int m, bool zfold = false) {
    static const std::string t12(kTemplateM12), t12z(kTemplateM12Z), none;
    return m != 12 ? none : (zfold ? t12z : t12);
}

// ---- nvJitLink (optional, dlopen) --------------------------------------------------------------
typedef int (*PFN_jlCreate)(void**, uint32_t, const char**);
typedef int (*PFN_jlDestroy)(void**);
typedef int (*PFN_jlAddData)(void*, int, const void*, size_t, const char*);
typedef int (*PFN_jlComplete)(void*);
typedef int (*PFN_jlGetSize)(void*, size_t*);
typedef int (*PFN_jlGet)(void*, void*);
struct JitLinkApi {
    void* lib = nullptr;
    PFN_jlCreate create = nullptr; PFN_jlDestroy destroy = nullptr; PFN_jlAddData add = nullptr; PFN_jlComplete complete = nullptr;
    PFN_jlGetSize cubin_size = nullptr, log_size = nullptr; PFN_jlGet cubin = nullptr, log = nullptr;
    unsigned ver[2] = { 0, 0 };                 // nvJitLinkVersion (part of the disk-cache key: another assembler, another cubin)
    bool ok = false;
};
static const JitLinkApi& jitlink() {
    static JitLinkApi A; static std::once_flag once;
    std::call_once(once, [] {
        // the toolkit's own copy first: a process that has imported torch already holds torch's bundled libnvJitLink (an older
        // minor version) under the same soname, and a bare soname would resolve to that one
        const char* names[] = { "/usr/local/cuda/lib64/libnvJitLink.so.12", "libnvJitLink.so.12", "/usr/local/cuda/lib64/libnvJitLink.so.13",
                                "libnvJitLink.so.13", "libnvJitLink.so" };
        const char* be = getenv("QM_JIT_BACKEND");            // "driver": skip nvJitLink, hand the PTX text to the driver's compiler
        if (be && strcmp(be, "driver") == 0) return;
        const char* env = getenv("QM_NVJITLINK");
        if (env && *env) A.lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        for (size_t i = 0; !A.lib && i < sizeof(names) / sizeof(names[0]); ++i) A.lib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
        if (!A.lib) return;
        A.create = (PFN_jlCreate)dlsym(A.lib, "nvJitLinkCreate"); A.destroy = (PFN_jlDestroy)dlsym(A.lib, "nvJitLinkDestroy");
        A.add = (PFN_jlAddData)dlsym(A.lib, "nvJitLinkAddData"); A.complete = (PFN_jlComplete)dlsym(A.lib, "nvJitLinkComplete");
        A.cubin_size = (PFN_jlGetSize)dlsym(A.lib, "nvJitLinkGetLinkedCubinSize"); A.cubin = (PFN_jlGet)dlsym(A.lib, "nvJitLinkGetLinkedCubin");
        A.log_size = (PFN_jlGetSize)dlsym(A.lib, "nvJitLinkGetErrorLogSize"); A.log = (PFN_jlGet)dlsym(A.lib, "nvJitLinkGetErrorLog");
        A.ok = A.create && A.destroy && A.add && A.complete && A.cubin_size && A.cubin;
        typedef int (*PFN_jlVersion)(unsigned*, unsigned*);
        if (PFN_jlVersion v = (PFN_jlVersion)dlsym(A.lib, "nvJitLinkVersion")) v(&A.ver[0], &A.ver[1]);
    });
    return A;
}

bool jit_ptx_for(int m, const V4Program& prog, std::string& ptx, bool zfold) {
    const std::string& t = template_for(m, zfold);
    ptx.clear();
    if (t.empty()) return false;
    return jit_build_ptx(t, prog, ptx);
}

bool jit_compile_cubin(const std::string& ptx, std::vector<char>& cubin, std::string& log) {
    const JitLinkApi& A = jitlink();
    cubin.clear(); log.clear();
    if (!A.ok) { log = "libnvJitLink not found"; return false; }
    void* h = nullptr;
    const char* opts[] = { "-arch=sm_100a", "-O3" };
    if (A.create(&h, 2, opts) != 0 || !h) { log = "nvJitLinkCreate failed"; return false; }
    bool ok = A.add(h, 2 /* NVJITLINK_INPUT_PTX */, ptx.data(), ptx.size(), "qm_jit_pass") == 0 && A.complete(h) == 0;
    size_t sz = 0;
    if (ok) ok = A.cubin_size(h, &sz) == 0 && sz > 0;
    if (ok) { cubin.resize(sz); ok = A.cubin(h, cubin.data()) == 0; }
    if (!ok && A.log_size && A.log) {
        size_t ls = 0;
        if (A.log_size(h, &ls) == 0 && ls > 0) { log.resize(ls + 1); A.log(h, &log[0]); }
    }
    A.destroy(&h);
    return ok;
}

// ---- disk cache of assembled kernels (optional) ---------------------------------------------------
// QM_JIT_CACHE_DIR / qm_jit_cache_dir(): cubins are kept as files keyed by a 128-bit hash of the COMPLETE PTX text handed to
// the assembler (template + generated groups: a rebuilt library or a changed generator can never meet a stale file) and of
// the assembler's version.  A later process that plans the same circuit structure loads them instead of assembling again
// (~0.25 s of one core per kernel; C3 has 110).  Files are written under a temporary name and renamed, so concurrent
// processes (one per GPU) can share a directory; a file that does not verify (size, hashes, checksum) is ignored.
static std::mutex g_dir_mu;
static std::string g_dir; static bool g_dir_set = false;
static std::atomic<uint64_t> g_disk_hits{0}, g_disk_stores{0};

static std::string disk_dir() {
    std::lock_guard<std::mutex> lk(g_dir_mu);
    if (!g_dir_set) { const char* e = getenv("QM_JIT_CACHE_DIR"); g_dir = e ? e : ""; g_dir_set = true; }
    return g_dir;
}
void jit_set_cache_dir(const char* path) {
    std::lock_guard<std::mutex> lk(g_dir_mu);
    g_dir = path ? path : ""; g_dir_set = true;
}
void jit_disk_stats(uint64_t out[2]) { out[0] = g_disk_hits.load(); out[1] = g_disk_stores.load(); }

static void hash128(const void* data, size_t n, uint64_t h[2]) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x9e3779b97f4a7c15ull;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w; memcpy(&w, p + i, 8);
        h1 = (h1 ^ w) * 0x100000001b3ull; h1 ^= h1 >> 29;
        h2 ^= w + 0x9e3779b97f4a7c15ull + (h2 << 6) + (h2 >> 2); h2 *= 0xff51afd7ed558ccdull; h2 ^= h2 >> 33;
    }
    for (; i < n; ++i) { h1 = (h1 ^ p[i]) * 0x100000001b3ull; h2 = (h2 ^ p[i]) * 0xc4ceb9fe1a85ec53ull; h2 ^= h2 >> 31; }
    h[0] = h1; h[1] = h2 ^ ((uint64_t)n << 40);
}

struct DiskHeader { uint64_t magic, ptx_h[2], ptx_len, ver, cubin_len, cubin_h[2]; };
static const uint64_t kDiskMagic = 0x31434a4d51ull;      // "QMJC1"

static std::string disk_path(const std::string& dir, const uint64_t h[2], uint64_t ver) {
    char name[96];
    snprintf(name, sizeof(name), "/qmjit-%016llx%016llx-%llx.cubin", (unsigned long long)h[0], (unsigned long long)h[1], (unsigned long long)ver);
    return dir + name;
}

static bool disk_load(const std::string& ptx, std::vector<char>& cubin) {
    const std::string dir = disk_dir();
    if (dir.empty()) return false;
    uint64_t h[2]; hash128(ptx.data(), ptx.size(), h);
    const JitLinkApi& A = jitlink();
    const uint64_t ver = ((uint64_t)A.ver[0] << 32) | A.ver[1];
    FILE* f = fopen(disk_path(dir, h, ver).c_str(), "rb");
    if (!f) return false;
    DiskHeader H; bool ok = fread(&H, sizeof(H), 1, f) == 1 && H.magic == kDiskMagic && H.ptx_h[0] == h[0] && H.ptx_h[1] == h[1] &&
                            H.ptx_len == ptx.size() && H.ver == ver && H.cubin_len > 0 && H.cubin_len < (64u << 20);
    if (ok) { cubin.resize((size_t)H.cubin_len); ok = fread(cubin.data(), 1, cubin.size(), f) == cubin.size() && fgetc(f) == EOF; }
    fclose(f);
    if (ok) { uint64_t c[2]; hash128(cubin.data(), cubin.size(), c); ok = c[0] == H.cubin_h[0] && c[1] == H.cubin_h[1]; }
    if (!ok) cubin.clear(); else g_disk_hits++;
    return ok;
}

static void disk_store(const std::string& ptx, const std::vector<char>& cubin) {
    const std::string dir = disk_dir();
    if (dir.empty() || cubin.empty()) return;
    mkdir(dir.c_str(), 0777);                              // one level; an existing directory is fine
    DiskHeader H; memset(&H, 0, sizeof(H));
    H.magic = kDiskMagic; hash128(ptx.data(), ptx.size(), H.ptx_h); H.ptx_len = ptx.size();
    const JitLinkApi& A = jitlink();
    H.ver = ((uint64_t)A.ver[0] << 32) | A.ver[1];
    H.cubin_len = cubin.size(); hash128(cubin.data(), cubin.size(), H.cubin_h);
    const std::string path = disk_path(dir, H.ptx_h, H.ver);
    static std::atomic<uint64_t> serial{0};
    char suffix[64]; snprintf(suffix, sizeof(suffix), ".tmp.%ld.%llu", (long)getpid(), (unsigned long long)serial++);
    const std::string tmp = path + suffix;
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                         // a directory that cannot be written is not an error: no cache
    const bool ok = fwrite(&H, sizeof(H), 1, f) == 1 && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    if (fclose(f) != 0 || !ok || rename(tmp.c_str(), path.c_str()) != 0) { remove(tmp.c_str()); return; }
    g_disk_stores++;
}

bool jit_cubin_for(const std::string& ptx, std::vector<char>& cubin, std::string& log, bool* from_disk) {
    if (from_disk) *from_disk = false;
    if (!jitlink().ok) { cubin.clear(); log = "libnvJitLink not found"; return false; }
    if (disk_load(ptx, cubin)) { if (from_disk) *from_disk = true; log.clear(); return true; }
    if (!jit_compile_cubin(ptx, cubin, log)) return false;
    disk_store(ptx, cubin);
    return true;
}

// ---- driver API (entry points through the runtime: the library never links libcuda) --------------
typedef CUresult (*PFN_cuModuleLoadData)(CUmodule*, const void*);
typedef CUresult (*PFN_cuModuleLoadDataEx)(CUmodule*, const void*, unsigned int, CUjit_option*, void**);
typedef CUresult (*PFN_cuModuleGetFunction)(CUfunction*, CUmodule, const char*);
typedef CUresult (*PFN_cuFuncSetAttribute)(CUfunction, CUfunction_attribute, int);
typedef CUresult (*PFN_cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
struct DriverApi {
    PFN_cuModuleLoadData load = nullptr; PFN_cuModuleLoadDataEx load_ex = nullptr; PFN_cuModuleGetFunction get_fn = nullptr;
    PFN_cuFuncSetAttribute set_attr = nullptr; PFN_cuLaunchKernel launch = nullptr;
    bool ok = false;
};
static const DriverApi& driver() {
    static DriverApi D; static std::once_flag once;
    std::call_once(once, [] {
        auto ep = [](const char* name) -> void* {
            void* p = nullptr; cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
            return p;
        };
        D.load = (PFN_cuModuleLoadData)ep("cuModuleLoadData"); D.load_ex = (PFN_cuModuleLoadDataEx)ep("cuModuleLoadDataEx");
        D.get_fn = (PFN_cuModuleGetFunction)ep("cuModuleGetFunction"); D.set_attr = (PFN_cuFuncSetAttribute)ep("cuFuncSetAttribute");
        D.launch = (PFN_cuLaunchKernel)ep("cuLaunchKernel");
        D.ok = D.load && D.load_ex && D.get_fn && D.set_attr && D.launch;
    });
    return D;
}

struct JitKernel { CUmodule mod = nullptr; CUfunction fn = nullptr; };

static size_t smem_for(int m) {
    const int stages = m == 12 ? 3 : 6;
    return (size_t)stages * (16u << m) + kV4TabBytes + sizeof(V4Rec) * kV4MaxRecs + 128 + 3 * stages * 8 + 1024;
}

// ---- cache ---------------------------------------------------------------------------------------
struct JitKey {
    int device, m; uint64_t h[2];
    bool operator<(const JitKey& o) const { return memcmp(this, &o, sizeof(JitKey)) < 0; }
};
struct JitEntry { int seen = 0; int state = 0 /* 0 not compiled, 1 ready, 2 failed */; JitKernel k; std::vector<uint32_t> sig; };
static std::mutex g_mu;
static std::map<JitKey, JitEntry> g_cache;
static JitStats g_stats;
static const size_t kMaxKernels = 4096;

JitStats jit_stats() { std::lock_guard<std::mutex> lk(g_mu); JitStats s = g_stats; s.cached = 0; for (auto& e : g_cache) s.cached += e.second.state == 1; return s; }

static JitKey make_key(int device, int m, const V4Program& prog, std::vector<uint32_t>& sig, bool zfold) {
    jit_structure_bytes(prog, m, sig);
    if (zfold) sig.push_back(0x5a464f4cu);          // "ZFOL": another kernel for the same opcode sequence
    JitKey k; memset(&k, 0, sizeof(k)); k.device = device; k.m = m; jit_hash(sig, k.h);
    return k;
}

// cubin (or PTX text when nvJitLink is missing) -> module + function, on the calling thread's current context
static bool load_kernel(const std::vector<char>& image, bool is_ptx, int m, JitKernel& out, bool zfold) {
    const DriverApi& D = driver();
    if (!D.ok) return false;
    CUmodule mod = nullptr;
    CUresult r = is_ptx ? D.load_ex(&mod, image.data(), 0, nullptr, nullptr) : D.load(&mod, image.data());
    if (r != CUDA_SUCCESS || !mod) return false;
    const std::string entry = jit_entry_name(template_for(m, zfold));
    CUfunction fn = nullptr;
    if (D.get_fn(&fn, mod, entry.c_str()) != CUDA_SUCCESS || !fn) return false;
    if (D.set_attr(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_for(m)) != CUDA_SUCCESS) return false;
    out.mod = mod; out.fn = fn;
    return true;
}

void jit_get_many(int device, int m, const std::vector<const V4Program*>& progs, int threshold, std::vector<JitKernel*>& out, bool zfold) {
    out.assign(progs.size(), nullptr);
    if (threshold <= 0 || template_for(m, zfold).empty() || !driver().ok) return;
    // which programs need compiling now (one compile per distinct opcode sequence)
    struct Job { JitKey key; const V4Program* prog; std::vector<char> image; std::string ptx; bool is_ptx = false, ok = false, from_disk = false; };
    std::vector<Job> jobs;
    std::vector<JitKey> keys(progs.size());
    {
        std::lock_guard<std::mutex> lk(g_mu);
        // sequences that were seen but never came back (one-shot circuits) must not pile up in a long-running process
        if (g_cache.size() > 65536)
            for (auto it = g_cache.begin(); it != g_cache.end();) { if (it->second.state == 0) it = g_cache.erase(it); else ++it; }
        std::vector<uint32_t> sig;
        for (size_t i = 0; i < progs.size(); ++i) {
            keys[i] = make_key(device, m, *progs[i], sig, zfold);
            JitEntry& e = g_cache[keys[i]];
            if (e.sig.empty()) e.sig = sig;
            else if (e.sig != sig) { keys[i].m = -1; continue; }        // hash collision: leave this program to the interpreter
            e.seen++;
            if (e.state == 0 && e.seen >= threshold && g_stats.compiled + jobs.size() < kMaxKernels) {
                bool dup = false;
                for (const Job& j : jobs) if (!(j.key < keys[i]) && !(keys[i] < j.key)) { dup = true; break; }
                if (!dup) { Job j; j.key = keys[i]; j.prog = progs[i]; jobs.push_back(std::move(j)); }
            }
        }
    }
    if (!jobs.empty()) {
        const auto t0 = std::chrono::steady_clock::now();
        const bool have_jl = jitlink().ok;
        int nth = (int)std::thread::hardware_concurrency(); if (nth < 1) nth = 1; if (nth > 32) nth = 32;
        if ((size_t)nth > jobs.size()) nth = (int)jobs.size();
        auto work = [&](int t) {
            for (size_t i = (size_t)t; i < jobs.size(); i += (size_t)nth) {
                Job& j = jobs[i];
                std::string& ptx = j.ptx; std::string log;
                if (!jit_build_ptx(template_for(m, zfold), *j.prog, ptx)) continue;
                if (have_jl) j.ok = jit_cubin_for(ptx, j.image, log, &j.from_disk);
                else { j.image.assign(ptx.begin(), ptx.end()); j.image.push_back('\0'); j.is_ptx = true; j.ok = true; }
            }
        };
        if (nth <= 1) work(0);
        else { std::vector<std::thread> th; for (int t = 0; t < nth; ++t) th.emplace_back(work, t); for (auto& x : th) x.join(); }
        // modules are loaded on this thread (the context of the caller's device is current here)
        for (Job& j : jobs) {
            JitKernel k; bool ok = j.ok && load_kernel(j.image, j.is_ptx, m, k, zfold);
            if (!ok && j.ok && j.from_disk) {        // a file that verified but does not load (another driver?): assemble it here
                std::string log;
                ok = jit_compile_cubin(j.ptx, j.image, log) && load_kernel(j.image, false, m, k, zfold);
            }
            std::lock_guard<std::mutex> lk(g_mu);
            JitEntry& e = g_cache[j.key];
            if (ok) { e.k = k; e.state = 1; g_stats.compiled++; } else { e.state = 2; g_stats.failed++; }
            g_stats.backend = have_jl ? 1 : 2;
        }
        std::lock_guard<std::mutex> lk(g_mu);
        g_stats.compile_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < progs.size(); ++i) {
        if (keys[i].m < 0) continue;
        auto it = g_cache.find(keys[i]);
        if (it != g_cache.end() && it->second.state == 1) out[i] = &it->second.k;      // std::map nodes never move
    }
}

JitKernel* jit_get(int device, int m, const V4Program& prog, int threshold, bool zfold) {
    std::vector<const V4Program*> one(1, &prog);
    std::vector<JitKernel*> out;
    jit_get_many(device, m, one, threshold, out, zfold);
    return out[0];
}

int jit_launch(JitKernel* k, cudaStream_t stream, int sms, int m, const CUtensorMap& tm, const V4Program& prog, const double* d_slots,
               const V4Launch& lp) {
    const DriverApi& D = driver();
    if (!k || !k->fn || !D.ok) return set_err(QM_ERR_STATE, "jit_launch without a kernel (internal error)");
    const unsigned threads = (m == 12 ? 2u * 256u : 4u * 128u) + (unsigned)kV4ProducerThreads;
    const uint64_t cap = (uint64_t)(lp.sm_cap > 0 && lp.sm_cap < sms ? lp.sm_cap : sms);
    const unsigned grid = (unsigned)(lp.ntiles < cap ? lp.ntiles : cap);
    void* params[4] = { (void*)&tm, (void*)&prog, (void*)&d_slots, (void*)&lp };
    const CUresult r = D.launch(k->fn, grid, 1, 1, threads, 1, 1, (unsigned)smem_for(m), (CUstream)stream, params, nullptr);
    if (r != CUDA_SUCCESS) return set_err(QM_ERR_CUDA, "cuLaunchKernel of a compiled pass failed (%d)", (int)r);
    return QM_OK;
}

}  // namespace qm
