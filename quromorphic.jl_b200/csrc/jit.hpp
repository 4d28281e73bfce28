// jit.hpp — pass programs whose opcode sequence repeats are compiled into straight-line kernels (see jit_codegen.hpp for the
// why and the what).  This header is the engine-side interface: a process-wide cache keyed by (device, tile size, opcode
// sequence), the compiler (nvJitLink when the shared library is there — it runs on several host threads at once — else the
// driver's own PTX compiler) and the launch through the driver API.  Nothing here is a fallback for a missing GPU: a kernel
// that cannot be compiled means the pass runs through the interpreter of the shipped binary, which is always there.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "fused_tma.cuh"

namespace qm {

struct JitKernel;     // opaque: a loaded module + its entry

struct JitStats {
    uint64_t compiled = 0;       // kernels compiled since the process started
    uint64_t failed = 0;         // ... that could not be compiled or loaded
    uint64_t cached = 0;         // kernels in the cache
    double   compile_ms = 0;     // wall time spent compiling (all threads' work, as seen by the caller)
    int      backend = 0;        // 0 none yet, 1 nvJitLink, 2 driver PTX compiler
};

// The kernel for this program's opcode sequence, or nullptr.  `threshold`: compile (synchronously, on this thread) when the
// sequence has been asked for this many times including now (1 = at first sight); below the threshold only the count is kept.
// zfold: the ZFOLD flavour of the kernel (the pass also takes the all-qubit <Z> sums, fused_tma_kernel.cuh) — a kernel, a count
// and a cache entry of its own.
JitKernel* jit_get(int device, int m, const V4Program& prog, int threshold, bool zfold = false);

// The same for many programs at once, compiled on several host threads (a cached plan of a deep circuit: hundreds of passes).
// out[i] = kernel or nullptr.
void jit_get_many(int device, int m, const std::vector<const V4Program*>& progs, int threshold, std::vector<JitKernel*>& out,
                  bool zfold = false);

// launch: same geometry and parameters as launch_v4_m12
int jit_launch(JitKernel* k, cudaStream_t stream, int sms, int m, const CUtensorMap& tm, const V4Program& prog, const double* d_slots,
               const V4Launch& lp);

JitStats jit_stats();

// host-only pieces, also used by the tests: PTX text of the kernel for this program (empty when there is no template for m
// or the program cannot be emitted), and the text compiled to a cubin with nvJitLink (false when the library is missing)
bool jit_ptx_for(int m, const V4Program& prog, std::string& ptx, bool zfold = false);
bool jit_compile_cubin(const std::string& ptx, std::vector<char>& cubin, std::string& log);

// The same through the optional disk cache (QM_JIT_CACHE_DIR or jit_set_cache_dir; empty = none): cubins are files keyed by a
// hash of the complete PTX text and the assembler's version.  *from_disk (optional) says whether the file was there.
bool jit_cubin_for(const std::string& ptx, std::vector<char>& cubin, std::string& log, bool* from_disk);
void jit_set_cache_dir(const char* path);
void jit_disk_stats(uint64_t out[2]);            // {cubins read from the directory, cubins written to it} since the process started

}  // namespace qm
