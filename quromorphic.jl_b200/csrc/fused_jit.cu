// fused_jit.cu — compiled to PTX TEXT only (Makefile: nvcc -ptx), never linked: the JIT template of the fused TMA pass.
// See fused_tma_kernel.cuh (JIT = true) and jit.cpp.
#include "fused_tma_kernel.cuh"
namespace qm {
#ifdef QM_JIT_ZFOLD      // the same kernel with the all-qubit <Z> sums folded into the pass (fused_tma_kernel.cuh, ZFOLD)
template __global__ void k_fused_tma<12, 2, 3, true, true>(const __grid_constant__ CUtensorMap, const __grid_constant__ V4Program, const double* __restrict__, const V4Launch);
#else
template __global__ void k_fused_tma<12, 2, 3, true>(const __grid_constant__ CUtensorMap, const __grid_constant__ V4Program, const double* __restrict__, const V4Launch);
#endif
}  // namespace qm
