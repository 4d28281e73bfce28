// fused_m12.cu — the fused TMA pass for 2^12-amplitude tiles: 2 teams x 256 threads, 3 stages of 64 KiB.
#include "fused_tma_kernel.cuh"
namespace qm {
int launch_v4_m12(cudaStream_t stream, int sms, const CUtensorMap& tm, const V4Program& prog, const double* d_slots, const V4Launch& lp) {
    return launch_v4_t<12, 2, 3>(stream, sms, tm, prog, d_slots, lp);
}
}  // namespace qm
