// fused_m11.cu — the fused TMA pass for 2^11-amplitude tiles: 4 teams x 128 threads, 6 stages of 32 KiB.
#include "fused_tma_kernel.cuh"
namespace qm {
int launch_v4_m11(cudaStream_t stream, int sms, const CUtensorMap& tm, const V4Program& prog, const double* d_slots, const V4Launch& lp) {
    return launch_v4_t<11, 4, 6>(stream, sms, tm, prog, d_slots, lp);
}
}  // namespace qm
