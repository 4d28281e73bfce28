// fused_s11.cu — the split-buffer fused pass for 2^11-amplitude tiles: 3 teams x 128 threads, an inbound and an
// outbound 32 KiB buffer per team (fused_tma_kernel.cuh k_fused_split).
#include "fused_tma_kernel.cuh"
namespace qm {
int launch_v4_s11(cudaStream_t stream, int sms, const CUtensorMap& tm, const V4Program& prog, const double* d_slots, const V4Launch& lp) {
    return launch_split_t<11, 3>(stream, sms, tm, prog, d_slots, lp);
}
}  // namespace qm
