// Micro-benchmark: FP64 FMA issue rate and latency on sm_100a.
//   chains = independent accumulators per thread (ILP), each updated by a dependent DFMA per round.
// Prints DFMA lanes per clock per SM for several (warps per SM, ILP) points, and the dependent-issue latency.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters, double c0, double c1, long long* cyc) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], c0, c1);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* cyc) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) k<ILP><<<148, warps * 32>>>(out, iters, 0.999999, 1e-9, cyc);
    cudaDeviceSynchronize();
    long long c[148]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
    double dfma_lanes = (double)warps * 32 * iters * 4 * ILP;
    printf("warps/SM %2d ILP %2d: %.1f DFMA lanes/clk/SM   (%.2f cycles per warp-DFMA per SMSP-warp, dependent-issue interval %.1f cycles)  %s\n",
           warps, ILP, dfma_lanes / avg, avg / (iters * 4.0 * ILP) , avg / (iters * 4.0), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {4, 8, 16, 32}) {
        run<1>(warps, out, cyc); run<2>(warps, out, cyc); run<4>(warps, out, cyc); run<8>(warps, out, cyc); run<16>(warps, out, cyc);
    }
    return 0;
}
