// Micro-benchmark: throughput of uniform-address record fetches on sm_100a, per SM:
//   (a) LDC.64 from the kernel-parameter constant bank with a register index
//   (b) LDS.64 / LDS.128 from shared memory (uniform address: broadcast)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldc_bench ldc_bench.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
struct Prog { double c[1024]; };

template <int MODE>
__global__ void k(const __grid_constant__ Prog P, double* out, int iters, long long* cyc) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = P.c[i];
    __syncthreads();
    uint64_t pa;
    asm("cvta.to.param.u64 %0, %1;" : "=l"(pa) : "l"(&P));
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3;
    uint32_t off = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double c0, c1, c2, c3, c4, c5, c6, c7;
        if (MODE == 0) {
            asm volatile("ld.param.f64 %0, [%8];\n\tld.param.f64 %1, [%8+8];\n\tld.param.f64 %2, [%8+16];\n\tld.param.f64 %3, [%8+24];\n\t"
                         "ld.param.f64 %4, [%8+32];\n\tld.param.f64 %5, [%8+40];\n\tld.param.f64 %6, [%8+48];\n\tld.param.f64 %7, [%8+56];"
                         : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3), "=d"(c4), "=d"(c5), "=d"(c6), "=d"(c7) : "l"(pa + off));
        } else if (MODE == 1) {
            asm volatile("ld.shared.f64 %0, [%8];\n\tld.shared.f64 %1, [%8+8];\n\tld.shared.f64 %2, [%8+16];\n\tld.shared.f64 %3, [%8+24];\n\t"
                         "ld.shared.f64 %4, [%8+32];\n\tld.shared.f64 %5, [%8+40];\n\tld.shared.f64 %6, [%8+48];\n\tld.shared.f64 %7, [%8+56];"
                         : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3), "=d"(c4), "=d"(c5), "=d"(c6), "=d"(c7) : "r"(sa + off));
        } else {
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%8];\n\tld.shared.v2.f64 {%2, %3}, [%8+16];\n\t"
                         "ld.shared.v2.f64 {%4, %5}, [%8+32];\n\tld.shared.v2.f64 {%6, %7}, [%8+48];"
                         : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3), "=d"(c4), "=d"(c5), "=d"(c6), "=d"(c7) : "r"(sa + off));
        }
        a0 = fma(a0, c0, c1); a1 = fma(a1, c2, c3); a2 = fma(a2, c4, c5); a3 = fma(a3, c6, c7);
        off = (off + 64) & 8191 & ~63u;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    Prog h; for (int i = 0; i < 1024; ++i) h.c[i] = 1.0 / (i + 2);
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(h, out, iters, cyc);
                if (mode == 1) k<1><<<148, warps * 32>>>(h, out, iters, cyc);
                if (mode == 2) k<2><<<148, warps * 32>>>(h, out, iters, cyc);
            }
            cudaDeviceSynchronize();
            long long c[148]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
            // per SM: warps * iters * 64 bytes fetched per warp
            printf("mode %d (%s) warps/SM %2d: %.1f cycles/iter/SM-total, %.2f cycles per warp-level 64-byte record, err=%s\n", mode,
                   mode == 0 ? "LDC.64 x8 param bank" : mode == 1 ? "LDS.64 x8 uniform" : "LDS.128 x4 uniform", warps,
                   avg / iters, avg / iters / warps, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
