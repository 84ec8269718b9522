// Instruction-cache footprint probe: a loop whose body is KB kilobytes of straight-line dependent FP64 adds (16 B per
// SASS instruction), run by 12 de-synchronised warps per SM (3 CTAs x 128 threads, the thread kernel's occupancy).
// Prints cycles per warp instruction against body size: where it leaves the ~8-cycle DADD latency floor is where
// the loop stops being served by the SM's instruction caches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>

#define R1(x) asm volatile("add.f64 %0, %0, %1;" : "+d"(x) : "d"(inc));
#define R4(x) R1(x) R1(x) R1(x) R1(x)
#define R16(x) R4(x) R4(x) R4(x) R4(x)
#define R64(x) R16(x) R16(x) R16(x) R16(x)     // 1 KB of code
#define R256(x) R64(x) R64(x) R64(x) R64(x)    // 4 KB
#define R1K(x) R256(x) R256(x) R256(x) R256(x) // 16 KB

// BR = 1: every 16 executed instructions are followed by a taken forward branch over 16 cold ones (the step loop's
// shape: rare paths interleaved with the hot one), so sequential prefetch is broken every 256 bytes
#define B32(x) R16(x) if (x < 0.0) { R16(x) }
#define B128(x) B32(x) B32(x) B32(x) B32(x)
#define B512(x) B128(x) B128(x) B128(x) B128(x)   // ~8 KB of code, 4 KB of it executed

template <int KB4, int BR>  // body size in units of 4 KB
__global__ void __launch_bounds__(128, 3) body(double* out, long long* cyc, int iters, double inc) {
    double x = threadIdx.x;
    // de-synchronise the warps of an SM: each spins a different number of cycles first
    const long long t0 = clock64();
    const int w = (threadIdx.x >> 5) + 4 * (blockIdx.x % 3);
    while (clock64() - t0 < 3000 * w) {}
    const long long t1 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < KB4; ++k) {
            if (BR) { B512(x) } else { R256(x) }
        }
    }
    const long long t2 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)cyc, (unsigned long long)(t2 - t1));
}

template <int KB4, int BR = 0>
void run(double* out, long long* cyc) {
    const int grid = 148 * 3, iters = 6144 / KB4;  // ~1.5 M instructions per warp whatever the body size
    cudaMemset(cyc, 0, 8);
    body<KB4, BR><<<grid, 128>>>(out, cyc, iters, 1.0);
    cudaMemset(cyc, 0, 8);
    body<KB4, BR><<<grid, 128>>>(out, cyc, iters, 1.0);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (grid * 4.0) / ((double)iters * KB4 * 256);
    std::printf("body %4d KB %s: %6.2f cycles per executed warp instruction (12 warps/SM)\n", KB4 * 4 * (BR ? 2 : 1),
                BR ? "with a taken branch every 256 B " : "straight-line ", per);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 3 * 128 * 8);
    cudaMalloc(&cyc, 8);
    run<2>(out, cyc);
    run<4>(out, cyc);
    run<6>(out, cyc);
    run<8>(out, cyc);
    run<10>(out, cyc);
    run<12>(out, cyc);
    run<14>(out, cyc);
    run<16>(out, cyc);
    run<20>(out, cyc);
    run<24>(out, cyc);
    run<32>(out, cyc);
    run<48>(out, cyc);
    run<2, 1>(out, cyc);
    run<3, 1>(out, cyc);
    run<4, 1>(out, cyc);
    run<5, 1>(out, cyc);
    run<6, 1>(out, cyc);
    run<7, 1>(out, cyc);
    run<8, 1>(out, cyc);
    run<10, 1>(out, cyc);
    run<12, 1>(out, cyc);
    run<16, 1>(out, cyc);
    run<24, 1>(out, cyc);
    std::printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
