// Dependent-chain latency microbenchmark for the operations on the LU / IDA critical paths (B200, sm_100a).
// One warp, one CTA: cycles per dependent operation = (clock64 delta) / chain length.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 2048
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int nwarps_active) {
    __shared__ double sm[64];
    double x = a + threadIdx.x * 1e-9;
    int xi = threadIdx.x + 1;
    unsigned ux = threadIdx.x * 2654435761u;
    sm[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __dadd_rn(x, b);
        if (OP == 1) x = __dmul_rn(x, b);
        if (OP == 2) x = __fma_rn(x, b, a);
        if (OP == 3) x = b / x;
        if (OP == 4) x = __drcp_rn(x) + b;   // rcp + add so the value stays in range
        if (OP == 5) x = sqrt(x) + b;
        if (OP == 6) ux = __reduce_max_sync(0xffffffffu, ux) + threadIdx.x;
        if (OP == 7) x = __shfl_xor_sync(0xffffffffu, x, 1);
        if (OP == 8) xi = __shfl_xor_sync(0xffffffffu, xi, 1);
        if (OP == 9) { sm[threadIdx.x & 63] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 63]; __syncwarp(); }
        if (OP == 10) ux = __ballot_sync(0xffffffffu, ux & 1) + threadIdx.x;
        if (OP == 11) { x = __dmul_rn(x, b); x = __dadd_rn(x, a); }  // un-fused mul+add pair
        if (OP == 12) x = (x != 0.0) ? __dadd_rn(x, b) : x;          // DSETP + select chain
        if (OP == 13) { long long k = __double_as_longlong(x); k = (k > (long long)xi) ? k + 1 : k - 1; x = __longlong_as_double(k); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + xi + ux;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const char* names[] = {"DADD", "DMUL", "DFMA", "b/x (div)", "drcp+DADD", "sqrt+DADD", "REDUX.max.u32+IADD", "SHFL f64", "SHFL i32",
                           "STS+syncwarp+LDS+syncwarp", "VOTE.ballot+IADD", "DMUL+DADD pair", "DSETP+sel+DADD", "i64 cmp/sel"};
    for (int nthreads = 32; nthreads <= 256; nthreads *= 8) {
        printf("threads per CTA = %d (1 CTA)\n", nthreads);
        for (int op = 0; op < 14; ++op) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
#define C(o) case o: chain<o><<<1, nthreads>>>(out, cyc, 1.000001, 0.999999, 1); break;
                    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13)
                }
                cudaDeviceSynchronize();
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("  %-28s %8.1f cycles/op\n", names[op], (double)h / N);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
