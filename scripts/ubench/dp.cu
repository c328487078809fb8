// Micro-benchmark (experiment, not product): FP64 pipe latency and throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void lat(long long* out, double* sink, double seed) {
    double a = seed + threadIdx.x, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int n = 0; n < N; ++n) {
        if (OP == 0) a = fma(a, b, c);
        if (OP == 1) a = a + c;
        if (OP == 2) a = a * b;
        if (OP == 3) { a = fma(a, b, c); a = a * b; }
        if (OP == 4) { int lo = __double2loint(a), hi = __double2hiint(a);
                       lo = __shfl_up_sync(0xffffffffu, lo, 1); hi = __shfl_up_sync(0xffffffffu, hi, 1);
                       a = __hiloint2double(hi, lo); a = fma(a, b, c); a = a * b; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[OP] = t1 - t0;
    sink[threadIdx.x] = a;
}
// throughput: W warps per block, each with 8 independent DFMA chains
__global__ void thr(long long* out, double* sink, double seed) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    sink[threadIdx.x] = s;
}
int main() {
    long long* out; double* sink; cudaMallocManaged(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 8192 * 8);
    const char* names[] = {"DFMA", "DADD", "DMUL", "DFMA+DMUL", "2xSHFL+DFMA+DMUL"};
    for (int rep = 0; rep < 2; ++rep) {
        lat<0><<<1, 32>>>(out, sink, 1); lat<1><<<1, 32>>>(out, sink, 1); lat<2><<<1, 32>>>(out, sink, 1);
        lat<3><<<1, 32>>>(out, sink, 1); lat<4><<<1, 32>>>(out, sink, 1);
        cudaDeviceSynchronize();
    }
    for (int o = 0; o < 5; ++o) printf("%-20s %.1f cycles/iter (dependent chain)\n", names[o], (double)out[o] / N);
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) { thr<<<1, warps * 32>>>(out, sink, 1); cudaDeviceSynchronize(); }
        double cyc = (double)out[0];
        printf("throughput %2d warps/SM: %.2f DFMA warp-instr/clk/SM (%.1f lanes/clk)\n", warps, warps * 8.0 * N / cyc, warps * 8.0 * N * 32 / cyc);
    }
    return 0;
}
