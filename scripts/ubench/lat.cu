// Micro-benchmark (experiment, not product): dependent-chain latencies on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP> __global__ void k(long long* out, int* sink, int seed) {
    int lane = threadIdx.x & 31;
    float f = seed * 0.001f + lane; int i = seed + lane; int j = seed * 3 + 1;
    __shared__ int sm[64];
    sm[threadIdx.x & 63] = (threadIdx.x + 1) & 31;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int n = 0; n < N; ++n) {
        if (OP == 0) f = __shfl_up_sync(0xffffffffu, f, 1);
        if (OP == 1) f = f + 1.25f;
        if (OP == 2) i = i * 8388608 + j;
        if (OP == 3) i = max(i, j) + 1;            // VIMNMX + IADD (2 ops)
        if (OP == 4) i = sm[i & 31];                // LDS pointer chase
        if (OP == 5) { f = __shfl_up_sync(0xffffffffu, f, 1); f = f + 1.25f; }
        if (OP == 6) { i = __shfl_up_sync(0xffffffffu, i, 1); i = (lane == 0) ? j : i; }
        if (OP == 7) { f = f * 1.0001f; }
        if (OP == 8) { i = max(max(i, j), seed) ; i = max(i - j, -100); }   // VIMNMX3 + VIADDMNMX
        if (OP == 9) { asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(i) : "r"((unsigned)__cvta_generic_to_shared(&sm[i & 31]))); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[OP] = t1 - t0;
    sink[threadIdx.x] = i + (int)f;
}
int main() {
    long long* out; int* sink; cudaMallocManaged(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 4096);
    const char* names[] = {"SHFL.UP", "FADD", "IMAD", "VIMNMX+IADD (2 ops)", "LDS chase", "SHFL+FADD", "SHFL+SEL(int)", "FMUL", "max3+addclamp(2 ops)", "LDS volatile chase"};
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(out, sink, 1); k<1><<<1, 32>>>(out, sink, 1); k<2><<<1, 32>>>(out, sink, 1); k<3><<<1, 32>>>(out, sink, 1);
        k<4><<<1, 32>>>(out, sink, 1); k<5><<<1, 32>>>(out, sink, 1); k<6><<<1, 32>>>(out, sink, 1); k<7><<<1, 32>>>(out, sink, 1);
        k<8><<<1, 32>>>(out, sink, 1); k<9><<<1, 32>>>(out, sink, 1);
        cudaDeviceSynchronize();
    }
    for (int o = 0; o < 10; ++o) printf("%-24s %.1f cycles/iter\n", names[o], (double)out[o] / N);
    return 0;
}
