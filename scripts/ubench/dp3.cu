// Micro-benchmark (experiment): the walker's per-step pattern (P=2) with its memory-side features added one at a
// time: F&1 emission loads from shared memory (one step ahead), F&2 lane-0 halo select, F&4 halo publish store,
// F&8 history store (STS.128 of high words), F&16 history store to global (STG.128), F&32 128-bit emission loads
// (two frames per load).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define N 2048
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, int4 v) { asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
template <int F>
__global__ void stepk(long long* out, double* sink, int4* hist, double seed) {
    constexpr int P = 2;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 10 + 1024; i += blockDim.x) sm[i] = 0.3 + 0.001 * i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t cb = base + 0, c0 = base + (1 + lane % 40) * 80, c1 = base + (2 + (lane * 7) % 40) * 80;
    const uint32_t halo = base + 64 * 80, hst = base + 64 * 80 + 512 + lane * 16;
    double bm[P], lm[P], fb[P], fls[P], flb[P];
    for (int p = 0; p < P; ++p) { bm[p] = seed + threadIdx.x + p; lm[p] = seed * 0.5 + p; fb[p] = 0.25; fls[p] = 0.125; flb[p] = 0.5; }
    double pm = 0.1;
    double yb = 0.3, yl[P] = {0.4, 0.41};
    int4* hg = hist + (size_t)blockIdx.x * 8 * 32 + lane;
    __syncthreads();
    long long t0c = clock64();
#pragma unroll 1
    for (int n = 0; n < N / 8; ++n) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double nyb = 0.3, nyl[P] = {0.4, 0.41}, hv = 0.0;
            if (F & 1) { const int nj = (j + 1) & 7; nyb = lds_f64(cb + nj * 8); nyl[0] = lds_f64(c0 + nj * 8); nyl[1] = lds_f64(c1 + nj * 8); }
            if (F & 2) hv = lds_f64(halo + j * 8);
            double t0[P], sb[P], sl[P];
#pragma unroll
            for (int p = P - 1; p >= 0; --p) { const double prev = p == 0 ? pm : lm[p - 1]; t0[p] = fma(bm[p], flb[p], lm[p]); sb[p] = fma(prev, fb[p], bm[p]); }
#pragma unroll
            for (int p = P - 1; p >= 0; --p) { const double prev = p == 0 ? pm : lm[p - 1]; sl[p] = fma(prev, fls[p], t0[p]); }
#pragma unroll
            for (int p = P - 1; p >= 0; --p) lm[p] = sl[p] * yl[p];
            double pn = __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(lm[P - 1]), 1), __shfl_up_sync(0xffffffffu, __double2loint(lm[P - 1]), 1));
            if ((F & 4) && lane == 31) sts_f64(halo + 256 + j * 8, lm[P - 1]);
#pragma unroll
            for (int p = P - 1; p >= 0; --p) bm[p] = sb[p] * yb;
            if (F & 8) sts_v4(hst + j * 512, make_int4(__double2hiint(bm[0]), __double2hiint(lm[0]), __double2hiint(bm[1]), __double2hiint(lm[1])));
            if (F & 16) hg[j * 32] = make_int4(__double2hiint(bm[0]), __double2hiint(lm[0]), __double2hiint(bm[1]), __double2hiint(lm[1]));
            pm = pn;
            if ((F & 2) && lane == 0) pm = hv;
            if (F & 1) { yb = nyb; yl[0] = nyl[0]; yl[1] = nyl[1]; }
        }
    }
    long long t1c = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1c - t0c;
    double s = 0; for (int p = 0; p < P; ++p) s += bm[p] + lm[p];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int F> void run(const char* name, long long* out, double* sink, int4* hist) {
    for (int warps : {1, 2}) {
        for (int rep = 0; rep < 2; ++rep) { stepk<F><<<64, warps * 32, 16384>>>(out, sink, hist, 1.0); cudaDeviceSynchronize(); }
        printf("F=%2d %-44s warps/cta=%d: %.1f clk/step\n", F, name, warps, (double)out[0] / N);
    }
}
int main() {
    long long* out; double* sink; int4* hist; cudaMallocManaged(&out, 64); cudaMalloc(&sink, 64 * 64 * 8 * 2); cudaMalloc(&hist, 64 * 8 * 32 * 16);
    run<0>("bare", out, sink, hist);
    run<1>("+emission LDS", out, sink, hist);
    run<3>("+emission LDS +halo select", out, sink, hist);
    run<7>("+emission LDS +halo select +publish", out, sink, hist);
    run<15>("+... +history STS.128", out, sink, hist);
    run<23>("+... +history STG.128", out, sink, hist);
    run<8>("bare +history STS.128", out, sink, hist);
    run<16>("bare +history STG.128", out, sink, hist);
    run<2>("bare +halo select", out, sink, hist);
    return 0;
}
