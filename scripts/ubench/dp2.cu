// Micro-benchmark (experiment): the walker's per-step FP64 dependency pattern, P pairs per lane,
// W warps per CTA, C CTAs -- does the step time depend on how many warps share the SM?
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int P, bool SHFL>
__global__ void stepk(long long* out, double* sink, double seed) {
    double bm[P], lm[P], fb[P], fls[P], flb[P];
    for (int p = 0; p < P; ++p) { bm[p] = seed + threadIdx.x + p; lm[p] = seed * 0.5 + p; fb[p] = 0.25; fls[p] = 0.125; flb[p] = 0.5; }
    double pm = 0.1; const double yb = 0.3, yl = 0.4;
    __syncthreads();
    long long t0c = clock64();
#pragma unroll 8
    for (int n = 0; n < N; ++n) {
        double t0[P], sb[P], sl[P];
#pragma unroll
        for (int p = P - 1; p >= 0; --p) { const double prev = p == 0 ? pm : lm[p - 1]; t0[p] = fma(bm[p], flb[p], lm[p]); sb[p] = fma(prev, fb[p], bm[p]); }
#pragma unroll
        for (int p = P - 1; p >= 0; --p) { const double prev = p == 0 ? pm : lm[p - 1]; sl[p] = fma(prev, fls[p], t0[p]); }
#pragma unroll
        for (int p = P - 1; p >= 0; --p) lm[p] = sl[p] * yl;
        double pn = lm[P - 1];
        if (SHFL) pn = __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(pn), 1), __shfl_up_sync(0xffffffffu, __double2loint(pn), 1));
#pragma unroll
        for (int p = P - 1; p >= 0; --p) bm[p] = sb[p] * yb;
        pm = pn;
    }
    long long t1c = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1c - t0c;
    double s = 0; for (int p = 0; p < P; ++p) s += bm[p] + lm[p];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int P, bool S> void run(const char* name, long long* out, double* sink) {
    for (int ctas : {1, 148, 296}) for (int warps : {1, 2, 3, 4, 8}) {
        for (int rep = 0; rep < 2; ++rep) { stepk<P, S><<<ctas, warps * 32>>>(out, sink, 1.0); cudaDeviceSynchronize(); }
        printf("%s P=%d ctas=%3d warps/cta=%d: %.1f clk/step\n", name, P, ctas, warps, (double)out[0] / N);
    }
}
int main() {
    long long* out; double* sink; cudaMallocManaged(&out, 64); cudaMalloc(&sink, 296 * 256 * 8 * 2);
    run<1, true>("shfl", out, sink); run<2, true>("shfl", out, sink); run<4, true>("shfl", out, sink);
    run<2, false>("noshfl", out, sink);
    return 0;
}
