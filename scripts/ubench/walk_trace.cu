// Experiment (not product): clock64 trace of k_walk's group phases for one utterance.
#define CTCB_TRACE 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <cmath>
#include "../../gluon_e2e_asr_b200/csrc/ctcb_kernels.cuh"
using namespace ctcb;
#ifndef TP
#define TP 2
#endif
#ifndef THIST
#define THIST true
#endif
#ifndef TNW
#define TNW 2
#endif
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 32, T = argc > 2 ? atoi(argv[2]) : 500, L = argc > 3 ? atoi(argv[3]) : 120;
    const int V = 46, W = 46, Lp = (L + 3) / 4 * 4, NB = (T + kG - 1) / kG, PAIRS = TNW * TP * 32;
    Workspace w{};
    cudaMalloc(&w.Tb, B * 4); cudaMalloc(&w.Lb, B * 4); cudaMalloc(&w.flags, B * 4);
    cudaMalloc(&w.lab, (size_t)B * Lp * 4);
    cudaMalloc(&w.E, (size_t)B * NB * W * kEC * 8); cudaMalloc(&w.hA, (size_t)B * NB * kG * PAIRS * 16); cudaMalloc(&w.hB, (size_t)B * NB * kG * PAIRS * 16);
    cudaMalloc(&w.oA, (size_t)B * NB * PAIRS * 8); cudaMalloc(&w.oB, (size_t)B * NB * PAIRS * 8);
    cudaMalloc(&w.fr, (size_t)B * T * 8); cudaMemset(w.fr, 0, (size_t)B * T * 8);
    w.Lp = Lp; w.W = W; w.NB = NB; w.dense = 1; w.P = TP; w.NW = TNW;
    std::vector<int> Tb(B, T), Lb(B, L), fl(B, 0), lab((size_t)B * Lp);
    for (auto& x : lab) x = 1 + rand() % (V - 1);
    std::vector<double> E((size_t)B * NB * W * kEC);
    for (auto& e : E) e = exp(-3.0 * rand() / RAND_MAX);
    cudaMemcpy(w.Tb, Tb.data(), B * 4, cudaMemcpyHostToDevice); cudaMemcpy(w.Lb, Lb.data(), B * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(w.flags, fl.data(), B * 4, cudaMemcpyHostToDevice); cudaMemcpy(w.lab, lab.data(), lab.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(w.E, E.data(), E.size() * 8, cudaMemcpyHostToDevice);
    float* loss; cudaMalloc(&loss, B * 4);
    const size_t tn = 64 * 4096;
    long long* trace; cudaMalloc(&trace, tn * 8);
    std::vector<long long> ht(tn);
    int stages = 13;
    size_t smem = walk_smem_bytes(W, TNW, stages);
    auto fn = k_walk<TP, TNW, THIST, false>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    WalkArgs a{Problem{}, w, T, stages, 0, loss, nullptr, trace};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(trace, 0, tn * 8);
        cudaEventRecord(e0);
        fn<<<dim3(B, 2), (TNW + 1) * 32, smem>>>(a);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %s  %.1f us\n", rep, cudaGetErrorString(err), ms * 1e3);
    }
    cudaMemcpy(ht.data(), trace, tn * 8, cudaMemcpyDeviceToHost);
    std::vector<float> hl(B); cudaMemcpy(hl.data(), loss, B * 4, cudaMemcpyDeviceToHost);
    printf("loss[0] = %f\n", hl[0]);
    const int NQ = (T + kG - 1) / kG;
    for (int dir = 0; dir < 1; ++dir)
        for (int wp = 0; wp < TNW; ++wp) {
            long long* t = ht.data() + (size_t)(dir * 32 + wp) * 4096;
            double acc[7] = {0};
            int cnt = 0, slow = 0;
            for (int n = 0; n < NQ; ++n) slow += t[n * 8 + 7] != 0;
            for (int n = 2; n < NQ - 2; ++n) {
                for (int i = 0; i < 6; ++i) acc[i] += (double)(t[n * 8 + i + 1] - t[n * 8 + i]);
                acc[6] += (double)(t[(n + 1) * 8] - t[n * 8]);
                ++cnt;
            }
            printf("dir %d warp %d: waitLR %.0f  boundary %.0f  mbar+emis0 %.0f  steps %.0f (%.1f/step)  publish %.0f  arrive %.0f  | group total %.0f  slow paths %d/%d  start %lld end %lld\n", dir, wp,
                   acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[3] / cnt / kG, acc[4] / cnt, acc[5] / cnt, acc[6] / cnt, slow, NQ,
                   t[0] - ht[0], t[(NQ - 1) * 8 + 6] - ht[0]);
        }
    { long long* t = ht.data(); printf("warp0 group starts:"); for (int n = 0; n < NQ; ++n) printf(" %lld", t[n*8]-t[0]); printf("\n"); }
    return 0;
}
