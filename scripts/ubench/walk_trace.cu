// Experiment (not product): clock64 trace of k_walk's group phases for one utterance.
#define CTCB_TRACE 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include "../../gluon_e2e_asr_b200/csrc/ctcb_kernels.cuh"
using namespace ctcb;
#ifndef TP
#define TP 1
#endif
#ifndef TNW
#define TNW 4
#endif
#ifndef TG
#define TG 16
#endif
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 32, T = argc > 2 ? atoi(argv[2]) : 500, L = argc > 3 ? atoi(argv[3]) : 120;
    const int W = (L + 2) / 2 * 2, HP = L + 1, Lp = (L + 3) / 4 * 4;
    Workspace w{};
    cudaMalloc(&w.Tb, B * 4); cudaMalloc(&w.Lb, B * 4); cudaMalloc(&w.flags, B * 4);
    cudaMalloc(&w.lab, (size_t)B * Lp * 4);
    cudaMalloc(&w.E, (size_t)B * T * W * 8); cudaMalloc(&w.hA, (size_t)B * T * HP * 16); cudaMalloc(&w.hB, (size_t)B * T * HP * 16);
    w.Lp = Lp; w.W = W; w.HP = HP;
    std::vector<int> Tb(B, T), Lb(B, L), fl(B, 0), lab((size_t)B * Lp);
    for (auto& x : lab) x = 1 + rand() % 45;
    std::vector<int2> E((size_t)B * T * W);
    for (auto& e : E) { float m = 0.75f + (rand() % 1000) / 2000.0f; memcpy(&e.x, &m, 4); e.y = -(rand() % 8); }
    cudaMemcpy(w.Tb, Tb.data(), B * 4, cudaMemcpyHostToDevice); cudaMemcpy(w.Lb, Lb.data(), B * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(w.flags, fl.data(), B * 4, cudaMemcpyHostToDevice); cudaMemcpy(w.lab, lab.data(), lab.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(w.E, E.data(), E.size() * 8, cudaMemcpyHostToDevice);
    float* loss; cudaMalloc(&loss, B * 4);
    long long* trace; cudaMallocManaged(&trace, 64 * 4096 * 8); 
    int stages = 4;
    size_t smem = (size_t)stages * TG * W * 8 + 2 * kStages * 8 + (size_t)TNW * 2 * TG * 8 + TNW * 4 + 16;
    auto fn = k_walk<TP, TNW, TG, true>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    WalkArgs a{w, T, stages, loss, nullptr, trace};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        memset(trace, 0, 64 * 4096 * 8);
        cudaEventRecord(e0);
        fn<<<dim3(B, 2), (TNW + 1) * 32, smem>>>(a);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %s  %.1f us\n", rep, cudaGetErrorString(err), ms * 1e3);
    }
    const int NQ = (T + TG - 1) / TG;
    for (int dir = 0; dir < 1; ++dir)
        for (int wp = 0; wp < TNW; ++wp) {
            long long* t = trace + (size_t)(dir * 32 + wp) * 4096;
            double acc[7] = {0};
            int cnt = 0;
            for (int n = 2; n < NQ - 2; ++n) {
                for (int i = 0; i < 6; ++i) acc[i] += (double)(t[n * 8 + i + 1] - t[n * 8 + i]);
                acc[6] += (double)(t[(n + 1) * 8] - t[n * 8]);
                ++cnt;
            }
            printf("dir %d warp %d: waitLR %.0f  mbar %.0f  emis0 %.0f  steps %.0f (%.1f/step)  publish %.0f  renorm+arrive %.0f  | group total %.0f  start %lld end %lld\n", dir, wp,
                   acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[3] / cnt / TG, acc[4] / cnt, acc[5] / cnt, acc[6] / cnt,
                   t[0] - trace[0], t[(NQ - 1) * 8 + 6] - trace[0]);
        }
    { long long* t = trace; printf("warp0 group starts:"); for (int n = 0; n < NQ; ++n) printf(" %lld", t[n*8]-t[0]); printf("\n");
      printf("warp0 group phases n=0: "); for (int i=0;i<7;i++) printf(" %lld", t[i]-t[0]); printf("\n");
      printf("warp0 group phases n=31: "); for (int i=0;i<7;i++) printf(" %lld", t[(NQ-1)*8+i]-t[0]); printf("\n"); }
    return 0;
}
