// Experiment (not product): clock64 trace of the fused k_walk's group phases for one utterance.
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
#ifndef TNW
#define TNW 2
#endif
#ifndef THIST
#define THIST true
#endif
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 32, T = argc > 2 ? atoi(argv[2]) : 500, L = argc > 3 ? atoi(argv[3]) : 120;
    const int V = 46, W = 46, Lp = (L + 3) / 4 * 4, NB = (T + kG - 1) / kG, PAIRS = TNW * TP * 32;
    Workspace w{};
    cudaMalloc(&w.Tb, B * 4); cudaMalloc(&w.Lb, B * 4); cudaMalloc(&w.flags, B * 4); cudaMalloc(&w.nd, B * 4);
    cudaMalloc(&w.rank, (size_t)B * Lp * 4); cudaMalloc(&w.dl, (size_t)B * (Lp + 1) * 8); cudaMalloc(&w.gprog, B * 16); cudaMalloc(&w.runv, (size_t)B * 64 * 8); cudaMalloc(&w.pinfo, (size_t)B * 8);
    cudaMalloc(&w.hA, (size_t)B * NB * kG * PAIRS * 8); cudaMalloc(&w.hB, (size_t)B * NB * kG * PAIRS * 8);
    cudaMalloc(&w.oA, (size_t)B * NB * PAIRS * 8); cudaMalloc(&w.oB, (size_t)B * NB * PAIRS * 8);
    cudaMalloc(&w.fr, (size_t)B * T * 8);
    w.Lp = Lp; w.W = W; w.NB = NB; w.dense = 1; w.P = TP; w.NW = TNW; w.fused = 1;
    std::vector<float> lab((size_t)B * L), x((size_t)B * T * V), tl(B, (float)T), ll(B, (float)L);
    for (auto& v : lab) v = (float)(1 + rand() % (V - 1));
    for (auto& v : x) v = 3.0f * rand() / RAND_MAX;
    Problem p{};
    p.T = T; p.B = B; p.V = V; p.Lmax = L; p.blank = 0; p.label_pad = 0;
    float* dx; cudaMalloc(&dx, x.size() * 4); cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    float* dl_; cudaMalloc(&dl_, lab.size() * 4); cudaMemcpy(dl_, lab.data(), lab.size() * 4, cudaMemcpyHostToDevice);
    float *dtl, *dll; cudaMalloc(&dtl, B * 4); cudaMalloc(&dll, B * 4);
    cudaMemcpy(dtl, tl.data(), B * 4, cudaMemcpyHostToDevice); cudaMemcpy(dll, ll.data(), B * 4, cudaMemcpyHostToDevice);
    p.logits = dx; p.st_t = V; p.st_b = (long long)T * V;
    p.labels = dl_; p.label_dtype = DT_F32; p.lst_b = L; p.lst_l = 1;
    p.data_len = dtl; p.data_len_dtype = DT_F32; p.label_len = dll; p.label_len_dtype = DT_F32;
    float* loss; cudaMalloc(&loss, B * 4); p.loss = loss;
    const size_t tn = 64 * 4096;
    long long* trace; cudaMalloc(&trace, tn * 8);
    std::vector<long long> ht(tn);
    int stages = argc > 4 ? atoi(argv[4]) : 8;
    size_t smem = walk_smem_bytes(W, TNW, stages, Lp);
    auto fn = k_walk<TP, TNW, THIST, true>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    WalkArgs a{p, w, T, stages, 0, loss, nullptr, trace, 0, 1};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(trace, 0, tn * 8);
        cudaEventRecord(e0);
        fn<<<dim3(B, THIST ? 2 : 1), (TNW + kFusedProducers + 1) * 32, smem>>>(a);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %s  %.1f us\n", rep, cudaGetErrorString(err), ms * 1e3);
    }
    cudaMemcpy(ht.data(), trace, tn * 8, cudaMemcpyDeviceToHost);
    std::vector<float> hl(B); cudaMemcpy(hl.data(), loss, B * 4, cudaMemcpyDeviceToHost);
    printf("loss[0] = %f\n", hl[0]);
    const int NQ = (T + kG - 1) / kG;
    for (int dir = 0; dir < (THIST ? 2 : 1); ++dir)
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
            printf("dir %d warp %d: wait+drift+left %.0f  meta/vote/renorm/pub %.0f  halo addr %.0f  steps %.0f (%.1f/step)  publish+arrive %.0f  tail %.0f  | group total %.0f  renorms %d/%d  first group at %lld, end at %lld (clk after warp0 group 0)\n", dir, wp,
                   acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[3] / cnt / kG, acc[4] / cnt, acc[5] / cnt, acc[6] / cnt, slow, NQ,
                   t[0] - ht[0], t[(NQ - 1) * 8 + 6] - ht[0]);
        }
    return 0;
}
