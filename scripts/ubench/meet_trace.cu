// Experiment (not product): clock64 trace of k_meet's warp roles for utterance 0, and the kernel's time at batch size B.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTP=4 -o meet_trace meet_trace.cu
#define CTCB_TRACE 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../gluon_e2e_asr_b200/csrc/ctcb_meet.cuh"
using namespace ctcb;
#ifndef TP
#define TP 4
#endif
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 32, T = argc > 2 ? atoi(argv[2]) : 500, L = argc > 3 ? atoi(argv[3]) : 120;
    const int V = 46, Lp = (L + 3) / 4 * 4, NB = (T + kG - 1) / kG, PW = 32 * TP;
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
    float* dg; cudaMalloc(&dg, x.size() * 4); p.grad = dg; p.gst_t = V; p.gst_b = (long long)T * V;
    p.labels = dl_; p.label_dtype = DT_F32; p.lst_b = L; p.lst_l = 1;
    p.data_len = dtl; p.data_len_dtype = DT_F32; p.label_len = dll; p.label_len_dtype = DT_F32;
    float* loss; cudaMalloc(&loss, B * 4); p.loss = loss;
    int2* hist; cudaMalloc(&hist, (size_t)B * NB * kHR * PW * 8);
    const size_t tn = 8 * 256 * 8;
    long long* trace; cudaMalloc(&trace, tn * 8);
    std::vector<long long> ht(tn);
    const size_t smem = meet_smem_layout(TP, V, Lp).total;
    auto fn = k_meet<TP>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MeetArgs a{p, hist, NB, trace};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("B=%d T=%d L=%d P=%d smem %zu\n", B, T, L, TP, smem);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(trace, 0, tn * 8);
        cudaEventRecord(e0);
        fn<<<B, 256, smem>>>(a);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %s  %.1f us\n", rep, cudaGetErrorString(err), ms * 1e3);
        if (err != cudaSuccess) return 1;
    }
    cudaMemcpy(ht.data(), trace, tn * 8, cudaMemcpyDeviceToHost);
    std::vector<float> hl(B); cudaMemcpy(hl.data(), loss, B * 4, cudaMemcpyDeviceToHost);
    printf("loss[0] = %f\n", hl[0]);
    const int NQ = NB, NA = (NQ + 1) / 2;
    auto T_ = [&](int w, int n, int id) { return ht[((size_t)w * 256 + n) * 8 + id]; };
    const long long t00 = T_(0, 0, 0);
    const char* names[8] = {"walker a", "walker b", "helper a0", "helper b0", "helper a1", "helper b1", "producer a", "producer b"};
    for (int w = 0; w < 8; ++w) {
        const int dir = w & 1;
        const int N1 = dir ? NQ - NA : NA;
        const bool helper = w >= 2 && w <= 5;
        const int h = (w >= 4) ? 1 : 0;
        for (int ph = 0; ph < 2; ++ph) {
            double acc[6] = {0}; int cnt = 0; double per = 0; int pc = 0;
            long long first = 0, last = 0;
            const int lo = ph ? N1 : 0, hi = ph ? NQ : N1;
            const int stp = helper ? 2 : 1;
            int n0 = lo; if (helper) while ((n0 & 1) != h) ++n0;
            for (int n = n0; n < hi && n < 256; n += stp) {
                if (T_(w, n, 0) == 0) continue;
                if (!first) first = T_(w, n, 0);
                for (int i = 0; i < 5; ++i) if (T_(w, n, i + 1) && T_(w, n, i)) acc[i] += (double)(T_(w, n, i + 1) - T_(w, n, i));
                if (n + stp < hi && T_(w, n + stp, 0)) { per += (double)(T_(w, n + stp, 0) - T_(w, n, 0)); ++pc; }
                for (int i = 5; i >= 0; --i) if (T_(w, n, i)) { last = T_(w, n, i); break; }
                ++cnt;
            }
            if (!cnt) continue;
            printf("%-10s phase %d: %3d blocks  seg %6.0f %6.0f %6.0f %6.0f %6.0f | period %6.0f clk/block-of-its-own  first %8lld last %8lld\n",
                   names[w], ph + 1, cnt, acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[4] / cnt, pc ? per / pc : 0.0, first - t00, last - t00);
        }
    }
    return 0;
}
