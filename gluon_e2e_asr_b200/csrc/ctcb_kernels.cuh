// ctcb_kernels.cuh -- sm_100a device code of the CTC training-loss path.
//
// Path restated: softmax over V (row a4), blank-extended lattice
// (a5), alpha recursion (a6), beta recursion + per-label accumulation + gradient (a7),
// head-gradient scaling (a8) of `mx.nd.contrib.ctc_loss` as called at
// /root/reference/scripts/swbd/loss.py:134-139.  Not a port: the reference operator works
// in fp32 log space; these kernels work in LINEAR space with an extended exponent
// (fp32 mantissa + int32 exponent per lattice state), which needs no exp/log in the
// T-sequential chain and is ~100x closer to the fp64 oracle.
//
// Kernels (one batch = three launches, all on the caller's stream):
//   k_emit<VEC>          per frame: max / log2-sum-exp of the logits row and the emission
//                        table E2[b][t][0..L_b] = (mantissa, exponent) of y_t(blank),
//                        y_t(l_1..l_L) in an utterance-major, 16-byte-aligned layout that TMA
//                        can stream; plus the per-utterance metadata (lengths, int labels,
//                        repeats, feasibility, same-label chains for the gradient scatter).
//   k_walk<P,NW,HIST>    grid (B, 2): the alpha walker and the (reversed) beta walker of one
//                        utterance run concurrently on different SMs; E2 is staged through a
//                        shared-memory ring with cp.async.bulk (TMA) + mbarriers by a producer
//                        warp; one (blank,label) state pair per lane slot, one warp shuffle
//                        per step, no CTA barrier (skewed wavefront across warps).
//   k_grad<VEC,CH>       per frame: posterior state occupancy normalised per frame
//                        (gamma = alpha*beta'/Z_t), scatter to label columns, fused
//                        grad = head * (softmax - occupancy) written once, coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctcb {

constexpr int kZeroE = -(1 << 28);     // exponent of the "zero" state (value 2^-268435456)
constexpr int kDClamp = -100;          // smallest relative exponent that is still added
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453094;
constexpr float kMinLog2 = -1048576.0f;  // clamp of one frame's log2-probability
constexpr int kStages = 4;             // emission ring depth (blocks of G frames), at most
constexpr int kFramesPerCta = 16;      // k_logsoftmax_gather / k_grad: 4 warps x 4 frames

enum : int { UTT_INFEASIBLE = 1, UTT_BAD_LABEL = 2, UTT_LEN_CLAMPED = 4 };
enum : int { DT_I32 = 0, DT_I64 = 1, DT_F32 = 2, DT_F64 = 3 };

struct Problem {          // device view of ctcb_problem_t
    int T, B, V, Lmax, blank, label_pad;
    const float* logits; long long st_t, st_b;
    float* grad; long long gst_t, gst_b;
    const void* labels; int label_dtype; long long lst_b, lst_l;
    const void* data_len; int data_len_dtype;
    const void* label_len; int label_len_dtype;
    const float* head;
    float* loss; double* loss_sum; int* status;
};

struct Workspace {        // carved out of the caller's workspace by the host (ctcb.cu)
    int* Tb; int* Lb; int* flags;     // (B,)
    int* lab;                         // (B, Lp) int32 labels
    int* nxt;                         // (B, Lp) next position with the same label, or -1
    int* first;                       // (B, Lp) 1 when no earlier position has this label
    float2* fr;                       // (B, T) {row max, log2 sum exp2((x-max)*log2e)}
    int2* E;                          // (B, T, W) split emissions {mantissa bits, exponent}: col 0 blank, col j label j
    int4* hA;                         // (B, T, HP) alpha  {blank m, blank e, label m, label e}
    int4* hB;                         // (B, T, HP) beta' in the reversed walker's coordinates
    int Lp, W, HP;
};

__device__ __forceinline__ long long load_as_int(const void* p, int dtype, long long i) {
    switch (dtype) {
        case DT_I32: return static_cast<const int*>(p)[i];
        case DT_I64: return static_cast<const long long*>(p)[i];
        case DT_F32: return static_cast<long long>(static_cast<const float*>(p)[i]);
        default:     return static_cast<long long>(static_cast<const double*>(p)[i]);
    }
}

__device__ __forceinline__ float fast_ex2(float x) {
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// m * 2^max(d, kDClamp) by adding to the exponent field.  Valid because every mantissa in
// flight is a positive normal float >= 2^-17 (see the invariant in k_walk) and d <= 0.
__device__ __forceinline__ float xscale(float m, int d) {
    d = max(d, kDClamp);
    return __int_as_float(__float_as_int(m) + d * (1 << 23));
}

// Same, but exactly 0 below 2^-64 (used off the critical chain, where "zero" states must
// not leak into a frame's normaliser).
__device__ __forceinline__ float xscale0(float m, int d) {
    return d < -64 ? 0.0f : __int_as_float(__float_as_int(m) + d * (1 << 23));
}

// log2-probability -> (mantissa in [2^-1/2, 2^1/2], integer exponent); one MUFU.EX2.
__device__ __forceinline__ void split_log2(float l, float& m, int& e) {
    const float magic = 12582912.0f;            // 1.5 * 2^23: rounds to nearest integer
    l = fmaxf(l, kMinLog2);
    float r = l + magic;
    e = __float_as_int(r) - __float_as_int(magic);
    m = fast_ex2(l - (r - magic));
}

// ---------------------------------------------------------------------------------------
// k_emit<VEC>: grid (ceil(T/16), B), block 128, dynamic smem Lp ints; one warp per frame.
//
// Rows a3/a4/a5 of SURVEY 8a in one launch: every CTA derives its utterance's lengths and
// int labels itself (no dependency on a prepare kernel), computes {row max, log2 sum} per
// frame and writes the emission table E2[b][t][0..L_b] = split(log2 y_t(blank | l_j)) as
// (mantissa, exponent) pairs, ready for the walkers.  CTA x == 0 of each utterance also
// publishes the per-utterance metadata (lengths, labels, repeats/feasibility, same-label
// chains for the gradient scatter).
// ---------------------------------------------------------------------------------------
template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void vec_get(const typename VecT<VEC>::type& v, float (&o)[VEC]);
template <> __device__ __forceinline__ void vec_get<1>(const float& v, float (&o)[1]) { o[0] = v; }
template <> __device__ __forceinline__ void vec_get<2>(const float2& v, float (&o)[2]) { o[0] = v.x; o[1] = v.y; }
template <> __device__ __forceinline__ void vec_get<4>(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

template <int VEC>
__global__ void __launch_bounds__(128) k_emit(Problem p, Workspace w) {
    using V_t = typename VecT<VEC>::type;
    extern __shared__ int slab[];                 // Lp ints: this utterance's labels
    __shared__ int s_L, s_rep, s_flags;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_L = p.Lmax; s_rep = 0; s_flags = 0; }
    __syncthreads();
    // operator parameter layer: lengths (trunc + clamp) and labels (trunc + clamp)
    int Tb = p.T, lenflags = 0;
    if (p.data_len) {
        long long t64 = load_as_int(p.data_len, p.data_len_dtype, b);
        if (t64 < 0) { t64 = 0; lenflags = UTT_LEN_CLAMPED; }
        if (t64 > p.T) { t64 = p.T; lenflags = UTT_LEN_CLAMPED; }
        Tb = (int)t64;
    }
    int L;
    if (p.label_len) {
        long long l64 = load_as_int(p.label_len, p.label_len_dtype, b);
        if (l64 < 0) { l64 = 0; lenflags = UTT_LEN_CLAMPED; }
        if (l64 > p.Lmax) { l64 = p.Lmax; lenflags = UTT_LEN_CLAMPED; }
        L = (int)l64;
    } else {
        for (int j = tid; j < p.Lmax; j += 128)
            if (load_as_int(p.labels, p.label_dtype, b * p.lst_b + j * p.lst_l) == p.label_pad) atomicMin(&s_L, j);
        __syncthreads();
        L = s_L;
    }
    int bad = 0;
    for (int j = tid; j < L; j += 128) {
        long long v = load_as_int(p.labels, p.label_dtype, b * p.lst_b + j * p.lst_l);
        if (v < 0 || v >= p.V || v == p.blank) bad = 1;
        slab[j] = (int)(v < 0 ? 0 : (v >= p.V ? p.V - 1 : v));
    }
    if (bad) atomicOr(&s_flags, UTT_BAD_LABEL);
    __syncthreads();

    const int nvec = p.V / VEC;
#pragma unroll 1
    for (int i = 0; i < kFramesPerCta / 4; ++i) {
        const int t = blockIdx.x * kFramesPerCta + warp * (kFramesPerCta / 4) + i;
        if (t >= Tb) break;
        const float* row = p.logits + b * p.st_b + t * p.st_t;
        const V_t* rowv = reinterpret_cast<const V_t*>(row);
        float mx = -INFINITY;
        for (int k = lane; k < nvec; k += 32) {
            float x[VEC]; vec_get<VEC>(__ldg(rowv + k), x);
#pragma unroll
            for (int j = 0; j < VEC; ++j) mx = fmaxf(mx, x[j]);
        }
        for (int v = nvec * VEC + lane; v < p.V; v += 32) mx = fmaxf(mx, __ldg(row + v));
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int k = lane; k < nvec; k += 32) {          // second pass hits L1
            float x[VEC]; vec_get<VEC>(__ldg(rowv + k), x);
#pragma unroll
            for (int j = 0; j < VEC; ++j) sum += fast_ex2((x[j] - mx) * kLog2e);
        }
        for (int v = nvec * VEC + lane; v < p.V; v += 32) sum += fast_ex2((__ldg(row + v) - mx) * kLog2e);
        sum = warp_sum(sum);
        const float lg2s = log2f(sum);
        if (lane == 0) w.fr[(size_t)b * p.T + t] = make_float2(mx, lg2s);
        int2* e = w.E + ((size_t)b * p.T + t) * w.W;
        for (int j = lane; j <= L; j += 32) {
            const int v = j == 0 ? p.blank : slab[j - 1];
            float m; int ex;
            split_log2(fmaf(__ldg(row + v) - mx, kLog2e, -lg2s), m, ex);
            e[j] = make_int2(__float_as_int(m), ex);
        }
    }

    if (blockIdx.x != 0) return;
    // ---- per-utterance metadata (one CTA per utterance) ----
    int* lab = w.lab + (size_t)b * w.Lp;
    int* nxt = w.nxt + (size_t)b * w.Lp;
    int* fst = w.first + (size_t)b * w.Lp;
    int rep = 0;
    for (int j = tid; j < L; j += 128) {
        lab[j] = slab[j];
        if (j > 0 && slab[j - 1] == slab[j]) ++rep;
    }
    if (rep) atomicAdd(&s_rep, rep);
    // same-label chains: warp per position, ballot over 32 candidates at a time
    for (int j = warp; j < L; j += 4) {
        const int v = slab[j];
        int f = 1, n = -1;
        for (int c = 0; c <= (j >> 5) && f; ++c) {
            const int k = c * 32 + lane;
            if (__ballot_sync(0xffffffffu, k < j && slab[k] == v)) f = 0;
        }
        for (int c = j >> 5; c * 32 < L; ++c) {
            const int k = c * 32 + lane;
            const unsigned m = __ballot_sync(0xffffffffu, k > j && k < L && slab[k] == v);
            if (m) { n = c * 32 + __ffs(m) - 1; break; }
        }
        if (lane == 0) { nxt[j] = n; fst[j] = f; }
    }
    __syncthreads();
    if (tid == 0) {
        int flags = s_flags | lenflags;
        if (Tb <= 0 || L + s_rep > Tb) flags |= UTT_INFEASIBLE;
        w.Tb[b] = Tb; w.Lb[b] = L; w.flags[b] = flags;
        if (p.status) p.status[b] = flags;
        if (flags & UTT_INFEASIBLE) p.loss[b] = 0.0f;   // defined behaviour, SURVEY 7.3-6
    }
}

// ---------------------------------------------------------------------------------------
// mbarrier / TMA (1-D bulk copy) helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int4 lds128_volatile(uint32_t addr) {
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128_volatile(uint32_t addr, int4 v) {
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int lds32_volatile(uint32_t addr) {
    int v; asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ void sts32_volatile(uint32_t addr, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------
// k_walk<P,NW,HIST>: grid (B, ndir), block (NW+1)*32.  Rows a6/a7 (the two recursions).
//
// Slot g (= walker thread * P + p) owns the state pair (blank 2g, label 2g+1) of the walker's
// lattice.  dir 0 walks frames 0..T_b-1 over ext = [_, l1, _, ..., lL, _] and stores alpha_t
// (emission included); dir 1 walks frames T_b-1..0 over the REVERSED label sequence -- which
// is exactly the beta recursion -- and stores the sum BEFORE the emission is applied
// (beta'_t), so that  sum_s alpha_t(s) beta'_t(S-1-s) = P(l|x)  for every t.
//
// Number format: value = m * 2^e, m fp32, e int32.  Invariants between renormalisations
// (every block of KB <= 16 steps): emission mantissas are in [2^-1/2, 2^1/2]; a state's new
// mantissa is >= 2^-1/2 times the mantissa of the term with the largest exponent and
// <= 3 * 2^1/2 times the largest term, so with m in [1,2) after a renormalisation m stays in
// [2^-8, 2^35] -- always a normal fp32, which xscale() relies on.  "Zero" is (1.0, kZeroE):
// it never wins the max and enters sums scaled by 2^-100.  States beyond the utterance's
// lattice are not masked: probability only flows towards higher states, so whatever they
// hold never reaches a valid state, the loss or the stored history.
//
// Execution: NW walker warps + 1 producer warp; the step loop has NO CTA barrier and, in
// full groups of 8 steps, no branch.
//   * producer warp: streams E2 blocks (KB = 8 or 16 frames) into a shared-memory ring with
//     cp.async.bulk (TMA) + full/empty mbarriers;
//   * a single in-order warp per scheduler pays for every instruction and every branch, so
//     steps run in groups of 8, fully unrolled: emission and neighbour loads are issued one
//     step ahead, all control (ring hand-over, back-pressure, renormalisation) sits between
//     groups;
//   * walker warp w needs, per step, one value from warp w-1 (its last label state at the
//     previous step).  Warp w-1 stores it into a 32-deep ring of (m, e) slots and, after
//     each group, publishes its step count (release); warp w starts group j once warp w-1
//     has finished group j (acquire).  Warps therefore run as a wavefront skewed by one group,
//     each at the speed of its own dependent chain; warp w-1 never leads by more than two
//     groups (ring depth).
// ---------------------------------------------------------------------------------------
struct WalkArgs {
    Workspace w; int T; int stages; float* loss; double* loss_sum;
    long long* trace;      // debug only (scripts/ubench/walk_trace.cu); nullptr in the product
};

#ifdef CTCB_TRACE
#define CTCB_TP(id) do { if (a.trace && blockIdx.x == 0 && lane == 0) \
    a.trace[((size_t)(blockIdx.y * 32 + warp) * 4096 + (size_t)n * 8 + (id))] = clock64(); } while (0)
#else
#define CTCB_TP(id) do { } while (0)
#endif

__device__ __forceinline__ int2 lds64(uint32_t addr) {
    int2 v; asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, int2 v) {
    asm volatile("st.shared.v2.s32 [%0], {%1,%2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// G = steps per group = frames per emission block (8 or 16); the halo ring holds 2 groups.
template <int P, int NW, int G, int DIR, bool HIST>
__device__ __forceinline__ void walk_dir(const WalkArgs& a, unsigned char* smem_raw, int Tb, int Lb) {
    const Workspace& w = a.w;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = w.W, NS = a.stages;
    const int NQ = (Tb + G - 1) / G;
    const uint32_t row_bytes = (uint32_t)W * 8u;
    const uint32_t stage_bytes = (uint32_t)G * row_bytes;

    const uint32_t ring = smem_u32(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * stage_bytes);
    uint64_t* empty = full + kStages;
    const uint32_t halo = smem_u32(empty + kStages);               // [NW][2G] int2
    const uint32_t prog = halo + NW * 2 * G * 8;                   // [NW] int: steps completed

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < NW) sts32_volatile(prog + tid * 4, 0);
    __syncthreads();

    const int2* Eb = w.E + (size_t)b * a.T * W;
    if (warp == NW) {                           // ---- producer warp ----
        if (lane == 0) {
            int st = 0, ph = 0;                 // ph = (n / NS) & 1; the stage's previous use is ph ^ 1
            for (int n = 0; n < NQ; ++n) {      // block n covers walker steps [n*G, ...)
                if (n >= NS) mbar_wait(&empty[st], ph ^ 1);
                const int k0 = n * G, nf = min(G, Tb - k0);
                const int t0 = DIR ? Tb - k0 - nf : k0;
                const uint32_t bytes = (uint32_t)nf * row_bytes;
                mbar_expect_tx(&full[st], bytes);
                tma_load_1d(smem_raw + (size_t)st * stage_bytes, Eb + (size_t)t0 * W, bytes, &full[st]);
                if (++st == NS) { st = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---- walker warps ----
    const int* lab = w.lab + (size_t)b * w.Lp;
    const int g0 = tid * P;
    bool vb[P]; int skcap[P]; uint32_t coff[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int g = g0 + p;
        const bool vl = g < Lb;
        vb[p] = g <= Lb;
        const int cur = vl ? (DIR ? lab[Lb - 1 - g] : lab[g]) : -1;
        const int prv = (vl && g >= 1) ? (DIR ? lab[Lb - g] : lab[g - 1]) : -2;
        skcap[p] = (vl && g >= 1 && cur != prv) ? INT_MAX : kZeroE;   // pe2 = min(pe, skcap)
        coff[p] = (uint32_t)(vl ? (DIR ? Lb - g : g + 1) : 0) * 8u;
    }
    float bm[P], lm[P]; int be[P], le[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { bm[p] = 1.0f; lm[p] = 1.0f; be[p] = kZeroE; le[p] = kZeroE; }
    if (tid == 0) be[0] = 0;                    // virtual alpha_{-1} = delta(s = 0)

    int4* hptr = nullptr;
    long long hstep = 0;
    if (HIST) {
        int4* hist = (DIR ? w.hB : w.hA) + (size_t)b * a.T * w.HP;
        hptr = hist + (size_t)(DIR ? Tb - 1 : 0) * w.HP + g0;
        hstep = DIR ? -(long long)w.HP : (long long)w.HP;
    }

    const bool has_left = NW > 1 && warp > 0, has_right = NW > 1 && warp < NW - 1;
    const bool pub = has_right && lane == 31;
    const uint32_t my_halo = halo + warp * 2 * G * 8;
    const uint32_t nb_halo = halo + (warp > 0 ? warp - 1 : 0) * 2 * G * 8;
    const uint32_t nb_prog = prog + (warp > 0 ? warp - 1 : 0) * 4, rt_prog = prog + (warp + 1 < NW ? warp + 1 : warp) * 4;
    float hal_m = 1.0f; int hal_e = kZeroE;     // neighbour's state for the coming step

    float ymb, yml[P]; int yeb, yel[P];         // emissions of the current step
    auto lds_emis = [&](uint32_t row, float& mb, int& eb, float (&ml)[P], int (&el)[P]) {
        int2 v = lds64(row);
        mb = __int_as_float(v.x); eb = v.y;
#pragma unroll
        for (int p = 0; p < P; ++p) { v = lds64(row + coff[p]); ml[p] = __int_as_float(v.x); el[p] = v.y; }
    };

    // one recursion step.  next_row: emission row of the next step; slot: halo slot of this step
    auto step = [&](uint32_t next_row, uint32_t slot) {
        float nmb, nml[P]; int neb, nel[P];
        lds_emis(next_row, nmb, neb, nml, nel);
        int2 hv = make_int2(0, 0);
        if (has_left) hv = lds64(nb_halo + slot);
        float nm = __shfl_up_sync(0xffffffffu, lm[P - 1], 1);
        int ne = __shfl_up_sync(0xffffffffu, le[P - 1], 1);
        if (lane == 0) { nm = hal_m; ne = hal_e; }
#pragma unroll
        for (int p = P - 1; p >= 0; --p) {
            const float pm = p == 0 ? nm : lm[p - 1];
            const int pe = p == 0 ? ne : le[p - 1];
            const float obm = bm[p], olm = lm[p];
            const int obe = be[p], ole = le[p];
            const int Eb_ = max(obe, pe);
            const float sb = xscale(obm, obe - Eb_) + xscale(pm, pe - Eb_);
            const int pe2 = min(pe, skcap[p]);
            const int El = max(max(ole, obe), pe2);
            const float sl = xscale(olm, ole - El) + xscale(obm, obe - El) + xscale(pm, pe2 - El);
            bm[p] = sb * ymb; be[p] = Eb_ + yeb;
            lm[p] = sl * yml[p]; le[p] = El + yel[p];
            if (HIST && vb[p]) {
                if (DIR) hptr[p] = make_int4(__float_as_int(sb), Eb_, __float_as_int(sl), El);
                else     hptr[p] = make_int4(__float_as_int(bm[p]), be[p], __float_as_int(lm[p]), le[p]);
            }
        }
        if (HIST) hptr += hstep;
        if (pub) sts64(my_halo + slot, make_int2(__float_as_int(lm[P - 1]), le[P - 1]));
        if (has_left) { hal_m = __int_as_float(hv.x); hal_e = hv.y; }
        ymb = nmb; yeb = neb;
#pragma unroll
        for (int p = 0; p < P; ++p) { yml[p] = nml[p]; yel[p] = nel[p]; }
    };

    const uint32_t row_inc = DIR ? (0u - row_bytes) : row_bytes;
    int st = 0, ph = 0;
    uint32_t stage_base = ring;
#pragma unroll 1
    for (int n = 0; n < NQ; ++n) {
        const int k0 = n * G, ns = min(G, Tb - k0), kend = k0 + ns;
        CTCB_TP(0);
        // ---- between groups: everything that needs a branch ----
        if (has_left) {                                         // left neighbour finished this group?
            while (lds32_volatile(nb_prog) < kend) { }
            fence_cta();
        }
        if (has_right) {                                        // do not lap the halo ring (2 groups deep)
            while (lds32_volatile(rt_prog) < k0 - G) { }
        }
        CTCB_TP(1);
        mbar_wait(&full[st], ph);
        CTCB_TP(2);
        const uint32_t row0 = stage_base + (DIR ? (uint32_t)(ns - 1) * row_bytes : 0u);
        lds_emis(row0, ymb, yeb, yml, yel);
        const uint32_t slot0 = (uint32_t)(n & 1) * (G * 8u);
        CTCB_TP(3);
        if (ns == G) {
#pragma unroll
            for (int j = 0; j < G; ++j)
                step(j + 1 < G ? row0 + (uint32_t)(j + 1) * row_inc : row0, slot0 + (uint32_t)j * 8u);
        } else {
#pragma unroll 1
            for (int j = 0; j < ns; ++j)
                step(j + 1 < ns ? row0 + (uint32_t)(j + 1) * row_inc : row0, slot0 + (uint32_t)j * 8u);
        }
        CTCB_TP(4);
        if (NW > 1) {                                           // publish: halo slots, then the count
            __syncwarp();
            if (lane == 31) { fence_cta(); sts32_volatile(prog + warp * 4, kend); }
        }
        CTCB_TP(5);
        // renormalise mantissas to [1,2), hand the ring stage back to the producer
#pragma unroll
        for (int p = 0; p < P; ++p) {
            int bits = __float_as_int(bm[p]);
            be[p] = max(be[p] + (bits >> 23) - 127, kZeroE);
            bm[p] = __int_as_float((bits & 0x007fffff) | 0x3f800000);
            bits = __float_as_int(lm[p]);
            le[p] = max(le[p] + (bits >> 23) - 127, kZeroE);
            lm[p] = __int_as_float((bits & 0x007fffff) | 0x3f800000);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        CTCB_TP(6);
        stage_base += stage_bytes;
        if (++st == NS) { st = 0; ph ^= 1; stage_base = ring; }
    }

    if (DIR == 0) {
        // P(l|x) = alpha_{T-1}(2L) + alpha_{T-1}(2L-1) = the blank sum of slot L_b at a
        // virtual step T_b (hal_* already holds the neighbour's state after step T_b-1).
        float nm = __shfl_up_sync(0xffffffffu, lm[P - 1], 1);
        int ne = __shfl_up_sync(0xffffffffu, le[P - 1], 1);
        if (lane == 0) { nm = hal_m; ne = hal_e; }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (g0 + p == Lb) {
                const float pm = p == 0 ? nm : lm[p - 1];
                const int pe = p == 0 ? ne : le[p - 1];
                const int Eb_ = max(be[p], pe);
                const float sb = xscale(bm[p], be[p] - Eb_) + xscale(pm, pe - Eb_);
                const double nll = -kLn2 * ((double)Eb_ + (double)log2f(sb));
                a.loss[b] = (float)nll;
                if (a.loss_sum) atomicAdd(a.loss_sum, nll);
            }
        }
    }
}

template <int P, int NW, int G, bool HIST>
__global__ void __launch_bounds__((NW + 1) * 32) k_walk(WalkArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int flags = a.w.flags[b], Tb = a.w.Tb[b], Lb = a.w.Lb[b];
    if (flags & UTT_INFEASIBLE) return;
    if (blockIdx.y == 0) walk_dir<P, NW, G, 0, HIST>(a, smem_raw, Tb, Lb);
    else                 walk_dir<P, NW, G, 1, HIST>(a, smem_raw, Tb, Lb);
}

// ---------------------------------------------------------------------------------------
// k_grad<VEC,CH>: grid (ceil(T/16), B), block 128, one warp per frame.  Rows a7 (accumulation,
// gradient) and a8 (head-gradient scaling), written once in the caller's layout.
// CH = register-resident chunks of 32 state pairs per lane (pairs <= 32*CH); CH = 0 is the
// generic two-pass variant for longer label sequences.
// ---------------------------------------------------------------------------------------
struct GradArgs { Problem p; Workspace w; };

template <int VEC>
__device__ __forceinline__ void zero_row(float* row, int V, int lane) {
    using V_t = typename VecT<VEC>::type;
    const int nvec = V / VEC;
    V_t z; memset(&z, 0, sizeof(z));
    V_t* rv = reinterpret_cast<V_t*>(row);
    for (int k = lane; k < nvec; k += 32) rv[k] = z;
    for (int v = nvec * VEC + lane; v < V; v += 32) row[v] = 0.0f;
}

// alpha history of pair g and beta' history re-expressed in the forward pair coordinates:
// blank of pair g <-> reversed-walker blank of slot L-g; label of pair g <-> reversed-walker
// label of slot L-1-g.
__device__ __forceinline__ void load_pair(const int4* A, const int4* Bh, int g, int Lb,
                                          float& wb, int& eb, float& wl, int& el) {
    const int4 av = A[g];
    const int4 bb = Bh[Lb - g];
    wb = __int_as_float(av.x) * __int_as_float(bb.x); eb = av.y + bb.y;
    if (g < Lb) {
        const int4 bl = Bh[Lb - 1 - g];
        wl = __int_as_float(av.z) * __int_as_float(bl.z); el = av.w + bl.w;
    } else { wl = 0.0f; el = INT_MIN / 2; }
}

template <int VEC, int CH>
__global__ void __launch_bounds__(128) k_grad(GradArgs a) {
    using V_t = typename VecT<VEC>::type;
    const Problem& p = a.p; const Workspace& w = a.w;
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Tb = w.Tb[b], Lb = w.Lb[b];
    const bool infeasible = (w.flags[b] & UTT_INFEASIBLE) != 0;
    const float head = p.head ? p.head[b] : 1.0f;
    extern __shared__ float gbuf_all[];
    float* gbuf = gbuf_all + (size_t)warp * w.Lp;
    const int* lab = w.lab + (size_t)b * w.Lp;
    const int* nxt = w.nxt + (size_t)b * w.Lp;
    const int* fst = w.first + (size_t)b * w.Lp;
    const int nvec = p.V / VEC;
    constexpr int NCH = CH > 0 ? CH : 1;
#pragma unroll 1
    for (int i = 0; i < kFramesPerCta / 4; ++i) {
        const int t = blockIdx.x * kFramesPerCta + warp * (kFramesPerCta / 4) + i;
        if (t >= p.T) break;
        float* grow = p.grad + b * p.gst_b + t * p.gst_t;
        if (t >= Tb || infeasible) { zero_row<VEC>(grow, p.V, lane); continue; }
        const int4* A = w.hA + ((size_t)b * p.T + t) * w.HP;
        const int4* Bh = w.hB + ((size_t)b * p.T + t) * w.HP;
        float zb = 0.0f, zl = 0.0f;
        if (CH > 0) {
            // single pass: products and exponents stay in registers
            float wb[NCH], wl[NCH]; int eb[NCH], el[NCH];
            int emax = INT_MIN;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int g = c * 32 + lane;
                if (g <= Lb) { load_pair(A, Bh, g, Lb, wb[c], eb[c], wl[c], el[c]); emax = max(emax, max(eb[c], el[c])); }
                else { wb[c] = wl[c] = 0.0f; eb[c] = el[c] = INT_MIN / 2; }
            }
            emax = __reduce_max_sync(0xffffffffu, emax);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int g = c * 32 + lane;
                if (g <= Lb) {
                    zb += xscale0(wb[c], eb[c] - emax);
                    if (g < Lb) { const float v = xscale0(wl[c], el[c] - emax); zl += v; gbuf[g] = v; }
                }
            }
        } else {
            int emax = INT_MIN;
            for (int g = lane; g <= Lb; g += 32) {
                float wb, wl; int eb, el;
                load_pair(A, Bh, g, Lb, wb, eb, wl, el);
                emax = max(emax, max(eb, el));
            }
            emax = __reduce_max_sync(0xffffffffu, emax);
            for (int g = lane; g <= Lb; g += 32) {
                float wb, wl; int eb, el;
                load_pair(A, Bh, g, Lb, wb, eb, wl, el);
                zb += xscale0(wb, eb - emax);
                if (g < Lb) { const float v = xscale0(wl, el - emax); zl += v; gbuf[g] = v; }
            }
        }
        zb = warp_sum(zb); zl = warp_sum(zl);
        const float rZ = 1.0f / (zb + zl);
        const float gblank = zb * rZ;
        __syncwarp();
        // dense row: head * softmax, blank column corrected in place
        const float2 fr = w.fr[(size_t)b * p.T + t];
        const float* xrow = p.logits + b * p.st_b + t * p.st_t;
        const V_t* xv = reinterpret_cast<const V_t*>(xrow);
        V_t* gv = reinterpret_cast<V_t*>(grow);
        for (int k = lane; k < nvec; k += 32) {
            float x[VEC]; vec_get<VEC>(__ldg(xv + k), x);
            float y[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                y[j] = fast_ex2(fmaf(x[j] - fr.x, kLog2e, -fr.y));
                if (k * VEC + j == p.blank) y[j] -= gblank;
                y[j] *= head;
            }
            V_t o; memcpy(&o, y, sizeof(o));
            gv[k] = o;
        }
        for (int v = nvec * VEC + lane; v < p.V; v += 32) {
            float y = fast_ex2(fmaf(__ldg(xrow + v) - fr.x, kLog2e, -fr.y));
            if (v == p.blank) y -= gblank;
            grow[v] = y * head;
        }
        __syncwarp();
        // label columns: the first occurrence of each label value owns its column and sums
        // the occupancy of every later occurrence (deterministic, no atomics)
        for (int j = lane; j < Lb; j += 32) {
            if (!fst[j]) continue;
            float occ = 0.0f;
            for (int k = j; k >= 0; k = nxt[k]) occ += gbuf[k];
            const int v = lab[j];
            const float y = fast_ex2(fmaf(__ldg(xrow + v) - fr.x, kLog2e, -fr.y));
            grow[v] = head * (y - occ * rZ);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// k_greedy_decode: grid B, block 256.  train_ctc_ce.py:149-160 (next-row scope).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_greedy_decode(const float* logits, long long st_t, long long st_b,
                                                       const void* data_len, int dl_dtype, int T, int B, int V,
                                                       int blank, int* out_tokens, int* out_len) {
    extern __shared__ int path[];                 // T ints
    __shared__ int s_scan[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long n64 = data_len ? load_as_int(data_len, dl_dtype, b) : T;
    const int n = (int)(n64 < 0 ? 0 : (n64 > T ? T : n64));
    for (int t = warp; t < n; t += 8) {
        const float* row = logits + b * st_b + t * st_t;
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int v = lane; v < V; v += 32) {
            const float x = __ldg(row + v);
            if (x > best || (x == best && v < bi)) { best = x; bi = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) path[t] = bi;
    }
    __syncthreads();
    // keep[t] = path[t] != blank && (t == 0 || path[t] != path[t-1]); compact in order
    const int chunk = (n + 255) / 256;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int cnt = 0;
    for (int t = lo; t < hi; ++t) {
        const int c = path[t];
        cnt += (c != blank && (t == 0 || c != path[t - 1]));
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < warp; ++k) base += s_scan[k];
    int pos = base + incl - cnt;
    int* out = out_tokens + (size_t)b * T;
    for (int t = lo; t < hi; ++t) {
        const int c = path[t];
        if (c != blank && (t == 0 || c != path[t - 1])) out[pos++] = c;
    }
    if (tid == 255) out_len[b] = base + incl;
}

}  // namespace ctcb
