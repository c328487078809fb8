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
// Kernels (one batch = four launches, all on the caller's stream):
//   k_prepare            per utterance: lengths, int labels, repeats, feasibility, the
//                        same-label chains used by the gradient scatter.
//   k_logsoftmax_gather  per frame: max / log2-sum-exp of the logits row, and the emission
//                        table E[b][t][0..L_b] = log2 y_t(blank), log2 y_t(l_1..l_L) in an
//                        utterance-major, 16-byte-aligned layout that TMA can stream.
//   k_walk<P,NW>         grid (B, 2): the alpha walker and the (reversed) beta walker of one
//                        utterance run concurrently on different SMs; E is staged through a
//                        shared-memory ring with cp.async.bulk (TMA) + mbarrier; one
//                        (blank,label) state pair per lane slot, one warp shuffle per step.
//   k_grad<VEC>          per frame: posterior state occupancy normalised per frame
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
constexpr int kStages = 4;             // emission ring depth (blocks of KB frames)
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
    float* E;                         // (B, T, W) log2 emissions, col 0 blank, col j label j
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
// k_prepare: grid B, block 128.  Operator parameter layer.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prepare(Problem p, Workspace w) {
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_L, s_rep, s_flags;
    int* lab = w.lab + (size_t)b * w.Lp;
    int* nxt = w.nxt + (size_t)b * w.Lp;
    int* fst = w.first + (size_t)b * w.Lp;
    if (tid == 0) { s_L = p.Lmax; s_rep = 0; s_flags = 0; }
    __syncthreads();
    // label length: trunc(label_lengths[b]) or the first padding value in the row
    if (p.label_len) {
        if (tid == 0) {
            long long L = load_as_int(p.label_len, p.label_len_dtype, b);
            if (L < 0) { L = 0; s_flags |= UTT_LEN_CLAMPED; }
            if (L > p.Lmax) { L = p.Lmax; s_flags |= UTT_LEN_CLAMPED; }
            s_L = (int)L;
        }
    } else {
        for (int j = tid; j < p.Lmax; j += blockDim.x)
            if (load_as_int(p.labels, p.label_dtype, b * p.lst_b + j * p.lst_l) == p.label_pad)
                atomicMin(&s_L, j);
    }
    __syncthreads();
    const int L = s_L;
    int bad = 0;
    for (int j = tid; j < L; j += blockDim.x) {
        long long v = load_as_int(p.labels, p.label_dtype, b * p.lst_b + j * p.lst_l);
        if (v < 0 || v >= p.V || v == p.blank) bad = 1;
        v = v < 0 ? 0 : (v >= p.V ? p.V - 1 : v);
        lab[j] = (int)v;
    }
    if (bad) atomicOr(&s_flags, UTT_BAD_LABEL);
    __syncthreads();
    int rep = 0;
    for (int j = tid; j < L; j += blockDim.x) {
        const int v = lab[j];
        if (j > 0 && lab[j - 1] == v) ++rep;
        int n = -1;
        for (int k = j + 1; k < L; ++k) if (lab[k] == v) { n = k; break; }
        int f = 1;
        for (int k = j - 1; k >= 0; --k) if (lab[k] == v) { f = 0; break; }
        nxt[j] = n; fst[j] = f;
    }
    if (rep) atomicAdd(&s_rep, rep);
    __syncthreads();
    if (tid == 0) {
        long long Tb = p.T;
        int flags = s_flags;
        if (p.data_len) {
            Tb = load_as_int(p.data_len, p.data_len_dtype, b);
            if (Tb < 0) { Tb = 0; flags |= UTT_LEN_CLAMPED; }
            if (Tb > p.T) { Tb = p.T; flags |= UTT_LEN_CLAMPED; }
        }
        if (Tb <= 0 || L + s_rep > Tb) flags |= UTT_INFEASIBLE;
        w.Tb[b] = (int)Tb; w.Lb[b] = L; w.flags[b] = flags;
        if (p.status) p.status[b] = flags;
        if (flags & UTT_INFEASIBLE) p.loss[b] = 0.0f;   // defined behaviour, SURVEY 7.3-6
    }
}

// ---------------------------------------------------------------------------------------
// k_logsoftmax_gather<VEC>: grid (ceil(T/16), B), block 128; one warp per frame.
// Row a4, done once: {max, log2 sum} per frame + the gathered emission row.
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
__global__ void __launch_bounds__(128) k_logsoftmax_gather(Problem p, Workspace w) {
    using V_t = typename VecT<VEC>::type;
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Tb = w.Tb[b], Lb = w.Lb[b];
    if (w.flags[b] & UTT_INFEASIBLE) return;
    const int* lab = w.lab + (size_t)b * w.Lp;
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
        float* e = w.E + ((size_t)b * p.T + t) * w.W;
        for (int j = lane; j <= Lb; j += 32) {
            const int v = j == 0 ? p.blank : lab[j - 1];
            e[j] = fmaf(__ldg(row + v) - mx, kLog2e, -lg2s);
        }
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

// ---------------------------------------------------------------------------------------
// k_walk<P,NW>: grid (B, ndir), block NW*32.  Rows a6/a7 (the recursions).
//
// Lane slot g (= thread*P + p) owns the state pair (blank 2g, label 2g+1) of the walker's
// lattice.  dir 0 walks frames 0..T_b-1 over ext = [_, l1, _, ..., lL, _] and stores
// alpha_t (emission included); dir 1 walks frames T_b-1..0 over the REVERSED label sequence
// -- which is exactly the beta recursion -- and stores the sum BEFORE the emission is
// applied (beta'_t), so that  sum_s alpha_t(s) beta'_t(S-1-s) = P(l|x)  for every t.
//
// Number format: value = m * 2^e, m fp32, e int32.  Invariants between renormalisations
// (every block of KB <= 16 steps): emission mantissas are in [2^-1/2, 2^1/2]; a state's new
// mantissa is >= 2^-1/2 times the mantissa of the term with the largest exponent and
// <= 3 * 2^1/2 times the largest term, so after 16 steps m stays in [2^-8, 2^35] given
// m in [1,2) after a renormalisation -- always a normal fp32, which xscale() relies on.
// "Zero" is (1.0, kZeroE): it never wins the max, and enters sums scaled by 2^-100.
// ---------------------------------------------------------------------------------------
struct WalkArgs {
    Workspace w; int T; int KB; float* loss; double* loss_sum; int store_hist;
};

template <int P, int NW>
__global__ void __launch_bounds__(NW * 32) k_walk(WalkArgs a) {
    const Workspace& w = a.w;
    const int b = blockIdx.x, dir = blockIdx.y;
    if (w.flags[b] & UTT_INFEASIBLE) return;
    const int Tb = w.Tb[b], Lb = w.Lb[b];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = w.W, KB = a.KB;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * KB * W * sizeof(float));
    int2* halo = reinterpret_cast<int2*>(bars + kStages);          // [2][NW]

    // per-slot constants
    const int* lab = w.lab + (size_t)b * w.Lp;
    bool vb[P], vl[P], sk[P]; int col[P];
    const int g0 = tid * P;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int g = g0 + p;
        vb[p] = g <= Lb; vl[p] = g < Lb;
        const int cur = vl[p] ? (dir ? lab[Lb - 1 - g] : lab[g]) : -1;
        const int prv = (vl[p] && g >= 1) ? (dir ? lab[Lb - g] : lab[g - 1]) : -2;
        sk[p] = vl[p] && g >= 1 && cur != prv;
        col[p] = vl[p] ? (dir ? Lb - g : g + 1) : 0;
    }
    float bm[P], lm[P]; int be[P], le[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { bm[p] = 1.0f; lm[p] = 1.0f; be[p] = kZeroE; le[p] = kZeroE; }
    if (tid == 0) be[0] = 0;                    // virtual alpha_{-1} = delta(s = 0)

    const int NQ = (Tb + KB - 1) / KB;
    const float* Eb = w.E + (size_t)b * a.T * W;
    int4* hist = (dir ? w.hB : w.hA) + (size_t)b * a.T * w.HP;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * NW) halo[tid] = make_int2(__float_as_int(1.0f), kZeroE);
    __syncthreads();

    auto issue = [&](int n) {                   // block n covers walker steps [n*KB, ...)
        const int k0 = n * KB, nf = min(KB, Tb - k0);
        const int t0 = dir ? Tb - k0 - nf : k0;
        const uint32_t bytes = (uint32_t)nf * W * sizeof(float);
        uint64_t* bar = &bars[n % kStages];
        mbar_expect_tx(bar, bytes);
        tma_load_1d(ring + (size_t)(n % kStages) * KB * W, Eb + (size_t)t0 * W, bytes, bar);
    };
    if (tid == 0) for (int n = 0; n < kStages && n < NQ; ++n) issue(n);

    // emission of the current step, already split
    float ymb, yml[P]; int yeb, yel[P];
    float rawb, rawl[P];
    auto load_raw = [&](const float* row) {
        rawb = row[0];
#pragma unroll
        for (int p = 0; p < P; ++p) rawl[p] = row[col[p]];
    };
    auto convert = [&]() {
        split_log2(rawb, ymb, yeb);
#pragma unroll
        for (int p = 0; p < P; ++p) split_log2(rawl[p], yml[p], yel[p]);
    };

    mbar_wait(&bars[0], 0);
    {
        const int nf0 = min(KB, Tb);
        load_raw(ring + (size_t)(dir ? nf0 - 1 : 0) * W);
        convert();
    }

#pragma unroll 1
    for (int n = 0; n < NQ; ++n) {
        const int k0 = n * KB, nf = min(KB, Tb - k0);
        const float* blk = ring + (size_t)(n % kStages) * KB * W;
#pragma unroll 1
        for (int f = 0; f < nf; ++f) {
            const int k = k0 + f;
            const int t = dir ? Tb - 1 - k : k;
            // prefetch the next step's raw emissions (consumed after the chain below)
            if (f + 1 < nf) {
                load_raw(blk + (size_t)(dir ? nf - 2 - f : f + 1) * W);
            } else if (n + 1 < NQ) {
                mbar_wait(&bars[(n + 1) % kStages], ((n + 1) / kStages) & 1);
                const int nf1 = min(KB, Tb - k0 - KB);
                load_raw(ring + (size_t)((n + 1) % kStages) * KB * W + (size_t)(dir ? nf1 - 1 : 0) * W);
            }
            // left neighbour's label state at step k-1
            float nm = __shfl_up_sync(0xffffffffu, lm[P - 1], 1);
            int ne = __shfl_up_sync(0xffffffffu, le[P - 1], 1);
            if (lane == 0) {
                if (NW > 1 && warp > 0) {
                    const int2 h = halo[((k + 1) & 1) * NW + warp - 1];
                    nm = __int_as_float(h.x); ne = h.y;
                } else { nm = 1.0f; ne = kZeroE; }
            }
            int4* hrow = hist + (size_t)t * w.HP;
#pragma unroll
            for (int p = P - 1; p >= 0; --p) {
                const float pm = p == 0 ? nm : lm[p - 1];
                const int pe = p == 0 ? ne : le[p - 1];
                const float obm = bm[p], olm = lm[p];
                const int obe = be[p], ole = le[p];
                const int Eb_ = max(obe, pe);
                const float sb = xscale(obm, obe - Eb_) + xscale(pm, pe - Eb_);
                const int pe2 = sk[p] ? pe : kZeroE;
                const int El = max(max(ole, obe), pe2);
                const float sl = xscale(olm, ole - El) + xscale(obm, obe - El) + xscale(pm, pe2 - El);
                float nbm = sb * ymb, nlm = sl * yml[p];
                int nbe = Eb_ + yeb, nle = El + yel[p];
                if (!vb[p]) { nbm = 1.0f; nbe = kZeroE; }
                if (!vl[p]) { nlm = 1.0f; nle = kZeroE; }
                if (a.store_hist && vb[p]) {
                    int4 h;
                    if (dir) h = make_int4(__float_as_int(sb), Eb_, __float_as_int(sl), vl[p] ? El : kZeroE);
                    else     h = make_int4(__float_as_int(nbm), nbe, __float_as_int(nlm), nle);
                    hrow[g0 + p] = h;
                }
                bm[p] = nbm; be[p] = nbe; lm[p] = nlm; le[p] = nle;
            }
            if (NW > 1) {
                if (lane == 31) halo[(k & 1) * NW + warp] = make_int2(__float_as_int(lm[P - 1]), le[P - 1]);
                __syncthreads();
            }
            convert();
        }
        // renormalise mantissas to [1,2)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            int bits = __float_as_int(bm[p]);
            be[p] = max(be[p] + (bits >> 23) - 127, kZeroE);
            bm[p] = __int_as_float((bits & 0x007fffff) | 0x3f800000);
            bits = __float_as_int(lm[p]);
            le[p] = max(le[p] + (bits >> 23) - 127, kZeroE);
            lm[p] = __int_as_float((bits & 0x007fffff) | 0x3f800000);
        }
        if (NW == 1) __syncwarp();
        if (tid == 0 && n + kStages < NQ) issue(n + kStages);
    }

    if (dir == 0) {
        // P(l|x) = alpha_{T-1}(2L) + alpha_{T-1}(2L-1) = the blank sum of slot L_b at a
        // virtual step T_b.
        float nm = __shfl_up_sync(0xffffffffu, lm[P - 1], 1);
        int ne = __shfl_up_sync(0xffffffffu, le[P - 1], 1);
        if (lane == 0) {
            if (NW > 1 && warp > 0) {
                const int2 h = halo[((Tb - 1) & 1) * NW + warp - 1];
                nm = __int_as_float(h.x); ne = h.y;
            } else { nm = 1.0f; ne = kZeroE; }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (g0 + p == Lb) {
                const float pm = p == 0 ? nm : lm[p - 1];
                const int pe = p == 0 ? ne : le[p - 1];
                const int Eb_ = max(be[p], pe);
                const float sb = xscale(bm[p], be[p] - Eb_) + xscale(pm, pe - Eb_);
                const double l2 = (double)Eb_ + (double)log2f(sb);
                const double nll = -kLn2 * l2;
                a.loss[b] = (float)nll;
                if (a.loss_sum) atomicAdd(a.loss_sum, nll);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// k_grad<VEC>: grid (ceil(T/16), B), block 128, one warp per frame.  Rows a7 (accumulation,
// gradient) and a8 (head-gradient scaling), written once in the caller's layout.
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

template <int VEC>
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
#pragma unroll 1
    for (int i = 0; i < kFramesPerCta / 4; ++i) {
        const int t = blockIdx.x * kFramesPerCta + warp * (kFramesPerCta / 4) + i;
        if (t >= p.T) break;
        float* grow = p.grad + b * p.gst_b + t * p.gst_t;
        if (t >= Tb || infeasible) { zero_row<VEC>(grow, p.V, lane); continue; }
        const int4* A = w.hA + ((size_t)b * p.T + t) * w.HP;
        const int4* Bh = w.hB + ((size_t)b * p.T + t) * w.HP;
        // pass 1: frame-wide maximum exponent of alpha*beta'
        int emax = INT_MIN;
        for (int g = lane; g <= Lb; g += 32) {
            const int4 av = A[g];
            const int4 bb = Bh[Lb - g];
            emax = max(emax, av.y + bb.y);
            if (g < Lb) { const int4 bl = Bh[Lb - 1 - g]; emax = max(emax, av.w + bl.w); }
        }
        emax = __reduce_max_sync(0xffffffffu, emax);
        // pass 2: scaled products, per-frame normaliser Z_t
        float zb = 0.0f, zl = 0.0f;
        for (int g = lane; g <= Lb; g += 32) {
            const int4 av = A[g];
            const int4 bb = Bh[Lb - g];
            zb += xscale0(__int_as_float(av.x) * __int_as_float(bb.x), av.y + bb.y - emax);
            if (g < Lb) {
                const int4 bl = Bh[Lb - 1 - g];
                const float wl = xscale0(__int_as_float(av.z) * __int_as_float(bl.z), av.w + bl.w - emax);
                zl += wl; gbuf[g] = wl;
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
