"""Design probe (not product, not oracle): numpy emulation of the kernels' arithmetic --
fp32 mantissa + int32 exponent ("extended-exponent") linear-space alpha/beta with
deferred renormalisation -- compared against the fp64 oracle.  Used to choose the
number format before writing CUDA; numbers are quoted in DESIGN.md."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle.ctc_oracle import ctc_loss_grad

f32 = np.float32
ZERO_E = -(1 << 28)
RENORM = 16
DCLAMP = -100

def emis(x):  # x (T,V) fp32 -> log2-prob fp32 the way K1 does it
    m = x.max(1, keepdims=True)
    s = np.exp2(((x - m) * f32(1.4426950408889634)).astype(f32)).astype(f32).sum(1, keepdims=True, dtype=f32)
    lse2 = (m * f32(1.4426950408889634)).astype(f32) + np.log2(s).astype(f32)   # log2-domain lse
    l = ((x * f32(1.4426950408889634)).astype(f32) - lse2).astype(f32)
    return l, lse2

def split(l):
    r = np.rint(l).astype(f32)
    fr = (l - r).astype(f32)
    return np.exp2(fr).astype(f32), r.astype(np.int64)

def scale(m, d):
    d = np.maximum(d, DCLAMP)
    return (m * np.exp2(d.astype(f32)).astype(f32)).astype(f32)

def walk(my, ey, skip, Tb, S, store_pre):
    """my,ey: (Tb,S) emission mantissa/exponent along the walker's lattice."""
    m = np.ones(S, f32); e = np.full(S, ZERO_E, np.int64); e[0] = 0
    Hm = np.zeros((Tb, S), f32); He = np.zeros((Tb, S), np.int64)
    for k in range(Tb):
        m1 = np.concatenate([[f32(1)], m[:-1]]); e1 = np.concatenate([[ZERO_E], e[:-1]])
        m2 = np.concatenate([[f32(1)] * 2, m[:-2]]); e2 = np.concatenate([[ZERO_E] * 2, e[:-2]])
        e2 = np.where(skip, e2, ZERO_E)
        E = np.maximum(np.maximum(e, e1), e2)
        sm = (scale(m, e - E) + scale(m1, e1 - E)).astype(f32)
        sm = (sm + scale(m2, e2 - E)).astype(f32)
        if store_pre:
            Hm[k], He[k] = sm, E
        m = (sm * my[k]).astype(f32); e = E + ey[k]
        if not store_pre:
            Hm[k], He[k] = m, e
        if (k % RENORM) == RENORM - 1:
            mant, ex = np.frexp(m)
            m = (mant * 2).astype(f32); e = e + ex - 1
            e = np.maximum(e, ZERO_E)
    # final P = sum for the last blank at a virtual extra step
    eS = e[S - 1]; e1 = e[S - 2] if S > 1 else ZERO_E
    E = max(eS, e1)
    pm = scale(m[S - 1:S], np.array([eS - E]))[0] + (scale(m[S - 2:S - 1], np.array([e1 - E]))[0] if S > 1 else 0)
    return Hm, He, (pm, E)

def one(x, lab, blank=0):
    Tb, V = x.shape; L = len(lab); S = 2 * L + 1
    l, lse2 = emis(x)
    ext = np.full(S, blank); ext[1::2] = lab
    skip = np.zeros(S, bool); skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    my, ey = split(l[:, ext])
    extr = ext[::-1]; skipr = np.zeros(S, bool); skipr[2:] = (extr[2:] != blank) & (extr[2:] != extr[:-2])
    Am, Ae, (pm, pe) = walk(my, ey, skip, Tb, S, False)
    Bm, Be, _ = walk(my[::-1, ::-1], ey[::-1, ::-1], skipr, Tb, S, True)
    Bm = Bm[::-1, ::-1]; Be = Be[::-1, ::-1]
    loss = -(np.float64(pe) + np.log2(np.float64(pm))) * np.log(2.0)
    wm = (Am * Bm).astype(f32); we = Ae + Be
    Et = we.max(1, keepdims=True)
    w = scale(wm, we - Et)
    Z = w.sum(1, keepdims=True, dtype=f32)
    gam = (w / Z).astype(f32)
    occ = np.zeros((Tb, V), f32)
    np.add.at(occ, (np.arange(Tb)[:, None], ext[None, :]), gam)
    y = np.exp2(l).astype(f32)
    return f32(loss), (y - occ).astype(f32)

def gen(B, T, V, L, seed, peaky=False):
    rng = np.random.default_rng(seed)
    Lb = rng.integers((L + 1) // 2, L + 1, B); Lb[0] = L
    lab = rng.integers(1, V, (B, L))
    Tb = rng.integers(int(np.ceil(0.6 * T)), T + 1, B); Tb[0] = T
    for b in range(B):
        rep = int((lab[b, 1:Lb[b]] == lab[b, :Lb[b] - 1]).sum())
        Tb[b] = max(Tb[b], Lb[b] + rep)
    x = rng.standard_normal((T, B, V)).astype(f32)
    if peaky:
        for b in range(B):
            # random valid alignment
            S = 2 * Lb[b] + 1
            ext = np.zeros(S, int); ext[1::2] = lab[b, :Lb[b]]
            pos = np.sort(rng.choice(np.arange(Tb[b]), size=Lb[b], replace=False)) if Lb[b] * 2 <= Tb[b] else None
            path = np.zeros(Tb[b], int)
            if pos is not None:
                path[pos] = lab[b, :Lb[b]]
                # fix repeated labels adjacent: ensure a blank between is possible (pos non-adjacent not guaranteed) -> fine, still a peaky input
            x[np.arange(Tb[b]), b, path] += 8
    return x, lab, Tb, Lb

if __name__ == '__main__':
    cfgs = {'cfg1': (8, 200, 46, 50), 'cfg2': (8, 500, 46, 120), 'cfg3': (4, 500, 2000, 150), 'cfg4': (4, 2000, 46, 300)}
    for peaky in (False, True):
        for name, (B, T, V, L) in cfgs.items():
            x, lab, Tb, Lb = gen(B, T, V, L, 0, peaky)
            lo, go, ok = ctc_loss_grad(x, lab, Tb, Lb)
            worst = 0; viol = 0; n = 0; lerr = 0
            for b in range(B):
                l, g = one(x[:Tb[b], b], lab[b, :Lb[b]])
                ref = go[:Tb[b], b]
                err = np.abs(g - ref)
                worst = max(worst, err.max()); viol += (err > 1e-5 + 1e-4 * np.abs(ref)).sum(); n += err.size
                lerr = max(lerr, abs(l - lo[b]) / abs(lo[b]))
            print(name, 'peaky' if peaky else 'normal', 'grad max abs err %.3g  viol %d/%d  loss rel err %.3g  loss~%.0f' % (worst, viol, n, lerr, lo.mean()))
