"""Synthetic-input generator shared by tests and bench.py: the shapes
and dtypes the reference's data pipeline delivers -- float32 logits (B,T,V) in the model's NTC
layout (scripts/swbd/model.py:421-424), float32 0-padded labels (reader_kaldi_io.py:33-35,
train_ctc_ce.py:235), float32 lengths (gluonE2EASR/data/batchify.py:78-82)."""
import numpy as np

CONFIGS = {
    # name: (B, T, V, Lmax)   -- BASELINE.json "configs", in order
    "cfg1": (8, 200, 46, 50),
    "cfg2": (32, 500, 46, 120),
    "cfg3": (64, 500, 2000, 150),
    "cfg4": (16, 2000, 46, 300),
    "cfg5": (1024, 500, 46, 120),
}


def make_batch(B, T, V, L, seed=0, peaky=False, full_lengths=False, blank=0, scale=1.0):
    """Returns dict(pred (B,T,V) f32, label (B,L) f32 0-padded, pred_lengths (B,) f32,
    label_lengths (B,) f32).  Labels uniform over the non-blank symbols, L_b ~ U{ceil(L/2)..L}
    with one L_b = L, T_b ~ U{ceil(0.6T)..T} with one T_b = T, clamped feasible."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo, hi = (1, V) if blank == 0 else (0, V - 1)
    lab = rng.integers(lo, hi, (B, L))
    if full_lengths:
        Lb = np.full((B,), L)
        Tb = np.full((B,), T)
    else:
        Lb = rng.integers((L + 1) // 2, L + 1, B)
        Lb[rng.integers(0, B)] = L
        Tb = rng.integers(int(np.ceil(0.6 * T)), T + 1, B)
        Tb[rng.integers(0, B)] = T
    for b in range(B):
        rep = int((lab[b, 1:Lb[b]] == lab[b, :Lb[b] - 1]).sum()) if Lb[b] > 1 else 0
        Tb[b] = min(T, max(Tb[b], Lb[b] + rep))
        if Lb[b] + rep > Tb[b]:               # cannot be made feasible by T: shorten the label
            Lb[b] = max(0, Tb[b] // 2)
    x = (rng.standard_normal((B, T, V)) * scale).astype(np.float32)
    if peaky:
        for b in range(B):
            n, l = int(Tb[b]), int(Lb[b])
            if l == 0 or 2 * l > n:
                continue
            pos = np.sort(rng.choice(np.arange(0, n, 2), size=l, replace=False))
            path = np.full((n,), blank)
            path[pos] = lab[b, :l]
            x[b, np.arange(n), path] += 8.0
    labf = lab.astype(np.float32)
    for b in range(B):
        labf[b, Lb[b]:] = 0 if blank == 0 else -1
    return dict(pred=x, label=labf, pred_lengths=Tb.astype(np.float32),
                label_lengths=Lb.astype(np.float32))
