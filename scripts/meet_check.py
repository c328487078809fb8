"""k_meet (one-kernel path) against the two-kernel path and the fp64 C oracle; timings over a batch-size sweep.
Run on a GPU box: python scripts/meet_check.py [quick]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from gluon_e2e_asr_b200 import _lib, ctc_loss_and_grad, ops
from tests.synth import CONFIGS, make_batch
from oracle import ctc_ref

dev = torch.device("cuda:0")

def run(d, meet, head=None):
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    with _lib.options(meet=meet):
        ops._ws_cache.clear()
        h = None if head is None else torch.tensor(head, device=dev, dtype=torch.float32)
        l, g = ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], head_grad=h)
        torch.cuda.synchronize()
        return l.cpu().numpy().copy(), g.cpu().numpy().copy()

def check(name, d, head=None):
    lo, go, ok = ctc_ref.ctc_ref(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"], blank=0, head_grad=head,
                                 layout="NTC", dtype=np.float64)
    l1, g1 = run(d, 1, head)
    l0, g0 = run(d, 0, head)
    el = np.abs(l1 - lo) / (1e-5 + 1e-4 * np.abs(lo)); eg = np.abs(g1 - go) / (1e-5 + 1e-4 * np.abs(go))
    el0 = np.abs(l0 - lo) / (1e-5 + 1e-4 * np.abs(lo)); eg0 = np.abs(g0 - go) / (1e-5 + 1e-4 * np.abs(go))
    print("%-34s meet: loss err %.3g grad err %.3g (max abs %.3g) | two-kernel: %.3g %.3g  %s" % (
        name, el.max(), eg.max(), np.abs(g1 - go).max(), el0.max(), eg0.max(), "OK" if el.max() <= 1 and eg.max() <= 1 else "FAIL"), flush=True)
    return el.max() <= 1 and eg.max() <= 1

def timeit(d, meet, iters=50):
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    with _lib.options(meet=meet):
        ops._ws_cache.clear()
        g = torch.empty_like(t["pred"])
        for _ in range(5):
            ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], out_grad=g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], out_grad=g)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

ok = True
ok &= check("tiny B2 T12 V5 L3", make_batch(2, 12, 5, 3, seed=1))
ok &= check("B4 T40 V11 L6", make_batch(4, 40, 11, 6, seed=2))
ok &= check("B3 T7 V46 L2", make_batch(3, 7, 46, 2, seed=3))
ok &= check("B3 T9 V46 L4 (2 blocks)", make_batch(3, 9, 46, 4, seed=4))
ok &= check("B5 T100 V33 L40 (P=2)", make_batch(5, 100, 33, 40, seed=5))
ok &= check("B6 T64 V64 L31 (P=1,V=64)", make_batch(6, 64, 64, 31, seed=6))
ok &= check("cfg1", make_batch(*CONFIGS["cfg1"], seed=0))
ok &= check("cfg1 peaky head", make_batch(*CONFIGS["cfg1"], seed=1, peaky=True), head=np.linspace(0.5, 2.0, 8))
ok &= check("cfg2", make_batch(*CONFIGS["cfg2"], seed=0))
ok &= check("cfg2 peaky", make_batch(*CONFIGS["cfg2"], seed=1, peaky=True))
ok &= check("cfg2 scale 12", make_batch(*CONFIGS["cfg2"], seed=2, scale=12.0))
ok &= check("B8 T2000 V46 L127", make_batch(8, 2000, 46, 127, seed=7))
if len(sys.argv) < 2:
    ok &= check("cfg5", make_batch(*CONFIGS["cfg5"], seed=0))
    ok &= check("B300 T120 V46 L120", make_batch(300, 120, 46, 120, seed=60))
print("ALL OK" if ok else "SOME FAILED", flush=True)
for B in (8, 32, 64, 128, 148, 256, 296, 512, 1024):
    d = make_batch(B, 500, 46, 120, seed=0)
    print("B=%4d T=500 V=46 L<=120: meet %.1f us   two-kernel %.1f us" % (B, timeit(d, 1), timeit(d, 0)), flush=True)
d = make_batch(*CONFIGS["cfg1"], seed=0)
print("cfg1: meet %.1f us   two-kernel %.1f us" % (timeit(d, 1), timeit(d, 0)))
d = make_batch(1024, 500, 46, 120, seed=0, full_lengths=True)
print("cfg5 full lengths: meet %.1f us   two-kernel %.1f us" % (timeit(d, 1), timeit(d, 0)))
