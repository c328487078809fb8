"""k_grad2 (P(l|x)-normalised gradient kernel) against k_grad and the fp64 C oracle; timings over a batch-size sweep.
Run on a GPU box: python scripts/grad2_check.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gluon_e2e_asr_b200 import _lib, ctc_loss_and_grad, ops
from tests.synth import CONFIGS, make_batch
from oracle import ctc_ref

dev = torch.device("cuda:0")

def run(d, g2, head=None):
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    with _lib.options(grad2=g2):
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
    eg = np.abs(g1 - go) / (1e-5 + 1e-4 * np.abs(go)); eg0 = np.abs(g0 - go) / (1e-5 + 1e-4 * np.abs(go))
    good = eg.max() <= 1 and np.array_equal(l0, l1)
    print("%-34s grad2: grad err %.3g (max abs %.3g) | k_grad: %.3g  %s" % (name, eg.max(), np.abs(g1 - go).max(), eg0.max(), "OK" if good else "FAIL"), flush=True)
    return good

def timeit(d, g2, iters=50):
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    with _lib.options(grad2=g2):
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
ok &= check("B5 T100 V33 L40", make_batch(5, 100, 33, 40, seed=5))
ok &= check("B6 T64 V64 L31", make_batch(6, 64, 64, 31, seed=6))
ok &= check("cfg1 peaky head", make_batch(*CONFIGS["cfg1"], seed=1, peaky=True), head=np.linspace(0.5, 2.0, 8))
ok &= check("cfg2", make_batch(*CONFIGS["cfg2"], seed=0))
ok &= check("cfg2 peaky", make_batch(*CONFIGS["cfg2"], seed=1, peaky=True))
ok &= check("cfg4 (CH=16)", make_batch(*CONFIGS["cfg4"], seed=0))
ok &= check("B8 T300 V46 L200 (CH=8)", make_batch(8, 300, 46, 200, seed=7))
ok &= check("cfg5", make_batch(*CONFIGS["cfg5"], seed=0))
ok &= check("B300 T120 V46 L120", make_batch(300, 120, 46, 120, seed=60))
print("ALL OK" if ok else "SOME FAILED", flush=True)
for B in (8, 32, 64, 128, 148, 256, 296, 512, 1024):
    d = make_batch(B, 500, 46, 120, seed=0)
    print("B=%4d T=500 V=46 L<=120: grad2 %.1f us   k_grad %.1f us" % (B, timeit(d, 1), timeit(d, 0)), flush=True)
for cfg in ("cfg1", "cfg4"):
    d = make_batch(*CONFIGS[cfg], seed=0)
    print("%s: grad2 %.1f us   k_grad %.1f us" % (cfg, timeit(d, 1), timeit(d, 0)))
d = make_batch(1024, 500, 46, 120, seed=0, full_lengths=True)
print("cfg5 full lengths: grad2 %.1f us   k_grad %.1f us" % (timeit(d, 1), timeit(d, 0)))
