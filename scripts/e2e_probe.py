"""Experiment (GPU): where the end-to-end step's time goes (H2D, Python, kernels)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import CtcLoss, ops

dev = torch.device("cuda:0")
B, T, V, L = CONFIGS["cfg2"]
d = make_batch(B, T, V, L, seed=0)
h = {k: torch.from_numpy(d[k]).pin_memory() for k in d}
blk = CtcLoss()
loss_host = torch.empty((B,), dtype=torch.float32).pin_memory()

def timeit(fn, n=200, sync_each=True):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync_each: torch.cuda.current_stream().synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

def h2d():
    return [h[k].to(dev, non_blocking=True) for k in ("pred", "label", "pred_lengths", "label_lengths")]
print("H2D 4 tensors + sync: %.1f us" % timeit(h2d))
print("H2D 4 tensors, no per-step sync (CPU issue cost): %.1f us" % timeit(h2d, sync_each=False))
t = {k: v.to(dev) for k, v in h.items()}
def fwd_bwd():
    pred = t["pred"].detach().requires_grad_(True)
    loss = blk(pred, t["label"], t["pred_lengths"], t["label_lengths"])
    loss.mean().backward()
    return pred.grad
print("autograd fwd+bwd on device tensors + sync: %.1f us" % timeit(fwd_bwd))
print("autograd fwd+bwd, no per-step sync (CPU-bound rate): %.1f us" % timeit(fwd_bwd, sync_each=False))
head = torch.full((B,), 1.0 / B, device=dev)
def fused():
    return ops.ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], head_grad=head)
print("fused call (dlpack) + sync: %.1f us" % timeit(fused))
print("fused call (dlpack), no per-step sync: %.1f us" % timeit(fused, sync_each=False))
def fused_p():
    return ops.ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], head_grad=head, handoff="pointer")
print("fused call (pointer), no per-step sync: %.1f us" % timeit(fused_p, sync_each=False))
def full():
    x = h2d()
    pred = x[0].requires_grad_(True)
    loss = blk(pred, x[1], x[2], x[3])
    loss.mean().backward()
    loss_host.copy_(loss.detach(), non_blocking=True)
    return pred.grad
print("full e2e step + sync: %.1f us" % timeit(full))
print("full e2e step, no per-step sync: %.1f us" % timeit(full, sync_each=False))
