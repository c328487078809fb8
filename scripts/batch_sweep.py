"""Experiment driver (GPU): step time of the fused call against the batch size at the cfg5 shape
(T=500, V=46, L<=120, full lengths): where the latency-bound regime ends and what a sub-batched
schedule of a large batch would cost.  usage: batch_sweep.py [B,B,...]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth import make_batch
from gluon_e2e_asr_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
Bs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32, 64, 74, 96, 128, 148, 192, 256, 296, 512, 1024]
T, V, L = 500, 46, 120
for B in Bs:
    d = make_batch(B, T, V, L, seed=0, full_lengths=True)
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    loss = torch.empty((B,), device=dev); grad = torch.empty_like(t["pred"])
    call = ops._Call(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], False, True, False)
    ws = ops._alloc_ws(call, True)
    p = call.problem(loss, grad, None)
    for _ in range(3):
        _lib.check(lib.ctcb_loss_grad(ctypes.byref(p), ws.data_ptr(), ws.numel(), None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        _lib.check(lib.ctcb_loss_grad(ctypes.byref(p), ws.data_ptr(), ws.numel(), None))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print("B %4d  step us %7.1f  us per utterance %.3f  M frames/s %.0f  history MB %.0f" %
          (B, us, us / B, B * T / us, ws.numel() / 1e6), flush=True)
    del ws, grad, t
    torch.cuda.empty_cache()
