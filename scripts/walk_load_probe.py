"""k_walk alone over the batch size, with (keep=1: alpha and beta CTAs, history stores) and without history (keep=0:
alpha CTAs only): does a walker slow down because of its neighbours on the SM or because of its history stores?"""
import ctypes, sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests.synth import make_batch
from gluon_e2e_asr_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load()
for B in (32, 64, 128, 148, 256, 296, 512, 1024):
    d = make_batch(B, 500, 46, 120, seed=0)
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    loss = torch.empty((B,), device=dev); grad = torch.empty_like(t["pred"])
    call = ops._Call(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], False, True, False)
    ws = ops._alloc_ws(call, True)
    p = call.problem(loss, grad, None)
    out = []
    for keep in (1, 0):
        for _ in range(3):
            _lib.check(lib.ctcb_forward(ctypes.byref(p), keep, ws.data_ptr(), ws.numel(), None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(lib.ctcb_forward(ctypes.byref(p), keep, ws.data_ptr(), ws.numel(), None))
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / 20)
    print("B=%4d  k_walk with history (2B CTAs) %.1f us   without (B CTAs, alpha only) %.1f us" % (B, out[0], out[1]), flush=True)
