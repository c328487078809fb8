"""Step time over the batch size for the launch schedules (overlap on/off, walkers per SM) -- where each one wins."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gluon_e2e_asr_b200 import _lib, ctc_loss_and_grad, ops
from tests.synth import make_batch
dev = torch.device("cuda:0")

def timeit(d, iters=40, **opts):
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    with _lib.options(**opts):
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
    ops._ws_cache.clear()
    return e0.elapsed_time(e1) / iters * 1e3

T, V, L = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (500, 46, 120)
for B in (32, 48, 64, 96, 128, 148, 192, 256, 296, 400, 512, 1024):
    d = make_batch(B, T, V, L, seed=0)
    r = {"overlap auto": timeit(d), "serial": timeit(d, overlap=0), "overlap forced": timeit(d, overlap=1) if B <= 296 else float("nan"),
         "overlap, 3 walkers/SM": timeit(d, overlap=1, walk_per_sm=3) if B <= 222 else float("nan"),
         "overlap, 4 walkers/SM": timeit(d, overlap=1, walk_per_sm=4) if B <= 296 else float("nan")}
    print("B=%4d: " % B + "  ".join("%s %.1f" % kv for kv in r.items()), flush=True)
