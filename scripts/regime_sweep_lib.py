import torch
from gluon_e2e_asr_b200 import _lib, ctc_loss_and_grad, ops
dev = torch.device("cuda:0")

def timeit(d, iters=40, **opts):
    """Mean device time (us) of one fused loss+gradient step on batch d under the given libctcb options."""
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
