"""Experiment (GPU): does the gradient kernel of chunk c hide under the recursion kernel of chunk c+1 in the
throughput regime (cfg5)?  Forward (keep history) on stream A, backward on stream B, per chunk of the batch."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
B, T, V, L = CONFIGS["cfg5"]
d = make_batch(B, T, V, L, seed=0, full_lengths=True)
t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
head = torch.full((B,), 1.0 / B, device=dev)
loss = torch.empty((B,), device=dev); grad = torch.empty_like(t["pred"])
sA, sB = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)


def run(nchunk, two_streams, reps=10):
    Bc = B // nchunk
    calls, wss = [], []
    for c in range(nchunk):
        s = slice(c * Bc, (c + 1) * Bc)
        call = ops._Call(t["pred"][s], t["label"][s], t["pred_lengths"][s], t["label_lengths"][s], False, True, False)
        wss.append(ops._alloc_ws(call, True))
        calls.append((call, call.problem(loss[s], grad[s], head[s])))
    evs = [torch.cuda.Event() for _ in range(nchunk)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def once():
        for c, (call, p) in enumerate(calls):
            _lib.check(lib.ctcb_forward(ctypes.byref(p), 1, wss[c].data_ptr(), wss[c].numel(), sA.cuda_stream))
            if two_streams:
                evs[c].record(sA)
                sB.wait_event(evs[c])
                _lib.check(lib.ctcb_backward(ctypes.byref(p), wss[c].data_ptr(), wss[c].numel(), sB.cuda_stream))
            else:
                _lib.check(lib.ctcb_backward(ctypes.byref(p), wss[c].data_ptr(), wss[c].numel(), sA.cuda_stream))
        if two_streams:
            sA.wait_stream(sB)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    e0.record(sA)
    for _ in range(reps):
        once()
    e1.record(sA)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


ref_l, ref_g = ops.ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], head_grad=head)
ref_l, ref_g = ref_l.clone(), ref_g.clone()
for nchunk in (1, 2, 4, 8, 16):
    a = run(nchunk, False)
    b = run(nchunk, True)
    ok = torch.equal(loss, ref_l) and torch.allclose(grad, ref_g, rtol=1e-5, atol=1e-8)
    print("chunks %2d (B=%4d each): one stream %.1f us, forward/backward on two streams %.1f us, same results %s" % (nchunk, B // nchunk, a, b, ok), flush=True)
