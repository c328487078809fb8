"""Experiment driver (GPU): device time of ctcb_forward (emission + walkers, history kept), ctcb_backward and the
fused ctcb_loss_grad call for one workload, under the current CTCB_* environment.  usage: phase_times.py cfg3"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import _lib, ops

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dev = torch.device("cuda:0")
lib = _lib.load()
B, T, V, L = CONFIGS[name]
d = make_batch(B, T, V, L, seed=0, full_lengths=(name == "cfg5"))
t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
loss = torch.empty((B,), device=dev); grad = torch.empty_like(t["pred"])
call = ops._Call(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], False, True, False)
ws = ops._alloc_ws(call, True)
p = call.problem(loss, grad, None)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


fwd = lambda: _lib.check(lib.ctcb_forward(ctypes.byref(p), 1, ws.data_ptr(), ws.numel(), None))
bwd = lambda: _lib.check(lib.ctcb_backward(ctypes.byref(p), ws.data_ptr(), ws.numel(), None))
both = lambda: _lib.check(lib.ctcb_loss_grad(ctypes.byref(p), ws.data_ptr(), ws.numel(), None))
print(name, {k: os.environ[k] for k in os.environ if k.startswith("CTCB_")},
      "forward us %.1f  backward us %.1f  fused call us %.1f" % (timed(fwd), timed(bwd), timed(both)), flush=True)
