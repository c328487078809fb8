"""Experiment driver (GPU): per-kernel device times of the fused fwd+bwd call for every
BASELINE config with the library's default choices.  usage: breakdown.py [cfg,cfg,...]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
for name in names:
    # a BASELINE config, or B:T:V:L
    B, T, V, L = CONFIGS[name] if name in CONFIGS else tuple(int(x) for x in name.split(":"))
    d = make_batch(B, T, V, L, seed=0, full_lengths=(name == "cfg5"))
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    loss = torch.empty((B,), device=dev); grad = torch.empty_like(t["pred"])
    call = ops._Call(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], False, True, False)
    ws = ops._alloc_ws(call, True)
    p = call.problem(loss, grad, None)
    kbuf = (ctypes.c_float * 8)(); nk = ctypes.c_int32(0)
    acc = np.zeros(8); reps = 10
    for i in range(reps + 3):
        _lib.check(lib.ctcb_loss_grad_timed(ctypes.byref(p), ws.data_ptr(), ws.numel(), None, kbuf, ctypes.byref(nk)))
        if i >= 3:
            acc[:nk.value] += np.array(list(kbuf)[:nk.value])
    acc /= reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        _lib.check(lib.ctcb_loss_grad(ctypes.byref(p), ws.data_ptr(), ws.numel(), None))
    e1.record(); torch.cuda.synchronize()
    print(name, "walk cfg", _lib.last_walk_config(), _lib.last_grad_kernel(), "kernel us", [round(float(x) * 1e3, 1) for x in acc[:nk.value]],
          "back-to-back step us %.1f" % (e0.elapsed_time(e1) * 1e3 / 20), "ws MB %.0f" % (ws.numel() / 1e6), flush=True)
    del ws, grad, t
    torch.cuda.empty_cache()
