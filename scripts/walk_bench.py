"""Experiment driver (GPU): per-kernel times of the CTC step for several walker
configurations (CTCB_WALK_P / CTCB_WALK_NW) and workloads.  Not part of the product."""
import ctypes, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()

def timed(name, P, NW, need_grad=True, reps=10):
    os.environ["CTCB_WALK_P"], os.environ["CTCB_WALK_NW"] = str(P), str(NW)
    B, T, V, L = CONFIGS[name]
    d = make_batch(B, T, V, L, seed=0, full_lengths=(name == "cfg5"))
    pred = torch.tensor(d["pred"], device=dev); lab = torch.tensor(d["label"], device=dev)
    pl = torch.tensor(d["pred_lengths"], device=dev); ll = torch.tensor(d["label_lengths"], device=dev)
    loss = torch.empty((B,), device=dev); grad = torch.empty_like(pred) if need_grad else None
    call = ops._Call(pred, lab, pl, ll, False, True, False)
    ws = ops._alloc_ws(call, True)
    p = call.problem(loss, grad, None)
    kbuf = (ctypes.c_float * 8)(); nk = ctypes.c_int32(0)
    acc = np.zeros(4)
    for i in range(reps + 2):
        _lib.check(lib.ctcb_loss_grad_timed(ctypes.byref(p), ws.data_ptr(), ws.numel(), None, kbuf, ctypes.byref(nk)))
        if i >= 2:
            acc[:nk.value] += np.array(list(kbuf)[:nk.value])
    acc /= reps
    return _lib.last_walk_config(), [round(float(x) * 1e3, 1) for x in acc[:nk.value]]

if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg2"]
    cfgs = [(1,1),(2,1),(1,2),(1,3),(1,4),(2,2),(4,1),(1,5),(1,6),(1,8),(2,4),(4,2),(1,10),(1,12),(1,16),(2,8),(4,4)]
    for name in names:
        L = CONFIGS[name][3]
        for P, NW in cfgs:
            if P * NW * 32 < L + 1 or P * NW * 32 > 8 * (L + 1):
                continue
            for ng in (True, False):
                cfg, us = timed(name, P, NW, ng)
                print(name, "P=%d NW=%d" % cfg, "grad" if ng else "fwd ", "us per kernel [emit, walk, grad]:", us, flush=True)
