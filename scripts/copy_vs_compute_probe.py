"""Does a 2.37 MB host->device copy stream slow down beside the cfg2 step's kernels (and vice versa)?  Times each alone and
both together (two streams, no dependencies between them).  One JSON line."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluon_e2e_asr_b200 import _lib
from gluon_e2e_asr_b200.ops import _Call, _alloc_ws
from tests.synth import CONFIGS, make_batch

dev = torch.device("cuda:0")
lib = _lib.load()
B, T, V, L = CONFIGS["cfg2"]
N = 2_366_000
src = torch.empty(N, dtype=torch.uint8).pin_memory()
dst = torch.empty(N, dtype=torch.uint8, device=dev)
sets = []
for i in range(4):
    d = make_batch(B, T, V, L, seed=i)
    t = tuple(torch.tensor(d[k], device=dev) for k in ("pred", "label", "pred_lengths", "label_lengths"))
    call = _Call(t[0], t[1], t[2], t[3], False, True, False)
    loss = torch.empty((B,), device=dev); grad = torch.empty_like(t[0])
    ws = _alloc_ws(call, True)
    sets.append((call.problem(loss, grad), ws, t, loss, grad))
sc, sk = torch.cuda.Stream(), torch.cuda.Stream()
REP = 400


def copies():
    with torch.cuda.stream(sc):
        for _ in range(REP):
            dst.copy_(src, non_blocking=True)


def kernels():
    for i in range(REP):
        p, ws = sets[i % 4][0], sets[i % 4][1]
        _lib.check(lib.ctcb_loss_grad(ctypes.byref(p), ws.data_ptr(), ws.numel(), sk.cuda_stream))


def timed(do_c, do_k):
    torch.cuda.synchronize()
    ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ec0.record(sc); ek0.record(sk)
    # interleave the submissions so that both queues stay full
    if do_c: copies()
    if do_k: kernels()
    ec1.record(sc); ek1.record(sk)
    torch.cuda.synchronize()
    return (round(ec0.elapsed_time(ec1) / REP * 1e3, 2) if do_c else None, round(ek0.elapsed_time(ek1) / REP * 1e3, 2) if do_k else None)


kernels(); torch.cuda.synchronize()
out = {"copy_alone_us": timed(True, False)[0], "kernels_alone_us": timed(False, True)[1]}
c, k = timed(True, True)
out["copy_beside_kernels_us"], out["kernels_beside_copy_us"] = c, k
print(json.dumps(out))
