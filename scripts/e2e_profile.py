"""Experiment (GPU): where the end-to-end step's time goes (host side) for cfg2.
usage: e2e_profile.py [nsets]"""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import CtcLoss
from gluon_e2e_asr_b200.batch import PinnedBatch
dev = torch.device("cuda:0")
B, T, V, L = CONFIGS["cfg2"]
nsets = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pbs = []
for i in range(nsets):
    d = make_batch(B, T, V, L, seed=i)
    pb = PinnedBatch.from_arrays(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"])
    pb.load(dev)["pred"].requires_grad_(True)
    pbs.append(pb)
blk = CtcLoss()
loss_host = torch.empty((B,), dtype=torch.float32).pin_memory()
def step(i=0, upto=9):
    x = pbs[i % nsets].load(dev)
    if upto < 1: torch.cuda.current_stream().synchronize(); return
    pred = x["pred"]; pred.grad = None
    loss = blk(pred, x["label"], x["pred_lengths"], x["label_lengths"])
    if upto < 2: torch.cuda.current_stream().synchronize(); return
    m = loss.mean()
    if upto < 3: torch.cuda.current_stream().synchronize(); return
    m.backward()
    if upto < 4: torch.cuda.current_stream().synchronize(); return
    loss_host.copy_(loss.detach(), non_blocking=True)
    torch.cuda.current_stream().synchronize()
def timeit(upto, n=300):
    for i in range(20): step(i, upto)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): step(i, upto)
    return (time.perf_counter() - t0) / n * 1e6
for upto, name in ((0, "H2D + sync"), (1, "+ CtcLoss forward (loss+grad kernels)"), (2, "+ mean"), (3, "+ backward"), (9, "+ loss D2H = full step")):
    print("%-40s %.1f us" % (name, timeit(upto)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200): step(i)
e1.record(); torch.cuda.synchronize()
print("full step, CUDA events over 200 steps: %.1f us" % (e0.elapsed_time(e1) * 1e3 / 200))
