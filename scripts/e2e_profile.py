"""Experiment (GPU): cProfile of the end-to-end step (host side) for cfg2."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import CtcLoss
from gluon_e2e_asr_b200.batch import PinnedBatch
dev = torch.device("cuda:0")
B, T, V, L = CONFIGS["cfg2"]
d = make_batch(B, T, V, L, seed=0)
pb = PinnedBatch.from_arrays(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"])
blk = CtcLoss()
loss_host = torch.empty((B,), dtype=torch.float32).pin_memory()
t = pb.load(dev); t["pred"].requires_grad_(True)
def step():
    x = pb.load(dev)
    pred = x["pred"]; pred.grad = None
    loss = blk(pred, x["label"], x["pred_lengths"], x["label_lengths"])
    loss.mean().backward()
    loss_host.copy_(loss.detach(), non_blocking=True)
    torch.cuda.current_stream().synchronize()
for _ in range(20): step()
t0 = time.perf_counter()
for _ in range(300): step()
print("arena e2e step: %.1f us" % ((time.perf_counter() - t0) / 300 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
