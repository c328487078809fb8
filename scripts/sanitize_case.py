"""Small fused + unfused steps for compute-sanitizer (memcheck / racecheck).  usage: sanitize_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth import make_batch
from gluon_e2e_asr_b200 import ops, greedy_decode, edit_distance
dev = torch.device("cuda:0")
# small vocabulary (fused path), V = 200 (k_emit, register rows), V = 600 / 2000 (rows staged by bulk copies, chunk variants 1 and 8)
for (B, T, V, L) in ((4, 60, 46, 12), (3, 40, 200, 10), (2, 100, 46, 70), (3, 30, 600, 9), (2, 24, 2000, 150)):
    d = make_batch(B, T, V, L, seed=1)
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    loss, grad = ops.ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"])
    toks, lens = greedy_decode(t["pred"], t["pred_lengths"])
    dist = edit_distance(t["label"].to(torch.int32), t["label_lengths"].to(torch.int32), toks, lens)
    torch.cuda.synchronize()
    print(B, T, V, L, float(loss.sum()), int(dist.sum()))
