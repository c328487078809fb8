"""Runs a few fused fwd+bwd steps of one workload (for ncu captures).  usage: one_step.py cfg2 [nsteps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth import CONFIGS, make_batch
from gluon_e2e_asr_b200 import ops
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, T, V, L = CONFIGS[name]
dev = torch.device("cuda:0")
d = make_batch(B, T, V, L, seed=0, full_lengths=(name == "cfg5"))
t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
head = torch.full((B,), 1.0 / B, device=dev)
for i in range(n):
    loss, grad = ops.ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"], head_grad=head, handoff="pointer")
torch.cuda.synchronize()
print(name, "loss mean", loss.mean().item())
