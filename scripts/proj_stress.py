"""Stress run of the fused projection (ctcb_proj_*): random shapes, ragged and degenerate lengths, both operand types, with and
without a gradient, against the logits path on the product of the same (tf32 / bf16 representable) operands.  A hang shows up
as the caller's timeout, a mismatch as an assertion.  usage: proj_stress.py [cases] [seed]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluon_e2e_asr_b200 import proj_ctc_loss
from gluon_e2e_asr_b200.ops import ctc_loss

def tf32(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)


def run(cases=100, seed=0):
  rng = np.random.Generator(np.random.PCG64(seed))
  dev = torch.device("cuda:0")
  worst = 0.0
  for c in range(cases):
      B = int(rng.integers(1, 40)); T = int(rng.integers(1, 600)); V = int(rng.integers(65, 1200))
      K = 4 * int(rng.integers(1, 66)); L = int(rng.integers(0, min(T, 160) + 1))
      bf = bool(rng.integers(0, 2)); grad = bool(rng.integers(0, 2)); use_bias = bool(rng.integers(0, 2))
      if V <= L + 1:
          V = L + 2 + 63
      if bf:
          K = 8 * max(1, K // 8)
      Tb = rng.integers(0, T + 1, B); Tb[rng.integers(0, B)] = T
      Lb = rng.integers(0, L + 1, B)
      lab = rng.integers(1, V, (B, max(L, 1)))[:, :L] if L else np.zeros((B, 0), np.int64)
      g = torch.Generator().manual_seed(c)
      h = torch.randn((B, T, K), generator=g); w = torch.randn((V, K), generator=g) / K ** 0.5
      h, w = (h.bfloat16().float(), w.bfloat16().float()) if bf else (tf32(h), tf32(w))
      bias = torch.randn((V,), generator=g) if use_bias else None
      hd, wd = h.to(dev), w.to(dev)
      bd = bias.to(dev) if use_bias else None
      labd = torch.tensor(lab.astype(np.float32), device=dev); pl = torch.tensor(Tb.astype(np.float32), device=dev)
      ll = torch.tensor(Lb.astype(np.float32), device=dev)
      logits = (hd.double() @ wd.double().t() + (bd.double() if use_bias else 0.0)).float()
      with torch.no_grad():
          ref = ctc_loss(logits.transpose(0, 1), labd, pl, ll, True, True)
      hin, win = (hd.bfloat16(), wd.bfloat16()) if bf else (hd, wd)
      if grad:
          hin = hin.clone().requires_grad_(True)
          got = proj_ctc_loss(hin, win, bd, labd, pl, ll, fused_training=True)
          got.sum().backward()
          assert torch.isfinite(hin.grad.float()).all(), (c, "non-finite d hidden")
          got = got.detach()
      else:
          with torch.no_grad():
              got = proj_ctc_loss(hin, win, bd, labd, pl, ll)
      torch.cuda.synchronize()
      err = ((got - ref).abs() / ref.abs().clamp_min(1.0)).max().item() if B else 0.0
      worst = max(worst, err)
      assert err < 2e-4, (c, B, T, K, V, L, bf, grad, err)
  return worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    w_ = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print("proj_stress ok: %d cases, worst relative loss difference %.2e" % (n, w_))
