"""Generates tests/golden/*.json|*.npz.  Run here (CPU container): python tests/golden/make_golden.py

The reference itself cannot produce vectors: its CTC arithmetic is MXNet's
`contrib.ctc_loss` (scripts/swbd/loss.py:134-139) and `import mxnet` fails in this image
(no wheel, no network), and the reference ships no tests or fixtures.
So the committed vectors are

  kat.json        upstream MXNet / warp-ctc known-answer tests K1-K5:
                  inputs + the expected losses as printed in those upstream tests
                  (4-5 significant digits) -- typed in, not computed;
  torch_fp64.npz  seeded random cases evaluated by an implementation that is independent of
                  this repo: torch's CPU `F.ctc_loss(log_softmax(x))` in float64 + autograd.
                  Inputs, per-utterance losses and logit gradients are stored.

tests/test_oracle.py checks oracle/ against both; tests/test_parity_gpu.py checks the CUDA
path against both.
"""
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def kat():
    r0 = [1.2, 3.4, 1.2, -0.1, -2.34]
    r1 = [0.1, 0.2, 0.3, 0.22, 0.123]
    r2 = [-15, -14, -13, -12, -11]
    k = []
    k.append(dict(name="K1_mxnet_test_operator_ctc_loss", layout="TNC", blank_label="first",
                  data=[[r0, r0], [r1, r1], [r2, r2]], label=[[2, 3, 0], [2, 3, 0]],
                  expect=[4.04789, 4.04789], rtol=2e-5))
    acts2 = [[[-5, -4, -3, -2, -1], r0], [[-10, -9, -8, -7, -6], r1],
             [[-15, -14, -13, -12, -11], [-15, -14.2, -13.5, -12.2, -11.22]]]
    k.append(dict(name="K2_mxnet_test_operator_ctc_loss_varlen", layout="TNC", blank_label="first",
                  data=acts2, label=[[2, 3, 1], [2, 0, 0]], expect=[7.3557, 5.4091], rtol=2e-5))
    a3 = np.roll(np.array(acts2), -1, axis=2).tolist()
    k.append(dict(name="K3_mxnet_blank_last", layout="TNC", blank_label="last",
                  data=a3, label=[[1, 2, 0], [1, -1, -1]], expect=[7.3557, 5.4091], rtol=2e-5))
    k.append(dict(name="K4_gluon_test_loss_ctc", layout="NTC", blank_label="last",
                  data=np.ones((2, 20, 4)).tolist(), label=[[1, 0, -1, -1], [2, 1, 1, -1]],
                  expect=[18.82820702, 16.50581741], rtol=2e-7,
                  variants=["NTC", "TNC", "TN", "label_lengths", "pred_lengths"]))
    k.append(dict(name="K5_warpctc_small_test", layout="TNC", blank_label="first",
                  data=[[[0.1, 0.6, 0.1, 0.1, 0.1]], [[0.1, 0.1, 0.6, 0.1, 0.1]]], label=[[1, 2]],
                  expect=[2.46285844], rtol=2e-7,
                  note="single path: -log(softmax(r0)[1]*softmax(r1)[2])"))
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(k, f, indent=1)


def torch_case(name, B, T, V, L, seed, blank="first", force=None):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((T, B, V)) * 2.0).astype(np.float32)
    lo, hi = (1, V) if blank == "first" else (0, V - 1)
    lab = rng.integers(lo, hi, (B, max(L, 1))).astype(np.int64)
    Lb = rng.integers(0, L + 1, B)
    Tb = rng.integers(max(1, T // 2), T + 1, B)
    if force is not None:
        force(lab, Lb, Tb)
    for b in range(B):
        rep = int((lab[b, 1:Lb[b]] == lab[b, :max(Lb[b] - 1, 0)]).sum()) if Lb[b] > 1 else 0
        Tb[b] = min(T, max(Tb[b], Lb[b] + rep))
        assert Lb[b] + rep <= Tb[b]
    bl = 0 if blank == "first" else V - 1
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    loss = torch.nn.functional.ctc_loss(torch.log_softmax(xt, -1), torch.tensor(lab),
                                        torch.tensor(Tb), torch.tensor(Lb), blank=bl,
                                        reduction="none", zero_infinity=False)
    head = rng.uniform(0.5, 1.5, B)
    (loss * torch.tensor(head)).sum().backward()
    return {name + "/data": x, name + "/label": lab.astype(np.int32), name + "/T_b": Tb.astype(np.int32),
            name + "/L_b": Lb.astype(np.int32), name + "/head": head, name + "/loss": loss.detach().numpy(),
            name + "/grad": xt.grad.numpy().astype(np.float64), name + "/blank": np.array(bl)}


def main():
    kat()
    out = {}

    def reps(lab, Lb, Tb):
        lab[0, :] = lab[0, 0]            # all-repeated label row
        Lb[0] = lab.shape[1]
        Lb[1] = 0                        # empty label
        Tb[2] = 1; Lb[2] = 1             # T == L == 1 single path
    out.update(torch_case("small_first", 5, 24, 6, 7, 1, "first", reps))
    out.update(torch_case("small_last", 4, 17, 5, 5, 2, "last"))

    def tight(lab, Lb, Tb):
        Lb[:] = lab.shape[1]
        Tb[:] = lab.shape[1]             # clamped up by repeats below: T == L + repeats
    out.update(torch_case("tight", 3, 12, 9, 8, 3, "first", tight))
    out.update(torch_case("wide_vocab", 2, 30, 301, 11, 4, "first"))
    out.update(torch_case("long_label", 2, 150, 12, 70, 5, "first"))
    np.savez_compressed(os.path.join(HERE, "torch_fp64.npz"), **out)
    print("wrote", sorted({k.split('/')[0] for k in out}))


if __name__ == "__main__":
    main()
