"""Times the output projection fused with the loss (ctcb_proj_*) against the unfused pair it replaces
(a library TF32 GEMM writing the logits + ctcb_forward / ctcb_loss_grad reading them) at a BASELINE configs[2] shaped
step with H hidden units.  CUDA events on torch's current stream, inputs resident, L2 flushed between iterations by
rotating over buffer sets larger than L2.

    python scripts/proj_bench.py [--B 64 --T 500 --V 2000 --L 150 --H 512 --iters 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluon_e2e_asr_b200 import ctc_loss_and_grad, proj_ctc_loss  # noqa: E402
from gluon_e2e_asr_b200.ops import ctc_loss  # noqa: E402
from tests.synth import make_batch  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3        # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--T", type=int, default=500)
    ap.add_argument("--V", type=int, default=2000)
    ap.add_argument("--L", type=int, default=150)
    ap.add_argument("--H", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sets", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    sets = []
    for s in range(a.sets):
        d = make_batch(a.B, a.T, a.V, a.L, seed=s)
        g = torch.Generator(device="cpu").manual_seed(s)
        sets.append(dict(
            h=torch.randn((a.B, a.T, a.H), generator=g).to(dev),
            w=(torch.randn((a.V, a.H), generator=g) / a.H ** 0.5).to(dev),
            bias=torch.zeros((a.V,), device=dev),
            lab=torch.tensor(d["label"], device=dev), pl=torch.tensor(d["pred_lengths"], device=dev),
            ll=torch.tensor(d["label_lengths"], device=dev), frames=float(d["pred_lengths"].sum())))
    frames = sets[0]["frames"]
    torch.backends.cuda.matmul.allow_tf32 = True
    res = {"shape": vars(a), "valid_frames": frames}

    def fused_fwd(i):
        s = sets[i % a.sets]
        with torch.no_grad():
            return proj_ctc_loss(s["h"], s["w"], s["bias"], s["lab"], s["pl"], s["ll"])

    def unfused_fwd(i):
        s = sets[i % a.sets]
        with torch.no_grad():
            logits = torch.addmm(s["bias"], s["h"].view(-1, a.H), s["w"].t()).view(a.B, a.T, a.V)
            return ctc_loss(logits.transpose(0, 1), s["lab"], s["pl"], s["ll"], True, True)

    def gemm_only(i):
        s = sets[i % a.sets]
        return torch.addmm(s["bias"], s["h"].view(-1, a.H), s["w"].t())

    def fused_step(i):
        s = sets[i % a.sets]
        h, w, b = s["h"].requires_grad_(True), s["w"].requires_grad_(True), s["bias"].requires_grad_(True)
        h.grad = w.grad = b.grad = None
        proj_ctc_loss(h, w, b, s["lab"], s["pl"], s["ll"], fused_training=True).mean().backward()

    def plugin_step(i):
        s = sets[i % a.sets]
        h, w, b = s["h"].requires_grad_(True), s["w"].requires_grad_(True), s["bias"].requires_grad_(True)
        h.grad = w.grad = b.grad = None
        proj_ctc_loss(h, w, b, s["lab"], s["pl"], s["ll"]).mean().backward()

    def unfused_step(i):
        s = sets[i % a.sets]
        h, w, b = s["h"].requires_grad_(True), s["w"].requires_grad_(True), s["bias"].requires_grad_(True)
        h.grad = w.grad = b.grad = None
        logits = torch.addmm(b, h.view(-1, a.H), w.t()).view(a.B, a.T, a.V)
        from gluon_e2e_asr_b200 import CtcLoss
        CtcLoss()(logits, s["lab"], s["pl"], s["ll"]).mean().backward()

    def fused_fwd_keep(i):
        s = sets[i % a.sets]
        h = s["h"].requires_grad_(True)
        return proj_ctc_loss(h, s["w"], s["bias"], s["lab"], s["pl"], s["ll"], fused_training=True)

    def unfused_fwd_keep(i):
        s = sets[i % a.sets]
        from gluon_e2e_asr_b200 import CtcLoss
        logits = torch.addmm(s["bias"], s["h"].view(-1, a.H), s["w"].t()).view(a.B, a.T, a.V).requires_grad_(True)
        return CtcLoss()(logits, s["lab"], s["pl"], s["ll"])

    def backward_gemms(i):
        s = sets[i % a.sets]
        G2 = gbuf.view(-1, a.V)
        return (G2 @ s["w"]), (G2.t() @ s["h"].view(-1, a.H)), G2.sum(0)

    gbuf = torch.randn((a.B, a.T, a.V), device=dev)
    def fused_fwd_bf16(i):
        s = sets[i % a.sets]
        with torch.no_grad():
            return proj_ctc_loss(s["hb"], s["wb"], s["bias"], s["lab"], s["pl"], s["ll"])

    def unfused_fwd_bf16(i):      # library bf16 GEMM (bf16 logits, widened) + the logits path
        s = sets[i % a.sets]
        with torch.no_grad():
            logits = (s["hb"].view(-1, a.H) @ s["wb"].t()).float().add_(s["bias"]).view(a.B, a.T, a.V)
            return ctc_loss(logits.transpose(0, 1), s["lab"], s["pl"], s["ll"], True, True)

    for s_ in sets:
        s_["hb"], s_["wb"] = s_["h"].bfloat16(), s_["w"].bfloat16()
    lf, lu = fused_fwd(0), unfused_fwd(0)
    res["max_rel_loss_difference_fused_vs_unfused"] = float(((lf - lu).abs() / lu.abs().clamp_min(1)).max())
    for name, fn in (("gemm_only_tf32_cublas", gemm_only), ("fused_forward", fused_fwd), ("unfused_forward", unfused_fwd),
                     ("fused_forward_bf16", fused_fwd_bf16), ("unfused_forward_bf16", unfused_fwd_bf16),
                     ("fused_forward_keep_logits", fused_fwd_keep), ("unfused_forward_keep", unfused_fwd_keep),
                     ("backward_gemms_tf32_cublas", backward_gemms),
                     ("fused_step", fused_step), ("plugin_default_step", plugin_step), ("unfused_step", unfused_step)):
        res[name + "_us"] = round(timed(fn, a.iters), 1)
    flops = 2.0 * a.B * a.T * a.H * a.V
    res["gemm_tflops_cublas"] = round(flops / res["gemm_only_tf32_cublas_us"] / 1e6, 1)
    res["fused_forward_frames_per_s"] = frames / res["fused_forward_us"] * 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
