"""Times the fused forward call (proj_ctc_loss under no_grad: projection kernel + metadata + walkers) at cfg3 + H = 512 for
option sets given on the command line, e.g.
    python scripts/proj_kernel_time.py proj_ctas=1 proj_ctas=2 proj_ctas=2,proj_dbg=1 keep=1,proj_dbg=6 meet_fwd=1,proj_overlap=0
(keep=1: forward with the logits kept for a backward).  CUDA events, 20 calls; one JSON line."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluon_e2e_asr_b200 import _lib, proj_ctc_loss
from tests.synth import make_batch

B, T, V, L, H = 64, 500, 2000, 150, 512
dev = torch.device("cuda:0")
d = make_batch(B, T, V, L, seed=0)
g = torch.Generator().manual_seed(0)
h = torch.randn((B, T, H), generator=g).to(dev); w = (torch.randn((V, H), generator=g) / H ** 0.5).to(dev)
bias = torch.zeros((V,), device=dev)
lab, pl, ll = (torch.tensor(d[k], device=dev) for k in ("label", "pred_lengths", "label_lengths"))
out = {}
for spec in sys.argv[1:]:
    kw = dict((k, int(v)) for k, v in (kv.split("=") for kv in spec.split(",")))
    keep = kw.pop("keep", 0)
    with _lib.options(**kw):
        def run():
            if keep:
                return proj_ctc_loss(h.requires_grad_(True), w, bias, lab, pl, ll)
            with torch.no_grad():
                return proj_ctc_loss(h, w, bias, lab, pl, ll)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        out[spec] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
print(json.dumps(out))
