"""k_grad2: frame blocks per CTA (option grad2_blocks) and the register-capped build (grad2_occ) over the batch size."""
import sys
sys.path.insert(0, ".")
from scripts.regime_sweep_lib import timeit
from tests.synth import make_batch
for B, T, L in ((64, 500, 120), (128, 500, 120), (256, 500, 120), (512, 500, 120), (1024, 500, 120)):
    d = make_batch(B, T, 46, L, seed=0)
    print("B=%4d T=%d L<=%d: " % (B, T, L) + "  ".join("gb=%d %.1f" % (gb, timeit(d, grad2=1, grad2_blocks=gb)) for gb in (4, 8, 16)) +
          "  occ0 %.1f occ1 %.1f" % (timeit(d, grad2=1, grad2_occ=0), timeit(d, grad2=1, grad2_occ=1)) + "  old k_grad %.1f" % timeit(d, grad2=0), flush=True)
