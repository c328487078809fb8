"""Walker configuration (state pairs per lane x walker warps) over the batch size: step time."""
import sys
sys.path.insert(0, ".")
from scripts.regime_sweep_lib import timeit
from tests.synth import make_batch
for B in (64, 128, 256, 512, 1024):
    d = make_batch(B, 500, 46, 120, seed=0)
    print("B=%4d: " % B + "  ".join("P=%d NW=%d %.1f" % (P, NW, timeit(d, walk_p=P, walk_nw=NW)) for P, NW in ((2, 2), (4, 1), (1, 4))), flush=True)
