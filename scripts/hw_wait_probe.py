"""Producers' wait: nanosleep polls against mbarrier.try_wait (option walk_hw_wait), step time over the batch size."""
import sys
sys.path.insert(0, ".")
from scripts.regime_sweep_lib import timeit
from tests.synth import make_batch
for B, T, L in ((8, 200, 50), (32, 500, 120), (64, 500, 120), (128, 500, 120), (256, 500, 120), (512, 500, 120), (1024, 500, 120), (16, 2000, 300)):
    d = make_batch(B, T, 46, L, seed=0)
    print("B=%4d T=%d: polls %.1f us   try_wait %.1f us   auto %.1f us" % (B, T, timeit(d, walk_hw_wait=0), timeit(d, walk_hw_wait=1), timeit(d)), flush=True)
