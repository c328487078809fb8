"""Generates tests/golden/sampler.json and tests/golden/wer.json by running the REFERENCE's own code
(only possible where /root/reference exists; the fixtures are committed, this script documents them).

* gluonE2EASR/data/sampler.py (FixedBucketSampler / SortedSampler / SortedBucketSampler): its only
  MXNet dependency is the empty base class `mxnet.gluon.data.Sampler`, stubbed here so that the
  unmodified file can be loaded from where it lies.
* scripts/swbd/wer.py (compute_wer, _edit_distance): pure Python.

usage: python tests/golden/make_next_rows_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def stub_mxnet():
    mx = types.ModuleType("mxnet"); gl = types.ModuleType("mxnet.gluon"); da = types.ModuleType("mxnet.gluon.data")
    da.Sampler = type("Sampler", (object,), {})
    mx.gluon = gl; gl.data = da
    sys.modules.update({"mxnet": mx, "mxnet.gluon": gl, "mxnet.gluon.data": da})


def lengths_for(seed, n, two_keys):
    rng = np.random.RandomState(seed)
    T = rng.randint(40, 1500, n)
    if not two_keys:
        return [int(t) for t in T]
    L = np.maximum(1, (T * rng.uniform(0.05, 0.2, n)).astype(int))
    return [(int(t), int(l)) for t, l in zip(T, L)]


def main():
    stub_mxnet()
    S = load(os.path.join(REF, "gluonE2EASR/data/sampler.py"), "ref_sampler")
    cases = []
    grid = [
        dict(seed=0, n=200, two_keys=True, batch_size=32, num_buckets=5, ratio=0.0, shuffle=False, reverse=True),
        dict(seed=1, n=333, two_keys=True, batch_size=32, num_buckets=5, ratio=0.5, shuffle=True, reverse=True),
        dict(seed=2, n=97, two_keys=False, batch_size=8, num_buckets=10, ratio=0.0, shuffle=True, reverse=False),
        dict(seed=3, n=64, two_keys=True, batch_size=16, num_buckets=3, ratio=1.5, shuffle=False, reverse=False),
        dict(seed=4, n=10, two_keys=False, batch_size=4, num_buckets=20, ratio=0.0, shuffle=False, reverse=True),
    ]
    for g in grid:
        lengths = lengths_for(g["seed"], g["n"], g["two_keys"])
        s = S.FixedBucketSampler(lengths, g["batch_size"], num_buckets=g["num_buckets"], ratio=g["ratio"],
                                 shuffle=g["shuffle"], reverse=g["reverse"])
        np.random.seed(100 + g["seed"])
        epochs = [[list(map(int, b)) for b in s] for _ in range(2)]      # two epochs: shuffles compound
        cases.append(dict(kind="fixed", args=g, lengths=lengths, stats=s.stats(), n_batches=len(s), epochs=epochs))
    # explicit bucket keys
    lengths = lengths_for(5, 120, True)
    keys = [(400, 80), (800, 160), (1500, 300)]
    s = S.FixedBucketSampler(lengths, 16, num_buckets=None, bucket_keys=keys)
    cases.append(dict(kind="fixed_keys", lengths=lengths, bucket_keys=keys, batch_size=16, stats=s.stats(),
                      epochs=[[list(map(int, b)) for b in s]]))
    for seed, shuffle, reverse in ((6, False, True), (7, True, False)):
        keys_ = lengths_for(seed, 150, False)
        np.random.seed(200 + seed)
        s = S.SortedBucketSampler(keys_, 16, mult=3, reverse=reverse, shuffle=shuffle)
        cases.append(dict(kind="sorted_bucket", sort_keys=keys_, batch_size=16, mult=3, reverse=reverse, shuffle=shuffle,
                          seed=200 + seed, n_batches=len(s), epochs=[[list(map(int, b)) for b in s]]))
    keys_ = lengths_for(8, 40, False)
    cases.append(dict(kind="sorted", sort_keys=keys_, reverse=True, ids=list(S.SortedSampler(keys_))))
    cases.append(dict(kind="sorted", sort_keys=keys_, reverse=False, ids=list(S.SortedSampler(keys_, reverse=False))))
    with open(os.path.join(HERE, "sampler.json"), "w") as f:
        json.dump(cases, f)

    W = load(os.path.join(REF, "scripts/swbd/wer.py"), "ref_wer")
    rng = np.random.RandomState(11)
    wcases = []
    for n_pairs, vmax, lmax in ((6, 5, 12), (10, 46, 60), (4, 2000, 150), (3, 3, 1)):
        refs, hyps = [], []
        for _ in range(n_pairs):
            r = rng.randint(1, vmax + 1, rng.randint(1, lmax + 1)).tolist()
            h = list(r)
            for _ in range(rng.randint(0, max(1, len(r) // 2) + 1)):      # random edits
                op = rng.randint(3)
                pos = rng.randint(0, len(h) + 1)
                if op == 0 and h: h.pop(min(pos, len(h) - 1))
                elif op == 1: h.insert(pos, int(rng.randint(1, vmax + 1)))
                elif h: h[min(pos, len(h) - 1)] = int(rng.randint(1, vmax + 1))
            if rng.rand() < 0.15:
                h = []
            refs.append([int(x) for x in r]); hyps.append([int(x) for x in h])
        dist = [int(W._edit_distance(r, h)) for r, h in zip(refs, hyps)]
        wer = W.compute_wer([list(map(str, r)) for r in refs], [list(map(str, h)) for h in hyps])
        wcases.append(dict(refs=refs, hyps=hyps, dist=dist, wer=wer))
    with open(os.path.join(HERE, "wer.json"), "w") as f:
        json.dump(wcases, f)
    print("wrote sampler.json (%d cases), wer.json (%d cases)" % (len(cases), len(wcases)))


if __name__ == "__main__":
    main()
