"""Length-bucketed batch samplers with the semantics of the reference's data layer.

Restates the observable behaviour of /root/reference/gluonE2EASR/data/sampler.py --
``SortedSampler`` (:58-77), ``FixedBucketSampler`` (:80-248), ``SortedBucketSampler`` (:251-312)
-- as used by scripts/swbd/train_ctc_ce.py:237-256 with ``(T, L)`` length pairs: which bucket a
sample lands in (the bucket whose keys all cover it with the least total padding, :30-55), the
bucket keys generated from the length range (:150-163), the batch-size scale-up of short buckets
(:199-203), the batch order (:204-213) and, for ``shuffle=True``, the exact sequence of draws
from numpy's global generator (:216-224), so a run with the same ``np.random.seed`` yields the
same batches as the reference (tests/golden/sampler.json was produced by the reference's file).

These samplers decide how ragged a CTC batch is (how many padded frames the loss kernels skip)
and, with ``sharding.balanced_assignment``, which GPU gets which utterances.  Host-side index
logic only.
"""
from __future__ import annotations

import logging

import numpy as np

__all__ = ["SortedSampler", "FixedBucketSampler", "SortedBucketSampler"]

logger = logging.getLogger(__name__)


class SortedSampler:
    """Indices in the order of ``sort_keys`` (descending by default); stable for equal keys."""

    def __init__(self, sort_keys, reverse=True):
        assert len(sort_keys) > 0
        self._ids = sorted(range(len(sort_keys)), key=lambda i: sort_keys[i], reverse=reverse)

    def __iter__(self):
        return iter(self._ids)

    def __len__(self):
        return len(self._ids)


def _assign_buckets(keys, lengths):
    """Bucket index of every sample: among the buckets whose key covers the sample in every
    component, the one with the smallest total padding (first one on ties)."""
    keys = np.asarray(keys, dtype=np.int64)
    lens = np.asarray(lengths, dtype=np.int64)
    if keys.ndim == 1:
        keys, lens = keys[:, None], lens[:, None]
    slack = keys[None, :, :] - lens[:, None, :]                   # (samples, buckets, attrs)
    fits = (slack >= 0).all(axis=2)
    if not fits.any(axis=1).all():
        bad = np.nonzero(~fits.any(axis=1))[0]
        raise ValueError("Find elements in seq_lengths that cannot fit in the given buckets, seq_length=%s, "
                         "bucket_keys=%s. You must increase the bucket size."
                         % (str(np.asarray(lengths)[bad]), str([tuple(k) for k in keys.tolist()])))
    pad = np.where(fits, slack.sum(axis=2), np.iinfo(np.int64).max)
    return pad.argmin(axis=1)


class FixedBucketSampler:
    """Batches of sample indices drawn bucket by bucket (reference :80-248).

    lengths: ints or equal-length tuples of ints (the CTC script passes ``(T, L)``);
    num_buckets / bucket_keys: as in the reference (keys generated from the length range when not
    given); ratio: batch-size scale-up of the buckets with short keys; shuffle: reshuffle the
    batch order and every bucket at each ``__iter__`` with numpy's global generator;
    reverse: longest bucket first (the default) or shortest first.
    """

    def __init__(self, lengths, batch_size, num_buckets=10, bucket_keys=None, ratio=0, shuffle=False, reverse=True):
        assert len(lengths) > 0, "FixedBucketSampler does not support empty lengths."
        assert batch_size > 0, "Batch size must be larger than 0."
        assert ratio >= 0, "batch size scaling ratio cannot be negative."
        self._lengths = np.array(lengths, dtype=np.int32)
        assert self._lengths.ndim in (1, 2), \
            "Elements in lengths must be either int or tuple/list of int. Received lengths=%s" % str(lengths)
        self._single = self._lengths.ndim == 1
        self._batch_size, self._ratio, self._shuffle = batch_size, ratio, shuffle
        hi, lo = self._lengths.max(axis=0), self._lengths.min(axis=0)
        assert np.all(lo > 0), "Sequence lengths must all be larger than 0."
        if bucket_keys is None:
            assert num_buckets > 0, "num_buckets must be set when bucket_keys is None. Received num_buckets=%d" % num_buckets
            if self._single:
                width = max((hi - lo) // num_buckets, 1)
                bucket_keys = [max(hi - i * width, lo) for i in range(num_buckets)]
            else:
                widths = [max((h - l) // num_buckets, 1) for h, l in zip(hi, lo)]
                bucket_keys = [tuple(max(h - i * w, l) for h, l, w in zip(hi, lo, widths)) for i in range(num_buckets)]
        else:
            if num_buckets is not None:
                logger.warning("num_buckets will not be used if bucket_keys is not None. bucket_keys=%s, num_buckets=%d"
                               % (str(bucket_keys), num_buckets))
            assert len(bucket_keys) > 0
            if self._single:
                assert isinstance(bucket_keys[0], int)
            else:
                assert isinstance(bucket_keys[0], tuple) and len(bucket_keys[0]) == self._lengths.shape[1]
        bucket_keys = sorted(set(bucket_keys))
        which = _assign_buckets(bucket_keys, self._lengths)
        members = [np.nonzero(which == k)[0].tolist() for k in range(len(bucket_keys))]
        empty = [key for key, m in zip(bucket_keys, members) if not m]
        if empty:
            logger.warning("Some buckets are empty and will be removed. Unused bucket keys=%s" % str(empty))
        self._bucket_keys = [key for key, m in zip(bucket_keys, members) if m]
        self._bucket_sample_ids = [m for m in members if m]
        weight = [key if self._single else sum(key) for key in self._bucket_keys]
        top = max(weight)
        self._bucket_batch_sizes = [max(int(top / float(wk) * ratio * batch_size), batch_size) for wk in weight]
        # batch table: longest bucket first, a bucket's batches in order
        self._batch_infos = [(k, begin)
                             for k in range(len(self._bucket_keys) - 1, -1, -1)
                             for begin in range(0, len(self._bucket_sample_ids[k]), self._bucket_batch_sizes[k])]
        if not reverse:
            self._batch_infos = sorted(self._batch_infos, key=lambda kb: kb[0])

    def __iter__(self):
        if self._shuffle:
            np.random.shuffle(self._batch_infos)
            for ids in self._bucket_sample_ids:
                np.random.shuffle(ids)
        for k, begin in self._batch_infos:
            yield self._bucket_sample_ids[k][begin:begin + self._bucket_batch_sizes[k]]

    def __len__(self):
        return len(self._batch_infos)

    @property
    def bucket_keys(self):
        return list(self._bucket_keys)

    @property
    def bucket_batch_sizes(self):
        return list(self._bucket_batch_sizes)

    def padded_fraction(self):
        """Share of the padded (key-sized) frames of an epoch that are padding: what bucketing buys the
        loss kernels, which skip frames >= T_b."""
        first = self._lengths if self._single else self._lengths[:, 0]
        total = sum((key if self._single else key[0]) * len(m) for key, m in zip(self._bucket_keys, self._bucket_sample_ids))
        return 1.0 - float(first.sum()) / float(total)

    def stats(self):
        return ("{name}:\n  sample_num={n}, batch_num={b}\n  key={k}\n  cnt={c}\n  batch_size={s}"
                .format(name=self.__class__.__name__, n=len(self._lengths), b=len(self._batch_infos),
                        k=self._bucket_keys, c=[len(m) for m in self._bucket_sample_ids], s=self._bucket_batch_sizes))


class SortedBucketSampler:
    """Batches from sorted chunks of ``mult * batch_size`` samples (reference :251-312)."""

    def __init__(self, sort_keys, batch_size, mult=100, reverse=True, shuffle=False):
        assert len(sort_keys) > 0
        assert batch_size > 0
        assert mult >= 1, "Bucket size multiplier must be larger than 1"
        self._keys, self._batch_size, self._mult = sort_keys, batch_size, mult
        self._reverse, self._shuffle = reverse, shuffle

    def __iter__(self):
        n = len(self._keys)
        ids = np.random.permutation(n) if self._shuffle else list(range(n))
        chunk = int(self._mult * self._batch_size)
        for lo in range(0, n, chunk):
            part = sorted(ids[lo:lo + chunk], key=lambda i: self._keys[i], reverse=self._reverse)
            begins = list(range(0, len(part), self._batch_size))
            if self._shuffle:
                np.random.shuffle(begins)
            for b in begins:
                yield part[b:b + self._batch_size]

    def __len__(self):
        return (len(self._keys) + self._batch_size - 1) // self._batch_size
