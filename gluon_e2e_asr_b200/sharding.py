"""Batch sharding for the CTC path.

``split_and_load`` keeps the semantics of /root/reference/scripts/swbd/utils.py:25-33
(contiguous ``n // k`` chunks, remainder on the last device, everything on device 0 when
``n < k``).  Utterances are independent, so sharding needs no tensor exchange; the only
collective on the path is the scalar loss-sum all-reduce, which replaces the reference's
host-side ``+=`` of ``.asscalar()`` values (train_ctc_ce.py:367-368).

Host-side index logic only -- no arithmetic of the path lives here.
"""
from __future__ import annotations

import torch

__all__ = ["split_slices", "split_and_load", "balanced_assignment", "shard_for_rank",
           "loss_sum_allreduce"]


def split_slices(n, k):
    """utils.py:25-33 as index slices."""
    if k <= 0:
        raise ValueError("need at least one device")
    if n < k:
        return [slice(0, n)]
    m = n // k
    return [slice(i * m, (i + 1) * m) for i in range(k - 1)] + [slice((k - 1) * m, n)]


def split_and_load(data, ctx):
    """``utils.split_and_load(data, ctx)``: ctx is a list of torch devices."""
    return [data[s].to(ctx[i], non_blocking=True) for i, s in enumerate(split_slices(data.shape[0], len(ctx)))]


def balanced_assignment(frame_lengths, label_lengths, k):
    """Cost-balanced alternative: greedy longest-processing-time on T_b * (2 L_b + 1).
    Length-bucketed batches (gluonE2EASR/data/sampler.py:205-213) put the longest utterances
    first, so contiguous slicing overloads device 0.  Returns k index lists; per-utterance
    results are identical to any other assignment."""
    cost = [float(t) * (2.0 * float(l) + 1.0) for t, l in zip(frame_lengths, label_lengths)]
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    loads = [0.0] * k
    out = [[] for _ in range(k)]
    for i in order:
        j = min(range(k), key=lambda r: (loads[r], r))
        out[j].append(i)
        loads[j] += cost[i]
    return [sorted(x) for x in out]


def shard_for_rank(n, rank, world_size):
    """Reference-mode slice of this rank (empty when n < world_size and rank > 0)."""
    s = split_slices(n, world_size)
    return s[rank] if rank < len(s) else slice(0, 0)


def loss_sum_allreduce(values, group=None, async_op=False):
    """Sum a small tensor of per-rank partial sums (loss sum, frame count, utterance count)
    over all ranks with one all-reduce (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
