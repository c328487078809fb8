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
           "loss_sum_allreduce", "PeerLossSum"]


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


class PeerLossSum:
    """The loss-sum exchange without a collective kernel on the step's path (``ctcb_mailbox_*``).

    One process per GPU (``torch.distributed`` initialised): every rank owns a small mailbox in
    device memory that its peers map through CUDA IPC; ``exchange(values, out)`` enqueues one tiny
    kernel on the current stream that (1) writes into ``out`` the all-rank sum of the values handed
    to the exchange ``lag`` calls earlier and (2) stores ``values`` into every rank's mailbox over
    NVLink peer access and zeroes them.  Ranks never rendezvous: a rank only waits for what its peers
    stored ``lag`` exchanges earlier, so a slow rank costs the others nothing until it is ``lag`` steps
    behind (the slack absorbs the jitter of data-dependent step times).
    ``flush(out)`` returns the sum of the last exchange.  Replaces the reference's host-side ``+=`` of
    two ``.asscalar()`` values per shard and step (train_ctc_ce.py:367-368).  CUDA only.
    """

    def __init__(self, device, group=None, lag=4):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        self._lib, self._ct = _lib, ctypes
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PeerLossSum has no CPU path: a CUDA device is required")
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        lib = _lib.load()
        self._h = ctypes.c_void_p()
        self.lag = int(lag)
        _lib.check(lib.ctcb_mailbox_create(self.device.index or 0, self.rank, self.world, self.lag, ctypes.byref(self._h)))
        if self.world > 1:
            mine = (ctypes.c_ubyte * 64)()
            _lib.check(lib.ctcb_mailbox_handle(self._h, mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine), group=group)
            buf = (ctypes.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
            rc = lib.ctcb_mailbox_connect(self._h, buf)
            err = lib.ctcb_last_error().decode("utf-8", "replace") if rc else ""
            # every rank learns whether every rank connected: a one-sided failure must not leave the others waiting
            oks = [None] * self.world
            dist.all_gather_object(oks, rc == 0, group=group)
            if not all(oks):
                self.close()
                raise RuntimeError("PeerLossSum: peer mapping failed on rank(s) %s %s" %
                                   ([r for r, ok in enumerate(oks) if not ok], err))

    def exchange(self, values, out):
        """values, out: float64 CUDA tensors of the same length (<= 6), contiguous."""
        self._check(values); self._check(out)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._lib.check(self._lib.load().ctcb_mailbox_exchange(self._h, values.data_ptr(), values.numel(), out.data_ptr(), stream))

    def exchange_with_next(self, values, out):
        """The same exchange as part of the next ``ctc_loss_and_grad`` / forward / backward call of this
        thread (``ctcb_mailbox_exchange_with_next``): the exchange kernel is launched behind that
        call's gradient kernel as its programmatic dependent and runs beside its last wave.
        ``values`` must be the PREVIOUS step's partial sums, not the tensor that call accumulates
        into."""
        self._check(values); self._check(out)
        self._lib.check(self._lib.load().ctcb_mailbox_exchange_with_next(self._h, values.data_ptr(), values.numel(), out.data_ptr()))

    def flush(self, out):
        self._check(out)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._lib.check(self._lib.load().ctcb_mailbox_flush(self._h, out.numel(), out.data_ptr(), stream))

    def _check(self, t):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 1 <= t.numel() <= 6):
            raise ValueError("expected a contiguous float64 CUDA tensor of 1..6 elements")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.load().ctcb_mailbox_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
