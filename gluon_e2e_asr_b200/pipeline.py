"""Prefetching host entry of the CTC path: ``HostPipeline`` over the C ABI's ``ctcb_pipe_*``.

The reference's training loop takes collated host batches from a ``DataLoader`` whose workers
prepare the next batch while the current one trains (scripts/swbd/train_ctc_ce.py:348-355;
gluonE2EASR/data/batchify.py:51 collates into shared host memory), then copies the four arrays
to the device and calls the loss.  ``HostPipeline`` is that step for host-resident batches with
the copy of batch i+1 overlapping the kernels of batch i: ``submit`` enqueues one pinned batch
and returns a ticket at once, ``wait`` returns when that batch's loss is on the host and its
gradient is ready on the device.  Host-side plumbing only: every byte of arithmetic is in
libctcb.so's kernels, and there is no CPU fallback (a missing library or GPU raises).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .batch import PinnedBatch

__all__ = ["HostPipeline"]

_DT = {torch.int32: _lib.DT_I32, torch.int64: _lib.DT_I64, torch.float32: _lib.DT_F32, torch.float64: _lib.DT_F64}


class _DevView:
    """``__cuda_array_interface__`` holder for a library-owned device buffer."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class HostPipeline:
    """``depth`` batches in flight on ``device`` (``ctcb_pipe_create``).

    ``submit(batch, loss_out, blank_label='first')`` takes a ``PinnedBatch`` (NTC logits, NT labels,
    both length vectors) and a pinned float32 ``(B,)`` tensor for the loss; ``wait(ticket)`` blocks
    until that batch is done and returns its gradient as a CUDA tensor *view* of library memory
    (valid until ``depth`` more submits).  One thread per pipeline.
    """

    def __init__(self, device=0, depth=2):
        self._lib = _lib.load()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline has no CPU path: a CUDA device is required")
        self.depth = int(depth)
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.ctcb_pipe_create(self.device.index, self.depth, ctypes.byref(self._h)))
        self._inflight = {}

    def problem(self, batch: PinnedBatch, loss_out, blank_label="first", head_grad=None):
        """The ``ctcb_problem_t`` (HOST pointers) of one pinned batch; reusable across submits."""
        B, T, V, Lmax = batch.shape
        if loss_out.dtype != torch.float32 or loss_out.numel() != B or loss_out.is_cuda:
            raise ValueError("loss_out must be a host float32 tensor of shape (%d,)" % B)
        q = _lib.Problem()
        last = blank_label == "last"
        q.T, q.B, q.V, q.Lmax = T, B, V, Lmax
        q.blank, q.label_pad = (V - 1, -1) if last else (0, 0)
        q.logits, q.logits_stride_t, q.logits_stride_b = batch.pred.data_ptr(), V, T * V
        if getattr(batch, "packed", False):          # valid frames only, back to back: ctcb_problem_t.logits_row_offsets
            q.logits_row_offsets = batch.row_offsets.data_ptr()
        q.labels, q.label_dtype = batch.label.data_ptr(), _DT[batch.label.dtype]
        q.label_stride_b, q.label_stride_l = Lmax, 1
        q.data_lengths, q.data_lengths_dtype = batch.pred_lengths.data_ptr(), _DT[batch.pred_lengths.dtype]
        q.label_lengths, q.label_lengths_dtype = batch.label_lengths.data_ptr(), _DT[batch.label_lengths.dtype]
        if head_grad is not None:
            if head_grad.dtype != torch.float32 or head_grad.numel() != B or head_grad.is_cuda:
                raise ValueError("head_grad must be a host float32 tensor of shape (%d,)" % B)
            q.head_grad = head_grad.data_ptr()
        q.loss = loss_out.data_ptr()
        q._keep = (batch, loss_out, head_grad)
        return q

    def submit_problem(self, q):
        t = ctypes.c_int64(-1)
        _lib.check(self._lib.ctcb_pipe_submit(self._h, ctypes.byref(q), ctypes.byref(t)))
        self._inflight[t.value] = q
        stale = t.value - self.depth
        self._inflight.pop(stale, None)
        return t.value

    def submit(self, batch, loss_out, blank_label="first", head_grad=None):
        return self.submit_problem(self.problem(batch, loss_out, blank_label, head_grad))

    def wait(self, ticket, as_tensor=True):
        g = ctypes.c_void_p()
        _lib.check(self._lib.ctcb_pipe_wait(self._h, ticket, ctypes.byref(g)))
        q = self._inflight.get(ticket)
        if not as_tensor or q is None:
            return g.value
        return torch.as_tensor(_DevView(g.value, (q.B, q.T, q.V)), device=self.device)

    def last_h2d_bytes(self):
        """(bytes the last submit moved host->device, whether the logits were pulled by the GPU: valid frames only)."""
        b, pulled = ctypes.c_int64(0), ctypes.c_int32(0)
        _lib.check(self._lib.ctcb_pipe_last_h2d_bytes(self._h, ctypes.byref(b), ctypes.byref(pulled)))
        return b.value, bool(pulled.value)

    def close(self):
        if self._h:
            self._lib.ctcb_pipe_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
