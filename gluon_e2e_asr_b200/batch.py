"""Pinned-memory collation of one CTC batch and its single-copy transfer to the GPU.

The reference collates a batch with ``btf.Tuple(btf.Pad(...), btf.Stack(), btf.Pad(pad_val=0),
btf.Stack())`` into *shared* host memory (gluonE2EASR/data/batchify.py:51, :135;
scripts/swbd/train_ctc_ce.py:233-236) and moves the four arrays to the device one by one with
``split_and_load`` -> ``as_in_context`` (scripts/swbd/utils.py:25-33, train_ctc_ce.py:352-355).
Here the four arrays of the CTC call -- logits ``(B, T, V)`` float32 (the model's NTC output),
0-padded labels ``(B, Lmax)``, and the two length vectors -- live in ONE page-locked arena, so
the host->device step of the path is one ``cudaMemcpyAsync`` on the caller's stream into a
persistent device arena of the same layout; the per-field device tensors are views created
once.  Host-side layout/plumbing only: no arithmetic of the path lives here.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["PinnedBatch"]

_ALIGN = 256


def _round_up(n, a=_ALIGN):
    return (n + a - 1) // a * a


class PinnedBatch:
    """Arena for ``(pred, label, pred_lengths, label_lengths)`` of a fixed shape.

    ``host`` holds numpy-writable pinned views (``.pred``, ``.label``, ``.pred_lengths``,
    ``.label_lengths`` as torch CPU tensors sharing the arena); ``load(device)`` issues one
    asynchronous copy and returns the matching device views.
    """

    FIELDS = ("pred", "label", "pred_lengths", "label_lengths")

    def __init__(self, B, T, V, Lmax, label_dtype=torch.float32, length_dtype=torch.float32, pin=True):
        self.shape = (B, T, V, Lmax)
        specs = (("pred", (B, T, V), torch.float32), ("label", (B, Lmax), label_dtype),
                 ("pred_lengths", (B,), length_dtype), ("label_lengths", (B,), length_dtype))
        self._layout, off = [], 0
        for name, shp, dt in specs:
            n = int(np.prod(shp)) * torch.empty((), dtype=dt).element_size()
            self._layout.append((name, shp, dt, off, n))
            off = _round_up(off + n)
        self.nbytes = off
        self.arena = torch.empty((self.nbytes,), dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            self.arena = self.arena.pin_memory()
        self.host = self._views(self.arena)
        for k, v in self.host.items():
            setattr(self, k, v)
        self._dev = {}
        # bytes the transfer moves (padding between fields included)
        self.h2d_bytes = self.nbytes

    def _views(self, arena):
        return {name: arena[off:off + n].view(dt).view(shp) for name, shp, dt, off, n in self._layout}

    def fill(self, pred, label, pred_lengths, label_lengths):
        """Copy one collated batch (numpy arrays or tensors) into the arena; returns self."""
        for name, src in zip(self.FIELDS, (pred, label, pred_lengths, label_lengths)):
            dst = self.host[name]
            dst.copy_(torch.as_tensor(src).to(dst.dtype).reshape(dst.shape))
        return self

    @classmethod
    def from_arrays(cls, pred, label, pred_lengths, label_lengths, pin=True):
        pred = np.asarray(pred)
        label = np.asarray(label)
        B, T, V = pred.shape
        out = cls(B, T, V, label.shape[1], label_dtype=torch.as_tensor(label).dtype,
                  length_dtype=torch.as_tensor(np.asarray(pred_lengths)).dtype, pin=pin)
        return out.fill(pred, label, pred_lengths, label_lengths)

    def load(self, device, non_blocking=True):
        """One host->device copy of the whole arena on the current stream of ``device``; returns the
        device views ``{pred, label, pred_lengths, label_lengths}`` (the same tensor objects on every
        call: the device arena is persistent, stream order protects it)."""
        device = torch.device(device)
        slot = self._dev.get(device)
        if slot is None:
            darena = torch.empty((self.nbytes,), dtype=torch.uint8, device=device)
            slot = self._dev[device] = (darena, self._views(darena))
        slot[0].copy_(self.arena, non_blocking=non_blocking)
        return slot[1]
