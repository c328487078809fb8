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

    def __init__(self, B, T, V, Lmax, label_dtype=torch.float32, length_dtype=torch.float32, pin=True, packed_frames=None):
        """``packed_frames`` (= sum of the valid frame counts): the logits field holds only the VALID frames of
        every utterance, back to back, plus an int64 ``row_offsets`` field (``ctcb_problem_t.logits_row_offsets``):
        the padded frames of a length-bucketed batch never cross PCIe.  For the C ABI's host entries
        (``HostPipeline``); ``load()`` -- dense device views for the torch plugin -- needs the dense layout."""
        self.shape = (B, T, V, Lmax)
        self.packed = packed_frames is not None
        specs = (("pred", (packed_frames, V) if self.packed else (B, T, V), torch.float32), ("label", (B, Lmax), label_dtype),
                 ("pred_lengths", (B,), length_dtype), ("label_lengths", (B,), length_dtype))
        if self.packed:
            specs = specs + (("row_offsets", (B,), torch.int64),)
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
        """Copy one collated batch (numpy arrays or tensors; ``pred`` dense ``(B, T, V)``) into the arena; returns self."""
        for name, src in zip(self.FIELDS[1:], (label, pred_lengths, label_lengths)):
            dst = self.host[name]
            dst.copy_(torch.as_tensor(src).to(dst.dtype).reshape(dst.shape))
        if not self.packed:
            self.host["pred"].copy_(torch.as_tensor(pred).to(torch.float32).reshape(self.host["pred"].shape))
            return self
        B, T, V, _ = self.shape
        x = torch.as_tensor(pred)
        n = np.clip(np.asarray(pred_lengths).astype(np.int64), 0, T)
        off = np.concatenate([[0], np.cumsum(n)[:-1]])
        if int(n.sum()) != self.host["pred"].shape[0]:
            raise ValueError("packed arena holds %d frames, the batch has %d valid ones" % (self.host["pred"].shape[0], int(n.sum())))
        for b in range(B):
            self.host["pred"][off[b]:off[b] + n[b]].copy_(x[b, :n[b]])
        self.host["row_offsets"].copy_(torch.as_tensor(off * V))
        return self

    @classmethod
    def from_arrays(cls, pred, label, pred_lengths, label_lengths, pin=True, packed=False):
        pred = np.asarray(pred)
        label = np.asarray(label)
        B, T, V = pred.shape
        frames = int(np.clip(np.asarray(pred_lengths).astype(np.int64), 0, T).sum()) if packed else None
        out = cls(B, T, V, label.shape[1], label_dtype=torch.as_tensor(label).dtype,
                  length_dtype=torch.as_tensor(np.asarray(pred_lengths)).dtype, pin=pin, packed_frames=frames)
        return out.fill(pred, label, pred_lengths, label_lengths)

    def load(self, device, non_blocking=True):
        """One host->device copy of the whole arena on the current stream of ``device``; returns the
        device views ``{pred, label, pred_lengths, label_lengths}`` (the same tensor objects on every
        call: the device arena is persistent, stream order protects it)."""
        if self.packed:
            raise RuntimeError("PinnedBatch.load needs the dense layout (packed batches go through HostPipeline)")
        device = torch.device(device)
        slot = self._dev.get(device)
        if slot is None:
            darena = torch.empty((self.nbytes,), dtype=torch.uint8, device=device)
            slot = self._dev[device] = (darena, self._views(darena))
        slot[0].copy_(self.arena, non_blocking=non_blocking)
        return slot[1]
