"""Word error rate with the reference's call surface, computed on the GPU.

``compute_wer(reference_corpus, translation_corpus, lower_case=False)`` keeps the name, arguments
and return value of /root/reference/scripts/swbd/wer.py:9-43 (total edit distance / total
reference length over the corpus); the O(N*M) dynamic programme of ``_edit_distance`` (:45-68) runs
in libctcb's ``k_edit_distance`` (one CTA per sentence pair) instead of a Python double loop.
``wer_from_tokens`` is the form the validation loop wants: hypotheses straight from
``greedy_decode`` (train_ctc_ce.py:149-168), nothing per token on the host.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["compute_wer", "wer_from_tokens"]


def _pack(rows, device):
    n = max(1, max((len(r) for r in rows), default=1))
    t = torch.zeros((len(rows), n), dtype=torch.int32)
    for i, r in enumerate(rows):
        if r:
            t[i, :len(r)] = torch.tensor(r, dtype=torch.int32)
    lens = torch.tensor([len(r) for r in rows], dtype=torch.int32)
    return t.to(device), lens.to(device)


def compute_wer(reference_corpus, translation_corpus, lower_case=False, device=None):
    """wer.py:9-43: sum of edit distances over sum of reference lengths.  Tokens (any hashable,
    strings in the reference) are mapped to integer ids on the host; the distances come from the GPU."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("gluon_e2e_asr_b200.wer has no CPU path: a CUDA device is required")
        device = torch.device("cuda", torch.cuda.current_device())
    ids = {}

    def enc(sent):
        if lower_case:
            sent = [str.lower(w) for w in sent]
        return [ids.setdefault(w, len(ids) + 1) for w in sent]

    pairs = list(zip(reference_corpus, translation_corpus))
    refs = [enc(r) for r, _ in pairs]
    hyps = [enc(h) for _, h in pairs]
    ref, ref_len = _pack(refs, device)
    hyp, hyp_len = _pack(hyps, device)
    totals = torch.zeros((2,), dtype=torch.int64, device=device)
    ops.edit_distance(ref, ref_len, hyp, hyp_len, totals=totals)
    d, n = totals.tolist()
    return d / n


def wer_from_tokens(ref, ref_lengths, hyp, hyp_lengths):
    """(wer, distances) for integer token tensors already on the device."""
    totals = torch.zeros((2,), dtype=torch.int64, device=ref.device)
    dist = ops.edit_distance(ref, ref_lengths, hyp, hyp_lengths, totals=totals)
    d, n = totals.tolist()
    return d / n, dist
