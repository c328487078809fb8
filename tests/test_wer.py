"""Edit distance / WER (next row): the oracle against fixtures produced by the
reference's scripts/swbd/wer.py (tests/golden/make_next_rows_golden.py), and -- on the GPU -- the
kernel against both, bit-exact (integer work)."""
import json
import os

import numpy as np
import pytest

from oracle import ctc_oracle as O


@pytest.fixture(scope="module")
def wcases(golden_dir):
    with open(os.path.join(golden_dir, "wer.json")) as f:
        return json.load(f)


def test_oracle_matches_the_reference_fixtures(wcases):
    for c in wcases:
        assert [O.edit_distance(r, h) for r, h in zip(c["refs"], c["hyps"])] == c["dist"]
        assert O.compute_wer([list(map(str, r)) for r in c["refs"]], [list(map(str, h)) for h in c["hyps"]]) == c["wer"]
    assert O.edit_distance([], [1, 2]) == 2 and O.edit_distance([3], []) == 1
    assert O.compute_wer([["A", "b"]], [["a", "B"]], lower_case=True) == 0.0


@pytest.mark.gpu
def test_gpu_edit_distance_and_wer(wcases):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gluon_e2e_asr_b200 import edit_distance, greedy_decode
    from gluon_e2e_asr_b200.wer import compute_wer, wer_from_tokens, _pack
    dev = torch.device("cuda:0")
    for c in wcases:
        ref, rl = _pack(c["refs"], dev)
        hyp, hl = _pack(c["hyps"], dev)
        assert edit_distance(ref, rl, hyp, hl).cpu().tolist() == c["dist"]
        wer = compute_wer([list(map(str, r)) for r in c["refs"]], [list(map(str, h)) for h in c["hyps"]])
        assert wer == c["wer"]
    # long, ragged sequences against the oracle; hypotheses straight from the greedy decoder
    from tests.synth import make_batch
    d = make_batch(12, 700, 46, 160, seed=31, peaky=True)
    pred = torch.tensor(d["pred"], device=dev); pl = torch.tensor(d["pred_lengths"], device=dev)
    toks, lens = greedy_decode(pred, pl)
    ref = torch.tensor(d["label"], device=dev).to(torch.int32); rl = torch.tensor(d["label_lengths"], device=dev).to(torch.int32)
    wer, dist = wer_from_tokens(ref, rl, toks, lens)
    hyps = [toks[b, :lens[b]].cpu().tolist() for b in range(12)]
    refs = [d["label"][b, :int(d["label_lengths"][b])].astype(int).tolist() for b in range(12)]
    assert dist.cpu().tolist() == [O.edit_distance(r, h) for r, h in zip(refs, hyps)]
    assert wer == O.compute_wer(refs, hyps)
    # empty hypothesis / empty reference rows
    z = torch.zeros((2, 4), dtype=torch.int32, device=dev)
    out = edit_distance(z, torch.tensor([0, 3], dtype=torch.int32, device=dev), z + 1, torch.tensor([2, 0], dtype=torch.int32, device=dev))
    assert out.cpu().tolist() == [2, 3]
    with pytest.raises(RuntimeError):
        edit_distance(z.cpu(), lens.cpu(), z.cpu(), lens.cpu())
