"""CPU tests of the pinned-arena collation (gluon_e2e_asr_b200/batch.py): layout and content of
the arena that replaces the reference's four host arrays (data/batchify.py:51, :135;
train_ctc_ce.py:233-236, :352-355).  The single-copy transfer itself is a GPU test."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gluon_e2e_asr_b200.batch import PinnedBatch
from tests.synth import make_batch


def test_arena_layout_and_content():
    d = make_batch(5, 33, 9, 7, seed=3)
    pb = PinnedBatch.from_arrays(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"], pin=False)
    assert pb.shape == (5, 33, 9, 7)
    offs = [off for _, _, _, off, _ in pb._layout]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert pb.nbytes >= 4 * (5 * 33 * 9 + 5 * 7 + 5 + 5) and pb.nbytes % 256 == 0
    for k in PinnedBatch.FIELDS:
        np.testing.assert_array_equal(pb.host[k].numpy(), d[k])
        assert pb.host[k].dtype == torch.float32              # the reference's dtypes (reader_kaldi_io.py:33-35)
    # the views alias the arena: writing a field changes the arena bytes that get copied
    pb.label_lengths[0] = 3
    x = pb.load("cpu")
    assert x["label_lengths"][0].item() == 3
    assert x["pred"].shape == (5, 33, 9) and x["label"].shape == (5, 7)
    np.testing.assert_array_equal(x["pred"].numpy(), d["pred"])


def test_int_labels_keep_their_dtype():
    d = make_batch(3, 20, 6, 4, seed=4)
    pb = PinnedBatch.from_arrays(d["pred"], d["label"].astype(np.int32), d["pred_lengths"], d["label_lengths"], pin=False)
    assert pb.label.dtype == torch.int32
    pb.fill(d["pred"] * 2, d["label"], d["pred_lengths"], d["label_lengths"])
    np.testing.assert_array_equal(pb.pred.numpy(), d["pred"] * 2)
    np.testing.assert_array_equal(pb.label.numpy(), d["label"].astype(np.int32))
