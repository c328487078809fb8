"""CPU tests of the oracle itself: oracle/ctc_oracle.py (numpy fp64) and oracle/ctc_ref.c
(C restatement, fp32 + fp64) against the committed golden vectors (upstream KATs K1-K5,
torch-fp64 fixtures), closed forms, finite differences and each other."""
import json
import os

import numpy as np
import pytest

from oracle import ctc_oracle as O
from oracle import ctc_ref as R
from tests.synth import make_batch


def _kats(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        return json.load(f)


def test_kat_numpy_oracle(golden_dir):
    for k in _kats(golden_dir):
        orc = O.CtcLossOracle(layout=k["layout"], label_layout="NT", blank_label=k["blank_label"])
        loss, _, ok = orc(np.array(k["data"], np.float32), np.array(k["label"]))
        assert ok.all()
        np.testing.assert_allclose(loss, k["expect"], rtol=k["rtol"], err_msg=k["name"])


def test_kat_k4_variants(golden_dir):
    k = [k for k in _kats(golden_dir) if k["name"].startswith("K4")][0]
    data = np.array(k["data"], np.float32)
    lab = np.array(k["label"])
    exp = np.array(k["expect"])
    ntc = O.CtcLossOracle("NTC", "NT", "last")
    tnc = O.CtcLossOracle("TNC", "NT", "last")
    tn = O.CtcLossOracle("NTC", "TN", "last")
    np.testing.assert_allclose(tnc(data.swapaxes(0, 1), lab)[0], exp, rtol=2e-7)
    np.testing.assert_allclose(tn(data, lab.T)[0], exp, rtol=2e-7)
    np.testing.assert_allclose(ntc(data, lab, None, np.array([2, 3]))[0], exp, rtol=2e-7)
    np.testing.assert_allclose(ntc(data, lab, np.array([20, 20]), np.array([2, 3]))[0], exp, rtol=2e-7)
    # shorter pred_lengths change the answer and only read the first frames
    a = ntc(data, lab, np.array([10, 10]), np.array([2, 3]))[0]
    b = ntc(data[:, :10], lab, None, np.array([2, 3]))[0]
    np.testing.assert_allclose(a, b, rtol=1e-12)


def test_kat_c_port(golden_dir):
    for k in _kats(golden_dir):
        data = np.array(k["data"], np.float64)
        if k["layout"] == "NTC":
            data = data.swapaxes(0, 1)
        lab = np.array(k["label"])
        T = data.shape[0]
        V = data.shape[2]
        blank = 0 if k["blank_label"] == "first" else V - 1
        Tb, Lb = O.infer_lengths(lab, None, None, T, False, False, k["blank_label"])
        for dt, rt in ((np.float64, k["rtol"]), (np.float32, max(k["rtol"], 2e-6))):
            loss, _, ok = R.ctc_ref(data, lab, Tb, Lb, blank=blank, dtype=dt)
            assert ok.all()
            np.testing.assert_allclose(loss, k["expect"], rtol=rt, err_msg=k["name"])


def _golden_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "torch_fp64.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return [(n, {k.split("/")[1]: z[k] for k in z.files if k.startswith(n + "/")}) for n in names]


def test_torch_fp64_fixtures_numpy_and_c(golden_dir):
    for name, c in _golden_cases(golden_dir):
        blank = int(c["blank"])
        loss, grad, ok = O.ctc_loss_grad(c["data"], c["label"], c["T_b"], c["L_b"], blank, c["head"])
        assert ok.all(), name
        np.testing.assert_allclose(loss, c["loss"], rtol=1e-12, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(grad, c["grad"], rtol=1e-10, atol=1e-12, err_msg=name)
        l2, g2, ok2 = R.ctc_ref(c["data"], c["label"], c["T_b"], c["L_b"], blank, c["head"], dtype=np.float64)
        assert ok2.all()
        np.testing.assert_allclose(l2, c["loss"], rtol=1e-11, atol=1e-11, err_msg=name)
        np.testing.assert_allclose(g2, c["grad"], rtol=1e-8, atol=1e-11, err_msg=name)


def test_closed_forms():
    rng = np.random.default_rng(7)
    T, V = 9, 6
    x = rng.standard_normal((T, 1, V))
    logy = x - np.log(np.exp(x).sum(-1, keepdims=True))
    # L = 0: the only path is all blanks
    loss, g, ok = O.ctc_loss_grad(x, np.zeros((1, 1)), [T], [0])
    np.testing.assert_allclose(loss[0], -logy[:, 0, 0].sum(), rtol=1e-13)
    # T == L, no repeats: single path
    lab = np.array([[1, 2, 3, 4, 5, 1, 2, 3, 4]])
    loss, g, ok = O.ctc_loss_grad(x, lab, [T], [T])
    np.testing.assert_allclose(loss[0], -logy[np.arange(T), 0, lab[0]].sum(), rtol=1e-13)
    # single-path gradient: softmax - onehot
    y = np.exp(logy[:, 0])
    oh = np.zeros_like(y); oh[np.arange(T), lab[0]] = 1
    np.testing.assert_allclose(g[:, 0], y - oh, atol=1e-13)


def test_infeasible_defined_behaviour():
    x = np.random.default_rng(0).standard_normal((3, 2, 5))
    lab = np.array([[1, 1, 2], [1, 2, 3]])        # row 0 needs 4 frames (one repeat), row 1 fits
    loss, g, ok = O.ctc_loss_grad(x, lab, [3, 3], [3, 3])
    assert list(ok) == [False, True]
    assert loss[0] == 0 and np.all(g[:, 0] == 0) and loss[1] > 0
    l2, g2, ok2 = R.ctc_ref(x, lab, [3, 3], [3, 3], dtype=np.float64)
    assert list(ok2) == [False, True] and l2[0] == 0 and np.all(g2[:, 0] == 0)


def test_finite_difference_gradient():
    rng = np.random.default_rng(3)
    T, B, V, L = 7, 2, 5, 3
    x = rng.standard_normal((T, B, V))
    lab = np.array([[1, 1, 2], [3, 4, 0]])
    Tb, Lb = np.array([7, 5]), np.array([3, 2])
    head = np.array([0.7, 1.3])
    _, g, _ = O.ctc_loss_grad(x, lab, Tb, Lb, head_grad=head)
    eps = 1e-6
    num = np.zeros_like(x)
    for idx in np.ndindex(*x.shape):
        xp = x.copy(); xp[idx] += eps
        xm = x.copy(); xm[idx] -= eps
        lp = (O.ctc_loss_grad(xp, lab, Tb, Lb)[0] * head).sum()
        lm = (O.ctc_loss_grad(xm, lab, Tb, Lb)[0] * head).sum()
        num[idx] = (lp - lm) / (2 * eps)
    np.testing.assert_allclose(g, num, atol=2e-8)
    assert np.all(g[5:, 1] == 0)                 # padded frames


def test_properties_batch_vs_single_and_row_sums():
    d = make_batch(6, 40, 11, 9, seed=5)
    x = d["pred"].swapaxes(0, 1)
    Tb, Lb = d["pred_lengths"], d["label_lengths"]
    loss, g, ok = O.ctc_loss_grad(x, d["label"], Tb, Lb)
    assert ok.all()
    for b in range(6):
        l1, g1, _ = O.ctc_loss_grad(x[:, b:b + 1], d["label"][b:b + 1], Tb[b:b + 1], Lb[b:b + 1])
        np.testing.assert_allclose(l1[0], loss[b], rtol=1e-13)
        np.testing.assert_allclose(g1[:, 0], g[:, b], atol=1e-14)
        np.testing.assert_allclose(g[:int(Tb[b]), b].sum(-1), 0, atol=1e-12)   # sum_v G = 0
        assert np.all(g[int(Tb[b]):, b] == 0)


def test_length_inference():
    lab = np.array([[3, 2, 0, 0], [1, 2, 3, 4], [0, 0, 0, 0]], np.float32)
    Tb, Lb = O.infer_lengths(lab, None, None, 10, False, False, "first")
    assert list(Tb) == [10, 10, 10] and list(Lb) == [2, 4, 0]
    lab2 = np.array([[3, 0, -1, -1], [-1, 2, 3, 4]], np.float32)
    assert list(O.infer_lengths(lab2, None, None, 5, False, False, "last")[1]) == [2, 0]
    Tb, Lb = O.infer_lengths(lab, np.array([3.9, 2.0, 7.0], np.float32), np.array([1.0, 3.2, 0.0]), 10, True, True)
    assert list(Tb) == [3, 2, 7] and list(Lb) == [1, 3, 0]


def test_c_fp32_port_is_reference_class_not_oracle():
    """The fp32 port reproduces the log-space fp32 error the survey measured:
    close in loss (relative), far outside rtol 1e-4/atol 1e-5 on the gradient at cfg2 size."""
    d = make_batch(4, 500, 46, 120, seed=0)
    x = d["pred"].swapaxes(0, 1)
    lo, go, _ = O.ctc_loss_grad(x, d["label"], d["pred_lengths"], d["label_lengths"])
    l32, g32, _ = R.ctc_ref(x, d["label"], d["pred_lengths"], d["label_lengths"], dtype=np.float32)
    np.testing.assert_allclose(l32, lo, rtol=1e-5)
    assert np.abs(g32 - go).max() < 2e-2


def test_greedy_decode_and_split():
    logits = np.zeros((1, 6, 4)); path = [0, 2, 2, 0, 2, 3]
    logits[0, np.arange(6), path] = 5
    assert O.greedy_decode(logits, [6]) == [[2, 2, 3]]
    assert O.greedy_decode(logits, [3]) == [[2]]
    s = O.split_and_load_slices(10, 4)
    assert [(x.start, x.stop) for x in s] == [(0, 2), (2, 4), (4, 6), (6, 10)]
    assert [(x.start, x.stop) for x in O.split_and_load_slices(3, 4)] == [(0, 3)]


def test_c_fp64_matches_numpy_at_size():
    """The C restatement in fp64 is the checker of the full-batch-size GPU parity tests
    (tests/test_parity_gpu.py::test_baseline_configs_at_full_batch_size): pin it to the numpy oracle
    on BASELINE-shaped utterances (NTC strides, ragged lengths, repeats, head gradient, peaky logits)."""
    for seed, peaky, shape in ((5, True, (6, 300, 46, 80)), (6, False, (4, 500, 46, 120)), (7, False, (2, 120, 300, 40))):
        d = make_batch(*shape, seed=seed, peaky=peaky)
        head = np.random.default_rng(seed).uniform(0.5, 1.5, shape[0])
        lo, go, ok = O.CtcLossOracle("NTC", "NT")(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"], head_grad=head)
        lc, gc, okc = R.ctc_ref(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"], head_grad=head,
                                layout="NTC", dtype=np.float64)
        assert ok.all() and okc.all()
        np.testing.assert_allclose(lc, lo, rtol=1e-12, atol=1e-11)
        np.testing.assert_allclose(gc, go, rtol=1e-9, atol=1e-12)


def test_greedy_decode_unk_rule():
    """decode_ctc.py:120-140 restated: <unk> -> second best; `prev` stays the raw best of the previous frame."""
    unk = 3
    x = np.zeros((1, 7, 5))
    #        best:  2    unk(2nd 2)  unk(2nd 1)  1    0(blank)  unk(2nd 0)  4
    best = [2, unk, unk, 1, 0, unk, 4]
    second = [1, 2, 1, 2, 1, 0, 2]
    for j, (a, b) in enumerate(zip(best, second)):
        x[0, j, a] = 5.0; x[0, j, b] = 4.0
    # j=1: curr=2 vs prev=2 (raw) -> dropped; j=2: curr=1 vs prev=unk -> kept; j=3: curr=1 vs prev=unk -> kept (the
    # reference compares with the RAW previous symbol, so the repeated 1 is NOT collapsed); j=5: curr=0 blank dropped
    assert O.greedy_decode_unk(x, [7], unk) == [[2, 1, 1, 4]]
    assert O.greedy_decode_unk(x, [7], -1) == O.greedy_decode(x, [7])
