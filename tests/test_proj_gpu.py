"""GPU parity of the output projection fused with the loss (model.py:394-398, :424 +
loss.py:121-139) against  oracle ∘ fp64 matmul.

The tensor cores take tf32 inputs (the low 13 mantissa bits of hidden and weight do not take part) and accumulate in
fp32.  Parity proper is therefore stated on tf32-representable inputs, where the products are exact and the fused path
must meet the loss/gradient bar of the logits path (rtol 1e-4 / atol 1e-5 against the fp64 oracle applied to the fp64
product); a second test documents what arbitrary fp32 inputs cost (the tf32 rounding of the product, as with any
TF32 GEMM)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ctc_ref
from tests.synth import make_batch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-5
TF32_REL = 4e-3                           # d hidden / d weight: tf32 library GEMMs of the (fp32) logit gradient, norm-wise


def _close_tf32(actual, desired, what):
    """max |actual - desired| <= TF32_REL * max |desired|: the norm-wise statement one can make about a tf32 GEMM whose
    factors (the logit gradient) are not tf32-representable -- 2^-11 relative per factor, summed over the contraction."""
    err, scale = np.abs(actual - desired).max(), np.abs(desired).max()
    assert err <= TF32_REL * scale, "%s: max error %.3e against scale %.3e" % (what, err, scale)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _tf32(x):
    """x with the 13 low mantissa bits cleared: representable in tf32."""
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _problem(B, T, K, V, L, seed, exact=True, bias=True, full_lengths=False):
    d = make_batch(B, T, V, L, seed=seed, full_lengths=full_lengths)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    h = rng.standard_normal((B, T, K)).astype(np.float32)
    w = (rng.standard_normal((V, K)) / np.sqrt(K)).astype(np.float32)
    bv = (rng.standard_normal((V,)) * 0.5).astype(np.float32) if bias else None
    if exact:
        h, w = _tf32(h), _tf32(w)
    return d, h, w, bv


def _oracle(d, h, w, bv, head):
    logits = h.astype(np.float64) @ w.astype(np.float64).T
    if bv is not None:
        logits = logits + bv.astype(np.float64)
    lo, G, feas = ctc_ref.ctc_ref(logits, d["label"], d["pred_lengths"], d["label_lengths"], blank=0, head_grad=head,
                                  layout="NTC", dtype=np.float64)
    assert feas.all()
    for b in range(G.shape[0]):                      # padded frames: exact zeros (the C restatement leaves them untouched)
        G[b, int(d["pred_lengths"][b]):] = 0.0
    G2 = G.reshape(-1, G.shape[2])
    dh = (G2 @ w.astype(np.float64)).reshape(h.shape)
    dw = G2.T @ h.astype(np.float64).reshape(-1, h.shape[2])
    db = G2.sum(0)
    return logits, lo, G, dh, dw, db


def _run(dev, d, h, w, bv, head, need_grad=True, fused_training=True):
    from gluon_e2e_asr_b200 import proj_ctc_loss
    th = torch.tensor(h, device=dev, requires_grad=need_grad)
    tw = torch.tensor(w, device=dev, requires_grad=need_grad)
    tb = torch.tensor(bv, device=dev, requires_grad=need_grad) if bv is not None else None
    lab = torch.tensor(d["label"], device=dev)
    pl = torch.tensor(d["pred_lengths"], device=dev)
    ll = torch.tensor(d["label_lengths"], device=dev)
    if not need_grad:
        with torch.no_grad():
            return proj_ctc_loss(th, tw, tb, lab, pl, ll).cpu().numpy()
    loss = proj_ctc_loss(th, tw, tb, lab, pl, ll, fused_training=fused_training)
    (loss * torch.tensor(head, device=dev, dtype=torch.float32)).sum().backward()
    return (loss.detach().cpu().numpy(), th.grad.cpu().numpy(), tw.grad.cpu().numpy(),
            tb.grad.cpu().numpy() if tb is not None else None)


SHAPES = [
    # B, T, K, V, L        -- what each one exercises
    (3, 150, 64, 300, 20),       # two frame tiles (second partial), two vocabulary tiles (second partial), two K blocks
    (2, 128, 32, 256, 10),       # exactly one tile each way, one K block
    (4, 300, 100, 520, 33),      # K not a multiple of 32 (zero-filled tail), three vocabulary tiles
    (5, 97, 256, 1000, 40),      # deep K loop: the ring wraps several times
]


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_fused_projection_matches_oracle_of_fp64_product(dev, shape, ctas):
    """ctas = 2: CTA pairs (tcgen05 cta_group::2, the default); 1: one CTA per 128-frame tile."""
    from gluon_e2e_asr_b200 import _lib
    with _lib.options(proj_ctas=ctas):
        _fused_projection_case(dev, shape)


def _fused_projection_case(dev, shape):
    B, T, K, V, L = shape
    d, h, w, bv = _problem(B, T, K, V, L, seed=B)
    head = np.linspace(0.5, 1.5, B)
    logits, lo, G, dh, dw, db = _oracle(d, h, w, bv, head)
    loss, gh, gw, gb = _run(dev, d, h, w, bv, head)
    np.testing.assert_allclose(loss, lo, rtol=RTOL, atol=ATOL, err_msg="loss")
    # the contractions of the backward are tf32 library GEMMs (like the forward product) of a gradient that meets the logits
    # path's bar; the gradient itself is not tf32-representable, hence the tf32-GEMM tolerance (2^-11 per factor)
    _close_tf32(gh, dh, "d hidden")
    _close_tf32(gw, dw, "d weight")
    np.testing.assert_allclose(gb, db, rtol=2e-3, atol=2e-4, err_msg="d bias")
    # forward only: the logits are never stored; the same losses
    loss_fw = _run(dev, d, h, w, bv, head, need_grad=False)
    np.testing.assert_array_equal(loss_fw, loss)


def test_default_training_path_is_the_library_product(dev):
    """With a gradient the plugin forms the logits with a library GEMM (measured faster than the fused epilogue's store)
    and runs the loss on them: same bar."""
    B, T, K, V, L = 3, 150, 64, 300, 20
    d, h, w, bv = _problem(B, T, K, V, L, seed=9)
    head = np.ones(B)
    _, lo, _, dh, dw, db = _oracle(d, h, w, bv, head)
    loss, gh, gw, gb = _run(dev, d, h, w, bv, head, fused_training=False)
    np.testing.assert_allclose(loss, lo, rtol=RTOL, atol=ATOL)
    _close_tf32(gh, dh, "d hidden")
    _close_tf32(gw, dw, "d weight")
    np.testing.assert_allclose(gb, db, rtol=2e-3, atol=2e-4)


def test_logits_and_dlogits_through_the_c_abi(dev):
    """ctcb_proj_loss_grad: the projection's output and d loss / d logits themselves, at the logits path's bar."""
    import ctypes
    from gluon_e2e_asr_b200 import _lib
    from gluon_e2e_asr_b200.ops import _Call, _alloc_ws, _stream_ptr
    from gluon_e2e_asr_b200.proj import _proj_struct
    B, T, K, V, L = 3, 200, 96, 600, 25
    d, h, w, bv = _problem(B, T, K, V, L, seed=7)
    head = np.array([1.0, 0.25, 2.0])
    logits64, lo, G, _, _, _ = _oracle(d, h, w, bv, head)
    th, tw, tb = (torch.tensor(x, device=dev) for x in (h, w, bv))
    logits = torch.full((B, T, V), float("nan"), device=dev)
    grad = torch.empty_like(logits)
    loss = torch.empty((B,), device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    call = _Call(logits, torch.tensor(d["label"], device=dev), torch.tensor(d["pred_lengths"], device=dev),
                 torch.tensor(d["label_lengths"], device=dev), False, True, False)
    ws = _alloc_ws(call, True)
    p = call.problem(loss, grad, torch.tensor(head, device=dev, dtype=torch.float32), status=status)
    pj = _proj_struct(th, tw, tb)
    _lib.check(_lib.load().ctcb_proj_loss_grad(ctypes.byref(pj), ctypes.byref(p), ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
    torch.cuda.synchronize()
    assert _lib.last_launch_count() == 4          # k_proj_emit, metadata, k_walk, k_grad
    np.testing.assert_allclose(loss.cpu().numpy(), lo, rtol=RTOL, atol=ATOL)
    Tb = d["pred_lengths"].astype(int)
    lg = logits.cpu().numpy()
    for b in range(B):
        np.testing.assert_allclose(lg[b, :Tb[b]], logits64[b, :Tb[b]], rtol=1e-5, atol=1e-5, err_msg="logits")
    np.testing.assert_allclose(grad.cpu().numpy(), G, rtol=RTOL, atol=ATOL, err_msg="d logits")
    assert (status.cpu().numpy() == 0).all()


def test_matches_the_unfused_path_on_the_same_logits(dev):
    """fr and E from the tensor-memory epilogue against k_emit's from the stored logits: same loss, same gradient
    (to rounding of the softmax normaliser's summation order)."""
    from gluon_e2e_asr_b200 import ctc_loss_and_grad, proj_ctc_loss
    B, T, K, V, L = 4, 260, 64, 512, 30
    d, h, w, bv = _problem(B, T, K, V, L, seed=3, bias=False)
    th, tw = torch.tensor(h, device=dev), torch.tensor(w, device=dev)
    lab, pl, ll = (torch.tensor(d[k], device=dev) for k in ("label", "pred_lengths", "label_lengths"))
    with torch.no_grad():
        fused = proj_ctc_loss(th, tw, None, lab, pl, ll)
    logits = (th.double() @ tw.double().t()).float()
    ref, _ = ctc_loss_and_grad(logits, lab, pl, ll)
    np.testing.assert_allclose(fused.cpu().numpy(), ref.cpu().numpy(), rtol=2e-6, atol=1e-5)


def test_arbitrary_fp32_inputs_cost_the_tf32_rounding(dev):
    """fp32 hidden / weight that are not tf32-representable: the product carries tf32's 2^-11 input rounding, the loss
    follows to ~1e-3 relative -- the documented price of the tensor-core product, not a parity claim."""
    B, T, K, V, L = 3, 150, 128, 300, 20
    d, h, w, bv = _problem(B, T, K, V, L, seed=5, exact=False)
    head = np.ones(B)
    _, lo, _, _, _, _ = _oracle(d, h, w, bv, head)
    loss = _run(dev, d, h, w, bv, head, need_grad=False)
    np.testing.assert_allclose(loss, lo, rtol=5e-3)


@pytest.mark.parametrize("blank_label", ["first", "last"])
def test_inferred_lengths_and_blank_last(dev, blank_label):
    """No length vectors (use_*_lengths=False in the reference's operator): every utterance has T frames and its labels end
    at the first padding value (0 for blank 'first', -1 for 'last'); blank 'last' puts the blank in the LAST vocabulary
    tile's partial half -- the epilogue infers L itself (it runs before the metadata kernel's result exists)."""
    from gluon_e2e_asr_b200 import proj_ctc_loss
    from oracle import ctc_oracle as O
    B, T, K, V, L = 4, 170, 64, 300, 18
    blank = 0 if blank_label == "first" else V - 1
    d = make_batch(B, T, V, L, seed=31, blank=blank)
    rng = np.random.Generator(np.random.PCG64(32))
    h = _tf32(rng.standard_normal((B, T, K)).astype(np.float32))
    w = _tf32((rng.standard_normal((V, K)) / np.sqrt(K)).astype(np.float32))
    logits = h.astype(np.float64) @ w.astype(np.float64).T
    lo, _, ok = O.CtcLossOracle("NTC", "NT", blank_label)(logits, d["label"], None, None)
    assert ok.all()
    with torch.no_grad():
        loss = proj_ctc_loss(torch.tensor(h, device=dev), torch.tensor(w, device=dev), None,
                             torch.tensor(d["label"], device=dev), None, None, blank_label=blank_label)
    np.testing.assert_allclose(loss.cpu().numpy(), lo, rtol=RTOL, atol=ATOL)


def test_long_label_rows_park_in_the_emission_table(dev):
    """Label rows whose parked columns do not fit shared memory (here 321 columns x 128 frames x 4 B = 164 KB): the
    columns are parked in the emission table itself even though no logits are stored."""
    B, T, K, V, L = 2, 700, 32, 600, 320
    d, h, w, bv = _problem(B, T, K, V, L, seed=41, full_lengths=True)
    head = np.ones(B)
    _, lo, _, _, _, _ = _oracle(d, h, w, bv, head)
    loss = _run(dev, d, h, w, bv, head, need_grad=False)
    np.testing.assert_allclose(loss, lo, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("K", [128, 96])
def test_bfloat16_operands(dev, K):
    """CTCB_PROJ_BF16: bfloat16 hidden / weight (a mixed-precision encoder), fp32 accumulation and logits.  The products of
    bfloat16 values are exact in fp32, so the loss meets the logits path's bar against the oracle of the fp64 product of
    the SAME bfloat16 values; K = 96 leaves half a K block to the hardware's zero fill."""
    from gluon_e2e_asr_b200 import proj_ctc_loss
    B, T, V, L = 3, 200, 600, 25
    d, h, w, bv = _problem(B, T, K, V, L, seed=21, exact=False)
    th, tw = torch.tensor(h, device=dev).bfloat16(), torch.tensor(w, device=dev).bfloat16()
    hb, wb = th.float().cpu().numpy(), tw.float().cpu().numpy()
    head = np.array([1.0, 0.5, 2.0])
    _, lo, G, dh, dw, db = _oracle(d, hb, wb, bv, head)
    lab, pl, ll = (torch.tensor(d[k], device=dev) for k in ("label", "pred_lengths", "label_lengths"))
    tb = torch.tensor(bv, device=dev)
    with torch.no_grad():
        loss = proj_ctc_loss(th, tw, tb, lab, pl, ll)
    np.testing.assert_allclose(loss.cpu().numpy(), lo, rtol=RTOL, atol=ATOL)
    th.requires_grad_(True); tw.requires_grad_(True); tb.requires_grad_(True)
    loss2 = proj_ctc_loss(th, tw, tb, lab, pl, ll)
    (loss2 * torch.tensor(head, device=dev, dtype=torch.float32)).sum().backward()
    np.testing.assert_allclose(loss2.detach().cpu().numpy(), lo, rtol=RTOL, atol=ATOL)
    assert th.grad.dtype == torch.bfloat16 and tw.grad.dtype == torch.bfloat16
    # gradients come back in the operands' dtype: bfloat16's 2^-9 on top of the tf32 contractions
    for got, want, what in ((th.grad, dh, "d hidden"), (tw.grad, dw, "d weight")):
        err, scale = np.abs(got.float().cpu().numpy() - want).max(), np.abs(want).max()
        assert err <= 1e-2 * scale, "%s: %.3e against %.3e" % (what, err, scale)
    np.testing.assert_allclose(tb.grad.cpu().numpy(), db, rtol=2e-3, atol=2e-4)


def test_cfg3_shape_and_unsupported_shapes(dev):
    """BASELINE configs[2]'s vocabulary and label row (V=2000, L<=150, T=500) with H=512, at a batch the oracle does in
    seconds; small vocabularies are refused (CTCB_UNSUPPORTED), CPU tensors raise."""
    from gluon_e2e_asr_b200 import _lib, proj_ctc_loss
    B, T, K, V, L = 6, 500, 512, 2000, 150
    d, h, w, bv = _problem(B, T, K, V, L, seed=11)
    head = np.full(B, 1.0 / B)
    _, lo, G, dh, dw, db = _oracle(d, h, w, bv, head)
    loss, gh, gw, gb = _run(dev, d, h, w, bv, head)
    np.testing.assert_allclose(loss, lo, rtol=RTOL, atol=ATOL)
    _close_tf32(gh, dh, "d hidden")
    _close_tf32(gw, dw, "d weight")
    d2, h2, w2, b2 = _problem(2, 64, 32, 46, 10, seed=1)
    with pytest.raises(_lib.CtcbError) as e:
        _run(dev, d2, h2, w2, b2, np.ones(2), need_grad=False)
    assert e.value.code == _lib.CTCB_UNSUPPORTED
    with pytest.raises(RuntimeError):
        proj_ctc_loss(torch.tensor(h2), torch.tensor(w2), None, torch.tensor(d2["label"]))


def test_random_shapes_against_the_logits_path(dev):
    """scripts/proj_stress.py: random (B, T, K, V, L), ragged / zero / infeasible lengths, both operand types, with and without
    a gradient (the walkers beside the projection kernel, resp. the logits store), against the logits path on the exact
    product.  A protocol error between the kernels would show up here as a timeout or a wrong loss."""
    from scripts.proj_stress import run
    assert run(60, seed=5) < 2e-4
