"""Loss evaluation with walkers that meet in the middle (WalkArgs::meet, option meet_fwd): the alpha
walker takes the first half of an utterance's frame blocks, the beta walker the second half from the end, and
P(l|x) = sum_s alpha_m(s) beta'_m(s) is formed at the meeting frame.  Same bar as every loss: rtol 1e-4 / atol 1e-5 against
the fp64 oracle -- for both schedules, both walker variants (fused emission producers / emission table), every walker
configuration the BASELINE shapes use, ragged and degenerate utterances."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ctc_ref
from tests.synth import make_batch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _oracle_loss(d, blank=0):
    lo, _, feas = ctc_ref.ctc_ref(d["pred"].astype(np.float64), d["label"], d["pred_lengths"], d["label_lengths"], blank=blank,
                                  layout="NTC", dtype=np.float64, need_grad=False)
    return lo, feas


def _gpu_loss(dev, d, meet, blank_label="first", loss_sum=False):
    from gluon_e2e_asr_b200 import _lib
    from gluon_e2e_asr_b200.ops import ctc_loss
    t = {k: torch.tensor(v, device=dev) for k, v in d.items()}
    with _lib.options(meet_fwd=meet), torch.no_grad():
        return ctc_loss(t["pred"].transpose(0, 1), t["label"], t["pred_lengths"], t["label_lengths"], True, True,
                        blank_label).cpu().numpy()


SHAPES = [
    # B, T, V, L, kwargs                      -- what each one exercises
    (8, 200, 46, 50, {}),                      # cfg1: fused producers, P=2 x 1 warp
    (6, 500, 46, 120, {}),                     # cfg2's lattice: two walker warps (halo hand-over at the meeting frame)
    (3, 700, 46, 300, {}),                     # cfg4's lattice: five walker warps
    (4, 123, 200, 30, {}),                     # wide vocabulary: emission table + TMA producer, partial last block
    (3, 260, 700, 150, {}),                    # cfg3's lattice (three warps), unfused
    (5, 40, 11, 6, {"peaky": True}),           # short utterances, trained-regime logits
    (4, 17, 9, 3, {}),                         # NQ = 3: alpha two blocks, beta one (partial)
    (4, 12, 9, 3, {}),                         # some utterances have a single block (no split: the alpha CTA does it all)
    (3, 64, 30, 0, {}),                        # empty label rows
    (4, 90, 20, 10, {"scale": 12.0}),          # wide logit range: renormalisations on both sides
]


@pytest.mark.parametrize("shape", SHAPES)
def test_meet_in_the_middle_matches_the_oracle(dev, shape):
    B, T, V, L, kw = shape
    d = make_batch(B, T, V, L, seed=B * 7 + T, **kw)
    lo, feas = _oracle_loss(d)
    got_meet = _gpu_loss(dev, d, 1)
    got_one = _gpu_loss(dev, d, 0)
    np.testing.assert_allclose(got_meet[feas], lo[feas], rtol=RTOL, atol=ATOL, err_msg="meet")
    np.testing.assert_allclose(got_one[feas], lo[feas], rtol=RTOL, atol=ATOL, err_msg="one walker")
    assert (got_meet[~feas] == 0).all()


def test_blank_last_and_full_lengths(dev):
    d = make_batch(5, 150, 46, 40, seed=3, blank=45, full_lengths=True)
    lo, feas = _oracle_loss(d, blank=45)
    np.testing.assert_allclose(_gpu_loss(dev, d, 1, "last"), lo, rtol=RTOL, atol=ATOL)


def test_repeatable_bits(dev):
    d = make_batch(16, 300, 46, 80, seed=9)
    a = _gpu_loss(dev, d, 1)
    for _ in range(3):
        np.testing.assert_array_equal(_gpu_loss(dev, d, 1), a)


def test_random_shapes_both_schedules_agree(dev):
    """80 random problems (both walker variants, one to eight walker warps, ragged / zero / infeasible lengths): the loss of
    the meeting walkers against the loss of the single alpha walker -- two different evaluations of the same lattice."""
    from gluon_e2e_asr_b200 import _lib
    from gluon_e2e_asr_b200.ops import ctc_loss
    rng = np.random.Generator(np.random.PCG64(77))
    for c in range(80):
        B = int(rng.integers(1, 24)); T = int(rng.integers(1, 400)); V = int(rng.integers(2, 300))
        L = int(rng.integers(0, min(T, 250) + 1))
        Tb = rng.integers(0, T + 1, B); Lb = rng.integers(0, L + 1, B)
        lab = rng.integers(1, V, (B, max(L, 1)))[:, :L].astype(np.float32) if L else np.zeros((B, 0), np.float32)
        x = torch.tensor((rng.standard_normal((T, B, V)) * rng.choice([1.0, 4.0])).astype(np.float32), device=dev)
        args = (x, torch.tensor(lab, device=dev), torch.tensor(Tb.astype(np.float32), device=dev),
                torch.tensor(Lb.astype(np.float32), device=dev), True, True)
        with torch.no_grad():
            with _lib.options(meet_fwd=1):
                a = ctc_loss(*args).cpu().numpy()
            with _lib.options(meet_fwd=0):
                b = ctc_loss(*args).cpu().numpy()
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-5, err_msg="case %d: B=%d T=%d V=%d L=%d" % (c, B, T, V, L))
