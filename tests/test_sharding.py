"""CPU tests of the batching/sharding layer: reference split semantics
(scripts/swbd/utils.py:25-33), the cost-balanced assignment, and the N>1 exchange step -- the
scalar loss-sum all-reduce -- on two gloo ranks.  The per-rank compute in the two-rank test is
the oracle (test infrastructure); on GPUs it is libctcb."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gluon_e2e_asr_b200 import sharding as S
from oracle import ctc_oracle as O
from tests.synth import make_batch


def test_split_slices_follow_the_reference():
    for n in (1, 3, 4, 7, 8, 32, 33, 1024):
        for k in (1, 2, 3, 4, 8):
            got = [(s.start, s.stop) for s in S.split_slices(n, k)]
            ref = [(s.start, s.stop) for s in O.split_and_load_slices(n, k)]
            assert got == ref
            # covers the batch exactly once, in order
            flat = [i for a, b in got for i in range(a, b)]
            assert flat == list(range(n))
    assert S.split_slices(3, 4) == [slice(0, 3)]               # n < k: everything on ctx[0]
    with pytest.raises(ValueError):
        S.split_slices(4, 0)


def test_shard_for_rank():
    assert S.shard_for_rank(10, 3, 4) == slice(6, 10)          # remainder on the last rank
    assert S.shard_for_rank(3, 0, 4) == slice(0, 3)
    assert S.shard_for_rank(3, 2, 4) == slice(0, 0)


def test_split_and_load_cpu_devices():
    x = torch.arange(10 * 3).reshape(10, 3)
    parts = S.split_and_load(x, [torch.device("cpu")] * 4)
    assert [p.shape[0] for p in parts] == [2, 2, 2, 4]
    assert torch.equal(torch.cat(parts), x)


def test_balanced_assignment_is_a_partition_and_balances():
    rng = np.random.default_rng(0)
    T = np.sort(rng.integers(100, 2000, 64))[::-1]             # length-sorted batch (bucketed sampler)
    L = np.maximum(1, T // 8)
    parts = S.balanced_assignment(T, L, 8)
    assert sorted(i for p in parts for i in p) == list(range(64))
    cost = T.astype(float) * (2 * L + 1)
    loads = np.array([cost[p].sum() for p in parts])
    contiguous = np.array([cost[s].sum() for s in S.split_slices(64, 8)])
    assert loads.max() / loads.mean() < 1.05
    assert loads.max() < contiguous.max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = make_batch(7, 40, 9, 6, seed=21)
    sl = S.shard_for_rank(7, rank, world)
    loss, _, _ = O.CtcLossOracle("NTC", "NT")(d["pred"][sl], d["label"][sl], d["pred_lengths"][sl],
                                               d["label_lengths"][sl])
    vals = torch.tensor([float(loss.sum()), float(d["pred_lengths"][sl].sum()), float(sl.stop - sl.start)],
                        dtype=torch.float64)
    S.loss_sum_allreduce(vals)
    q.put((rank, vals.tolist(), (sl.start, sl.stop)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_loss_sum_allreduce_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = make_batch(7, 40, 9, 6, seed=21)
    loss, _, _ = O.CtcLossOracle("NTC", "NT")(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"])
    want = [float(loss.sum()), float(d["pred_lengths"].sum()), 7.0]
    assert [o[2] for o in out] == [(0, 3), (3, 7)]
    for _, vals, _ in out:
        np.testing.assert_allclose(vals, want, rtol=1e-12)


def test_loss_sum_allreduce_is_a_noop_without_a_group():
    v = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert S.loss_sum_allreduce(v) is None
    assert v.tolist() == [1.0, 2.0]


def _peer_rank_main(rank, world, port, q, steps, lag):
    """One process per rank, all on cuda:0 (CUDA IPC maps a mailbox into another process whether or not
    it lives on another GPU): the exchange kernel's protocol -- gather exchange k-1, publish exchange k,
    two slots, rank-ordered sums -- against plain host arithmetic, with the ranks deliberately skewed."""
    import time
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    ps = S.PeerLossSum(dev, lag=lag)
    got = []
    vals = torch.zeros(3, dtype=torch.float64, device=dev)
    out = torch.zeros(3, dtype=torch.float64, device=dev)
    for k in range(steps):
        if (k + rank) % 3 == 0:
            time.sleep(0.01)                                  # skew: the ranks never meet on purpose
        vals.copy_(torch.tensor([0.1 * (k + 1) * (rank + 1), float(100 * k + rank), 1.0], dtype=torch.float64))
        ps.exchange(vals, out)
        torch.cuda.synchronize()
        assert vals.abs().sum().item() == 0.0                 # handed over and zeroed
        got.append(out.tolist())
    dist.barrier()
    ps.flush(out)
    torch.cuda.synchronize()
    got.append(out.tolist())
    # the exchange riding on a step: the loss sums of two alternating slots, handed over one step late, come back
    # as the all-rank sum of each step's losses (checked against the per-utterance losses gathered with gloo)
    from gluon_e2e_asr_b200 import ctc_loss_and_grad
    part = torch.zeros((2, 1), dtype=torch.float64, device=dev)
    red = torch.zeros((1,), dtype=torch.float64, device=dev)
    ride = []
    nstep = 5
    for k in range(nstep + lag + 1):
        d = make_batch(3, 30, 46, 5, seed=100 * rank + min(k, nstep - 1))
        t = {n: torch.tensor(v, device=dev) for n, v in d.items()}
        ps.exchange_with_next(part[(k + 1) % 2], red)          # the previous step's slot
        loss, _ = ctc_loss_and_grad(t["pred"], t["label"], t["pred_lengths"], t["label_lengths"],
                                    loss_sum=part[k % 2, 0] if k < nstep else None)
        torch.cuda.synchronize()
        mine = torch.tensor([loss.double().sum().item() if k < nstep else 0.0], dtype=torch.float64)
        dist.all_reduce(mine)
        ride.append((red.item(), mine.item()))
    got.append(ride)
    q.put((rank, got))
    dist.barrier()
    ps.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world,lag", [(2, 1), (3, 3)])
def test_peer_mailbox_loss_sum(world, lag):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    steps = 9
    procs = [ctx.Process(target=_peer_rank_main, args=(r, world, port, q, steps, lag)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sums = []
    for k in range(steps):
        rows = [[0.1 * (k + 1) * (r + 1), float(100 * k + r), 1.0] for r in range(world)]
        tot = [0.0, 0.0, 0.0]
        for r in range(world):                                # rank order, like the kernel
            tot = [a + b for a, b in zip(tot, rows[r])]
        sums.append(tot)
    # exchange k returns the sums of exchange k - lag (zeros before), the flush the sums of the last exchange
    want = [sums[k - lag] if k >= lag else [0.0, 0.0, 0.0] for k in range(steps)] + [sums[-1]]
    for r in range(world):
        ride = res[r].pop()
        assert res[r] == want, "rank %d" % r                  # bit-identical on every rank
        # exchange k returns the sum handed over by exchange k-lag, which carried step k-lag-1's losses
        assert len(ride) > lag + 1
        for k in range(lag + 1, len(ride)):
            np.testing.assert_allclose(ride[k][0], ride[k - lag - 1][1], rtol=1e-6)
    assert all(res[r] == res[0] for r in range(world))


def test_peer_loss_sum_needs_cuda():
    from gluon_e2e_asr_b200 import _lib
    import ctypes
    l = _lib.load()
    h = ctypes.c_void_p()
    assert l.ctcb_mailbox_create(0, 2, 2, 1, ctypes.byref(h)) == _lib.CTCB_INVALID_VALUE
    assert l.ctcb_mailbox_create(0, 0, 17, 1, ctypes.byref(h)) == _lib.CTCB_INVALID_VALUE
    assert l.ctcb_mailbox_create(0, 0, 2, 0, ctypes.byref(h)) == _lib.CTCB_INVALID_VALUE
    assert l.ctcb_mailbox_create(0, 0, 2, 9, ctypes.byref(h)) == _lib.CTCB_INVALID_VALUE
    assert l.ctcb_mailbox_exchange(None, None, 1, None, None) == _lib.CTCB_INVALID_VALUE
    assert l.ctcb_mailbox_destroy(None) == _lib.CTCB_OK
    with pytest.raises(RuntimeError):
        S.PeerLossSum("cpu")
    if not torch.cuda.is_available():
        assert l.ctcb_mailbox_create(0, 0, 1, 4, ctypes.byref(h)) == _lib.CTCB_UNSUPPORTED
