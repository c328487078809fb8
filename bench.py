#!/usr/bin/env python
"""bench.py -- CTC fwd+bwd valid frames/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

A *step* is one pass of the hot path (log-softmax, alpha/beta recursion, gradient to the
logits) over one synthetic batch.  N=1 runs BASELINE.json configs[1] (cfg2: B=32, T=500,
V=46, L<=120, variable lengths).  N>1 (torchrun) gives every rank its own cfg2-shaped shard of
utterances -- the path has no data-path collective; the per-step float64 loss-sum all-reduce
(NCCL) runs on a side stream -- and reports the aggregate ("scaling": "weak").

`value`  : device-resident inputs, the step's kernels replayed from CUDA graphs, CUDA events
           around exactly K steps, max over ranks.
`e2e`    : same metric through the drop-in boundary with HOST buffers: the C ABI's prefetching host
           entry (ctcb_pipe_submit / ctcb_pipe_wait) with pinned host batches -- every step's H2D of
           logits/labels/lengths, its kernels and the D2H of its loss vector are inside the timed loop
           and every loss is read on the host; batch i+1's copy overlaps batch i's kernels (2 in
           flight); the gradient stays on the device, where the model's backward consumes it.
`e2e_sync`: the same through the synchronous host entry (ctcb_loss_grad_host_resident): copy, kernels,
           loss back and a synchronisation inside every call, nothing overlapped.
`e2e_plugin`: the same step through the Python mirror of the reference's block,
           CtcLoss(...)(pred, ...).mean().backward() (torch autograd on the path), pinned host
           inputs in one arena (PinnedBatch), loss read back.
`roofline`: dominant kernel (k_walk) against the measured HBM copy peak.
`cpu_baseline`: the oracle's C restatement of the reference's CPU operator, timed on this
           box's host cores on a bounded sample (N=1, rank 0 only).
`--impl reference`: the same C restatement as the measured arm (the reference's MXNet
           operator is not installable here -- DESIGN.md section 3), all host threads.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

from tests.synth import CONFIGS, make_batch

METRIC = "ctc_fwd_bwd_valid_frames_per_sec"
UNIT = "frames/s"
L2_BYTES = 126 * 1024 * 1024


def algorithmic_bytes(V, T, B, Tb, Lb):
    """read valid logits once + write the dense gradient once + labels/lengths/loss."""
    return 4 * V * int(Tb.sum()) + 4 * V * B * T + 4 * int(Lb.sum()) + 12 * B


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """NVML samples of SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.max_mhz = period, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
            for bit, n in names.items():
                if r & bit:
                    self.reasons.add(n)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample()
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cuda_local_index():
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    return None if vis else 0


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's C restatement on the host cores
# ------------------------------------------------------------------------------------------
def cpu_step_fn(d):
    from oracle import ctc_ref
    B = d["pred"].shape[0]
    head = np.full((B,), 1.0 / B, np.float32)
    g = np.empty_like(d["pred"])
    lab = d["label"].astype(np.int32)

    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask explicitly)
    cores = min(len(os.sched_getaffinity(0)), B) if hasattr(os, "sched_getaffinity") else ctc_ref.max_threads()

    def step():
        # NTC logits addressed through strides (no swapaxes copy: a favour to the baseline),
        # softmax + alpha + beta + grad + head scaling, OpenMP over the minibatch
        return ctc_ref.ctc_ref(d["pred"], lab, d["pred_lengths"], d["label_lengths"], blank=0, head_grad=head,
                               layout="NTC", dtype=np.float32, out_grad=g, num_threads=cores)
    return step, cores


def time_cpu(d, steps, warmup, budget_s=None):
    step, cores = cpu_step_fn(d)
    for _ in range(warmup):
        step()
    ts = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s and len(ts) >= 3:
            break
    return ts, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = args.workload
    B, T, V, L = CONFIGS[name]
    d = make_batch(B, T, V, L, seed=0)
    frames = float(d["pred_lengths"].sum())
    ts, cores = time_cpu(d, args.steps, max(args.warmup, 1))
    ms = 1e3 * sum(ts) / len(ts)
    val = frames / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(ts), "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "B": B, "T": T, "V": V, "Lmax": L, "valid_frames_per_step": frames,
                   "note": "C restatement of the reference's CPU CTC operator (oracle/ctc_ref.c, fp32, OpenMP over the "
                           "minibatch); MXNet itself is not installable in this image"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d full %s steps (B=%d)" % (len(ts), name, B)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------
def run_cuda(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from gluon_e2e_asr_b200 import CtcLoss, _lib
    from gluon_e2e_asr_b200 import ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device and no CPU fallback for the measured arm")
    _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    B, T, V, L = CONFIGS[name]
    blk = CtcLoss(layout="NTC", label_layout="NT")
    head = torch.full((B,), 1.0 / B, device=dev)

    # input sets: rotate over enough distinct (logits, grad) buffers to exceed L2
    per_set = 2 * 4 * B * T * V
    nset = max(2, min(64, (2 * L2_BYTES + per_set - 1) // per_set + 1))
    sets = []
    for i in range(nset):
        d = make_batch(B, T, V, L, seed=1000 * rank + i)
        sets.append({
            "np": d,
            "pred": torch.tensor(d["pred"], device=dev), "label": torch.tensor(d["label"], device=dev),
            "pl": torch.tensor(d["pred_lengths"], device=dev), "ll": torch.tensor(d["label_lengths"], device=dev),
            "loss": torch.empty((B,), device=dev), "grad": torch.empty((B, T, V), device=dev),
            "frames": float(d["pred_lengths"].sum()),
        })
    # Per-step loss sum (float64, accumulated by the walkers) and its all-reduce.  Two slots: step i
    # accumulates into slot i % 2 while the side branch of the same graph reduces slot (i-1) % 2 --
    # the previous step's sum -- over NCCL, so the collective never sits on the kernels' critical path
    # and costs no host call per step (it is part of the captured graph).
    part = torch.zeros((2, 3), dtype=torch.float64, device=dev)         # per slot: partial {loss sum, frames, utterances}
    loss_sums = part[:, 0]                                               # the walkers accumulate straight into the partials
    red_buf = torch.zeros((2, 3), dtype=torch.float64, device=dev)      # {loss sum, frames, utterances} per slot
    if nset % 2:
        nset -= 1
        sets = sets[:nset]
    graph_allreduce = world > 1 and not args.no_graph_allreduce and not args.no_allreduce
    force_x = world == 1 and os.environ.get("CTCB_BENCH_FORCE_EXCHANGE") == "1"    # experiment: the exchange's own cost on one GPU
    graph_allreduce = graph_allreduce or force_x

    def step_eager(s, slot=0):
        ops.ctc_loss_and_grad(s["pred"], s["label"], s["pl"], s["ll"], head_grad=head, loss_sum=loss_sums[slot],
                              out_loss=s["loss"], out_grad=s["grad"], handoff="pointer")

    # The exchange: by default the library's peer mailbox (ctcb_mailbox_*: one tiny kernel that stores the partial sums
    # into every rank's mailbox over NVLink peer access and picks up the previous exchange -- no collective kernel and
    # no rendezvous on the step's path); --collective nccl keeps torch.distributed's all-reduce.
    peer = None
    collective = "none"
    if force_x:
        from gluon_e2e_asr_b200 import PeerLossSum
        peer = PeerLossSum(dev, lag=args.peer_lag)
        collective = "peer"
    if world > 1 and args.no_allreduce and os.environ.get("CTCB_BENCH_CONNECT_ONLY") == "1":
        from gluon_e2e_asr_b200 import PeerLossSum
        _connected_only = PeerLossSum(dev, lag=args.peer_lag)      # experiment: peer mailboxes mapped, never used
    if world > 1 and not args.no_allreduce:
        collective = "nccl"
        if args.collective == "peer":
            try:
                from gluon_e2e_asr_b200 import PeerLossSum
                peer = PeerLossSum(dev, lag=args.peer_lag)
                collective = "peer"
            except Exception as exc:  # noqa: BLE001
                if rank == 0:
                    print("bench.py: peer mailbox unavailable (%s); using the NCCL all-reduce" % str(exc)[:160], file=sys.stderr)

    def reduce_slot(slot):
        if peer is not None:
            # ONE kernel: red_buf <- all-rank sum of the previous exchange; this slot's partials stored to every rank, zeroed
            peer.exchange(part[slot], red_buf[slot])
        else:
            red_buf[slot].copy_(part[slot], non_blocking=True)
            part[slot].zero_()
            dist.all_reduce(red_buf[slot])

    stream = torch.cuda.Stream(dev)
    comm_stream = torch.cuda.Stream(dev)
    graphs = []
    with torch.cuda.stream(stream):
        for i, s in enumerate(sets[:2]):
            step_eager(s, i % 2)
        launches_per_step = _lib.last_launch_count()
        if world > 1 and not args.no_allreduce:
            reduce_slot(0); reduce_slot(1)              # warm-up of the exchange outside any capture (same count on every rank)
        torch.cuda.synchronize()
        walk_cfg = _lib.last_walk_config()
        for i, s in enumerate(sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                if graph_allreduce and peer is not None:
                    # the previous step's slot: a one-warp kernel behind the step's gradient kernel, as its programmatic dependent
                    peer.exchange_with_next(part[(i + 1) % 2], red_buf[(i + 1) % 2])
                elif graph_allreduce:
                    comm_stream.wait_stream(stream)
                    with torch.cuda.stream(comm_stream):
                        reduce_slot((i + 1) % 2)        # the previous step's slot
                step_eager(s, i % 2)
                if graph_allreduce and peer is None:
                    stream.wait_stream(comm_stream)
            graphs.append(g)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(k, graph=True, allreduce=True):
        frames = 0.0
        with torch.cuda.stream(stream):
            for i in range(k):
                s = sets[i % nset]
                if graph:
                    graphs[i % nset].replay()
                else:
                    step_eager(s, i % 2)
                frames += s["frames"]
                if world > 1 and allreduce and not args.no_allreduce and not (graph and graph_allreduce):
                    # scalar loss-sum all-reduce on a side stream: never blocks the next step
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    comm_stream.wait_event(ev)
                    with torch.cuda.stream(comm_stream):
                        reduce_slot(i % 2)
            stream.wait_stream(comm_stream)
        return frames

    # ---- value: K steps, device-resident inputs ------------------------------------------
    barrier()
    run_steps(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    frames = run_steps(args.steps)
    e1.record(stream)
    sampler.sample()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total, frames], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, frames_all = tmax[0].item(), tsum[1].item()
    else:
        frames_all = frames
    value = frames_all / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps

    # eager (no CUDA graph) timing of the same steps, for the record
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_steps(min(args.warmup, 5), graph=False, allreduce=False)
    ee0.record(stream)
    n_eager = min(args.steps, 200)
    run_steps(n_eager, graph=False, allreduce=False)
    ee1.record(stream)
    torch.cuda.synchronize()
    eager_ms = ee0.elapsed_time(ee1) / n_eager

    # ---- e2e: public API, pinned host inputs, H2D + D2H inside the timed region ----------
    # Each host batch is collated in ONE pinned arena (gluon_e2e_asr_b200.batch.PinnedBatch: what the
    # reference's batchify + split_and_load do with four arrays), so the step's H2D is a single copy.
    from gluon_e2e_asr_b200.batch import PinnedBatch
    hsets = []
    for s in sets[:min(nset, 8)]:
        d = s["np"]
        hsets.append(PinnedBatch.from_arrays(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"]))
    h2d = hsets[0].h2d_bytes
    loss_host = torch.empty((B,), dtype=torch.float32).pin_memory()
    d2h = int(loss_host.numel() * 4)
    for h in hsets:
        h.load(dev)["pred"].requires_grad_(True)

    def e2e_step(h):
        x = h.load(dev)                                 # one cudaMemcpyAsync: logits, labels, both length vectors
        pred = x["pred"]
        pred.grad = None
        loss = blk(pred, x["label"], x["pred_lengths"], x["label_lengths"])
        loss.mean().backward()
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the step's result is on the host
        return pred.grad

    import gc
    gc.collect()
    gc.freeze()        # the bench holds thousands of long-lived objects (46 input sets, graphs): keep the collector off them
    e2e_steps = max(10, min(args.steps, 200))
    for i in range(max(3, min(args.warmup, 10))):
        e2e_step(hsets[i % len(hsets)])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_frames = 0.0
    for i in range(e2e_steps):
        e2e_step(hsets[i % len(hsets)])
        e2e_frames += sets[i % len(hsets)]["frames"]
    f1.record()
    torch.cuda.synchronize()
    e2e_ms = f0.elapsed_time(f1)
    te = torch.tensor([e2e_ms, e2e_frames], dtype=torch.float64, device=dev)
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_ms, e2e_frames = a[0].item(), b[1].item()
    e2e_value = e2e_frames / (e2e_ms * 1e-3)

    # ---- e2e through the C ABI's host entry (no torch on the path): same pinned host buffers ----
    import ctypes
    e2e_cabi = None
    try:
        lib = _lib.load()
        probs = []
        for h in hsets:
            q = _lib.Problem()
            q.T, q.B, q.V, q.Lmax, q.blank, q.label_pad = T, B, V, L, 0, 0
            q.logits, q.logits_stride_t, q.logits_stride_b = h.pred.data_ptr(), V, T * V
            q.labels, q.label_dtype, q.label_stride_b, q.label_stride_l = h.label.data_ptr(), _lib.DT_F32, L, 1
            q.data_lengths, q.data_lengths_dtype = h.pred_lengths.data_ptr(), _lib.DT_F32
            q.label_lengths, q.label_lengths_dtype = h.label_lengths.data_ptr(), _lib.DT_F32
            q.loss = loss_host.data_ptr()
            probs.append(q)
        dgrad = ctypes.c_void_p()
        for i in range(5):
            _lib.check(lib.ctcb_loss_grad_host_resident(ctypes.byref(probs[i % len(probs)]), local_rank, ctypes.byref(dgrad)))
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            _lib.check(lib.ctcb_loss_grad_host_resident(ctypes.byref(probs[i % len(probs)]), local_rank, ctypes.byref(dgrad)))
        cabi_ms = (time.perf_counter() - t0) * 1e3
        tc = torch.tensor([cabi_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        cabi_ms = tc[0].item()
        cabi_h2d = sum(int(getattr(hsets[0], k).numel() * getattr(hsets[0], k).element_size()) for k in PinnedBatch.FIELDS)
        e2e_cabi = {"value": e2e_frames / (cabi_ms * 1e-3), "unit": UNIT, "ms_per_step": cabi_ms / e2e_steps,
                    "h2d_bytes_per_step": cabi_h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "ctcb_loss_grad_host_resident(problem with pinned HOST pointers): H2D of logits/labels/lengths, "
                           "kernels, loss back to the host, gradient left on the device; host wall clock around the "
                           "synchronous calls"}
    except Exception as exc:  # noqa: BLE001
        e2e_cabi = {"error": str(exc)[:200]}

    # ---- e2e through the C ABI's prefetching host entry (ctcb_pipe_*): the same pinned host batches, batch i+1's
    # H2D copy in flight while batch i's kernels run; every step's loss is read on the host before the step counts ----
    e2e_pipe = None
    try:
        depth = args.pipe_depth
        ph = ctypes.c_void_p()
        _lib.check(lib.ctcb_pipe_create(local_rank, depth, ctypes.byref(ph)))
        loss_bufs = [torch.zeros((B,), dtype=torch.float32).pin_memory() for _ in hsets]
        loss_np = [b.numpy() for b in loss_bufs]
        pprobs = []
        for q0, lb in zip(probs, loss_bufs):
            q = _lib.Problem()
            ctypes.memmove(ctypes.byref(q), ctypes.byref(q0), ctypes.sizeof(q))
            q.loss = lb.data_ptr()
            pprobs.append(q)
        npb = len(pprobs)
        tk = ctypes.c_int64(-1)

        moved = {}                                      # bytes each host batch moves host -> device (asked from the library)
        hb, pulled = ctypes.c_int64(0), ctypes.c_int32(0)

        def pipe_run(k, record=False):
            """k batches through the pipe, `depth` in flight; returns the sum of all losses read on the host."""
            acc, pending = 0.0, []
            for i in range(k):
                _lib.check(lib.ctcb_pipe_submit(ph, ctypes.byref(pprobs[i % npb]), ctypes.byref(tk)))
                if record:
                    _lib.check(lib.ctcb_pipe_last_h2d_bytes(ph, ctypes.byref(hb), ctypes.byref(pulled)))
                    moved[i % npb] = (hb.value, pulled.value)
                pending.append((tk.value, i % npb))
                if len(pending) >= depth:
                    t_, j = pending.pop(0)
                    _lib.check(lib.ctcb_pipe_wait(ph, t_, None))
                    acc += float(loss_np[j].sum())
            for t_, j in pending:
                _lib.check(lib.ctcb_pipe_wait(ph, t_, None))
                acc += float(loss_np[j].sum())
            return acc

        pipe_run(max(2 * depth, npb), record=True)
        # same bits as the synchronous host entry
        _lib.check(lib.ctcb_loss_grad_host_resident(ctypes.byref(probs[0]), local_rank, ctypes.byref(dgrad)))
        ref_loss = loss_host.clone()
        pipe_run(1)
        if not torch.equal(ref_loss, loss_bufs[0]):
            raise RuntimeError("pipelined host entry disagrees with the synchronous one")
        pipe_steps = max(e2e_steps, min(args.steps, 1000))
        barrier()
        t0 = time.perf_counter()
        acc = pipe_run(pipe_steps)
        pipe_ms = (time.perf_counter() - t0) * 1e3
        pipe_frames = sum(sets[i % npb]["frames"] for i in range(pipe_steps))
        tp_ = torch.tensor([pipe_ms, pipe_frames], dtype=torch.float64, device=dev)
        if world > 1:
            a = tp_.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = tp_.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            pipe_ms, pipe_frames = a[0].item(), b[1].item()
        e2e_pipe = {"value": pipe_frames / (pipe_ms * 1e-3), "unit": UNIT, "ms_per_step": pipe_ms / pipe_steps,
                    "h2d_bytes_per_step": sum(moved[i % npb][0] for i in range(pipe_steps)) / pipe_steps,
                    "d2h_bytes_per_step": d2h, "steps": pipe_steps,
                    "in_flight": depth, "loss_checksum": acc,
                    "h2d": ("the GPU pulls the VALID frames of the pinned logits itself (k_pull_valid, zero-copy loads over PCIe: "
                            "padded frames never cross the bus); labels and lengths in one copy"
                            if moved and all(v[1] for v in moved.values()) else "one cudaMemcpyAsync of the batch's pinned arena"),
                    "h2d_bytes_per_step_dense": hsets[0].h2d_bytes,
                    "api": "ctcb_pipe_submit / ctcb_pipe_wait (pinned HOST batches): every step's "
                           "inputs cross PCIe and its loss is read on the host inside the timed region; batch i+1's "
                           "transfer overlaps batch i's kernels (%d batches in flight); gradient left on the device; "
                           "host wall clock around the loop including the drain" % depth}
        lib.ctcb_pipe_destroy(ph)
    except Exception as exc:  # noqa: BLE001
        e2e_pipe = {"error": str(exc)[:200]}

    # ---- roofline of the dominant kernel: per-kernel CUDA events inside the library -------
    kms = np.zeros((8,), np.float64)
    nrep = 20
    kbuf = (ctypes.c_float * 8)()
    nk = ctypes.c_int32(0)
    alg = 0.0
    with torch.cuda.stream(stream):
        for i in range(nrep + 3):
            s = sets[i % nset]
            call = ops._Call(s["pred"], s["label"], s["pl"], s["ll"], False, True, False)
            ws = ops._ws_cache[(dev.index, stream.cuda_stream, call.T, call.B, call.V, call.Lmax)]
            p = call.problem(s["loss"], s["grad"], head)
            _lib.check(_lib.load().ctcb_loss_grad_timed(ctypes.byref(p), ws.data_ptr(), ws.numel(), stream.cuda_stream,
                                                        kbuf, ctypes.byref(nk)))
            if i >= 3:
                kms += np.array(list(kbuf))
                alg += algorithmic_bytes(V, T, B, s["np"]["pred_lengths"], s["np"]["label_lengths"])
    kms = kms[:nk.value] / nrep
    alg /= nrep
    # small dense vocabularies run the fused walker (no k_emit launch): two kernels per step
    knames = ["k_emit", "k_walk", "k_grad"] if nk.value == 3 else ["k_walk", "k_grad"]
    dom = int(np.argmax(kms))
    peak, peak_src = measured_peak()
    achieved = alg / (kms[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(name, {}).get(knames[dom])
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": knames[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg,
        "kernel_ms": {k: float(v) for k, v in zip(knames, kms)},
        "step_achieved": alg / (ms_per_step * 1e-3) / 1e9, "step_frac": alg / (ms_per_step * 1e-3) / 1e9 / peak,
        "note": "fp32 logits/gradient in HBM, fp64 linear-domain lattice recursion; the V=46 path is recursion-latency "
                "bound (T dependent steps per utterance, 2B CTAs), not HBM bound: SURVEY.md 8d / DESIGN.md sections 5-6; "
                "kernel_ms are the kernels timed one after the other, in the step k_grad runs concurrently with k_walk",
    }

    # ---- other workloads, same run (N=1 only): context numbers, not the headline ----------
    others = []
    if world == 1 and not args.no_others:
        for oname in ("cfg1", "cfg3", "cfg4", "cfg5"):
            if oname == name:
                continue
            try:
                others.append(measure_other(torch, ops, dev, oname, peak))
            except Exception as exc:  # noqa: BLE001
                others.append({"workload": oname, "error": str(exc)[:200]})

    cpu_baseline = None
    if rank == 0 and world == 1:
        ts, cores = time_cpu(sets[0]["np"], 1000, 2, budget_s=12.0)
        cms = 1e3 * sum(ts) / len(ts)
        cpu_baseline = {"value": sets[0]["frames"] / (cms * 1e-3), "unit": UNIT, "cores": cores, "kind": "port",
                        "ms_per_step": cms,
                        "sample": "%d full %s steps (B=%d) of oracle/ctc_ref.c fp32, OpenMP over the minibatch" % (len(ts), name, B)}

    e2e_plugin = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                  "api": "PinnedBatch.load(dev) -> CtcLoss(layout='NTC',label_layout='NT')(pred,label,pred_lengths,"
                         "label_lengths).mean().backward() -> loss to pinned host (torch autograd on the path)"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name if world == 1 else "%s per rank (global B=%d)" % (name, B * world),
                       "B_per_gpu": B, "T": T, "V": V, "Lmax": L, "layout": "NTC", "lengths": "variable",
                       "valid_frames_per_step": frames_all / args.steps,
                       "utterances_per_sec": B * world / (ms_per_step * 1e-3),
                       "l2": "inputs rotate over %d buffer sets (%.0f MB logits+grad > 126 MB L2)" % (nset, nset * per_set / 1e6),
                       "launch": "%d kernels per step (k_grad a programmatic dependent of k_walk) replayed from a CUDA graph; "
                                 "eager_ms_per_step=%.4f" % (launches_per_step, eager_ms),
                       "walker": {"pairs_per_lane": walk_cfg[0], "warps": walk_cfg[1]},
                       "collective": ("none" if world == 1 or collective == "none" else
                                      "none on the data path; float64 loss-sum exchange of the previous step's sum inside each step's CUDA graph: " +
                                      ("ctcb_mailbox_exchange_with_next: a one-warp kernel stores the partial sums into every rank's mailbox over "
                                       "NVLink peer memory and picks up the sums of %d steps before (no collective kernel, no rendezvous); it is the " % args.peer_lag +
                                       "programmatic dependent of the step's gradient kernel and runs beside that kernel's last wave"
                                       if collective == "peer" else "NCCL all-reduce on a side branch") if graph_allreduce else
                                      "none on the data path; float64 loss-sum all-reduce (%s) per step on a side stream" % collective)},
            "clocks": clocks,
            # headline end-to-end number: the C ABI's host entry (the drop-in boundary itself, HOST buffers in,
            # loss back on the host); the same step through the Python plugin + torch autograd is reported beside it
            "e2e": (e2e_pipe if e2e_pipe and "value" in e2e_pipe else
                    e2e_cabi if e2e_cabi and "value" in e2e_cabi else e2e_plugin),
            "e2e_sync": e2e_cabi,
            "e2e_plugin": e2e_plugin,
            "gpu_launches": (launches_per_step + (1 if (peer is not None and graph_allreduce) else 0)) * args.steps,
            "roofline": roofline,
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if others:
            line["other_workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured graphs hold NCCL work: tearing the communicator down under them can block, and a
        # rank that lingers would stall the launcher.  Everything is measured and printed; leave together.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def measure_other(torch, ops, dev, oname, peak, steps=20, warmup=3):
    B, T, V, L = CONFIGS[oname]
    per_set = 2 * 4 * B * T * V
    nset = max(2, min(8, (2 * L2_BYTES + per_set - 1) // per_set + 1))
    sets = []
    for i in range(nset):
        d = make_batch(B, T, V, L, seed=50 + i, full_lengths=(oname == "cfg5"))
        sets.append((torch.tensor(d["pred"], device=dev), torch.tensor(d["label"], device=dev),
                     torch.tensor(d["pred_lengths"], device=dev), torch.tensor(d["label_lengths"], device=dev),
                     torch.empty((B,), device=dev), torch.empty((B, T, V), device=dev),
                     float(d["pred_lengths"].sum()), algorithmic_bytes(V, T, B, d["pred_lengths"], d["label_lengths"])))
    head = torch.full((B,), 1.0 / B, device=dev)

    def step(s):
        ops.ctc_loss_and_grad(s[0], s[1], s[2], s[3], head_grad=head, out_loss=s[4], out_grad=s[5], handoff="pointer")
    for i in range(warmup):
        step(sets[i % nset])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    frames = alg = 0.0
    for i in range(steps):
        s = sets[i % nset]
        step(s)
        frames += s[6]; alg += s[7]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ops._ws_cache.clear()
    del sets
    torch.cuda.empty_cache()
    return {"workload": oname, "B": B, "T": T, "V": V, "Lmax": L, "ms_per_step": ms / steps,
            "value": frames / (ms * 1e-3), "unit": UNIT, "step_achieved_gbs": alg / (ms * 1e-3) / 1e9,
            "step_frac": alg / (ms * 1e-3) / 1e9 / peak, "launch": "eager"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--pipe-depth", type=int, default=2, help="batches in flight in the prefetching host entry (e2e)")
    ap.add_argument("--no-others", action="store_true", help="skip the context measurements of the other configs")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: loss-sum exchange through the library's peer mailbox (default) or torch.distributed's NCCL all-reduce")
    ap.add_argument("--peer-lag", type=int, default=4, help="slack between the ranks of the peer mailbox exchange, in steps")
    ap.add_argument("--no-allreduce", action="store_true", help="experiment: no loss-sum collective at all (N>1)")
    ap.add_argument("--no-graph-allreduce", action="store_true",
                    help="N>1: issue the loss-sum all-reduce from the host every step instead of from the captured graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 50:
            args.steps = 50          # each step is a full CPU pass (tens of ms); keep the run to minutes
        run_reference(args, rank, world)
        return
    run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
