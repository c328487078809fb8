"""Where the end-to-end step of the prefetching host entry goes: host time inside ctcb_pipe_submit, host time blocked in
ctcb_pipe_wait, and the loop's period, for cfg2 packed batches (the bench's e2e leg).  usage: pipe_probe.py [depth] [steps]"""
import ctypes, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluon_e2e_asr_b200 import _lib
from gluon_e2e_asr_b200.batch import PinnedBatch
from gluon_e2e_asr_b200.ops import _DT
from tests.synth import CONFIGS, make_batch

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
B, T, V, L = CONFIGS["cfg2"]
lib = _lib.load()
nset = 8
probs, keep = [], []
for i in range(nset):
    d = make_batch(B, T, V, L, seed=i)
    pk = PinnedBatch.from_arrays(d["pred"], d["label"], d["pred_lengths"], d["label_lengths"], packed=True)
    lb = torch.empty((B,), dtype=torch.float32).pin_memory()
    q = _lib.Problem()
    q.T, q.B, q.V, q.Lmax, q.blank, q.label_pad = T, B, V, L, 0, 0
    q.logits, q.logits_stride_t, q.logits_stride_b = pk.pred.data_ptr(), V, T * V
    q.labels, q.label_dtype, q.label_stride_b, q.label_stride_l = pk.label.data_ptr(), _DT[pk.label.dtype], L, 1
    q.data_lengths, q.data_lengths_dtype = pk.pred_lengths.data_ptr(), _DT[pk.pred_lengths.dtype]
    q.label_lengths, q.label_lengths_dtype = pk.label_lengths.data_ptr(), _DT[pk.label_lengths.dtype]
    q.logits_row_offsets = pk.row_offsets.data_ptr()
    q.loss = lb.data_ptr()
    probs.append(q); keep.append((pk, lb))
ph = ctypes.c_void_p()
_lib.check(lib.ctcb_pipe_create(0, depth, ctypes.byref(ph)))
tk = ctypes.c_int64(-1)
pc = time.perf_counter


def run(k, measure=False):
    t_sub = t_wait = 0.0
    pending = []
    for i in range(k):
        a = pc()
        _lib.check(lib.ctcb_pipe_submit(ph, ctypes.byref(probs[i % nset]), ctypes.byref(tk)))
        b = pc()
        pending.append(tk.value)
        if len(pending) >= depth:
            _lib.check(lib.ctcb_pipe_wait(ph, pending.pop(0), None))
        c = pc()
        t_sub += b - a; t_wait += c - b
    for t_ in pending:
        _lib.check(lib.ctcb_pipe_wait(ph, t_, None))
    return t_sub, t_wait


run(50)
t0 = pc()
ts, tw = run(steps)
tot = pc() - t0
print(json.dumps({"depth": depth, "steps": steps, "period_us": round(tot / steps * 1e6, 2), "submit_us": round(ts / steps * 1e6, 2),
                  "wait_us": round(tw / steps * 1e6, 2)}))
lib.ctcb_pipe_destroy(ph)
