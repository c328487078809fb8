"""Loss evaluation (ctcb_forward, keep_for_backward = 0) per BASELINE shape with one walker per utterance (meet_fwd=0) and
with walkers that meet in the middle (meet_fwd=1).  Raw C-ABI calls (the Python plugin's ~35 us per call would hide the small
shapes), CUDA events over 100 back-to-back calls, two buffer sets; one JSON line."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gluon_e2e_asr_b200 import _lib
from gluon_e2e_asr_b200.ops import _Call, _alloc_ws, _stream_ptr
from tests.synth import CONFIGS, make_batch

dev = torch.device("cuda:0")
lib = _lib.load()
out = {}
for name in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    B, T, V, L = CONFIGS[name]
    res = {}
    for meet in (0, 1):
        with _lib.options(meet_fwd=meet):
            probs = []
            for i in range(2):
                d = make_batch(B, T, V, L, seed=i, full_lengths=(name == "cfg5"))
                t = tuple(torch.tensor(d[k], device=dev) for k in ("pred", "label", "pred_lengths", "label_lengths"))
                call = _Call(t[0], t[1], t[2], t[3], False, True, False)
                loss = torch.empty((B,), device=dev)
                ws = _alloc_ws(call, False)
                probs.append((call.problem(loss), ws, t, loss))
            st = _stream_ptr(dev)

            def run(i):
                p, ws = probs[i % 2][0], probs[i % 2][1]
                _lib.check(lib.ctcb_forward(ctypes.byref(p), 0, ws.data_ptr(), ws.numel(), st))
            for i in range(5): run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(100): run(i)
            e1.record(); torch.cuda.synchronize()
            res["meet_fwd=%d" % meet] = round(e0.elapsed_time(e1) / 100 * 1e3, 1)
            del probs
    out[name] = res
    torch.cuda.empty_cache()
print(json.dumps(out))
