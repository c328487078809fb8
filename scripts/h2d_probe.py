"""H2D copy rate of a packed cfg2 batch (2.37 MB) from ordinary pinned memory against write-combined pinned memory
(cudaHostAllocWriteCombined), and of two half copies on two streams.  CUDA events; prints one JSON line."""
import ctypes, json, sys
import torch

rt = ctypes.CDLL("libcudart.so.12")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_370_000
dev = torch.device("cuda:0")
dst = torch.empty(N, dtype=torch.uint8, device=dev)
out = {"bytes": N}


def time_copy(src_ptr, reps=200, streams=1):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    part = N // streams
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i, s in enumerate(ss):
            s.wait_stream(torch.cuda.current_stream())
            rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + i * part), ctypes.c_void_p(src_ptr + i * part), ctypes.c_size_t(part), 1,
                               ctypes.c_void_p(s.cuda_stream))
        for s in ss:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return {"us": round(us, 2), "GBps": round(N / us / 1e3, 1)}


for name, flags in (("pinned_default", 0), ("pinned_write_combined", 4)):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 1, N)
    out[name] = time_copy(p.value)
    out[name + "_2streams"] = time_copy(p.value, streams=2)
    rt.cudaFreeHost(p)
print(json.dumps(out))
