"""Host<->device copy rates of the box (pinned memory), alone and both directions at once."""
import torch, time, json
dev = torch.device("cuda", 0)
n = 424 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory(); h.fill_(1.0)
h2 = torch.empty(n // 3, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev); d2 = torch.ones(n // 3, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps
def up():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): up(); down()
def up_chunks():
    k = n // 96
    with torch.cuda.stream(s1):
        for i in range(96): d[i * k:(i + 1) * k].copy_(h[i * k:(i + 1) * k], non_blocking=True)
r = {}
t = timed(up); r["h2d_GBps"] = n * 4 / t / 1e9
t = timed(down); r["d2h_GBps"] = n // 3 * 4 / t / 1e9
t = timed(both); r["both_ms"] = t * 1e3; r["both_h2d_GBps"] = n * 4 / t / 1e9
t = timed(up_chunks); r["h2d_96chunks_GBps"] = n * 4 / t / 1e9
# cudaHostRegister'ed numpy memory (what the C-ABI does with the caller's arrays)
import numpy as np, ctypes
a = np.ones(n, dtype=np.float32)
rt = ctypes.CDLL("libcudart.so.12")
rc = rt.cudaHostRegister(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes), 0)
ta = torch.from_numpy(a)
def up_reg():
    with torch.cuda.stream(s1): rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes), 1, ctypes.c_void_p(s1.cuda_stream))
t = timed(up_reg); r["h2d_registered_numpy_GBps"] = n * 4 / t / 1e9; r["register_rc"] = rc
print(json.dumps(r))
