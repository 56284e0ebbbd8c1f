# H2D rate by kind of host memory: cudaHostAlloc (torch pin_memory) vs cudaHostRegister over numpy / mmap memory (4 KB pages, THP-advised),
# alone and with a concurrent D2H into cudaHostAlloc memory.
import ctypes, json, mmap, time, numpy as np, torch
cudart = ctypes.CDLL("libcudart.so.12") if False else None
n = 1 << 30
dev = torch.device("cuda:0")
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.ones(n, dtype=torch.uint8, device=dev)
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
rt = torch.cuda.cudart()
def bench(h_t, label, out):
    def run(h2d, d2h, reps=8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_t, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t0) / 1e9
    run(True, True, 2)
    out[label] = {"h2d_alone": round(run(True, False), 1), "h2d+d2h_each": round(run(True, True), 1)}
out = {"thp": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()}
bench(torch.empty(n, dtype=torch.uint8).pin_memory().fill_(1), "cudaHostAlloc", out)
a = np.ones(n, np.uint8)
t = torch.from_numpy(a)
assert int(rt.cudaHostRegister(a.ctypes.data, n, 3)) == 0
bench(t, "numpy + cudaHostRegister", out)
rt.cudaHostUnregister(a.ctypes.data)
m = mmap.mmap(-1, n + (2 << 20))
try:
    m.madvise(mmap.MADV_HUGEPAGE)
    out["madvise"] = "ok"
except Exception as e:
    out["madvise"] = repr(e)
b = np.frombuffer(m, np.uint8)
off = (-b.ctypes.data) % (2 << 20)
b = b[off:off + n]; b[:] = 1
t2 = torch.from_numpy(b)
assert int(rt.cudaHostRegister(b.ctypes.data, n, 3)) == 0
bench(t2, "mmap + MADV_HUGEPAGE + cudaHostRegister", out)
try:
    out["AnonHugePages_kB"] = [l for l in open("/proc/self/smaps_rollup") if "AnonHuge" in l][0].split()[1]
except Exception as e:
    out["AnonHugePages_kB"] = repr(e)
print(json.dumps(out))
