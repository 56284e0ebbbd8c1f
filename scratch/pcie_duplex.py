# One-box measurement: PCIe copy rates with pinned host memory -- H2D alone, D2H alone, both at once (two streams), for the batch sizes the
# filo_scan_series pipeline uses.  Gives the ceiling the e2e number is read against (profiles/r2/r2_e2e_stages.md).
import json, sys, time, torch
dev = torch.device("cuda:0")
out = {}
for mb in (32, 128, 512, 2048):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = max(4, (8 << 30) // n)
    def run(h2d, d2h):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); return time.perf_counter() - t0
    run(True, True)
    a, b, c = run(True, False), run(False, True), run(True, True)
    out["%d MiB" % mb] = {"h2d_GBps": round(reps * n / a / 1e9, 1), "d2h_GBps": round(reps * n / b / 1e9, 1),
                          "duplex_total_GBps": round(2 * reps * n / c / 1e9, 1)}
    del h_in, h_out, d_in, d_out
print(json.dumps(out))
