"""H2D / D2H rate of 400 MB pinned buffers: one stream vs chunks on several streams."""
import torch, time, json
n = 50243409
h = torch.empty(n, dtype=torch.float64).pin_memory()
h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
d2 = torch.empty(n, dtype=torch.float64, device="cuda")
out = []
def timed(label, fn, bytes_):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    rec = {"case": label, "ms": round(dt * 1e3, 2), "GBps": round(bytes_ / dt / 1e9, 1)}
    out.append(rec); print(json.dumps(rec), flush=True)
timed("h2d 1 stream", lambda: d.copy_(h, non_blocking=True), 8 * n)
timed("d2h 1 stream", lambda: h.copy_(d, non_blocking=True), 8 * n)
for k in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    def chunked(dst, src):
        c = (n + k - 1) // k
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * c:(i + 1) * c].copy_(src[i * c:(i + 1) * c], non_blocking=True)
    timed(f"h2d {k} streams", lambda: chunked(d, h), 8 * n)
    timed(f"d2h {k} streams", lambda: chunked(h, d), 8 * n)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
timed("h2d + d2h concurrently", both, 16 * n)
json.dump(out, open("gpurun_out/probe_pcie.json", "w"), indent=1)
