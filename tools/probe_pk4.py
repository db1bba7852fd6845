"""PCG on C2 with block-ordered rows: tile size x CTA geometry (bench-like loop, per-kernel timing)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
N = int(os.environ.get("CELLS", 185))
out = []
def measure(label):
    t0 = time.perf_counter()
    A = capi.Matrix.brick(2, N, 1)
    capi.synchronize()
    build_s = time.perf_counter() - t0
    n = A.rows
    xs, b, x = capi.Vector(n), capi.Vector(n), capi.Vector(n)
    capi.brick_solution_device(2, N, 1, xs)
    A.spmv_dev(xs, b)
    def solve():
        x.fill(0.0)
        return A.pcg_dev(x, b, 100, 0.0, True)
    solve()
    capi.set_option("timing", 1); capi.timing_reset()
    capi.synchronize(); capi.timer_start()
    r = solve()
    ms = capi.timer_stop()
    sp = capi.timing("spmv")
    capi.set_option("timing", 0)
    rec = {"case": label, "build_s": round(build_s, 3), "it_per_s": round(100 / (ms / 1e3), 1), "spmv_ms": round(sp[0] / max(sp[1], 1), 4),
           "layout": A.layout(), "resid": r[2]}
    out.append(rec); print(json.dumps(rec), flush=True)
    A.close()
capi.set_option("sell.block_order", 0)
measure("caller order")
capi.set_option("sell.block_order", 1)
for tile in (4, 6, 8, 12, 16, 24):
    capi.set_option("sell.block_tile", tile)
    for variant, thr, spc in ((0, 128, 16), (9, 512, 64), (9, 512, 128), (8, 1024, 256), (5, 256, 32)):
        if tile not in (8, 12) and variant not in (0, 9): continue
        capi.set_option("spmv.pkvariant", variant); capi.set_option("spmv.pkdotvariant", variant)
        capi.set_option("spmv.pk_threads", thr); capi.set_option("spmv.pk_slices_per_cta", spc)
        measure(f"tile{tile} v{variant} t{thr} spc{spc}")
json.dump(out, open("gpurun_out/probe_pk4.json", "w"), indent=1)
