"""PCG on C2: rows of a sigma window grouped by stencil type (sell.type_sort) or by length only."""
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
    xs, b, x, y = capi.Vector(n), capi.Vector(n), capi.Vector(n), capi.Vector(n)
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
    for _ in range(3): A.spmv_dev(xs, y)
    capi.synchronize(); capi.timer_start()
    for _ in range(20): A.spmv_dev(xs, y)
    plain = capi.timer_stop() / 20
    rec = {"case": label, "build_s": round(build_s, 3), "it_per_s": round(100 / (ms / 1e3), 1), "spmv_dot_ms": round(sp[0] / max(sp[1], 1), 4),
           "plain_spmv_ms": round(plain, 4), "resid": r[2]}
    out.append(rec); print(json.dumps(rec), flush=True)
    A.close(); xs.close(); b.close(); x.close(); y.close()
for rep in range(2):
    capi.set_option("sell.type_sort", 0); measure("length only")
    capi.set_option("sell.type_sort", 1); measure("length, stencil type")
for tile in (4, 6, 12, 16):
    capi.set_option("sell.block_tile", tile); measure(f"length, stencil type, tile {tile}")
capi.set_option("sell.block_tile", 8)
capi.set_option("sell.block_order", 0); measure("length, stencil type, caller order")
capi.set_option("sell.type_sort", 0); measure("length only, caller order")
capi.set_option("sell.block_order", 1)
json.dump(out, open("gpurun_out/probe_type.json", "w"), indent=1)
