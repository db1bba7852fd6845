"""PCG on C2: anisotropic cells of the graph-block order (sell.block_tile_a/b/c), per-kernel timing."""
import json, sys, os, time, itertools
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
    rec = {"case": label, "build_s": round(build_s, 3), "it_per_s": round(100 / (ms / 1e3), 1), "spmv_ms": round(sp[0] / max(sp[1], 1), 4), "resid": r[2]}
    out.append(rec); print(json.dumps(rec), flush=True)
    A.close(); xs.close(); b.close(); x.close()
def tiles(a, b, c):
    capi.set_option("sell.block_tile_a", a); capi.set_option("sell.block_tile_b", b); capi.set_option("sell.block_tile_c", c)
tiles(0, 0, 0)
measure("tile 8 8 8")
for t in ((32, 4, 4), (4, 32, 4), (4, 4, 32), (16, 8, 4), (4, 8, 16), (8, 16, 4), (16, 4, 8), (4, 16, 8), (8, 4, 16),
          (16, 16, 2), (2, 16, 16), (16, 2, 16), (64, 4, 4), (4, 64, 4), (4, 4, 64), (16, 6, 6), (6, 16, 6), (6, 6, 16)):
    tiles(*t)
    measure("tile %d %d %d" % t)
tiles(0, 0, 0)
json.dump(out, open("gpurun_out/probe_tile.json", "w"), indent=1)
