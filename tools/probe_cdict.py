"""Dictionary of the packed SpMV from the kernel's parameter block (spmv.pk_cdict) instead of shared memory: C2, inside PCG
and as a plain product."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
N = int(os.environ.get("CELLS", 185))
ITERS = 100
A = capi.Matrix.brick(2, N, 1)
n = A.rows
xs, b, x, y = capi.Vector(n), capi.Vector(n), capi.Vector(n), capi.Vector(n)
capi.brick_solution_device(2, N, 1, xs)
A.spmv_dev(xs, b)
out = []
def measure(label, **opts):
    for k, v in opts.items(): capi.set_option(k.replace("__", "."), v)
    def solve():
        x.fill(0.0)
        return A.pcg_dev(x, b, ITERS, 0.0, True)
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
    rec = {"case": label, "it_per_s": round(ITERS / (ms / 1e3), 1), "spmv_dot_ms": round(sp[0] / max(sp[1], 1), 4),
           "plain_spmv_ms": round(plain, 4), "resid": r[2]}
    out.append(rec); print(json.dumps(rec), flush=True)
def geo(thr, spc):
    capi.set_option("spmv.pk_threads", thr); capi.set_option("spmv.pk_slices_per_cta", spc)
    capi.set_option("spmv.pk_dot_threads", thr); capi.set_option("spmv.pk_dot_slices_per_cta", spc)
both = lambda v: dict(spmv__pk_cdict=v, spmv__pk_cdict_dot=v)
geo(384, 192)
measure("U14 384x3 (default)", **both(8))
measure("U15 384x3", **both(17))
for cv, name in ((108, "U14 384x3"), (117, "U15 384x3"), (118, "U16 384x3"), (109, "U12 384x3")):
    measure(name + " masked tail", **both(cv))
geo(512, 128)
for cv, name in ((101, "U8 512x3"), (103, "U10 512x3"), (120, "U7 512x3"), (121, "U9 512x3")):
    measure(name + " masked tail", **both(cv))
geo(512, 256); measure("U16 512x2 masked tail spc256", **both(102))
geo(256, 64); measure("U20 256x3 masked tail", **both(110))
geo(384, 192)
measure("U14 384x3 (default)", **both(8))
measure("U14 384x3 masked tail", **both(108))
assert len({r["resid"] for r in out}) == 1, "variants disagree"
json.dump(out, open("gpurun_out/probe_cdict.json", "w"), indent=1)
