"""x window of the packed row-ordered SpMV in shared memory (spmv.pk_xwin): variants x CTA geometry on C2, inside PCG."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
N = int(os.environ.get("CELLS", 185))
ITERS = 100
A = capi.Matrix.brick(2, N, 1)
n = A.rows
xs, b, x = capi.Vector(n), capi.Vector(n), capi.Vector(n)
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
    sp = capi.timing("spmv"); up = capi.timing("pcg.update"); di = capi.timing("pcg.direction")
    capi.set_option("timing", 0)
    rec = {"case": label, "it_per_s": round(ITERS / (ms / 1e3), 1), "spmv_ms": round(sp[0] / max(sp[1], 1), 4),
           "update_ms": round(up[0] / max(up[1], 1), 4), "direction_ms": round(di[0] / max(di[1], 1), 4), "resid": r[2]}
    out.append(rec); print(json.dumps(rec), flush=True)
capi.set_option("spmv.pk_cdict", 0); capi.set_option("spmv.pk_cdict_dot", 0)      # the window variants use the shared-memory dictionary
measure("no window", spmv__pk_xwin=0)
for xv, name in ((1, "U8/40r(spill)"), (5, "U6/40r"), (4, "U4/40r"), (2, "U8/64r x2")):
    for thr, spc in ((512, 128), (512, 64), (512, 256), (256, 128)):
        if xv != 5 and (thr, spc) not in ((512, 128),): continue
        measure(f"xwin {name} t{thr} spc{spc}", spmv__pk_xwin=xv, spmv__pk_threads=thr, spmv__pk_slices_per_cta=spc)
measure("xwin U8/1024 spc256", spmv__pk_xwin=3, spmv__pk_threads=1024, spmv__pk_slices_per_cta=256)
measure("xwin U8/1024 spc128", spmv__pk_xwin=3, spmv__pk_threads=1024, spmv__pk_slices_per_cta=128)
for halo in (8, 32):
    measure(f"xwin U6 halo{halo}", spmv__pk_xwin=5, spmv__pk_threads=512, spmv__pk_slices_per_cta=128, spmv__pk_xwin_halo=halo)
measure("no window again", spmv__pk_xwin=0, spmv__pk_xwin_halo=0)
assert len({r["resid"] for r in out}) == 1, "variants disagree"
json.dump(out, open("gpurun_out/probe_xw.json", "w"), indent=1)
