"""Times the PCG SpMV (q = A p fused with p.q) on C2 under the dot modes / register variants."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
fe, N, numbering = 2, int(os.environ.get("CELLS", 185)), 1
A = capi.Matrix.brick(fe, N, numbering)
n, nnz = A.rows, A.nnz
x, y = capi.Vector(n), capi.Vector(n)
x.fill_hash()
bytes_ = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
out = []
def run(label, fn, reps=30):
    for _ in range(5): fn()
    capi.set_option("timing", 1); capi.timing_reset()
    for _ in range(reps): fn()
    ms, cnt = capi.timing("spmv")
    vms, vcnt = capi.timing("vec")
    capi.set_option("timing", 0)
    t = (ms + vms) / reps
    rec = {"case": label, "ms": t, "gbs": bytes_ / t / 1e6, "spmv_ms": ms / reps, "vec_ms": vms / reps}
    out.append(rec); print(json.dumps(rec), flush=True)
for variant in (0, 2, 6, 7, 8, 1):
    capi.set_option("spmv.variant", variant)
    run(f"plain v{variant}", lambda: A.spmv_dev(x, y))
    for mode in (1, 2):
        capi.set_option("spmv.dotmode", mode)
        run(f"fused dot mode{mode} v{variant}", lambda: A.spmv_dot_dev(x, y, x))
capi.set_option("spmv.variant", 0)
run("unfused: spmv + dot", lambda: (A.spmv_dev(x, y), capi.dot(x, y)))
for thr, spc in ((128, 8), (128, 32), (256, 16), (64, 16)):
    capi.set_option("spmv.threads", thr); capi.set_option("spmv.slices_per_cta", spc)
    capi.set_option("spmv.dotmode", 2)
    run(f"mode2 v0 threads{thr} spc{spc}", lambda: A.spmv_dot_dev(x, y, x))
    run(f"plain v0 threads{thr} spc{spc}", lambda: A.spmv_dev(x, y))
json.dump(out, open("gpurun_out/probe_dot.json", "w"), indent=1)
