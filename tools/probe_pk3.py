"""Packed SpMV on C2 with the TMA-staged stream: ring geometry x slices per CTA."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
A = capi.Matrix.brick(2, int(os.environ.get("CELLS", 185)), 1)
n, nnz = A.rows, A.nnz
x, y = capi.Vector(n), capi.Vector(n)
x.fill_hash()
bytes4 = 4.0 * nnz + 4.0 * (n + 1) + 16.0 * n
out = []
def run(label, fn, reps=20):
    for _ in range(3): fn()
    capi.set_option("timing", 1); capi.timing_reset()
    for _ in range(reps): fn()
    ms, cnt = capi.timing("spmv")
    capi.set_option("timing", 0)
    t = ms / reps
    rec = {"case": label, "ms": round(t, 4), "gbs_format": round(bytes4 / t / 1e6, 1)}
    out.append(rec); print(json.dumps(rec), flush=True)
run("register-staged plain", lambda: A.spmv_dev(x, y))
run("register-staged dot", lambda: A.spmv_dot_dev(x, y, x))
capi.set_option("spmv.pktma", 1)
for geo in range(9):
    capi.set_option("spmv.pktma_geometry", geo)
    for spc in (8, 16, 32, 64):
        capi.set_option("spmv.pktma_slices_per_cta", spc)
        run(f"tma g{geo} spc{spc} plain", lambda: A.spmv_dev(x, y))
        run(f"tma g{geo} spc{spc} dot", lambda: A.spmv_dot_dev(x, y, x))
print("fault", capi.lib().drops_b200_spmv_tma_fault())
json.dump(out, open("gpurun_out/probe_pk3.json", "w"), indent=1)
