"""Packed SpMV on C2: software-pipelined variants and L2 prefetch distance."""
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
for pf in (0, 1, 2, 4, 8):
    capi.set_option("spmv.pk_prefetch", pf)
    for variant, thr, spc in ((0, 128, 16), (11, 128, 16), (12, 128, 16), (13, 128, 16), (9, 512, 128), (14, 512, 128), (1, 128, 16)):
        capi.set_option("spmv.pkvariant", variant); capi.set_option("spmv.pkdotvariant", variant)
        capi.set_option("spmv.pk_threads", thr); capi.set_option("spmv.pk_slices_per_cta", spc)
        run(f"pf{pf} plain v{variant} t{thr} spc{spc}", lambda: A.spmv_dev(x, y))
        run(f"pf{pf} dot   v{variant} t{thr} spc{spc}", lambda: A.spmv_dot_dev(x, y, x))
json.dump(out, open("gpurun_out/probe_pk2.json", "w"), indent=1)
