"""Times the register-staged and the TMA-staged SpMV on C2, with and without 16-bit columns."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
fe, N, numbering = 2, int(os.environ.get("CELLS", 185)), 1
out = []
for compress in (0, 1):
    capi.set_option("sell.compress", compress)
    A = capi.Matrix.brick(fe, N, numbering)
    n, nnz = A.rows, A.nnz
    fmt = A.format()
    x, y = capi.Vector(n), capi.Vector(n)
    x.fill_hash()
    bytes_ = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
    def run(label, fn, reps=20):
        for _ in range(3): fn()
        capi.set_option("timing", 1); capi.timing_reset()
        for _ in range(reps): fn()
        ms, cnt = capi.timing("spmv")
        capi.set_option("timing", 0)
        rec = {"case": label, "compress": compress, "ms": ms / reps, "gbs_algorithmic": bytes_ / (ms / reps) / 1e6,
               "wide_slices": fmt["wide_slices"], "nslices": fmt["nslices"], "fault": capi.spmv_tma_fault()}
        out.append(rec); print(json.dumps(rec), flush=True)
    capi.set_option("spmv.kernel", 0)
    run("ldg plain", lambda: A.spmv_dev(x, y))
    run("ldg dot", lambda: A.spmv_dot_dev(x, y, x))
    capi.set_option("spmv.kernel", 2)
    for geo in (0, 1, 2, 3, 4):
        capi.set_option("spmv.tma_geometry", geo)
        for spc in (16, 32, 64):
            capi.set_option("spmv.tma_slices_per_cta", spc)
            run(f"tma g{geo} spc{spc} plain", lambda: A.spmv_dev(x, y))
        capi.set_option("spmv.tma_slices_per_cta", 16)
        run(f"tma g{geo} spc16 dot", lambda: A.spmv_dot_dev(x, y, x))
    capi.set_option("spmv.kernel", 0)
    A.close()
json.dump(out, open("gpurun_out/probe_tma.json", "w"), indent=1)
