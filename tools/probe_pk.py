"""Times the SpMV of C2 on the dictionary-packed stream (spmv_pk.cu) against the 12-byte stream:
plain product and product fused with the PCG dot, per register/unroll variant and launch geometry."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
fe, N, numbering = 2, int(os.environ.get("CELLS", 185)), int(os.environ.get("NUMBERING", 1))
A = capi.Matrix.brick(fe, N, numbering)
n, nnz = A.rows, A.nnz
print(json.dumps({"rows": n, "nnz": nnz, "dict": A.value_dictionary(), "format": A.format()}), flush=True)
x, y = capi.Vector(n), capi.Vector(n)
x.fill_hash()
bytes12 = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
bytes4 = 4.0 * nnz + 4.0 * (n + 1) + 16.0 * n
out = []
def run(label, fn, nbytes, reps=30):
    for _ in range(5): fn()
    capi.set_option("timing", 1); capi.timing_reset()
    for _ in range(reps): fn()
    ms, cnt = capi.timing("spmv")
    vms, vcnt = capi.timing("vec")
    capi.set_option("timing", 0)
    t = (ms + vms) / reps
    rec = {"case": label, "ms": round(t, 4), "gbs_format": round(nbytes / t / 1e6, 1), "gbs_12B_equiv": round(bytes12 / t / 1e6, 1)}
    out.append(rec); print(json.dumps(rec), flush=True)
capi.set_option("spmv.pk", 0)
run("12B plain", lambda: A.spmv_dev(x, y), bytes12)
run("12B fused dot", lambda: A.spmv_dot_dev(x, y, x), bytes12)
capi.set_option("spmv.pk", 1)
for variant in range(11):
    capi.set_option("spmv.pkvariant", variant); capi.set_option("spmv.pkdotvariant", variant)
    run(f"pk plain v{variant}", lambda: A.spmv_dev(x, y), bytes4)
    run(f"pk fused dot v{variant}", lambda: A.spmv_dot_dev(x, y, x), bytes4)
for form in (0, 1):
    capi.set_option("spmv.pkdotform", form)
    for variant in (0, 1, 4, 6, 9):
        capi.set_option("spmv.pkdotvariant", variant)
        run(f"pk fused dot FORM{form} v{variant}", lambda: A.spmv_dot_dev(x, y, x), bytes4)
capi.set_option("spmv.pkdotform", 2)
for variant, thr, spc in ((8, 1024, 128), (8, 1024, 256), (8, 512, 64), (9, 512, 64), (9, 512, 128), (9, 256, 32), (10, 512, 64), (10, 256, 32)):
    capi.set_option("spmv.pkvariant", variant); capi.set_option("spmv.pkdotvariant", variant)
    capi.set_option("spmv.pk_threads", thr); capi.set_option("spmv.pk_slices_per_cta", spc)
    run(f"pk plain v{variant} threads{thr} spc{spc}", lambda: A.spmv_dev(x, y), bytes4)
    run(f"pk fused dot v{variant} threads{thr} spc{spc}", lambda: A.spmv_dot_dev(x, y, x), bytes4)
capi.set_option("spmv.pk_threads", 128); capi.set_option("spmv.pk_slices_per_cta", 16)
best = int(os.environ.get("PKBEST", 0))
capi.set_option("spmv.pkvariant", best); capi.set_option("spmv.pkdotvariant", best)
for thr, spc in ((128, 8), (128, 32), (128, 64), (256, 16), (256, 32), (64, 16), (64, 8)):
    capi.set_option("spmv.pk_threads", thr); capi.set_option("spmv.pk_slices_per_cta", spc)
    run(f"pk plain v{best} threads{thr} spc{spc}", lambda: A.spmv_dev(x, y), bytes4)
    run(f"pk fused dot v{best} threads{thr} spc{spc}", lambda: A.spmv_dot_dev(x, y, x), bytes4)
capi.set_option("spmv.pk_threads", 128); capi.set_option("spmv.pk_slices_per_cta", 16)
capi.set_option("spmv.pkdict", 1)
run(f"pk plain v0 dict via ldg", lambda: A.spmv_dev(x, y), bytes4)
capi.set_option("spmv.pkdict", 0)
for cps in (8, 16, 32):
    capi.set_option("spmv.ctas_per_sm", cps)
    run(f"pk plain v{best} persistent {cps} CTAs/SM", lambda: A.spmv_dev(x, y), bytes4)
capi.set_option("spmv.ctas_per_sm", 0)
run("unfused: pk spmv + dot kernel", lambda: (A.spmv_dev(x, y), capi.dot(x, y)), bytes4)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_pk.json", "w"), indent=1)
