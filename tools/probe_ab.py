"""PCG on C2 (line order on): A/B of packed-product options, interleaved rounds, medians of per-launch SpMV time."""
import json, sys, os, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
N = int(os.environ.get("CELLS", 185))
ROUNDS = int(os.environ.get("ROUNDS", 4))
for kv in os.environ.get("BUILD_OPTS", "").split(","):
    if kv:
        k, v = kv.split("="); capi.set_option(k, float(v))
A = capi.Matrix.brick(2, N, 1)
n = A.rows
xs, b, x, y = capi.Vector(n), capi.Vector(n), capi.Vector(n), capi.Vector(n)
capi.brick_solution_device(2, N, 1, xs)
A.spmv_dev(xs, b)
print(A.layout(), flush=True)
base = {"spmv.segments": 1, "spmv.pk_lean": 2, "spmv.scv_unroll": 8, "spmv.scv_threads": 256, "spmv.scv_slices_per_cta": 32, "spmv.pk_dot_threads": 384, "spmv.pk_dot_slices_per_cta": 192, "spmv.pk_cdict_dot": 1}
def run(opts):
    d = dict(base); d.update(opts)
    for k, v in d.items(): capi.set_option(k, v)
    x.fill(0.0)
    A.pcg_dev(x, b, 10, 0.0, True)
    x.fill(0.0)
    capi.set_option("timing", 1); capi.timing_reset()
    capi.synchronize(); capi.timer_start()
    r = A.pcg_dev(x, b, 60, 0.0, True)
    ms = capi.timer_stop()
    sp = capi.timing("spmv")
    capi.set_option("timing", 0)
    return sp[0] / max(sp[1], 1), 60 / (ms / 1e3), r[2]
cases = json.loads(os.environ["CASES"]) if "CASES" in os.environ else [
    ["base", {}], ["prefetch", {"spmv.pk_prefetch": 1}],
    ["wide1 32/8 320x3", {"spmv.pk_wide": 1, "spmv.pk_wide_threads": 320, "spmv.pk_dot_slices_per_cta": 160}],
    ["wide1+pf", {"spmv.pk_wide": 1, "spmv.pk_wide_threads": 320, "spmv.pk_dot_slices_per_cta": 160, "spmv.pk_prefetch": 1}],
    ["wide2 32/16 256x3", {"spmv.pk_wide": 2, "spmv.pk_wide_threads": 256, "spmv.pk_dot_slices_per_cta": 128}],
    ["wide3 24/8 384x3", {"spmv.pk_wide": 3, "spmv.pk_wide_threads": 384}],
    ["wide3+pf", {"spmv.pk_wide": 3, "spmv.pk_wide_threads": 384, "spmv.pk_prefetch": 1}],
    ["wide4 32/8 256x4", {"spmv.pk_wide": 4, "spmv.pk_wide_threads": 256, "spmv.pk_dot_slices_per_cta": 128}],
    ["wide4+pf", {"spmv.pk_wide": 4, "spmv.pk_wide_threads": 256, "spmv.pk_dot_slices_per_cta": 128, "spmv.pk_prefetch": 1}],
    ["wide5 32/4 384x3", {"spmv.pk_wide": 5, "spmv.pk_wide_threads": 384}],
    ["wide5+pf", {"spmv.pk_wide": 5, "spmv.pk_wide_threads": 384, "spmv.pk_prefetch": 1}],
]
res = {c[0]: [] for c in cases}
its = {c[0]: [] for c in cases}
resid = set()
for rnd in range(ROUNDS):
    for label, o in cases:
        sp, ips, rr = run(o)
        res[label].append(sp); its[label].append(ips); resid.add(rr)
out = [{"case": l, "spmv_ms_median": round(statistics.median(v), 4), "spmv_ms_min": round(min(v), 4), "spmv_ms_all": [round(t, 4) for t in v],
        "it_per_s_median": round(statistics.median(its[l]), 1)} for l, v in res.items()]
for r in out: print(json.dumps(r), flush=True)
print("residuals", resid)
json.dump(out, open(os.environ.get("OUT", "gpurun_out/r02_probe_ab.json"), "w"), indent=1)
