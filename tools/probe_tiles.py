"""PCG on C2: cell size of the graph blocks (line segments as units) x slices per CTA, interleaved rounds."""
import json, sys, os, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
N = int(os.environ.get("CELLS", 185))
ROUNDS = int(os.environ.get("ROUNDS", 3))
for kv in os.environ.get("OPTS", "spmv.pk_prefetch=1,spmv.pk_tail=1").split(","):
    k, v = kv.split("="); capi.set_option(k, float(v))
tiles = [int(t) for t in os.environ.get("TILES", "8,4,5,6").split(",")]
spcs = [int(t) for t in os.environ.get("SPCS", "192,128,256").split(",")]
n, _ = capi.brick_sizes(2, N)
xs, b, x = capi.Vector(n), capi.Vector(n), capi.Vector(n)
capi.brick_solution_device(2, N, 1, xs)
res = {}
for rnd in range(ROUNDS):
    for t in tiles:
        capi.set_option("sell.block_tile", t)
        A = capi.Matrix.brick(2, N, 1)
        A.spmv_dev(xs, b)
        for spc in spcs:
            capi.set_option("spmv.pk_dot_slices_per_cta", spc)
            x.fill(0.0); A.pcg_dev(x, b, 10, 0.0, True)
            x.fill(0.0)
            capi.set_option("timing", 1); capi.timing_reset()
            capi.synchronize(); capi.timer_start()
            r = A.pcg_dev(x, b, 60, 0.0, True)
            ms = capi.timer_stop()
            sp = capi.timing("spmv")
            capi.set_option("timing", 0)
            res.setdefault((t, spc), []).append(sp[0] / max(sp[1], 1))
        A.close()
out = [{"tile": t, "spc": s, "spmv_ms_median": round(statistics.median(v), 4), "all": [round(q, 4) for q in v]} for (t, s), v in res.items()]
for r in out: print(json.dumps(r), flush=True)
json.dump(out, open(os.environ.get("OUT", "gpurun_out/r02_probe_tiles.json"), "w"), indent=1)
