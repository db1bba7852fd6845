"""Workload for ncu captures inside a PCG solve on C2 (row-ordered SpMV with fused dot, update, direction)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
for k, v in (a.split("=") for a in sys.argv[1:]):
    capi.set_option(k, float(v))
capi.set_option("pcg.graph", 0)
N = int(os.environ.get("CELLS", 185))
A = capi.Matrix.brick(2, N, 1)
xs, b, x = capi.Vector(A.rows), capi.Vector(A.rows), capi.Vector(A.rows)
capi.brick_solution_device(2, N, 1, xs)
A.spmv_dev(xs, b)
x.fill(0.0)
print(A.pcg_dev(x, b, 6, 0.0, True))
capi.synchronize()
