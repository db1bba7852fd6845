"""Workload for one ncu capture of the packed SpMV (plain and fused dot) on C2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drops_b200 import capi
capi.init(0)
A = capi.Matrix.brick(2, int(os.environ.get("CELLS", 185)), 1)
x, y = capi.Vector(A.rows), capi.Vector(A.rows)
x.fill_hash()
for k, v in (a.split("=") for a in sys.argv[1:]):
    capi.set_option(k, float(v))
for _ in range(3):
    A.spmv_dev(x, y)
for _ in range(3):
    A.spmv_dot_dev(x, y, x)
capi.synchronize()
