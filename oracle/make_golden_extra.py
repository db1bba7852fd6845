"""BiCGStab / GCR fixtures from the COMPILED REFERENCE (oracle/_ref/libdrops_ref.so) on the systems
already stored under tests/golden/ -> tests/golden/krylov_extra.npz.

Run where /root/reference exists:  make -C oracle ref && python oracle/make_golden_extra.py
Keys: "<system>/<solver>" = [converged, iterations, residual], "<system>/<solver>_x" = solution.
Also the reference's own known-answer systems (src/tests/bicgstab.cpp, src/tests/gcr.cpp).
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = [("bicgstab_jac", lambda R, x0, b: R.bicgstab(x0, b, 400, 1e-10, True, 1, 1.0)),
         ("bicgstab_dummy", lambda R, x0, b: R.bicgstab(x0, b, 400, 1e-10, True, 0)),
         ("bicgstab_jac_abs_w08", lambda R, x0, b: R.bicgstab(x0, b, 400, 1e-9, False, 1, 0.8)),
         ("gcr_jac_m100", lambda R, x0, b: R.gcr(x0, b, 100, 400, 1e-10, True, 1, 1.0)),
         ("gcr_dummy_m100", lambda R, x0, b: R.gcr(x0, b, 100, 400, 1e-10, True, 0)),
         ("gcr_jac_m4_trunc", lambda R, x0, b: R.gcr(x0, b, 4, 60, 1e-10, True, 1, 1.0))]


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(GOLDEN, "brick_*.npz")) + glob.glob(os.path.join(GOLDEN, "random_spd*.npz"))):
        g = dict(np.load(path))
        name = os.path.basename(path)[:-4]
        n = int(g["rows"])
        R = ref.RefMatrix(n, n, g["rowbeg"], g["colind"], g["val"])
        for key, fn in CASES:
            c, it, res, x = fn(R, np.zeros(n), g["b"])
            out[f"{name}/{key}"] = np.array([c, it, res], float)
            out[f"{name}/{key}_x"] = x
    A2 = ref.RefMatrix.from_triplets(2, 2, [0, 1], [1, 0], [1., 1.])
    M = np.array([[-249., -453, -397, -52], [-453, -731, -601, -78], [-397, -601, -648, -91], [-52, -78, -91, -13]])
    ii, jj = np.nonzero(M)
    A4 = ref.RefMatrix.from_triplets(4, 4, ii, jj, M[ii, jj])
    b2, b4 = np.array([1., 2.]), np.array([-1277. / 2, -2015. / 2, -3907. / 4, -533. / 4])
    for key, r in [("kat/bicgstab2", A2.bicgstab(np.zeros(2), b2, 10, 1e-10, True, 0)),
                   ("kat/bicgstab4", A4.bicgstab(np.zeros(4), b4, 10, 1e-10, True, 0)),
                   ("kat/gcr2", A2.gcr(np.full(2, 3.0), b2, 10, 100, 1e-10, False, 0)),
                   ("kat/gcr4", A4.gcr(np.zeros(4), b4, 10, 100, 1e-10, False, 0))]:
        out[key] = np.array(r[:3], float)
        out[key + "_x"] = r[3]
    np.savez_compressed(os.path.join(GOLDEN, "krylov_extra.npz"), **out)
    for k in sorted(out):
        if not k.endswith("_x"):
            print(k, out[k])


if __name__ == "__main__":
    main()
