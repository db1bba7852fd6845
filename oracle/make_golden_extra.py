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


def lincomb_fixture():
    """SparseMatBaseCL::LinComb (spmat.h:1457-1551) by the compiled reference: the time-step matrix
    L = (1/dt) E + theta H of levelset/coupling.tpp:95-97 in miniature -- a "mass" matrix whose pattern is a
    strict subset of the stiffness pattern plus a few entries of its own, and the stiffness matrix."""
    g = dict(np.load(os.path.join(GOLDEN, "brick_fe2_N4_num1_nonsym.npz")))
    n = int(g["rows"])
    rng = np.random.default_rng(7)
    import scipy.sparse as sp
    H = sp.csr_matrix((g["val"], g["colind"].astype(np.int64), g["rowbeg"].astype(np.int64)), shape=(n, n))
    keep = rng.random(H.nnz) < 0.6
    keep[g["colind"].astype(np.int64) == np.repeat(np.arange(n), np.diff(g["rowbeg"].astype(np.int64)))] = True
    E = sp.csr_matrix((np.abs(H.data[keep]) * 0.01 + 0.1, H.indices[keep], np.r_[0, np.cumsum(np.add.reduceat(
        keep.astype(np.int64), H.indptr[:-1]))]), shape=(n, n))
    extra = sp.random(n, n, density=0.01, random_state=11, format="csr")
    extra.data[:] = 0.37
    E = (E + extra).tocsr()
    E.sort_indices()
    E.data[E.data == 0.0] = 0.25                              # keep every stored position nonzero
    A = ref.RefMatrix(n, n, E.indptr, E.indices, E.data)
    B = ref.RefMatrix(n, n, g["rowbeg"], g["colind"], g["val"])
    out = {"rows": n, "A_rowbeg": E.indptr.astype(np.uint64), "A_colind": E.indices.astype(np.uint64), "A_val": E.data,
           "B_rowbeg": g["rowbeg"], "B_colind": g["colind"], "B_val": g["val"], "x_in": g["x_in"], "b": g["b"]}
    for tag, (a, b) in {"dt": (1.0 / 3e-4, 1.0), "neg": (0.7, -1.3), "zero": (0.0, 2.0)}.items():
        Cm = ref.RefMatrix.lincomb(a, A, b, B)
        rb, ci, va = Cm.csr()
        out[f"{tag}_coef"] = np.array([a, b])
        out[f"{tag}_rowbeg"], out[f"{tag}_colind"], out[f"{tag}_val"] = rb, ci, va
        out[f"{tag}_y"] = Cm.spmv(g["x_in"])
    c, it, res, x = ref.RefMatrix.lincomb(1.0 / 3e-4, A, 1.0, B).gmres_jacobi(np.zeros(n), g["b"], 50, 200, 1e-10, True, 1.0, True)
    out["dt_gmres"] = np.array([c, it, res]); out["dt_gmres_x"] = x
    np.savez_compressed(os.path.join(GOLDEN, "lincomb.npz"), **out)
    print("lincomb: nnz A", E.nnz, "B", len(g["val"]), "C", len(out["dt_val"]), "gmres", out["dt_gmres"])


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
    lincomb_fixture()
    for k in sorted(out):
        if not k.endswith("_x"):
            print(k, out[k])


if __name__ == "__main__":
    main()
