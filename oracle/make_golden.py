"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libdrops_ref.so).

Run where /root/reference exists:   make -C oracle ref && python oracle/make_golden.py
The fixtures pin oracle/drops_oracle.c (CPU tests) and the CUDA path (GPU tests) to outputs of
the unmodified DROPS code on small seeded systems; /root/reference is not needed to USE them.
Every fixture stores the full CSR input, so it does not depend on our generator.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref  # noqa: E402
from drops_b200 import capi  # noqa: E402  (host-side brick generator only; no GPU needed)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hash_vector(n):
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)).astype(np.float64) / 4294967296.0 - 0.5


def random_sparse(rng, rows, cols, per_row, spd=False):
    """Ragged rows (some empty unless spd), ascending unique columns."""
    rb, ci, va = [0], [], []
    for i in range(rows):
        k = int(rng.integers(0, per_row + 1))
        c = set(rng.choice(cols, size=min(k, cols), replace=False).tolist())
        if spd:
            c.add(i)
        c = sorted(c)
        ci += c
        va += rng.standard_normal(len(c)).tolist()
        rb.append(len(ci))
    rb, ci, va = np.array(rb, np.uint64), np.array(ci, np.uint64), np.array(va)
    if spd:
        import scipy.sparse as sp
        A = sp.csr_matrix((va, ci.astype(np.int64), rb.astype(np.int64)), shape=(rows, cols))
        A = (A + A.T).tocsr()
        A = A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
        A = A.tocsr()
        A.sort_indices()
        rb, ci, va = A.indptr.astype(np.uint64), A.indices.astype(np.uint64), A.data.copy()
    return rb, ci, va


def solve_all(R, b, tag, out, restart=20, max_iter=400, tol=1e-10, nonsym=False):
    n = R.rows
    x0 = np.zeros(n)
    if not nonsym:
        c, it, res, x, h = R.pcg_jacobi(x0, b, max_iter, tol, True, 1.0, hist=True)
        out[f"{tag}_pcgjac"] = np.array([c, it, res]); out[f"{tag}_pcgjac_x"] = x; out[f"{tag}_pcgjac_hist"] = h[:it + 1]
        c, it, res, x, h = R.pcg_jacobi(x0, b, max_iter, tol, False, 0.8, hist=True)
        out[f"{tag}_pcgjac_w08"] = np.array([c, it, res]); out[f"{tag}_pcgjac_w08_x"] = x
        c, it, res, x, h = R.pcg_dummy(x0, b, max_iter, tol, True, hist=True)
        out[f"{tag}_pcgdummy"] = np.array([c, it, res]); out[f"{tag}_pcgdummy_x"] = x
        c, it, res, x, h = R.cg(x0, b, max_iter, tol, True, hist=True)
        out[f"{tag}_cg"] = np.array([c, it, res]); out[f"{tag}_cg_x"] = x; out[f"{tag}_cg_hist"] = h[:it + 1]
        c, it, res, x = R.minres(x0, b, max_iter, 1e-9)
        out[f"{tag}_minres"] = np.array([c, it, res]); out[f"{tag}_minres_x"] = x
        c, it, res, x = R.pminres_jacobi(x0, b, max_iter, 1e-9)
        out[f"{tag}_pminresjac"] = np.array([c, it, res]); out[f"{tag}_pminresjac_x"] = x
        # fixed work: 7 iterations, never converges (tol 0)
        c, it, res, x = R.pcg_jacobi(x0, b, 7, 0.0, True, 1.0)
        out[f"{tag}_pcgjac_fixed7"] = np.array([c, it, res]); out[f"{tag}_pcgjac_fixed7_x"] = x
    for left in (1, 0):
        c, it, res, x = R.gmres_jacobi(x0, b, restart, max_iter, tol, True, 1.0, bool(left))
        out[f"{tag}_gmresjac_l{left}"] = np.array([c, it, res]); out[f"{tag}_gmresjac_l{left}_x"] = x
    c, it, res, x = R.gmres_jacobi(x0, b, 5, max_iter, tol, True, 1.0, True)      # many restarts
    out[f"{tag}_gmresjac_m5"] = np.array([c, it, res]); out[f"{tag}_gmresjac_m5_x"] = x
    c, it, res, x = R.gmres_dummy(x0, b, restart, max_iter, tol, False, True)
    out[f"{tag}_gmresdummy"] = np.array([c, it, res]); out[f"{tag}_gmresdummy_x"] = x


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)

    # --- brick systems ---------------------------------------------------------------------
    for fe, N, numbering in [(1, 6, 0), (2, 4, 0), (2, 4, 1), (2, 6, 1)]:
        n, rb, ci, va = capi.brick_csr_host(fe, N, numbering)
        xs = capi.brick_solution_host(fe, N, numbering)
        R = ref.RefMatrix(n, n, rb, ci, va)
        out = {"rows": n, "cols": n, "rowbeg": rb, "colind": ci, "val": va, "xstar": xs}
        xin = hash_vector(n)
        out["x_in"] = xin
        out["y_Ax"] = R.spmv(xin)
        out["y_ATx"] = R.spmv_transpose(xin)
        out["diag"] = R.diag()
        out["jac_w1"] = R.jacobi_apply(xin, 1.0)
        out["jac_w07"] = R.jacobi_apply(xin, 0.7)
        b = R.spmv(xs)
        out["b"] = b
        solve_all(R, b, "sym", out)
        np.savez_compressed(os.path.join(OUT, f"brick_fe{fe}_N{N}_num{numbering}.npz"), **out)

    # --- nonsymmetric brick (GMRES stand-in: upper triangle scaled) -----------------------
    n, rb, ci, va = capi.brick_csr_host(2, 4, 1, 1.1)
    xs = capi.brick_solution_host(2, 4, 1)
    R = ref.RefMatrix(n, n, rb, ci, va)
    out = {"rows": n, "cols": n, "rowbeg": rb, "colind": ci, "val": va, "xstar": xs}
    xin = hash_vector(n)
    out["x_in"] = xin
    out["y_Ax"] = R.spmv(xin)
    out["y_ATx"] = R.spmv_transpose(xin)
    b = R.spmv(xs)
    out["b"] = b
    solve_all(R, b, "nonsym", out, nonsym=True)
    np.savez_compressed(os.path.join(OUT, "brick_fe2_N4_num1_nonsym.npz"), **out)

    # --- ragged random matrices: empty rows, rectangular, long rows ------------------------
    for tag, rows, cols, per_row in [("rect_tall", 150, 40, 9), ("rect_wide", 40, 150, 30), ("ragged", 97, 97, 70)]:
        rb, ci, va = random_sparse(rng, rows, cols, per_row)
        R = ref.RefMatrix(rows, cols, rb, ci, va)
        x = rng.standard_normal(cols)
        xt = rng.standard_normal(rows)
        np.savez_compressed(os.path.join(OUT, f"random_{tag}.npz"), rows=rows, cols=cols, rowbeg=rb, colind=ci,
                            val=va, x_in=x, y_Ax=R.spmv(x), xt_in=xt, y_ATx=R.spmv_transpose(xt))

    # --- SPD random system with irregular rows (C4 stand-in) ------------------------------
    rb, ci, va = random_sparse(rng, 211, 211, 12, spd=True)
    R = ref.RefMatrix(211, 211, rb, ci, va)
    b = rng.standard_normal(211)
    out = {"rows": 211, "cols": 211, "rowbeg": rb, "colind": ci, "val": va, "b": b}
    xin = rng.standard_normal(211)
    out["x_in"] = xin
    out["y_Ax"] = R.spmv(xin)
    out["y_ATx"] = R.spmv_transpose(xin)
    out["diag"] = R.diag()
    solve_all(R, b, "sym", out)
    np.savez_compressed(os.path.join(OUT, "random_spd211.npz"), **out)

    # --- vector kernels -----------------------------------------------------------------------
    n = 10007
    x, y, y2 = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    x[::97] *= 1e8
    y[::89] *= 1e-8
    np.savez_compressed(os.path.join(OUT, "vectors.npz"), x=x, y=y, y2=y2,
                        dot=ref.dot(x, y), norm=ref.norm(x), norm_sq=ref.norm_sq(x), supnorm=ref.supnorm(x),
                        kahan_dot=ref.kahan_dot(x, y), kahan_norm=ref.kahan_norm(x),
                        axpy=ref.axpy(0.37, x, y), z_xpay=ref.z_xpay(x, -1.25, y),
                        z_xpaypby2=ref.z_xpaypby2(x, 0.5, y, -3.0, y2))

    # --- the reference's own known-answer tests (src/tests/minres.cpp, blockmat.cpp) -------
    A2 = ref.RefMatrix.from_triplets(2, 2, [0, 1], [1, 0], [1., 1.])
    M = np.array([[-249., -453, -397, -52], [-453, -731, -601, -78], [-397, -601, -648, -91], [-52, -78, -91, -13]])
    ii, jj = np.nonzero(M)
    A4 = ref.RefMatrix.from_triplets(4, 4, ii, jj, M[ii, jj])
    b4 = np.array([-1277. / 2, -2015. / 2, -3907. / 4, -533. / 4])
    B = ref.RefMatrix.from_triplets(1, 2, [0, 0], [0, 1], [1., 1.])
    t2 = A2.minres(np.zeros(2), np.array([1., 2.]), 10, 1e-10)
    t3 = A4.minres(np.zeros(4), b4, 10, 1e-10)
    t5 = A4.pminres_dummy(np.zeros(4), b4, 10, 1e-10)
    Q, a, bb = A2.lanczos_basis(np.array([1., 2.]))
    Sb, STb = ref.blockmat_saddle(A2, B, np.array([1., 2., 3.]))
    rb4, ci4, va4 = A4.csr()
    np.savez_compressed(os.path.join(OUT, "reference_kats.npz"),
                        test2=np.array(t2[:3], float), test2_x=t2[3], test3=np.array(t3[:3], float), test3_x=t3[3],
                        test5=np.array(t5[:3], float), test5_x=t5[3], lanczos_Q=Q, lanczos_a=a, lanczos_b=bb,
                        transp_mul_B=B.spmv_transpose(np.array([3.])), block_Sb=Sb, block_STb=STb,
                        A4_rowbeg=rb4, A4_colind=ci4, A4_val=va4, b4=b4)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
