"""Fixtures for the SURVEY.md section 8(f) callers from the COMPILED REFERENCE (oracle/_ref) ->
tests/golden/ext_callers.npz:  BBTDiag, MaxEigenvalueScaledMatrix, ChebyshevPcCL / ChebyshevBBTPcCL,
PCG<ChebyshevPcCL>, PCGNE, IDRS, BlockMatrixCL / CompositeMatrixCL products.

Run where /root/reference exists:  make -C oracle ref && python oracle/make_golden_ext.py
Systems: A = P2 brick stiffness (tests/golden/brick_fe2_N4_num1.npz, 343 rows), N = its nonsymmetric twin,
B = every third row of A (115 x 343: full row rank, the shape of a discrete divergence), W = random_rect_wide.
Solver results are stored as "<key>" = [ret, iterations, residual] and "<key>_x".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def row_subset(g, step):
    rb = g["rowbeg"].astype(np.int64)
    rows = np.arange(0, int(g["rows"]), step)
    lens = rb[rows + 1] - rb[rows]
    idx = np.concatenate([np.arange(rb[r], rb[r + 1]) for r in rows])
    return len(rows), int(g["cols"]), np.r_[0, np.cumsum(lens)].astype(np.uint64), g["colind"][idx], g["val"][idx]


def main():
    gA, gN, gW = load("brick_fe2_N4_num1.npz"), load("brick_fe2_N4_num1_nonsym.npz"), load("random_rect_wide.npz")
    n = int(gA["rows"])
    A = ref.RefMatrix(n, n, gA["rowbeg"], gA["colind"], gA["val"])
    N = ref.RefMatrix(n, n, gN["rowbeg"], gN["colind"], gN["val"])
    m, _, Brb, Bci, Bva = row_subset(gA, 3)
    B = ref.RefMatrix(m, n, Brb, Bci, Bva)
    W = ref.RefMatrix(int(gW["rows"]), int(gW["cols"]), gW["rowbeg"], gW["colind"], gW["val"])   # has empty rows
    g6 = load("brick_fe2_N6_num1.npz")
    m2, n2, Crb, Cci, Cva = row_subset(g6, 2)
    Cm = ref.RefMatrix(m2, n2, Crb, Cci, Cva)                     # 666 x 1331
    rng = np.random.default_rng(20261017)
    out = {"B_rows": m, "B_cols": n, "B_rowbeg": Brb, "B_colind": Bci, "B_val": Bva}
    bA = gA["b"]
    bB = rng.standard_normal(m)
    bW = rng.standard_normal(W.rows)
    bC = rng.standard_normal(m2)
    out["bB"], out["bW"], out["bC"] = bB, bW, bC
    out.update({"C_rows": m2, "C_cols": n2, "C_rowbeg": Crb, "C_colind": Cci, "C_val": Cva})

    out["B_bbt"], out["W_bbt"] = B.bbt_diag(), W.bbt_diag()
    out["A_lambda"] = A.max_eig_scaled(A.diag(), False)
    out["B_lambda_ne"] = B.max_eig_scaled(B.bbt_diag(), True)
    for tag, (M, b, ne) in {"A": (A, bA, False), "B": (B, bB, True), "C": (Cm, bC, True)}.items():
        for deg in (1, 3, 5):
            x, _ = M.chebyshev_apply(b, ne=ne, degree=deg)
            out[f"{tag}_cheb_d{deg}"] = x
    x, _ = A.chebyshev_apply(bA, ne=False, scale=1.3, degree=4, ratio=12.0)
    out["A_cheb_s13_d4_r12"] = x

    def put(key, r):
        out[key] = np.array([float(r[0]), r[1], r[2]])
        assert np.all(np.isfinite(out[key])) and np.all(np.isfinite(r[3])), key
        out[key + "_x"] = r[3]
        print(key, out[key])

    put("A_pcg_cheb", A.pcg_chebyshev(np.zeros(n), bA, 200, 1e-10, True))
    put("A_pcg_cheb_d5_abs", A.pcg_chebyshev(np.zeros(n), bA, 200, 1e-8, False, 1.1, 5, 30.0))
    put("A_pcg_cheb_maxit", A.pcg_chebyshev(np.zeros(n), bA, 3, 1e-14, True))
    for pc in (0, 1, 2):
        put(f"B_pcgne_pc{pc}", B.pcgne(np.zeros(m), bB, 500, 1e-10, True, pc=pc, omega=1.0))
        put(f"C_pcgne_pc{pc}", Cm.pcgne(np.zeros(m2), bC, 500, 1e-9, False, pc=pc, omega=0.8))
    u0 = rng.standard_normal(m)
    out["B_u0"] = u0
    put("B_pcgne_pc2_guess", B.pcgne(u0, bB, 500, 1e-10, True, pc=2))
    put("B_pcgne_pc1_maxit", B.pcgne(np.zeros(m), bB, 4, 1e-14, True, pc=1))

    for s in (1, 4, 6):
        draws = rng.integers(0, ref.RAND_MAX, size=s * n).astype(np.int32)
        out[f"idrs_draws_s{s}"] = draws
        put(f"N_idrs_s{s}_jac", N.idrs(np.zeros(n), gN["b"], 300, 1e-10, draws, True, s=s, pc=1))
        put(f"N_idrs_s{s}_dummy_abs", N.idrs(np.zeros(n), gN["b"], 300, 1e-8, draws, False, s=s, pc=0))
    put("N_idrs_s4_maxit", N.idrs(np.zeros(n), gN["b"], 7, 1e-14, out["idrs_draws_s4"], True, s=4, pc=1))
    put("A_idrs_s4_ob09", A.idrs(np.zeros(n), bA, 300, 1e-10, out["idrs_draws_s4"], True, s=4, omega_bound=0.9, pc=1))

    xs = rng.standard_normal(n + m)
    out["x_saddle"] = xs
    cases = {"saddle": ([A, B, B, None], [0, 1, 0, 0]),            # ( A B^T ; B 0 ), tests/blockmat.cpp
             "upper": ([A, B, None, None], [0, 1, 0, 0]),          # second block row empty is illegal -> skipped below
             "nodiag": ([None, B, B, None], [0, 1, 0, 0]),
             "full": ([N, B, B, None], [1, 1, 0, 0])}
    for tag, (blocks, ops) in cases.items():
        if tag == "upper":
            continue
        out[f"block_{tag}"] = ref.block_mul(blocks, ops, xs)
        out[f"block_{tag}_T"] = ref.block_mul(blocks, ops, xs, transposed=True)
    xw = rng.standard_normal(W.rows)
    out["x_w"] = xw
    out["comp_BBt"] = ref.composite_mul(B, 1, B, 0, bB)              # B * (B^T x)
    out["comp_BtB"] = ref.composite_mul(B, 0, B, 1, gA["x_in"])      # B^T * (B x)
    out["comp_AB_T"] = ref.composite_mul(B, 1, A, 0, gA["x_in"], transposed=True)   # (A B^T)^T x = B (A^T x)
    out["comp_WWt"] = ref.composite_mul(W, 1, W, 0, xw)
    # DiagPcCL (precond.h:618-642): a vector that is NOT the Jacobi diagonal; no random draws (keeps the arrays above)
    D = 1.0 / (A.diag() * (1.0 + 0.05 * (np.arange(n) % 5)))
    out["A_diagpc_D"] = D
    out["A_diagpc_apply"] = ref.diagpc_apply(D, bA)
    put("A_pcg_diagpc", A.pcg_diagpc(np.zeros(n), bA, 200, 1e-10, True, D))
    put("A_pcg_diagpc_abs_maxit", A.pcg_diagpc(np.zeros(n), bA, 4, 1e-14, False, D))
    np.savez_compressed(os.path.join(GOLDEN, "ext_callers.npz"), **out)
    print("wrote ext_callers.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
