"""CPU tests of the oracle restatement of the SURVEY.md section 8(f) callers (oracle/drops_oracle_ext.c):
BBTDiag, MaxEigenvalueScaledMatrix, Chebyshev, PCG<ChebyshevPcCL>, PCGNE, IDRS, block / composite products.
Checked bit for bit against tests/golden/ext_callers.npz (written by the compiled reference,
oracle/make_golden_ext.py) and, where oracle/_ref is present, against the compiled reference live."""
import numpy as np
import pytest

from conftest import load_golden


def systems(mod, cls):
    gA, gN = load_golden("brick_fe2_N4_num1.npz"), load_golden("brick_fe2_N4_num1_nonsym.npz")
    E = load_golden("ext_callers.npz")
    n = int(gA["rows"])
    M = getattr(mod, cls)
    A = M(n, n, gA["rowbeg"], gA["colind"], gA["val"])
    N = M(n, n, gN["rowbeg"], gN["colind"], gN["val"])
    B = M(int(E["B_rows"]), int(E["B_cols"]), E["B_rowbeg"], E["B_colind"], E["B_val"])
    Cm = M(int(E["C_rows"]), int(E["C_cols"]), E["C_rowbeg"], E["C_colind"], E["C_val"])
    return E, gA, gN, A, N, B, Cm


def same(r, E, key):
    exp = E[key]
    assert (float(r[0]), r[1], r[2]) == (exp[0], int(exp[1]), exp[2]), (key, r[:3], exp)
    assert np.array_equal(r[3], E[key + "_x"]), key


def test_oracle_chebyshev_pcgne_match_reference_fixture(port):
    E, gA, gN, A, N, B, Cm = systems(port, "PortMatrix")
    n, m, m2 = A.rows, B.rows, Cm.rows
    assert np.array_equal(B.bbt_diag(), E["B_bbt"])
    assert A.max_eig_scaled(A.diag(), False) == float(E["A_lambda"])
    assert B.max_eig_scaled(B.bbt_diag(), True) == float(E["B_lambda_ne"])
    for tag, (M, b, ne) in {"A": (A, gA["b"], False), "B": (B, E["bB"], True), "C": (Cm, E["bC"], True)}.items():
        for deg in (1, 3, 5):
            assert np.array_equal(M.chebyshev_apply(b, ne=ne, degree=deg)[0], E[f"{tag}_cheb_d{deg}"]), (tag, deg)
    assert np.array_equal(A.chebyshev_apply(gA["b"], scale=1.3, degree=4, ratio=12.0)[0], E["A_cheb_s13_d4_r12"])
    same(A.pcg_chebyshev(np.zeros(n), gA["b"], 200, 1e-10, True), E, "A_pcg_cheb")
    same(A.pcg_chebyshev(np.zeros(n), gA["b"], 200, 1e-8, False, 1.1, 5, 30.0), E, "A_pcg_cheb_d5_abs")
    same(A.pcg_chebyshev(np.zeros(n), gA["b"], 3, 1e-14, True), E, "A_pcg_cheb_maxit")
    # DiagPcCL (precond.h:618-642)
    same(A.pcg_diagpc(np.zeros(n), gA["b"], 200, 1e-10, True, E["A_diagpc_D"]), E, "A_pcg_diagpc")
    same(A.pcg_diagpc(np.zeros(n), gA["b"], 4, 1e-14, False, E["A_diagpc_D"]), E, "A_pcg_diagpc_abs_maxit")
    assert np.array_equal(E["A_diagpc_D"] * gA["b"], E["A_diagpc_apply"])
    for pc in (0, 1, 2):
        same(B.pcgne(np.zeros(m), E["bB"], 500, 1e-10, True, pc=pc, omega=1.0), E, f"B_pcgne_pc{pc}")
        same(Cm.pcgne(np.zeros(m2), E["bC"], 500, 1e-9, False, pc=pc, omega=0.8), E, f"C_pcgne_pc{pc}")
    same(B.pcgne(E["B_u0"], E["bB"], 500, 1e-10, True, pc=2), E, "B_pcgne_pc2_guess")
    same(B.pcgne(np.zeros(m), E["bB"], 4, 1e-14, True, pc=1), E, "B_pcgne_pc1_maxit")
    # the generic driver with the plain preconditioners is the PCG of drops_oracle.c
    r0 = A.pcg_jacobi(np.zeros(n), gA["b"], 200, 1e-10, True, 0.9)
    r1 = A._pcg_ext(0, 1, 0.9, 1.0, 3, 30.0, np.zeros(n), gA["b"], 200, 1e-10, True, False)
    assert r0[:3] == r1[:3] and np.array_equal(r0[3], r1[3])


def test_oracle_idrs_matches_reference_fixture(port):
    from oracle import ref as refconst       # only RAND_MAX / shadow_from_draws (pure numpy), no library load
    E, gA, gN, A, N, B, Cm = systems(port, "PortMatrix")
    n = A.rows
    for s in (1, 4, 6):
        P = refconst.shadow_from_draws(E[f"idrs_draws_s{s}"], s)
        same(N.idrs(np.zeros(n), gN["b"], 300, 1e-10, P, True, s=s, pc=1), E, f"N_idrs_s{s}_jac")
        same(N.idrs(np.zeros(n), gN["b"], 300, 1e-8, P, False, s=s, pc=0), E, f"N_idrs_s{s}_dummy_abs")
    P4 = refconst.shadow_from_draws(E["idrs_draws_s4"], 4)
    same(N.idrs(np.zeros(n), gN["b"], 7, 1e-14, P4, True, s=4, pc=1), E, "N_idrs_s4_maxit")
    same(A.idrs(np.zeros(n), gA["b"], 300, 1e-10, P4, True, s=4, omega_bound=0.9, pc=1), E, "A_idrs_s4_ob09")
    # the reference's return convention (krylovsolver.h:1689-1696): "false" on convergence inside the loop,
    # and the loop over k does not look at max_iter: 7 allowed, 10 done
    assert E["N_idrs_s4_jac"][0] == 0.0 and E["N_idrs_s4_maxit"][0] == 1.0 and E["N_idrs_s4_maxit"][1] == 10
    # converged start vector: returns true with max_iter = 0 and the absolute residual (:1597-1601)
    xs = np.zeros(n)
    r = N.idrs(xs, np.zeros(n), 10, 1e-10, P4, True, s=4, pc=1)
    assert r[:3] == (1, 0, 0.0)


def test_oracle_block_and_composite_match_reference_fixture(port):
    E, gA, gN, A, N, B, Cm = systems(port, "PortMatrix")
    gW = load_golden("random_rect_wide.npz")
    W = port.PortMatrix(int(gW["rows"]), int(gW["cols"]), gW["rowbeg"], gW["colind"], gW["val"])
    xs = E["x_saddle"]
    cases = {"saddle": ([A, B, B, None], [0, 1, 0, 0]), "nodiag": ([None, B, B, None], [0, 1, 0, 0]),
             "full": ([N, B, B, None], [1, 1, 0, 0])}
    for tag, (blocks, ops) in cases.items():
        assert np.array_equal(port.block_mul(blocks, ops, xs), E[f"block_{tag}"]), tag
        assert np.array_equal(port.block_mul(blocks, ops, xs, transposed=True), E[f"block_{tag}_T"]), tag
    assert np.array_equal(port.composite_mul(B, 1, B, 0, E["bB"]), E["comp_BBt"])
    assert np.array_equal(port.composite_mul(B, 0, B, 1, gA["x_in"]), E["comp_BtB"])
    assert np.array_equal(port.composite_mul(B, 1, A, 0, gA["x_in"], transposed=True), E["comp_AB_T"])
    assert np.array_equal(port.composite_mul(W, 1, W, 0, E["x_w"]), E["comp_WWt"])
    # tests/blockmat.cpp + Testblockmat.ref: S = (A B^T; B 0) with A = [[0,1],[1,0]], B = (1 1)
    A2 = port.PortMatrix(2, 2, [0, 1, 2], [1, 0], [1., 1.])
    B2 = port.PortMatrix(1, 2, [0, 2], [0, 1], [1., 1.])
    b = np.array([1., 2., 3.])
    assert np.array_equal(port.block_mul([A2, B2, B2, None], [0, 1, 0, 0], b), [5., 4., 3.])
    assert np.array_equal(port.block_mul([A2, B2, B2, None], [0, 1, 0, 0], b, transposed=True), [5., 4., 3.])


def test_oracle_ext_matches_compiled_reference_live(port, refmod):
    """Fresh seeded inputs (not the fixture's) through both checkers: bit-identical."""
    Ep, gA, gN, A, N, B, Cm = systems(port, "PortMatrix")
    _, _, _, RA, RN, RB, RC = systems(refmod, "RefMatrix")
    rng = np.random.default_rng(99)
    n, m = A.rows, B.rows
    b, bb = rng.standard_normal(n), rng.standard_normal(m)
    for deg, ratio, scale in [(2, 30.0, 1.1), (6, 8.0, 1.0)]:
        assert np.array_equal(A.chebyshev_apply(b, False, scale, deg, ratio)[0], RA.chebyshev_apply(b, False, scale, deg, ratio)[0])
        assert np.array_equal(B.chebyshev_apply(bb, True, scale, deg, ratio)[0], RB.chebyshev_apply(bb, True, scale, deg, ratio)[0])
    o, r = A.pcg_chebyshev(np.zeros(n), b, 100, 1e-9, True, 1.1, 4, 20.0), RA.pcg_chebyshev(np.zeros(n), b, 100, 1e-9, True, 1.1, 4, 20.0)
    assert o[:3] == r[:3] and np.array_equal(o[3], r[3])
    for pc in (0, 1, 2):
        o, r = B.pcgne(np.zeros(m), bb, 300, 1e-9, True, pc=pc, omega=0.7), RB.pcgne(np.zeros(m), bb, 300, 1e-9, True, pc=pc, omega=0.7)
        assert o[:3] == r[:3] and np.array_equal(o[3], r[3]), pc
    s = 3
    draws = rng.integers(0, refmod.RAND_MAX, size=s * n).astype(np.int32)
    o = N.idrs(np.zeros(n), b, 200, 1e-9, refmod.shadow_from_draws(draws, s), True, s=s, pc=1, omega=0.9)
    r = RN.idrs(np.zeros(n), b, 200, 1e-9, draws, True, s=s, pc=1, omega=0.9)
    assert o[:3] == r[:3] and np.array_equal(o[3], r[3])
    x = rng.standard_normal(n + m)
    assert np.array_equal(port.block_mul([A, B, B, None], [0, 1, 0, 0], x), refmod.block_mul([RA, RB, RB, None], [0, 1, 0, 0], x))
    assert np.array_equal(port.block_mul([N, B, B, None], [1, 1, 0, 0], x, True), refmod.block_mul([RN, RB, RB, None], [1, 1, 0, 0], x, True))
