"""GPU parity tests of the SURVEY.md section 8(f) callers, through the C ABI, against the compiled reference's
results in tests/golden/ext_callers.npz (oracle/make_golden_ext.py) and against the oracle on larger seeded systems:
BlockMatrixBaseCL / CompositeMatrixBaseCL products and BBTDiag (bit-identical), ChebyshevPcCL / ChebyshevBBTPcCL,
PCG<ChebyshevPcCL>, PCGNE, IDR(s) (iteration counts +-1 for the CG family, solutions 1e-10 / 1e-8 relative)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def gmat(gpu, g, prefix=""):
    return gpu.Matrix(int(g[prefix + "rows"]), int(g[prefix + "cols"]), g[prefix + "rowbeg"], g[prefix + "colind"],
                      g[prefix + "val"])


@pytest.fixture(scope="module")
def sysm(gpu):
    gA, gN = load_golden("brick_fe2_N4_num1.npz"), load_golden("brick_fe2_N4_num1_nonsym.npz")
    E = load_golden("ext_callers.npz")
    return E, gA, gN, gmat(gpu, gA), gmat(gpu, gN), gmat(gpu, E, "B_"), gmat(gpu, E, "C_")


def test_block_composite_bbt_bit_identical(gpu, sysm):
    E, gA, gN, A, N, B, Cm = sysm
    gW = load_golden("random_rect_wide.npz")
    W = gmat(gpu, gW)
    assert np.array_equal(B.bbt_diag(), E["B_bbt"]) and np.array_equal(W.bbt_diag(), E["W_bbt"])
    xs = E["x_saddle"]
    cases = {"saddle": ([A, B, B, None], [0, 1, 0, 0]), "nodiag": ([None, B, B, None], [0, 1, 0, 0]),
             "full": ([N, B, B, None], [1, 1, 0, 0])}
    for tag, (blocks, ops) in cases.items():
        assert gpu.block_dims(blocks, ops) == (A.rows + B.rows, A.rows + B.rows)
        assert np.array_equal(gpu.block_spmv(blocks, ops, xs), E[f"block_{tag}"]), tag
        assert np.array_equal(gpu.block_spmv(blocks, ops, xs, transposed=True), E[f"block_{tag}_T"]), tag
    assert np.array_equal(gpu.composite_spmv(B, 1, B, 0, E["bB"]), E["comp_BBt"])
    assert np.array_equal(gpu.composite_spmv(B, 0, B, 1, gA["x_in"]), E["comp_BtB"])
    assert np.array_equal(gpu.composite_spmv(B, 1, A, 0, gA["x_in"], transposed=True), E["comp_AB_T"])
    assert np.array_equal(gpu.composite_spmv(W, 1, W, 0, E["x_w"]), E["comp_WWt"])
    # the reference's own KAT: tests/blockmat.cpp + Testblockmat.ref
    A2 = gpu.Matrix(2, 2, [0, 1, 2], [1, 0], [1., 1.])
    B2 = gpu.Matrix(1, 2, [0, 2], [0, 1], [1., 1.])
    b = np.array([1., 2., 3.])
    assert np.array_equal(gpu.block_spmv([A2, B2, B2, None], [0, 1, 0, 0], b), [5., 4., 3.])
    assert np.array_equal(gpu.block_spmv([A2, B2, B2, None], [0, 1, 0, 0], b, transposed=True), [5., 4., 3.])
    # device-vector form, incompatible blocks, all-null block row
    x, y = gpu.Vector.from_host(xs), gpu.Vector(A.rows + B.rows)
    hb, ob, _k = gpu._block_args([A, B, B, None], [0, 1, 0, 0])
    gpu.check(gpu.lib().drops_b200_block_spmv_dev(hb, ob, x.h, y.h))
    assert np.array_equal(y.download(), E["block_saddle"])
    with pytest.raises(gpu.DropsB200Error):
        gpu.block_dims([A, B, B, None], [0, 0, 0, 0])
    with pytest.raises(gpu.DropsB200Error):
        gpu.block_dims([A, B, None, None], [0, 1, 0, 0])


def test_chebyshev_apply_and_lambda(gpu, sysm):
    E, gA, gN, A, N, B, Cm = sysm
    # lambda_max: the reference's Norm is a Kahan sum of rounded squares, ours an exactly rounded sum
    assert abs(A.chebyshev_lambda_max(False) - 1.1 * float(E["A_lambda"])) <= 1e-14 * float(E["A_lambda"]) * 8
    assert abs(B.chebyshev_lambda_max(True) - float(E["B_lambda_ne"])) <= 1e-14 * float(E["B_lambda_ne"]) * 8
    for tag, (M, b, ne) in {"A": (A, gA["b"], False), "B": (B, E["bB"], True), "C": (Cm, E["bC"], True)}.items():
        for deg in (1, 3, 5):
            M.set_chebyshev(ne, None, deg, 30.0)
            x = M.precond_apply(b, gpu.PC_CHEBYSHEV, 1.0, ne)
            assert rel(x, E[f"{tag}_cheb_d{deg}"]) <= 1e-13, (tag, deg, rel(x, E[f"{tag}_cheb_d{deg}"]))
    A.set_chebyshev(False, 1.3, 4, 12.0)
    assert rel(A.precond_apply(gA["b"], gpu.PC_CHEBYSHEV), E["A_cheb_s13_d4_r12"]) <= 1e-13
    A.set_chebyshev(False, 1.1, 0, 30.0)                       # degree 0: x untouched (precond.h:1362)
    assert not A.precond_apply(gA["b"], gpu.PC_CHEBYSHEV).any()
    A.set_chebyshev(False)
    B.set_chebyshev(True)
    Cm.set_chebyshev(True)
    # the other kinds through the same entry point
    assert np.array_equal(A.precond_apply(gA["b"], gpu.PC_JACOBI, 0.7), A.jacobi_apply(gA["b"], 0.7))
    assert np.array_equal(B.precond_apply(E["bB"], gpu.PC_JACOBI, 0.8, True), 0.8 * E["bB"] / E["B_bbt"])
    assert np.array_equal(A.precond_apply(gA["b"], gpu.PC_DUMMY), gA["b"])


def near(out, E, key, x_rtol, dit=1):
    exp = E[key]
    assert out[0] == bool(exp[0]) and abs(out[1] - int(exp[1])) <= dit, (key, out[:3], exp)
    if out[1] == int(exp[1]):
        assert rel(out[3], E[key + "_x"]) <= x_rtol, (key, rel(out[3], E[key + "_x"]))
        assert abs(out[2] - exp[2]) <= 1e-3 * exp[2] + 1e-16, (key, out[2], exp[2])


def test_pcg_chebyshev_and_pcgne_against_reference(gpu, sysm):
    E, gA, gN, A, N, B, Cm = sysm
    n, m, m2 = A.rows, B.rows, Cm.rows
    A.set_chebyshev(False)
    near(A.pcg(np.zeros(n), gA["b"], 200, 1e-10, True, gpu.PC_CHEBYSHEV), E, "A_pcg_cheb", 1e-10)
    A.set_chebyshev(False, 1.1, 5, 30.0)
    near(A.pcg(np.zeros(n), gA["b"], 200, 1e-8, False, gpu.PC_CHEBYSHEV), E, "A_pcg_cheb_d5_abs", 1e-9)
    A.set_chebyshev(False)
    out = A.pcg(np.zeros(n), gA["b"], 3, 1e-14, True, gpu.PC_CHEBYSHEV)
    assert out[0] is False and out[1] == 3 and rel(out[3], E["A_pcg_cheb_maxit_x"]) <= 1e-12
    assert abs(out[2] - E["A_pcg_cheb_maxit"][2]) <= 1e-10 * E["A_pcg_cheb_maxit"][2]
    for pc in (gpu.PC_DUMMY, gpu.PC_JACOBI, gpu.PC_CHEBYSHEV):
        near(B.pcgne(np.zeros(m), E["bB"], 500, 1e-10, True, pc, 1.0), E, f"B_pcgne_pc{pc}", 1e-8)
        near(Cm.pcgne(np.zeros(m2), E["bC"], 500, 1e-9, False, pc, 0.8), E, f"C_pcgne_pc{pc}", 1e-7)
    near(B.pcgne(E["B_u0"], E["bB"], 500, 1e-10, True, gpu.PC_CHEBYSHEV), E, "B_pcgne_pc2_guess", 1e-8)
    out = B.pcgne(np.zeros(m), E["bB"], 4, 1e-14, True, gpu.PC_JACOBI)
    assert out[0] is False and out[1] == 4 and rel(out[3], E["B_pcgne_pc1_maxit_x"]) <= 1e-12
    # residual history of PCGNE = what the reference prints ("PCGNE: iter: i resid: r")
    out = B.pcgne(np.zeros(m), E["bB"], 500, 1e-10, True, gpu.PC_DUMMY, hist=True)
    assert np.isfinite(out[4][:out[1] + 1]).all() and out[4][out[1]] == out[2] and out[4][0] == 1.0
    # the generic driver with DummyPcCL / JACPcCL gives the results of the fused fast path
    for pc, om in ((gpu.PC_DUMMY, 1.0), (gpu.PC_JACOBI, 0.9)):
        a = A.pcg(np.zeros(n), gA["b"], 200, 1e-10, True, pc, om)
        b = A.pcg_general(np.zeros(n), gA["b"], 200, 1e-10, True, pc, om)
        assert a[:3] == b[:3] and np.array_equal(a[3], b[3])
    # edge cases: zero rhs, converged guess, max_iter = 0
    out = B.pcgne(np.zeros(m), np.zeros(m), 10, 1e-10, True, gpu.PC_CHEBYSHEV)
    assert out[0] and out[1] == 0 and out[2] == 0.0
    out = B.pcgne(np.zeros(m), E["bB"], 0, 1e-10, True, gpu.PC_JACOBI)
    assert out[0] is False and out[1] == 0 and out[2] == 1.0
    with pytest.raises(gpu.DropsB200Error):
        Cm.pcg(np.zeros(m2), E["bC"], 10, 1e-10, True, gpu.PC_CHEBYSHEV)       # not square


def test_pcg_diagpc_against_reference(gpu, sysm):
    """PCGSolverCL<DiagPcCL> and DiagPcCL::Apply (precond.h:618-642) against the compiled reference's fixtures."""
    E, gA, gN, A, N, B, Cm = sysm
    n = A.rows
    D = E["A_diagpc_D"]
    with pytest.raises(gpu.DropsB200Error):
        A.pcg(np.zeros(n), gA["b"], 10, 1e-10, True, gpu.PC_DIAG)             # no vector set
    A.set_diag_pc(D)
    near(A.pcg(np.zeros(n), gA["b"], 200, 1e-10, True, gpu.PC_DIAG), E, "A_pcg_diagpc", 1e-10)
    out = A.pcg(np.zeros(n), gA["b"], 4, 1e-14, False, gpu.PC_DIAG)
    assert out[0] is False and out[1] == 4 and rel(out[3], E["A_pcg_diagpc_abs_maxit_x"]) <= 1e-12
    assert abs(out[2] - E["A_pcg_diagpc_abs_maxit"][2]) <= 1e-10 * E["A_pcg_diagpc_abs_maxit"][2]
    # D = omega / diag(A) is the Jacobi preconditioner: same iterates as the fused JACPcCL path
    A.set_diag_pc(0.9 / A.diag())
    a = A.pcg(np.zeros(n), gA["b"], 200, 1e-10, True, gpu.PC_DIAG)
    b = A.pcg(np.zeros(n), gA["b"], 200, 1e-10, True, gpu.PC_JACOBI, 0.9)
    assert a[0] and abs(a[1] - b[1]) <= 1 and rel(a[3], b[3]) <= 1e-10
    A.set_diag_pc(None)
    with pytest.raises(gpu.DropsB200Error):
        A.pcg(np.zeros(n), gA["b"], 10, 1e-10, True, gpu.PC_DIAG)
    with pytest.raises(gpu.DropsB200Error):
        A.gmres(np.zeros(n), gA["b"], 10, 20, 1e-10, True, gpu.PC_DIAG)       # PCG only
    dv, bv, xv = gpu.Vector.from_host(D), gpu.Vector.from_host(gA["b"]), gpu.Vector(n)
    gpu.diagpc_apply(dv, bv, xv)
    assert np.array_equal(xv.download(), E["A_diagpc_apply"])                  # bit-identical


def test_idrs_against_reference(gpu, sysm):
    from oracle import ref as refconst       # RAND_MAX / shadow_from_draws: numpy only
    E, gA, gN, A, N, B, Cm = sysm
    n = A.rows
    for s in (1, 4, 6):
        P = refconst.shadow_from_draws(E[f"idrs_draws_s{s}"], s)
        for key, args in ((f"N_idrs_s{s}_jac", (300, 1e-10, P, True, s, 0.7, gpu.PC_JACOBI)),
                          (f"N_idrs_s{s}_dummy_abs", (300, 1e-8, P, False, s, 0.7, gpu.PC_DUMMY))):
            out = N.idrs(np.zeros(n), gN["b"], *args)
            exp = E[key]
            # IDR(s) is sensitive to the rounding of its s+2 dot products per step (the reference sums them left
            # to right, the device in a tree): same return value, iteration count within 3, and a solution as
            # good as the reference's
            assert out[0] == int(exp[0]) and abs(out[1] - int(exp[1])) <= 3, (key, out[:3], exp)
            tol = args[1]
            normb = np.linalg.norm(gN["b"]) if args[3] else 1.0
            assert out[2] / normb <= tol, (key, out[2])
            assert rel(out[3], E[key + "_x"]) <= 1e-6, (key, rel(out[3], E[key + "_x"]))
    P4 = refconst.shadow_from_draws(E["idrs_draws_s4"], 4)
    out = N.idrs(np.zeros(n), gN["b"], 7, 1e-14, P4, True, 4, 0.7, gpu.PC_JACOBI)
    exp = E["N_idrs_s4_maxit"]
    assert out[0] == 1 and out[1] == 10 and abs(out[2] - exp[2]) <= 1e-6 * exp[2]        # 7 allowed, 10 done (:1627-1671)
    assert rel(out[3], E["N_idrs_s4_maxit_x"]) <= 1e-8
    out = A.idrs(np.zeros(n), gA["b"], 300, 1e-10, P4, True, 4, 0.9, gpu.PC_JACOBI)
    exp = E["A_idrs_s4_ob09"]
    assert out[0] == 0 and abs(out[1] - int(exp[1])) <= 3 and rel(out[3], E["A_idrs_s4_ob09_x"]) <= 1e-7
    # converged start: true, 0 iterations, absolute residual
    out = N.idrs(np.zeros(n), np.zeros(n), 10, 1e-10, P4, True, 4)
    assert out[:3] == (1, 0, 0.0)
    # shadow space drawn inside (srand(time)/rand like the reference): still solves the system
    out = N.idrs(np.zeros(n), gN["b"], 300, 1e-10, None, True, 4)
    assert out[0] == 0 and out[2] / np.linalg.norm(gN["b"]) <= 1e-10
    y = N.spmv(out[3])
    assert np.linalg.norm(y - gN["b"]) / np.linalg.norm(gN["b"]) <= 1e-9


def test_ext_live_against_oracle_larger(gpu):
    """P2 brick N=10 (6859 rows): the saddle-point block product, PCG<ChebyshevPcCL> and PCGNE<ChebyshevBBTPcCL> on
    B = every second row, against the oracle restatement."""
    from oracle import port as oport
    n, rb, ci, va = gpu.brick_csr_host(2, 10, 1)
    rbi = rb.astype(np.int64)
    rows = np.arange(0, n, 2)
    idx = np.concatenate([np.arange(rbi[r], rbi[r + 1]) for r in rows])
    Brb = np.r_[0, np.cumsum(rbi[rows + 1] - rbi[rows])].astype(np.uint64)
    m = len(rows)
    PA, PB = oport.PortMatrix(n, n, rb, ci, va), oport.PortMatrix(m, n, Brb, ci[idx], va[idx])
    GA, GB = gpu.Matrix(n, n, rb, ci, va), gpu.Matrix(m, n, Brb, ci[idx], va[idx])
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n + m)
    assert np.array_equal(gpu.block_spmv([GA, GB, GB, None], [0, 1, 0, 0], x), oport.block_mul([PA, PB, PB, None], [0, 1, 0, 0], x))
    assert np.array_equal(gpu.block_spmv([GA, GB, GB, None], [0, 1, 0, 0], x, True),
                          oport.block_mul([PA, PB, PB, None], [0, 1, 0, 0], x, True))
    b = PA.spmv(gpu.brick_solution_host(2, 10, 1))
    r, o = PA.pcg_chebyshev(np.zeros(n), b, 300, 1e-10, True), GA.pcg(np.zeros(n), b, 300, 1e-10, True, gpu.PC_CHEBYSHEV)
    assert r[0] and o[0] and abs(o[1] - r[1]) <= 1 and rel(o[3], r[3]) <= 1e-10, (r[:3], o[:3])
    bb = rng.standard_normal(m)
    r, o = PB.pcgne(np.zeros(m), bb, 500, 1e-10, True, pc=2), GB.pcgne(np.zeros(m), bb, 500, 1e-10, True, gpu.PC_CHEBYSHEV)
    assert r[0] and o[0] and abs(o[1] - r[1]) <= 1 and rel(o[3], r[3]) <= 1e-8, (r[:3], o[:3])
    # value update invalidates the cached lambda_max / BBTDiag
    l0 = GB.chebyshev_lambda_max(True)
    GB.update_values(va[idx] * 2.0, 2)
    assert np.array_equal(GB.bbt_diag(), 4.0 * PB.bbt_diag())
    assert abs(GB.chebyshev_lambda_max(True) - l0) <= 1e-12 * l0       # D^-1 B B^T is scale invariant
