"""CPU tests of the oracle: the plain-C restatement (oracle/drops_oracle.c) is pinned against
  (1) the reference's own known-answer tests (src/tests/minres.cpp + TestMinres.ref,
      src/tests/blockmat.cpp + Testblockmat.ref, src/tests/vectest.cpp + Testvectest.ref),
  (2) golden fixtures produced by the compiled reference (oracle/make_golden.py), bit for bit,
  (3) the compiled reference itself (oracle/_ref), bit for bit, where it is present.
"""
import numpy as np
import pytest

from conftest import golden_files, load_golden

M4 = np.array([[-249., -453, -397, -52], [-453, -731, -601, -78], [-397, -601, -648, -91], [-52, -78, -91, -13]])
B4 = np.array([-1277. / 2, -2015. / 2, -3907. / 4, -533. / 4])


def dense_csr(M):
    rb, ci, va = [0], [], []
    for row in M:
        nz = np.nonzero(row)[0]
        ci += nz.tolist()
        va += row[nz].tolist()
        rb.append(len(ci))
    return np.array(rb, np.uint64), np.array(ci, np.uint64), np.array(va)


# ---- (1) the reference's own KATs, values from the .ref specification files ---------------------
def test_kat_lanczos_lucky_breakdown(port):
    """tests/minres.cpp TEST1 / TestMinres.ref: 'lucky breakdown after 2 vectors'."""
    A = port.PortMatrix(2, 2, *dense_csr(np.array([[0., 1.], [1., 0.]])))
    Q, a, b = A.lanczos_basis(np.array([1., 2.]))
    assert len(Q) == 2
    assert abs(Q[0] @ Q[1]) < 1e-15 and abs(Q[0] @ Q[0] - 1) < 1e-15


def test_kat_minres_2x2(port):
    """TEST2: x = 2 1, residual 4.213e-16 (abs 1e-10)."""
    A = port.PortMatrix(2, 2, *dense_csr(np.array([[0., 1.], [1., 0.]])))
    conv, it, res, x = A.minres(np.zeros(2), np.array([1., 2.]), 10, 1e-10)
    assert conv and np.allclose(x, [2., 1.], atol=1e-12)
    assert abs(res - 4.213e-16) < 1e-10


def test_kat_minres_4x4(port):
    """TEST3: x = 0.25 0.5 0.75 1, residual 2.87588e-11 (abs 1e-10)."""
    A = port.PortMatrix(4, 4, *dense_csr(M4))
    conv, it, res, x = A.minres(np.zeros(4), B4, 10, 1e-10)
    assert conv and np.allclose(x, [0.25, 0.5, 0.75, 1.0], atol=1e-9)
    assert abs(res - 2.87588e-11) < 1e-10
    assert f"{res:.5e}" == "2.87588e-11"       # the printed reference value, to all its digits


def test_kat_pminres_dummy_4x4(port):
    """TEST5 (PLanczosONBCL<DummyPcCL>): residual 2.87596e-11."""
    A = port.PortMatrix(4, 4, *dense_csr(M4))
    conv, it, res, x = A.pminres_dummy(np.zeros(4), B4, 10, 1e-10)
    assert conv and np.allclose(x, [0.25, 0.5, 0.75, 1.0], atol=1e-9)
    assert f"{res:.5e}" == "2.87596e-11"


def test_kat_transp_mul(port):
    """tests/blockmat.cpp / Testblockmat.ref: transp_mul(B, (3)) = 3 3; S*b = 5 4 3."""
    B = port.PortMatrix(1, 2, *dense_csr(np.array([[1., 1.]])))
    assert np.array_equal(B.spmv_transpose(np.array([3.])), [3., 3.])
    A = port.PortMatrix(2, 2, *dense_csr(np.array([[0., 1.], [1., 0.]])))
    b = np.array([1., 2., 3.])
    # S = [A B^T; B 0] applied blockwise with the same two kernels
    Sb = np.concatenate([A.spmv(b[:2]) + B.spmv_transpose(b[2:]), B.spmv(b[:2])])
    assert np.array_equal(Sb, [5., 4., 3.])


def test_kat_vectest_expression(port):
    """tests/vectest.cpp / Testvectest.ref: v1 + 3 v2 + a v3 equals the raw loop to 1e-16."""
    n = 100000
    rng = np.random.default_rng(1)
    v1, v2, v3 = rng.random(n), rng.random(n), rng.random(n)
    a = 0.1
    z = port.z_xpaypby2(v1, 3.0, v2, a, v3)
    assert np.abs(z - ((v1 + 3.0 * v2) + a * v3)).max() <= 1e-16


# ---- (2) golden fixtures from the compiled reference: bit for bit --------------------------------
SOLVES = {
    "pcgjac": lambda A, b: A.pcg_jacobi(np.zeros_like(b), b, 400, 1e-10, True, 1.0, hist=True),
    "pcgjac_w08": lambda A, b: A.pcg_jacobi(np.zeros_like(b), b, 400, 1e-10, False, 0.8),
    "pcgdummy": lambda A, b: A.pcg_dummy(np.zeros_like(b), b, 400, 1e-10, True),
    "cg": lambda A, b: A.cg(np.zeros_like(b), b, 400, 1e-10, True, hist=True),
    "minres": lambda A, b: A.minres(np.zeros_like(b), b, 400, 1e-9),
    "pminresjac": lambda A, b: A.pminres_jacobi(np.zeros_like(b), b, 400, 1e-9),
    "pcgjac_fixed7": lambda A, b: A.pcg_jacobi(np.zeros_like(b), b, 7, 0.0, True, 1.0),
    "gmresjac_l1": lambda A, b: A.gmres_jacobi(np.zeros_like(b), b, 20, 400, 1e-10, True, 1.0, True),
    "gmresjac_l0": lambda A, b: A.gmres_jacobi(np.zeros_like(b), b, 20, 400, 1e-10, True, 1.0, False),
    "gmresjac_m5": lambda A, b: A.gmres_jacobi(np.zeros_like(b), b, 5, 400, 1e-10, True, 1.0, True),
    "gmresdummy": lambda A, b: A.gmres_dummy(np.zeros_like(b), b, 20, 400, 1e-10, False, True),
}


def _matrix(mod_cls, g):
    return mod_cls(int(g["rows"]), int(g["cols"]), g["rowbeg"], g["colind"], g["val"])


@pytest.mark.parametrize("path", golden_files("brick_*.npz") + golden_files("random_spd*.npz"),
                         ids=lambda p: p.split("/")[-1][:-4])
def test_port_matches_golden_systems(port, path):
    g = dict(np.load(path))
    A = _matrix(port.PortMatrix, g)
    assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"])
    assert np.array_equal(A.spmv_transpose(g["x_in"]), g["y_ATx"])
    if "diag" in g:
        assert np.array_equal(A.diag(), g["diag"])
    if "jac_w07" in g:
        assert np.array_equal(A.jacobi_apply(g["x_in"], 0.7), g["jac_w07"])
    tag = "nonsym" if "nonsym" in path else "sym"
    checked = 0
    for name, fn in SOLVES.items():
        key = f"{tag}_{name}"
        if key not in g:
            continue
        out = fn(A, g["b"])
        conv, it, res = bool(g[key][0]), int(g[key][1]), g[key][2]
        assert (out[0], out[1]) == (conv, it), name
        assert out[2] == res, name
        assert np.array_equal(out[3], g[key + "_x"]), name
        if key + "_hist" in g:
            assert np.array_equal(out[4][1:it + 1], g[key + "_hist"][1:]), name
        checked += 1
    assert checked >= 4


@pytest.mark.parametrize("name", ["random_rect_tall.npz", "random_rect_wide.npz", "random_ragged.npz"])
def test_port_matches_golden_ragged(port, name):
    g = load_golden(name)
    A = _matrix(port.PortMatrix, g)
    assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"])
    assert np.array_equal(A.spmv_transpose(g["xt_in"]), g["y_ATx"])


def test_port_matches_golden_vectors(port):
    g = load_golden("vectors.npz")
    x, y, y2 = g["x"], g["y"], g["y2"]
    assert port.dot(x, y) == g["dot"] and port.norm(x) == g["norm"] and port.norm_sq(x) == g["norm_sq"]
    assert port.supnorm(x) == g["supnorm"]
    assert port.kahan_dot(x, y) == g["kahan_dot"] and port.kahan_norm(x) == g["kahan_norm"]
    assert np.array_equal(port.axpy(0.37, x, y), g["axpy"])
    assert np.array_equal(port.z_xpay(x, -1.25, y), g["z_xpay"])
    assert np.array_equal(port.z_xpaypby2(x, 0.5, y, -3.0, y2), g["z_xpaypby2"])


def test_port_matches_golden_kats(port):
    g = load_golden("reference_kats.npz")
    A4 = port.PortMatrix(4, 4, g["A4_rowbeg"], g["A4_colind"], g["A4_val"])
    t3 = A4.minres(np.zeros(4), g["b4"], 10, 1e-10)
    assert np.array_equal(np.array(t3[:3], float), g["test3"]) and np.array_equal(t3[3], g["test3_x"])
    t5 = A4.pminres_dummy(np.zeros(4), g["b4"], 10, 1e-10)
    assert np.array_equal(np.array(t5[:3], float), g["test5"]) and np.array_equal(t5[3], g["test5_x"])
    A2 = port.PortMatrix(2, 2, *dense_csr(np.array([[0., 1.], [1., 0.]])))
    Q, a, b = A2.lanczos_basis(np.array([1., 2.]))
    assert np.array_equal(Q, g["lanczos_Q"]) and np.array_equal(a, g["lanczos_a"]) and np.array_equal(b, g["lanczos_b"])


# ---- (3) the compiled reference, live, where it exists -------------------------------------------
def test_port_equals_compiled_reference_live(port, refmod):
    import scipy.sparse as sp
    rng = np.random.default_rng(7)
    n = 400
    Mx = sp.random(n, n, density=0.03, random_state=rng, format="csr")
    A = (Mx + Mx.T).tocsr()
    A = (A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 0.5)).tocsr()
    A.sort_indices()
    R = refmod.RefMatrix(n, n, A.indptr, A.indices, A.data)
    P = port.PortMatrix(n, n, A.indptr, A.indices, A.data)
    x, b = rng.standard_normal(n), rng.standard_normal(n)
    assert np.array_equal(R.spmv(x), P.spmv(x))
    assert np.array_equal(R.spmv_transpose(x), P.spmv_transpose(x))
    for name, args in [("pcg_jacobi", (np.zeros(n), b, 300, 1e-11, True)),
                       ("cg", (np.zeros(n), b, 300, 1e-11, False)),
                       ("gmres_jacobi", (np.zeros(n), b, 10, 300, 1e-11)),
                       ("minres", (np.zeros(n), b, 300, 1e-10)),
                       ("pminres_jacobi", (np.zeros(n), b, 300, 1e-10))]:
        r, p = getattr(R, name)(*args), getattr(P, name)(*args)
        assert r[:3] == p[:3], name
        assert np.array_equal(r[3], p[3]), name


def test_reference_kats_live(refmod):
    """The compiled reference reproduces its own .ref values here (TestMinres.ref, Testblockmat.ref)."""
    A2 = refmod.RefMatrix.from_triplets(2, 2, [0, 1], [1, 0], [1., 1.])
    conv, it, res, x = A2.minres(np.zeros(2), np.array([1., 2.]), 10, 1e-10)
    assert np.allclose(x, [2., 1.], atol=1e-12) and abs(res - 4.213e-16) < 1e-10
    ii, jj = np.nonzero(M4)
    A4 = refmod.RefMatrix.from_triplets(4, 4, ii, jj, M4[ii, jj])
    conv, it, res, x = A4.minres(np.zeros(4), B4, 10, 1e-10)
    assert f"{res:.5e}" == "2.87588e-11" and np.allclose(x, [0.25, 0.5, 0.75, 1.0], atol=1e-9)
    B = refmod.RefMatrix.from_triplets(1, 2, [0, 0], [0, 1], [1., 1.])
    assert np.array_equal(B.spmv_transpose(np.array([3.])), [3., 3.])
    Sb, STb = refmod.blockmat_saddle(A2, B, np.array([1., 2., 3.]))
    assert np.array_equal(Sb, [5., 4., 3.]) and np.array_equal(STb, [5., 4., 3.])


def test_oracle_bicgstab_and_gcr_match_reference_fixtures(port):
    """oracle/drops_oracle.c BiCGStab (krylovsolver.h:1189-1258) and GCR (:1410-1483, including the
    stale vacc[min_idx] of the truncation branch) against tests/golden/krylov_extra.npz, produced by
    the compiled reference (oracle/make_golden_extra.py): bit-identical, and the reference's own
    known-answer systems (src/tests/bicgstab.cpp + TestBiCGStab.ref, src/tests/gcr.cpp + TestGCR.ref)."""
    E = load_golden("krylov_extra.npz")
    systems = sorted({k.split("/")[0] for k in E if not k.startswith("kat/")})
    assert len(systems) >= 6
    calls = {"bicgstab_jac": lambda P, x0, b: P.bicgstab(x0, b, 400, 1e-10, True, 1, 1.0),
             "bicgstab_dummy": lambda P, x0, b: P.bicgstab(x0, b, 400, 1e-10, True, 0),
             "bicgstab_jac_abs_w08": lambda P, x0, b: P.bicgstab(x0, b, 400, 1e-9, False, 1, 0.8),
             "gcr_jac_m100": lambda P, x0, b: P.gcr(x0, b, 100, 400, 1e-10, True, 1, 1.0),
             "gcr_dummy_m100": lambda P, x0, b: P.gcr(x0, b, 100, 400, 1e-10, True, 0),
             "gcr_jac_m4_trunc": lambda P, x0, b: P.gcr(x0, b, 4, 60, 1e-10, True, 1, 1.0)}
    for name in systems:
        g = load_golden(name + ".npz")
        n = int(g["rows"])
        P = port.PortMatrix(n, n, g["rowbeg"], g["colind"], g["val"])
        for key, fn in calls.items():
            c, it, res, x = fn(P, np.zeros(n), g["b"])
            exp = E[f"{name}/{key}"]
            assert (c, it) == (bool(exp[0]), int(exp[1])), (name, key)
            assert res == exp[2] and np.array_equal(x, E[f"{name}/{key}_x"]), (name, key)
    # reference test programs: residuals 0 +- 1e-10 (TestBiCGStab.ref, TestGCR.ref)
    A2 = port.PortMatrix(2, 2, [0, 1, 2], [1, 0], [1., 1.])
    M = np.array([[-249., -453, -397, -52], [-453, -731, -601, -78], [-397, -601, -648, -91], [-52, -78, -91, -13]])
    A4 = port.PortMatrix(4, 4, np.arange(0, 17, 4), np.tile(np.arange(4), 4), M.reshape(-1))
    b2, b4 = np.array([1., 2.]), np.array([-1277. / 2, -2015. / 2, -3907. / 4, -533. / 4])
    for key, r in [("kat/bicgstab2", A2.bicgstab(np.zeros(2), b2, 10, 1e-10, True, 0)),
                   ("kat/bicgstab4", A4.bicgstab(np.zeros(4), b4, 10, 1e-10, True, 0)),
                   ("kat/gcr2", A2.gcr(np.full(2, 3.0), b2, 10, 100, 1e-10, False, 0)),
                   ("kat/gcr4", A4.gcr(np.zeros(4), b4, 10, 100, 1e-10, False, 0))]:
        assert r[0] and r[1] == int(E[key][1]) and r[2] == E[key][2] and abs(r[2]) < 1e-10, key
        assert np.array_equal(r[3], E[key + "_x"]), key
    assert np.allclose(E["kat/gcr2_x"], [2., 1.]) and np.allclose(E["kat/gcr4_x"], [0.25, 0.5, 0.75, 1.0])


def test_oracle_lincomb_matches_reference_fixture(port):
    """SparseMatBaseCL::LinComb (spmat.h:1457-1551): pattern merge and values, bit-identical."""
    L = load_golden("lincomb.npz")
    n = int(L["rows"])
    A = (L["A_rowbeg"], L["A_colind"], L["A_val"])
    B = (L["B_rowbeg"], L["B_colind"], L["B_val"])
    for tag in ("dt", "neg", "zero"):
        a, b = L[f"{tag}_coef"]
        rb, ci, va = port.lincomb(n, a, A, b, B)
        assert np.array_equal(rb, L[f"{tag}_rowbeg"]) and np.array_equal(ci, L[f"{tag}_colind"])
        assert np.array_equal(va, L[f"{tag}_val"])
        P = port.PortMatrix(n, n, rb, ci, va)
        assert np.array_equal(P.spmv(L["x_in"]), L[f"{tag}_y"])
    assert len(L["dt_val"]) > max(len(A[2]), len(B[2]))        # the patterns really differ
