"""GPU parity tests: the CUDA path, called through the C ABI, against
  * the golden fixtures produced by the compiled reference (tests/golden, oracle/make_golden.py),
  * the oracle (oracle/drops_oracle.c) on seeded brick systems it finishes in seconds,
  * size-independent properties at BASELINE.json's full size (C2: P2, N=185, 50.2 M rows).
Tolerances (BASELINE.json north_star): SpMV 1e-14 relative per entry -- measured against
(|A||x|)_i, and in the default exact mode the result is required to be BIT-IDENTICAL; Krylov
iteration counts within +-1; solutions within 1e-10 relative.
"""
import numpy as np
import pytest

from conftest import absAx, golden_files, hash_vector, load_golden

pytestmark = pytest.mark.gpu

X_RTOL = 1e-10      # final solutions, relative 2-norm
SPMV_RTOL = 1e-14   # per entry, relative to (|A||x|)_i


def gmat(gpu, g):
    return gpu.Matrix(int(g["rows"]), int(g["cols"]), g["rowbeg"], g["colind"], g["val"])


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ---------------------------------------------------------------------------------------------
# SpMV, transpose, diagonal, Jacobi against the reference's outputs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", golden_files("brick_*.npz") + golden_files("random_*.npz"),
                         ids=lambda p: p.split("/")[-1][:-4])
def test_spmv_bit_identical_to_reference(gpu, path):
    g = dict(np.load(path))
    A = gmat(gpu, g)
    assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"])
    xt = g["xt_in"] if "xt_in" in g else g["x_in"]
    assert np.array_equal(A.spmv_transpose(xt), g["y_ATx"])
    if "diag" in g:
        assert np.array_equal(A.diag(), g["diag"])
    if "jac_w07" in g:
        assert np.array_equal(A.jacobi_apply(g["x_in"], 0.7), g["jac_w07"])
        assert np.array_equal(A.jacobi_apply(g["x_in"], 1.0), g["jac_w1"])
    rb, ci, va = A.download()
    assert np.array_equal(rb, g["rowbeg"]) and np.array_equal(ci, g["colind"]) and np.array_equal(va, g["val"])


@pytest.mark.parametrize("kernel,lanes", [(1, 2), (1, 8), (1, 32)])
def test_spmv_fast_modes_within_tolerance(gpu, kernel, lanes):
    """The tree-summed CSR-vector comparison kernel: |y - y_ref|_i <= 1e-14 (|A||x|)_i."""
    try:
        gpu.set_option("spmv.kernel", kernel)
        gpu.set_option("spmv.lanes", lanes or 8)
        for path in golden_files("brick_fe2_N6*.npz") + golden_files("random_ragged.npz"):
            g = dict(np.load(path))
            A = gmat(gpu, g)
            y = A.spmv(g["x_in"])
            scale = absAx(g["rowbeg"], g["colind"], g["val"], g["x_in"], int(g["rows"]))
            assert np.all(np.abs(y - g["y_Ax"]) <= SPMV_RTOL * scale + 1e-300)
    finally:
        gpu.set_option("spmv.kernel", 0)


@pytest.mark.parametrize("sigma,threads,ctas,spc,variant,prefetch",
                         [(0, 64, 1, 64, 0, 0), (64, 128, 0.25, 64, 1, 1), (512, 512, 2, 64, 2, 0),
                          (4096, 256, 0, 1, 3, 1), (4096, 32, 0, 7, 4, 0), (1024, 256, 0, 1000000, 5, 1)])
def test_spmv_layout_and_launch_options_do_not_change_results(gpu, sigma, threads, ctas, spc, variant, prefetch):
    g = load_golden("brick_fe2_N6_num1.npz")
    try:
        gpu.set_option("spmv.pk", 0)       # the 12-byte stream; the packed one: test_value_dictionary_*
        gpu.set_option("sell.sigma", sigma)
        gpu.set_option("spmv.threads", threads)
        gpu.set_option("spmv.ctas_per_sm", ctas)
        gpu.set_option("spmv.slices_per_cta", spc)
        A = gmat(gpu, g)
        assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"])
        assert np.array_equal(A.spmv_transpose(g["x_in"]), g["y_ATx"])
        out = A.pcg_jacobi(np.zeros(A.rows), g["b"], 400, 1e-10, True)
        assert out[1] == int(g["sym_pcgjac"][1]) and rel(out[3], g["sym_pcgjac_x"]) <= 1e-12
    finally:
        gpu.set_option("spmv.pk", 1)
        gpu.set_option("sell.sigma", 4096)
        gpu.set_option("spmv.threads", 128)
        gpu.set_option("spmv.ctas_per_sm", 0)
        gpu.set_option("spmv.slices_per_cta", 16)


CDICT_DOT_DEFAULT = 1  # library default of "spmv.pk_cdict_dot"
CDICT_DEFAULT = 1     # library default of "spmv.pk_cdict"


PK_DEFAULTS = {"spmv.pk": 1, "pcg.reorder": 1, "pcg.defer_x": 1, "pcg.fused_dot": 1, "krylov.reorder": 1, "spmv.pk_threads": 512,
               "spmv.pk_slices_per_cta": 128, "spmv.pkdict": 0, "spmv.dotmode": 0, "sell.sigma": 4096,
               "sell.valdict": 1,
               "spmv.pk_cdict": CDICT_DEFAULT, "spmv.pk_cdict_dot": CDICT_DOT_DEFAULT,
               "spmv.pk_dot_threads": 384, "spmv.pk_dot_slices_per_cta": 192}


def _restore_pk(gpu):
    for k, v in PK_DEFAULTS.items():
        gpu.set_option(k, v)


def test_value_dictionary_qualifies_only_when_lossless_packing_fits(gpu):
    """Brick systems hold a handful of distinct values and get the packed 4-byte stream; matrices with
    too many distinct values, or whose column + index bits exceed 32, keep the 12-byte stream only."""
    g = load_golden("brick_fe2_N6_num1.npz")
    A = gmat(gpu, g)
    d = A.value_dictionary()
    assert d["packed"] and d["n_distinct"] == len(np.unique(g["val"].view(np.uint64)))
    assert (1 << d["column_bits"]) >= A.cols
    r = load_golden("random_ragged.npz")
    R = gmat(gpu, r)
    nd = len(np.unique(r["val"].view(np.uint64)))
    assert R.value_dictionary()["packed"] == (nd <= 65535) and (not R.value_dictionary()["packed"] or R.value_dictionary()["n_distinct"] == nd)
    # 70 000 distinct values: no dictionary
    m = 70000
    W = gpu.Matrix(m, m, np.arange(m + 1), np.arange(m), np.arange(1, m + 1, dtype=np.float64))
    assert W.value_dictionary() == {"n_distinct": 0, "packed": False, "six_byte": False, "column_bits": 0}
    xw = hash_vector(m)
    assert np.array_equal(W.spmv(xw), np.arange(1, m + 1, dtype=np.float64) * xw)
    # 2^20 columns (20 bits) and 5000 values (13 bits) do not fit one word together
    cols = 1 << 20
    rows = 5000
    V = gpu.Matrix(rows, cols, np.arange(rows + 1), (np.arange(rows) * 209) % cols, np.arange(1, rows + 1) / 7.0)
    assert not V.value_dictionary()["packed"] and V.value_dictionary()["six_byte"] and V.value_dictionary()["n_distinct"] == rows
    xv = hash_vector(cols)
    assert np.array_equal(V.spmv(xv), (np.arange(1, rows + 1) / 7.0) * xv[(np.arange(rows) * 209) % cols])
    try:
        gpu.set_option("sell.valdict", 0)
        assert not gmat(gpu, g).value_dictionary()["packed"]
    finally:
        _restore_pk(gpu)


@pytest.mark.parametrize("variant,threads,spc,dictmode,dotmode,sigma",
                         [(0, 128, 16, 0, 0, 4096), (1, 128, 1, 1, 1, 4096), (2, 64, 7, 0, 2, 64), (3, 128, 1000000, 0, 1, 0),
                          (4, 32, 3, 0, 2, 512), (5, 256, 16, 0, 0, 4096), (6, 128, 2, 1, 2, 1024), (7, 96, 5, 0, 1, 128)])
def test_value_dictionary_stream_is_bit_identical(gpu, variant, threads, spc, dictmode, dotmode, sigma):
    """The packed stream reproduces every product, transpose product, fused dot and solve of the 12-byte
    stream bit for bit, under every launch geometry."""
    g = load_golden("brick_fe2_N6_num1.npz")
    n, rb, ci, va = gpu.brick_csr_host(2, 14, 1)
    xin = hash_vector(n)
    try:
        gpu.set_option("spmv.pk", 0)
        B0 = gpu.Matrix(n, n, rb, ci, va)
        y0, yt0 = B0.spmv(xin), B0.spmv_transpose(xin)
        X, Y = gpu.Vector.from_host(xin), gpu.Vector(n)
        d0 = B0.spmv_dot_dev(X, Y, X)
        s0 = B0.pcg_jacobi(np.zeros(n), y0, 60, 1e-10, True)
        gpu.set_option("spmv.pk", 1)
        gpu.set_option("spmv.pk_cdict", variant & 1)
        gpu.set_option("spmv.pk_cdict_dot", (variant >> 1) & 1)
        gpu.set_option("spmv.pk_threads", threads)
        gpu.set_option("spmv.pk_slices_per_cta", spc)
        gpu.set_option("spmv.pk_dot_threads", threads)
        gpu.set_option("spmv.pk_dot_slices_per_cta", spc)
        gpu.set_option("spmv.pkdict", dictmode)
        gpu.set_option("spmv.dotmode", dotmode)
        gpu.set_option("sell.sigma", sigma)
        A = gmat(gpu, g)
        assert A.value_dictionary()["packed"]
        assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"])
        assert np.array_equal(A.spmv_transpose(g["x_in"]), g["y_ATx"])
        assert A.value_dictionary(transposed=True)["packed"]
        out = A.pcg_jacobi(np.zeros(A.rows), g["b"], 400, 1e-10, True)
        assert out[1] == int(g["sym_pcgjac"][1]) and rel(out[3], g["sym_pcgjac_x"]) <= 1e-12
        B = gpu.Matrix(n, n, rb, ci, va)
        assert np.array_equal(B.spmv(xin), y0) and np.array_equal(B.spmv_transpose(xin), yt0)
        assert B.spmv_dot_dev(X, Y, X) == d0 and np.array_equal(Y.download(), y0)
        s1 = B.pcg_jacobi(np.zeros(n), y0, 60, 1e-10, True)
        # the packed solve runs in the row-ordered numbering: same products, reductions in another order
        assert s1[1] == s0[1] and abs(s1[2] - s0[2]) <= 1e-9 * abs(s0[2]) and rel(s1[3], s0[3]) <= 1e-12
        gpu.set_option("pcg.reorder", 0)
        s2 = B.pcg_jacobi(np.zeros(n), y0, 60, 1e-10, True)
        assert s2[1] == s0[1] and s2[2] == s0[2] and np.array_equal(s2[3], s0[3])
    finally:
        _restore_pk(gpu)


@pytest.mark.parametrize("reorder,defer_x,fused_dot,graph,lean",
                         [(1, 1, 1, 1, 2), (1, 1, 1, 0, 1), (1, 1, 1, 1, 0), (1, 0, 0, 0, 2), (0, 1, 0, 1, 2), (0, 0, 1, 0, 1), (1, 1, 0, 0, 0)])
def test_pcg_forms_agree_with_reference(gpu, reorder, defer_x, fused_dot, graph, lean):
    """Row-ordered solve, deferred x update, separate dot kernel, graph replay: every combination gives
    the reference's iteration count and solution, also when the solve stops inside a chunk, at
    max_iter, or before the first iteration."""
    g = load_golden("brick_fe2_N6_num1.npz")
    n, rb, ci, va = gpu.brick_csr_host(2, 10, 1)
    try:
        gpu.set_option("pcg.reorder", reorder)
        gpu.set_option("pcg.defer_x", defer_x)
        gpu.set_option("pcg.fused_dot", fused_dot)
        gpu.set_option("pcg.graph", graph)
        gpu.set_option("spmv.pk_lean", lean)
        A = gmat(gpu, g)
        out = A.pcg_jacobi(np.zeros(A.rows), g["b"], 400, 1e-10, True)
        assert out[0] and out[1] == int(g["sym_pcgjac"][1]) and rel(out[3], g["sym_pcgjac_x"]) <= 1e-12
        out = A.pcg_jacobi(np.zeros(A.rows), g["b"], 7, 1e-30, True)          # stops at max_iter
        assert not out[0] and rel(out[3], g["sym_pcgjac_fixed7_x"]) <= 1e-12
        out = A.pcg_jacobi(g["sym_pcgjac_x"].copy(), g["b"], 50, 1e-6, True)  # converged at iteration 0 or 1
        assert out[0] and out[1] <= 1
        for maxit in (1, 2, 5, 16, 17, 33):                                    # x = the reference's iterate after maxit steps
            gpu.set_option("pcg.reorder", reorder); gpu.set_option("pcg.defer_x", defer_x)
            out = A.pcg_jacobi(np.zeros(A.rows), g["b"], maxit, 1e-30, True)
            gpu.set_option("pcg.reorder", 0); gpu.set_option("pcg.defer_x", 0)
            ref = A.pcg_jacobi(np.zeros(A.rows), g["b"], maxit, 1e-30, True)
            assert not out[0] and out[1] == ref[1] and rel(out[3], ref[3]) <= 1e-12 and abs(out[2] - ref[2]) <= 1e-9 * ref[2]
        gpu.set_option("pcg.reorder", reorder); gpu.set_option("pcg.defer_x", defer_x)
        for check_every in (1, 2, 3, 16):
            gpu.set_option("pcg.check_every", check_every)
            B = gpu.Matrix(n, n, rb, ci, va)
            xs = gpu.brick_solution_host(2, 10, 1)
            b = B.spmv(xs)
            out = B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
            gpu.set_option("pcg.reorder", 0); gpu.set_option("pcg.defer_x", 0); gpu.set_option("pcg.fused_dot", 1)
            ref = B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
            gpu.set_option("pcg.reorder", reorder); gpu.set_option("pcg.defer_x", defer_x); gpu.set_option("pcg.fused_dot", fused_dot)
            assert out[0] and out[1] == ref[1] and rel(out[3], ref[3]) <= 1e-12 and rel(out[3], xs) <= 1e-8
    finally:
        gpu.set_option("pcg.check_every", 16)
        gpu.set_option("pcg.graph", 1)
        gpu.set_option("spmv.pk_lean", 2)
        _restore_pk(gpu)


@pytest.mark.parametrize("threads,spc,dotmode,graph", [(512, 128, 0, 1), (64, 3, 1, 0), (1024, 1000000, 2, 1)])
def test_dictionary_in_the_parameter_block_keeps_every_bit(gpu, threads, spc, dotmode, graph):
    """"spmv.pk_cdict": dictionaries of <= 47 values ride in the kernel's parameter block and a_ij comes from an
    indexed constant load instead of shared memory.  Same values, same order: products, transposed products,
    fused dots and whole solves keep their bits; matrices with more values stay on the shared-memory path."""
    g = load_golden("brick_fe2_N6_num1.npz")
    r = load_golden("random_ragged.npz")
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1)
    xin = hash_vector(n)
    xs = gpu.brick_solution_host(2, 12, 1)
    try:
        gpu.set_option("sell.block_min_rows", 0)
        gpu.set_option("pcg.graph", graph)
        gpu.set_option("spmv.pk_threads", threads)
        gpu.set_option("spmv.pk_slices_per_cta", spc)
        gpu.set_option("spmv.pk_dot_threads", threads)
        gpu.set_option("spmv.pk_dot_slices_per_cta", spc)
        gpu.set_option("spmv.dotmode", dotmode)
        B = gpu.Matrix(n, n, rb, ci, va)
        assert B.value_dictionary()["packed"] and B.value_dictionary()["n_distinct"] <= 47
        b = B.spmv(xs)
        X, Y = gpu.Vector.from_host(xin), gpu.Vector(n)
        A, R = gmat(gpu, g), gmat(gpu, r)
        res = {}
        for cd in (0, 1, 2, 3):       # 0: shared-memory dictionary; 1: parameter block, U8 kernel also for the fused dot; 3: + U15 dot kernel
            gpu.set_option("spmv.pk_cdict", cd & 1)
            gpu.set_option("spmv.pk_cdict_dot", cd >> 1)
            d = B.spmv_dot_dev(X, Y, X)
            res[cd] = (B.spmv(xin), B.spmv_transpose(xin), np.float64(d), Y.download(),
                       B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)[3], np.float64(B.pcg_jacobi(np.zeros(n), b, 9, 1e-30, True)[2]),
                       B.gmres_jacobi(np.zeros(n), b, 20, 300, 1e-10, True, 1.0, True)[3],
                       A.spmv(g["x_in"]), A.spmv_transpose(g["x_in"]), R.spmv(r["x_in"]))
        for cd in (1, 2, 3):
            for u, v in zip(res[0], res[cd]):
                assert np.array_equal(u, v)
        assert np.array_equal(res[1][7], g["y_Ax"]) and np.array_equal(res[1][8], g["y_ATx"]) and np.array_equal(res[1][9], r["y_Ax"])
    finally:
        gpu.set_option("sell.block_min_rows", 2.0e6)
        gpu.set_option("pcg.graph", 1)
        _restore_pk(gpu)


@pytest.mark.parametrize("graph,unroll2,unroll2_update,omega", [(1, 1, 0, 1.0), (0, 0, 1, 0.8), (1, 1, 1, 1.3)])
def test_diagonal_as_byte_index_keeps_every_bit(gpu, graph, unroll2, unroll2_update, omega):
    """"pcg.diag_index": on operators with a small value dictionary the row-ordered PCG reads the Jacobi diagonal as
    one byte per position + a value list in the kernels' parameter block.  d_i keeps its bits, so iterates,
    residuals and iteration counts equal the solve that reads the fp64 vector -- also after a value update
    (index rebuilt) and with a diagonal of more than 32 distinct values (index not used)."""
    g = load_golden("brick_fe2_N6_num1.npz")
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1)
    xs = gpu.brick_solution_host(2, 12, 1)
    try:
        gpu.set_option("sell.block_min_rows", 0)
        gpu.set_option("pcg.graph", graph)
        gpu.set_option("vec.unroll2", unroll2)
        gpu.set_option("vec.unroll2_update", unroll2_update)
        gpu.set_option("vec.unroll2_update_idx", unroll2_update)     # same per-thread order of the reduction in both solves
        gpu.set_option("pcg.diag_index_direction", unroll2)
        gpu.set_option("vec.one_wave", graph)
        B = gpu.Matrix(n, n, rb, ci, va)
        assert B.layout()["row_ordered_stream"]
        b = B.spmv(xs)
        A = gmat(gpu, g)
        # 40 distinct diagonal values, off-diagonals untouched: the matrix still packs (<= 47 values), the diagonal index does not apply
        va40 = va.copy()
        dpos = np.flatnonzero(ci == np.repeat(np.arange(n, dtype=ci.dtype), np.diff(rb).astype(np.int64)))
        va40[dpos] = va[dpos] * (1.0 + 0.125 * (np.arange(n) % 40))
        res = {}
        for on in (0, 1):
            gpu.set_option("pcg.diag_index", on)
            r = [B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True, omega), B.pcg_jacobi(np.zeros(n), b, 9, 1e-30, True, omega),
                 A.pcg_jacobi(np.zeros(A.rows), g["b"], 400, 1e-10, True)]
            B.update_values(va * 2.0, 10 + 3 * on)
            r.append(B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True, omega))
            B.update_values(va40, 11 + 3 * on)
            r.append(B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True, omega))
            B.update_values(va, 12 + 3 * on)
            res[on] = r
        for r0, r1 in zip(res[0], res[1]):
            assert r0[0] == r1[0] and r0[1] == r1[1] and r0[2] == r1[2] and np.array_equal(r0[3], r1[3])
        assert res[1][2][1] == int(g["sym_pcgjac"][1]) and rel(res[1][2][3], g["sym_pcgjac_x"]) <= 1e-12
        assert res[1][0][0] and rel(res[1][0][3], xs) <= 1e-8 and rel(res[1][3][3], 0.5 * xs) <= 1e-8
    finally:
        gpu.set_option("sell.block_min_rows", 2.0e6)
        gpu.set_option("pcg.graph", 1)
        gpu.set_option("vec.unroll2", 1)
        gpu.set_option("vec.unroll2_update", 0)
        gpu.set_option("vec.unroll2_update_idx", 1)
        gpu.set_option("pcg.diag_index_direction", 0)
        gpu.set_option("vec.one_wave", 0)
        gpu.set_option("pcg.diag_index", 1)
        _restore_pk(gpu)


@pytest.mark.parametrize("tile,sigma", [(1, 4096), (2, 64), (8, 4096), (3, 0)])
def test_block_order_is_transparent(gpu, tile, sigma):
    """Rows cut into blocks of the matrix graph (reorder.cu): any tile size, also on graphs with isolated
    rows and several components -- products keep their bits, solves their iteration counts."""
    g = load_golden("brick_fe2_N6_num1.npz")
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1)
    xin = hash_vector(n)
    xs = gpu.brick_solution_host(2, 12, 1)
    try:
        gpu.set_option("sell.block_order", 0)
        B0 = gpu.Matrix(n, n, rb, ci, va)
        assert not B0.layout()["block_ordered"]
        y0, yt0 = B0.spmv(xin), B0.spmv_transpose(xin)
        b = B0.spmv(xs)
        s0 = B0.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
        gm0 = B0.gmres_jacobi(np.zeros(n), b, 20, 300, 1e-10, True, 1.0, True)
        gpu.set_option("sell.block_order", 1)
        gpu.set_option("sell.block_min_rows", 0)
        gpu.set_option("sell.block_tile", tile)
        gpu.set_option("sell.sigma", sigma)
        A = gmat(gpu, g)
        assert A.layout()["block_ordered"]
        assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"]) and np.array_equal(A.spmv_transpose(g["x_in"]), g["y_ATx"])
        assert np.array_equal(A.diag(), g["diag"])
        out = A.pcg_jacobi(np.zeros(A.rows), g["b"], 400, 1e-10, True)
        assert out[1] == int(g["sym_pcgjac"][1]) and rel(out[3], g["sym_pcgjac_x"]) <= 1e-12
        B = gpu.Matrix(n, n, rb, ci, va)
        assert B.layout()["block_ordered"]
        assert np.array_equal(B.spmv(xin), y0) and np.array_equal(B.spmv_transpose(xin), yt0)
        s1 = B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
        assert s1[0] and s1[1] == s0[1] and rel(s1[3], s0[3]) <= 1e-12
        gm1 = B.gmres_jacobi(np.zeros(n), b, 20, 300, 1e-10, True, 1.0, True)
        assert gm1[1] == gm0[1] and rel(gm1[3], gm0[3]) <= 1e-10
        rb2, ci2, va2 = B.download()
        assert np.array_equal(rb2, rb) and np.array_equal(ci2, ci) and np.array_equal(va2, va)
        # two components, isolated rows (reachable from nowhere), an empty row
        m = 600
        rows = []
        for i in range(m):
            if i % 50 == 7:
                rows.append([])                                   # empty row
            elif i % 50 == 9:
                rows.append([i])                                  # isolated
            else:
                half = 0 if i < m // 2 else m // 2
                rows.append(sorted({i, half + (i - half + 1) % (m // 2), half + (i - half + 17) % (m // 2)}))
        rbd = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint64)
        cid = np.array([c for r in rows for c in r], dtype=np.uint64)
        vad = 1.0 + (np.arange(len(cid)) % 7) * 0.25
        xd = hash_vector(m)
        gpu.set_option("sell.block_order", 0)
        yd0 = gpu.Matrix(m, m, rbd, cid, vad).spmv(xd)
        gpu.set_option("sell.block_order", 1)
        D = gpu.Matrix(m, m, rbd, cid, vad)
        assert D.layout()["block_ordered"] and np.array_equal(D.spmv(xd), yd0)
    finally:
        gpu.set_option("sell.block_order", 1)
        gpu.set_option("sell.block_min_rows", 2.0e6)
        gpu.set_option("sell.block_tile", 8)
        _restore_pk(gpu)


@pytest.mark.parametrize("dictmode,dotmode,sigma,threads,spc", [(0, 0, 4096, 512, 128), (1, 1, 64, 128, 3), (0, 2, 0, 64, 1000000)])
def test_value_dictionary_six_byte_form_is_bit_identical(gpu, dictmode, dotmode, sigma, threads, spc):
    """"sell.pack_bits" = 0 forces the form large systems with many distinct values get: int32 column +
    16-bit dictionary index.  Same bits as the 12-byte stream in products, fused dots and solves."""
    g = load_golden("brick_fe2_N6_num1.npz")
    r = load_golden("random_ragged.npz")
    n, rb, ci, va = gpu.brick_csr_host(2, 14, 1)
    xin = hash_vector(n)
    try:
        gpu.set_option("spmv.pk", 0)
        B0 = gpu.Matrix(n, n, rb, ci, va)
        y0, yt0 = B0.spmv(xin), B0.spmv_transpose(xin)
        X, Y = gpu.Vector.from_host(xin), gpu.Vector(n)
        d0 = B0.spmv_dot_dev(X, Y, X)
        s0 = B0.pcg_jacobi(np.zeros(n), y0, 60, 1e-10, True)
        gm0 = B0.gmres_jacobi(np.zeros(n), y0, 20, 200, 1e-10, True, 1.0, True)
        gpu.set_option("spmv.pk", 1)
        gpu.set_option("sell.pack_bits", 0)
        gpu.set_option("spmv.pkdict", dictmode)
        gpu.set_option("spmv.dotmode", dotmode)
        gpu.set_option("sell.sigma", sigma)
        gpu.set_option("spmv.pk_threads", threads)
        gpu.set_option("spmv.pk_slices_per_cta", spc)
        gpu.set_option("spmv.pk_dot_threads", threads)
        gpu.set_option("spmv.pk_dot_slices_per_cta", spc)
        A = gmat(gpu, g)
        d = A.value_dictionary()
        assert d["six_byte"] and not d["packed"] and A.layout()["row_ordered_kind"] == 2     # twin: positions + the same indices
        assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"]) and np.array_equal(A.spmv_transpose(g["x_in"]), g["y_ATx"])
        out = A.pcg_jacobi(np.zeros(A.rows), g["b"], 400, 1e-10, True)
        assert out[1] == int(g["sym_pcgjac"][1]) and rel(out[3], g["sym_pcgjac_x"]) <= 1e-12
        R = gmat(gpu, r)
        if R.value_dictionary()["six_byte"]:
            assert np.array_equal(R.spmv(r["x_in"]), r["y_Ax"])
            assert np.array_equal(R.spmv_transpose(r["xt_in"] if "xt_in" in r else r["x_in"]), r["y_ATx"])
        B = gpu.Matrix(n, n, rb, ci, va)
        assert B.value_dictionary()["six_byte"]
        assert np.array_equal(B.spmv(xin), y0) and np.array_equal(B.spmv_transpose(xin), yt0)
        assert B.spmv_dot_dev(X, Y, X) == d0 and np.array_equal(Y.download(), y0)
        s1 = B.pcg_jacobi(np.zeros(n), y0, 60, 1e-10, True)
        assert s1[1] == s0[1] and rel(s1[3], s0[3]) <= 1e-12                           # row-ordered now: other reduction order
        gm1 = B.gmres_jacobi(np.zeros(n), y0, 20, 200, 1e-10, True, 1.0, True)
        assert gm1[1] == gm0[1] and rel(gm1[3], gm0[3]) <= 1e-12
        gpu.set_option("pcg.reorder", 0); gpu.set_option("krylov.reorder", 0)          # caller's order: same reductions, same bits
        s2 = B.pcg_jacobi(np.zeros(n), y0, 60, 1e-10, True)
        gm2 = B.gmres_jacobi(np.zeros(n), y0, 20, 200, 1e-10, True, 1.0, True)
        gpu.set_option("pcg.reorder", 1); gpu.set_option("krylov.reorder", 1)
        assert s2[1] == s0[1] and s2[2] == s0[2] and np.array_equal(s2[3], s0[3])
        assert gm2[1] == gm0[1] and np.array_equal(gm2[3], gm0[3])
    finally:
        gpu.set_option("sell.pack_bits", 32)
        _restore_pk(gpu)


def test_value_dictionary_padding_and_special_values(gpu):
    """Ragged rows (padding inside slices), +0.0 / -0.0 / denormal entries, a non-finite x_0 (padding
    points at column 0) and new values on the old pattern: always the bits of the 12-byte stream."""
    rng = np.random.default_rng(7)
    rows, cols = 3000, 2500
    lens = rng.integers(0, 40, rows)
    lens[::97] = 0
    rb = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ci = np.concatenate([np.sort(rng.choice(cols, l, replace=False)) for l in lens]).astype(np.uint64)
    pool = np.array([0.0, -0.0, 1.0, -1.0, 0.1, 1e-310, -2.5, 3.0e200, 7.0 / 3.0])
    va = pool[rng.integers(0, len(pool), len(ci))]
    x = rng.standard_normal(cols)
    try:
        for x0 in (x[0], np.inf, np.nan, -0.0):
            x[0] = x0
            gpu.set_option("spmv.pk", 0)
            A0 = gpu.Matrix(rows, cols, rb, ci, va)
            y0 = A0.spmv(x)
            gpu.set_option("spmv.pk", 1)
            A = gpu.Matrix(rows, cols, rb, ci, va)
            assert A.value_dictionary()["packed"] and A.value_dictionary()["n_distinct"] == len(pool)
            y = A.spmv(x)
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64))
            gpu.set_option("sell.pack_bits", 0)
            A6 = gpu.Matrix(rows, cols, rb, ci, va)
            gpu.set_option("sell.pack_bits", 32)
            assert A6.value_dictionary()["six_byte"]
            assert np.array_equal(A6.spmv(x).view(np.uint64), y0.view(np.uint64))
        x[0] = 0.25
        xt = rng.standard_normal(rows)
        gpu.set_option("spmv.pk", 0)
        yt0 = A0.spmv_transpose(xt)
        gpu.set_option("spmv.pk", 1)
        assert np.array_equal(A.spmv_transpose(xt).view(np.uint64), yt0.view(np.uint64))
        # new values on the old pattern: the dictionary is rebuilt; values that no longer compress drop it
        va2 = pool[::-1][rng.integers(0, len(pool), len(ci))] * 3.0
        A.update_values(va2, 2)
        A0.update_values(va2, 2)
        assert A.value_dictionary()["packed"]
        gpu.set_option("spmv.pk", 0)
        y0 = A0.spmv(x)
        gpu.set_option("spmv.pk", 1)
        assert np.array_equal(A.spmv(x).view(np.uint64), y0.view(np.uint64))
        va3 = rng.standard_normal(len(ci)) if len(ci) > 65535 else np.arange(len(ci), dtype=np.float64)
        A.update_values(va3, 3)
        if len(np.unique(va3)) > 65535:
            assert not A.value_dictionary()["packed"]
        A0.update_values(va3, 3)
        assert np.array_equal(A.spmv(x), A0.spmv(x))
    finally:
        _restore_pk(gpu)


@pytest.mark.parametrize("cells,lines,tiles,sigma,stride", [(12, 1, (8, 8, 8), 4096, 8), (40, 0, (8, 8, 8), 4096, 8), (40, 1, (8, 8, 8), 4096, 8),
                                                            (40, 1, (3, 5, 2), 512, 2), (40, 1, (16, 16, 16), 64, 1), (36, 1, (8, 8, 8), 0, 8)])
def test_line_segments_in_graph_blocks_are_transparent(gpu, cells, lines, tiles, sigma, stride):
    """"sell.line_order": rows that repeat each other translated along the caller's numbering are chained and cut into
    segments of 32 = one slice each (reorder.cu).  A permutation like any other: every product keeps its bits, PCG
    (row-ordered) its iteration count; the P2 brick has lines of cells - 3 ... cells - 1 same-shape rows."""
    n, rb, ci, va = gpu.brick_csr_host(2, cells, 1)
    xin = hash_vector(n)
    xs = gpu.brick_solution_host(2, cells, 1)
    try:
        gpu.set_option("sell.block_order", 0)
        B0 = gpu.Matrix(n, n, rb, ci, va)
        assert B0.layout()["line_segments"] == 0
        y0, yt0 = B0.spmv(xin), B0.spmv_transpose(xin)
        b = B0.spmv(xs)
        s0 = B0.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
        gpu.set_option("sell.block_order", 1)
        gpu.set_option("sell.block_min_rows", 0)
        gpu.set_option("sell.line_order", lines)
        gpu.set_option("sell.line_stride", stride)
        gpu.set_option("sell.block_tile_a", tiles[0]); gpu.set_option("sell.block_tile_b", tiles[1]); gpu.set_option("sell.block_tile_c", tiles[2])
        gpu.set_option("sell.sigma", sigma)
        B = gpu.Matrix(n, n, rb, ci, va)
        lay = B.layout()
        assert lay["block_ordered"] and lay["row_ordered_stream"]
        if lines and cells >= 36 and stride >= 2:
            # 8 DOF classes, (cells - 1)^2 lines each, at least one complete segment per interior line
            assert lay["line_segments"] >= 6 * (cells - 3) ** 2, lay
        elif not lines:
            assert lay["line_segments"] == 0
        assert np.array_equal(B.spmv(xin), y0) and np.array_equal(B.spmv_transpose(xin), yt0)
        s1 = B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
        assert s1[0] and s1[1] == s0[1] and rel(s1[3], s0[3]) <= 1e-12
        if lines and cells >= 36 and stride >= 2:
            # segment descriptors ("sell.segments"): most slices of the twin are one dictionary index + two runs of positions
            # per entry-row; the segment kernel gives the bits of the packed kernel in PCG, GMRES (Jacobi epilogue) and MINRES
            assert lay["segment_slices"] >= 0.8 * lay["line_segments"] and lay["segment_entries"] >= 0.4 * len(va), lay
            m1 = B.minres(np.zeros(n), b, 40, 1e-30, True)
            g1 = B.gmres_jacobi(np.zeros(n), b, 10, 25, 1e-30, True, 1.0, True)
            gpu.set_option("spmv.segments", 0)
            s2 = B.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
            m2 = B.minres(np.zeros(n), b, 40, 1e-30, True)
            g2 = B.gmres_jacobi(np.zeros(n), b, 10, 25, 1e-30, True, 1.0, True)
            gpu.set_option("spmv.segments", 1)
            assert s2[1] == s1[1] and s2[2] == s1[2] and np.array_equal(s2[3], s1[3])
            assert m2[1] == m1[1] and np.array_equal(m2[3], m1[3]) and g2[1] == g1[1] and np.array_equal(g2[3], g1[3])
        gm0 = B0.gmres_jacobi(np.zeros(n), b, 20, 60, 1e-10, True, 1.0, True)
        gm1 = B.gmres_jacobi(np.zeros(n), b, 20, 60, 1e-10, True, 1.0, True)
        assert gm1[1] == gm0[1] and rel(gm1[3], gm0[3]) <= 1e-12        # row-ordered in two different orders
        gpu.set_option("krylov.reorder", 0)
        gm2 = B.gmres_jacobi(np.zeros(n), b, 20, 60, 1e-10, True, 1.0, True)
        gm3 = B0.gmres_jacobi(np.zeros(n), b, 20, 60, 1e-10, True, 1.0, True)
        gpu.set_option("krylov.reorder", 1)
        assert gm2[1] == gm3[1] and np.array_equal(gm2[3], gm3[3])      # caller's order: same bits
    finally:
        gpu.set_option("spmv.segments", 1)
        gpu.set_option("sell.block_min_rows", 2.0e6)
        gpu.set_option("sell.line_order", 1)
        gpu.set_option("sell.line_stride", 8)
        gpu.set_option("sell.block_tile_a", 0); gpu.set_option("sell.block_tile_b", 0); gpu.set_option("sell.block_tile_c", 0)
        gpu.set_option("sell.sigma", 4096)


def test_line_segments_on_matrices_without_lines(gpu):
    """Random patterns, a tridiagonal matrix (ONE chain of all rows) and a matrix of empty rows: the chain search finds
    nothing or everything, the products keep their bits."""
    rng = np.random.default_rng(11)
    try:
        gpu.set_option("sell.block_min_rows", 0)
        n = 5000
        rows = [np.unique(np.r_[i, rng.choice(n, rng.integers(0, 12), replace=False)]) for i in range(n)]
        rb = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint64)
        ci = np.concatenate(rows)
        ci = ci.astype(np.uint64)
        va = rng.standard_normal(len(ci))
        x = rng.standard_normal(n)
        gpu.set_option("sell.block_order", 0)
        y0 = gpu.Matrix(n, n, rb, ci, va).spmv(x)
        gpu.set_option("sell.block_order", 1)
        A = gpu.Matrix(n, n, rb, ci, va)
        assert A.layout()["block_ordered"] and np.array_equal(A.spmv(x), y0)
        # tridiagonal: rows 1..n-2 are one chain
        n = 4099
        lens = np.full(n, 3); lens[0] = lens[-1] = 2
        rb = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        ci = np.concatenate([[0, 1]] + [[i - 1, i, i + 1] for i in range(1, n - 1)] + [[n - 2, n - 1]]).astype(np.uint64)
        va = np.concatenate([[2.0, -1.0]] + [[-1.0, 2.0, -1.0]] * (n - 2) + [[-1.0, 2.0]])
        x = rng.standard_normal(n)
        gpu.set_option("sell.block_order", 0)
        T0 = gpu.Matrix(n, n, rb, ci, va)
        y0 = T0.spmv(x)
        s0 = T0.pcg_jacobi(np.zeros(n), y0, 50, 1e-30, True)
        gpu.set_option("sell.block_order", 1)
        T = gpu.Matrix(n, n, rb, ci, va)
        assert T.layout()["line_segments"] == (n - 2) // 32
        assert np.array_equal(T.spmv(x), y0) and np.array_equal(T.spmv_transpose(x), T0.spmv_transpose(x))
        s1 = T.pcg_jacobi(np.zeros(n), y0, 50, 1e-30, True)
        assert s1[1] == s0[1] == 50 and rel(s1[3], s0[3]) <= 1e-12
    finally:
        gpu.set_option("sell.block_order", 1)
        gpu.set_option("sell.block_min_rows", 2.0e6)


@pytest.mark.parametrize("block_min_rows,graph", [(0, 1), (2.0e6, 0)])
def test_row_ordered_pcg_on_the_12_byte_stream(gpu, port, block_min_rows, graph):
    """Operators without a value dictionary (variable coefficients: > 65 535 distinct values, or "sell.valdict" = 0)
    get the row-ordered twin of their int32 columns, so PCG iterates on row-ordered vectors on the general
    12-byte stream as well.  Products keep their bits, the solve its iteration count."""
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1)
    rng = np.random.default_rng(7)
    # a symmetric positive definite operator with ~n distinct values: D A D with a smooth random diagonal D
    d = 1.0 + 0.3 * rng.random(n)
    rows = np.repeat(np.arange(n), np.diff(rb.astype(np.int64)))
    vv = va * d[rows] * d[ci.astype(np.int64)]
    P = port.PortMatrix(n, n, rb, ci, vv)
    xs = gpu.brick_solution_host(2, 12, 1)
    b = P.spmv(xs)
    ref = P.pcg_jacobi(np.zeros(n), b, 2000, 1e-10, True)
    xin = hash_vector(n)
    try:
        gpu.set_option("sell.block_min_rows", block_min_rows)
        gpu.set_option("pcg.graph", graph)
        G = gpu.Matrix(n, n, rb, ci, vv)
        assert G.value_dictionary()["n_distinct"] == 0 and G.layout()["row_ordered_kind"] == 2
        assert G.layout()["block_ordered"] == (block_min_rows == 0)
        assert np.array_equal(G.spmv(xin), P.spmv(xin)) and np.array_equal(G.spmv_transpose(xin), P.spmv_transpose(xin))
        out = G.pcg_jacobi(np.zeros(n), b, 2000, 1e-10, True)
        assert ref[0] and out[0] and abs(out[1] - ref[1]) <= 1 and rel(out[3], ref[3]) <= X_RTOL
        gpu.set_option("pcg.reorder", 0)                      # the same solve in the caller's numbering
        out0 = G.pcg_jacobi(np.zeros(n), b, 2000, 1e-10, True)
        assert out0[1] == out[1] and rel(out0[3], out[3]) <= 1e-12
        # a constant-coefficient matrix forced onto the 12-byte stream takes the same path
        gpu.set_option("pcg.reorder", 1)
        gpu.set_option("sell.valdict", 0)
        H = gpu.Matrix(n, n, rb, ci, va)
        assert H.layout()["row_ordered_kind"] == 2
        Pc = port.PortMatrix(n, n, rb, ci, va)
        bc = Pc.spmv(xs)
        rc = Pc.pcg_jacobi(np.zeros(n), bc, 2000, 1e-10, True)
        oc = H.pcg_jacobi(np.zeros(n), bc, 2000, 1e-10, True)
        assert oc[0] and abs(oc[1] - rc[1]) <= 1 and rel(oc[3], rc[3]) <= X_RTOL
    finally:
        gpu.set_option("sell.block_min_rows", 2.0e6)
        gpu.set_option("pcg.graph", 1)
        gpu.set_option("pcg.reorder", 1)
        gpu.set_option("sell.valdict", 1)


@pytest.mark.parametrize("pkdict", [0, 1])
def test_segment_records_with_value_indices(gpu, pkdict):
    """The form of a matrix DROPS assembled itself on a structured mesh (tests/test_drops_exe.py): columns that are
    translates, values equal only up to rounding -- here: a brick operator whose entries got one of 1500 random last-bit
    perturbations --, at a size where column bits + index bits need more than one word ("sell.pack_bits" = 0 forces that).
    Row-ordered twin of the six-byte form with segment records (k_spmv_sell_sci, 2.5 B per entry): the bits of the six-byte
    kernel in PCG, GMRES and MINRES; the products those of the 12-byte stream."""
    cells = 40
    n, rb, ci, va = gpu.brick_csr_host(2, cells, 1)
    rng = np.random.default_rng(5)
    rows = np.repeat(np.arange(n), np.diff(rb.astype(np.int64)))
    # symmetric rounding variants: entry (i, j) and (j, i) get the same factor out of 1500
    lo, hi = np.minimum(rows, ci.astype(np.int64)), np.maximum(rows, ci.astype(np.int64))
    fac = 1.0 + 2.0 ** -50 * ((lo * 7919 + hi * 104729) % 1500)
    vv = va * fac
    xs = gpu.brick_solution_host(2, cells, 1)
    xin = hash_vector(n)
    try:
        gpu.set_option("sell.block_min_rows", 0)
        gpu.set_option("spmv.pk", 0)
        G0 = gpu.Matrix(n, n, rb, ci, vv)
        y0 = G0.spmv(xin)
        gpu.set_option("spmv.pk", 1)
        gpu.set_option("sell.pack_bits", 0)
        gpu.set_option("spmv.pkdict", pkdict)
        G = gpu.Matrix(n, n, rb, ci, vv)
        d, lay = G.value_dictionary(), G.layout()
        assert d["six_byte"] and not d["packed"] and 1000 < d["n_distinct"] <= 65535, d
        assert lay["row_ordered_kind"] == 2 and lay["block_ordered"] and lay["segment_entries"] >= 0.4 * len(vv), lay
        assert np.array_equal(G.spmv(xin), y0)
        b = G.spmv(xs)
        seg = (G.pcg_jacobi(np.zeros(n), b, 3000, 1e-10, True), G.gmres_jacobi(np.zeros(n), b, 10, 25, 1e-30, True, 1.0, True),
               G.minres(np.zeros(n), b, 40, 1e-30, True))
        gpu.set_option("spmv.segments", 0)
        six = (G.pcg_jacobi(np.zeros(n), b, 3000, 1e-10, True), G.gmres_jacobi(np.zeros(n), b, 10, 25, 1e-30, True, 1.0, True),
               G.minres(np.zeros(n), b, 40, 1e-30, True))
        gpu.set_option("spmv.segments", 1)
        for a, c in zip(seg, six):
            assert a[1] == c[1] and a[2] == c[2] and np.array_equal(a[3], c[3])
        assert seg[0][0] and rel(seg[0][3], xs) <= 1e-7
    finally:
        gpu.set_option("spmv.segments", 1)
        gpu.set_option("spmv.pkdict", 0)
        gpu.set_option("sell.pack_bits", 32)
        gpu.set_option("sell.block_min_rows", 2.0e6)
        _restore_pk(gpu)


@pytest.mark.parametrize("m,left,maxit", [(8, True, 400), (30, False, 400), (200, True, 400), (30, True, 17), (12, True, 12)])
def test_gmres_device_side_givens_keeps_every_bit(gpu, m, left, maxit):
    """"gmres.device_givens" (classical Gram-Schmidt path): the rotations of every column run in a one-block kernel and the
    host enqueues the steps of a cycle without waiting for the columns.  The same operations in the same order: iteration
    count, residual and solution of the host-driven loop, bit for bit -- through restarts, at convergence inside a chunk,
    and when max_iter ends the solve."""
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1, 1.1)
    rows = np.repeat(np.arange(n), np.diff(rb.astype(np.int64)))
    va = va.copy()
    va[ci.astype(np.int64) == rows] *= 1.05
    xs = gpu.brick_solution_host(2, 12, 1)
    G = gpu.Matrix(n, n, rb, ci, va)
    b = G.spmv(xs)
    try:
        gpu.set_option("gmres.ortho", 1)
        outs = []
        for dg, chunk in ((0, 8), (1, 8), (1, 1), (1, 5)):
            gpu.set_option("gmres.device_givens", dg)
            gpu.set_option("gmres.check_every", chunk)
            outs.append(G.gmres_jacobi(np.zeros(n), b, m, maxit, 1e-10, True, 1.0, left))
        for o in outs[1:]:
            assert o[0] == outs[0][0] and o[1] == outs[0][1] and o[2] == outs[0][2] and np.array_equal(o[3], outs[0][3]), (o[:3], outs[0][:3])
        if maxit == 400:
            assert outs[0][0] and rel(outs[0][3], xs) <= 1e-7
    finally:
        gpu.set_option("gmres.ortho", 0)
        gpu.set_option("gmres.device_givens", 1)
        gpu.set_option("gmres.check_every", 8)


def test_polled_fetch_gives_the_same_solves(gpu):
    """"fetch.poll": Hessenberg columns and Lanczos scalars published to mapped pinned memory and polled, against the copy +
    stream synchronisation they replace -- the same bytes reach the host, so GMRES, MINRES and BiCGStab are bit for bit the same."""
    g = load_golden("brick_fe2_N6_num1.npz")
    A = gmat(gpu, g)
    n = A.rows
    try:
        outs = []
        for poll in (1, 0):
            gpu.set_option("fetch.poll", poll)
            outs.append((A.gmres_jacobi(np.zeros(n), g["b"], 20, 400, 1e-10, True, 1.0, True), A.minres(np.zeros(n), g["b"], 400, 1e-9),
                         A.pminres_jacobi(np.zeros(n), g["b"], 400, 1e-9), A.bicgstab(np.zeros(n), g["b"], 400, 1e-10, True)))
        for a, c in zip(*outs):
            assert a[0] == c[0] and a[1] == c[1] and a[2] == c[2] and np.array_equal(a[3], c[3])
    finally:
        gpu.set_option("fetch.poll", 1)


def test_segment_descriptors_on_the_12_byte_stream(gpu):
    """Variable coefficients on a structured mesh: the VALUES differ from row to row, the column structure is still
    translation invariant.  The twin of the int32 columns gets segment descriptors (0.5 instead of 4 bytes per column,
    8.5 instead of 12 per entry) and k_spmv_sell_scv gives the bits of k_spmv_sell in PCG, GMRES and MINRES."""
    cells = 40
    n, rb, ci, va = gpu.brick_csr_host(2, cells, 1)
    rng = np.random.default_rng(3)
    d = 1.0 + 0.3 * rng.random(n)
    rows = np.repeat(np.arange(n), np.diff(rb.astype(np.int64)))
    vv = va * d[rows] * d[ci.astype(np.int64)]                # D A D: symmetric positive definite, ~nnz distinct values
    xs = gpu.brick_solution_host(2, cells, 1)
    try:
        gpu.set_option("sell.block_min_rows", 0)
        G = gpu.Matrix(n, n, rb, ci, vv)
        lay = G.layout()
        assert G.value_dictionary()["n_distinct"] == 0 and lay["row_ordered_kind"] == 2 and lay["block_ordered"]
        assert lay["line_segments"] >= 6 * (cells - 3) ** 2 and lay["segment_entries"] >= 0.4 * len(vv), lay
        b = G.spmv(xs)
        seg = (G.pcg_jacobi(np.zeros(n), b, 3000, 1e-10, True), G.gmres_jacobi(np.zeros(n), b, 10, 25, 1e-30, True, 1.0, True),
               G.minres(np.zeros(n), b, 40, 1e-30, True))
        gpu.set_option("spmv.segments", 0)
        gen = (G.pcg_jacobi(np.zeros(n), b, 3000, 1e-10, True), G.gmres_jacobi(np.zeros(n), b, 10, 25, 1e-30, True, 1.0, True),
               G.minres(np.zeros(n), b, 40, 1e-30, True))
        gpu.set_option("spmv.segments", 1)
        for a, c in zip(seg, gen):
            assert a[1] == c[1] and a[2] == c[2] and np.array_equal(a[3], c[3])
        assert seg[0][0] and rel(seg[0][3], xs) <= 1e-7
        # new values on the old pattern keep the descriptors' structure
        G.update_values(vv * 0.5, 2)
        assert G.layout()["segment_entries"] == lay["segment_entries"]
        o2 = G.pcg_jacobi(np.zeros(n), 0.5 * b, 3000, 1e-10, True)
        assert o2[1] == seg[0][1] and rel(o2[3], seg[0][3]) <= 1e-12
    finally:
        gpu.set_option("spmv.segments", 1)
        gpu.set_option("sell.block_min_rows", 2.0e6)


@pytest.mark.parametrize("m,left", [(8, True), (30, True), (30, False), (200, True)])
def test_gmres_classical_gram_schmidt_option(gpu, port, m, left):
    """"gmres.ortho" = 1: classical Gram-Schmidt as in the reference's distributed ModGMRES (StandardGramSchmidt,
    krylovsolver.h:465-494, :862-866) -- all products of a column against the unmodified w in one fused kernel, then one
    fused update.  Same Krylov space: iteration counts within +-1 of the serial (modified Gram-Schmidt) reference on a
    well-conditioned nonsymmetric system (the level-set stand-in), solutions within 1e-8."""
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1, 1.1)
    rows = np.repeat(np.arange(n), np.diff(rb.astype(np.int64)))
    va = va.copy()
    va[ci.astype(np.int64) == rows] *= 1.25
    P = port.PortMatrix(n, n, rb, ci, va)
    G = gpu.Matrix(n, n, rb, ci, va)
    xs = gpu.brick_solution_host(2, 12, 1)
    b = P.spmv(xs)
    r = P.gmres_jacobi(np.zeros(n), b, m, 2000, 1e-9, True, 1.0, left)
    try:
        gpu.set_option("gmres.ortho", 1)
        l0 = gpu.launch_count()
        o = G.gmres_jacobi(np.zeros(n), b, m, 2000, 1e-9, True, 1.0, left)
        launches = gpu.launch_count() - l0
    finally:
        gpu.set_option("gmres.ortho", 0)
    assert r[0] and o[0]
    assert abs(o[1] - r[1]) <= 1, (o[1], r[1])
    assert rel(o[3], r[3]) <= 1e-8
    assert launches <= 8 * o[1] + 40                       # a handful of kernels per column, however long the basis


def test_wide_and_scalar_vector_kernels_agree(gpu):
    g = load_golden("brick_fe2_N6_num1.npz")
    A = gmat(gpu, g)
    outs = []
    for wide in (1, 0):
        gpu.set_option("vec.wide", wide)
        outs.append(A.pcg_jacobi(np.zeros(A.rows), g["b"], 400, 1e-10, True, hist=True))
    gpu.set_option("vec.wide", 1)
    assert outs[0][1] == outs[1][1]
    assert rel(outs[0][3], outs[1][3]) <= 1e-13
    assert np.allclose(outs[0][4], outs[1][4], rtol=1e-9, equal_nan=True)


def test_reference_kats_on_gpu(gpu):
    """src/tests/minres.cpp + TestMinres.ref and src/tests/blockmat.cpp + Testblockmat.ref."""
    g = load_golden("reference_kats.npz")
    A2 = gpu.Matrix(2, 2, [0, 1, 2], [1, 0], [1., 1.])
    conv, it, res, x = A2.minres(np.zeros(2), np.array([1., 2.]), 10, 1e-10)
    assert conv and it == int(g["test2"][1]) and np.allclose(x, [2., 1.], atol=1e-12) and abs(res - 4.213e-16) < 1e-10
    A4 = gpu.Matrix(4, 4, g["A4_rowbeg"], g["A4_colind"], g["A4_val"])
    conv, it, res, x = A4.minres(np.zeros(4), g["b4"], 10, 1e-10)
    assert conv and it == int(g["test3"][1]) and np.allclose(x, [0.25, 0.5, 0.75, 1.0], atol=1e-9)
    assert abs(res - 2.87588e-11) < 1e-10                      # TestMinres.ref "Residual3"
    conv, it, res, x = A4.pminres_dummy(np.zeros(4), g["b4"], 10, 1e-10)
    assert conv and it == int(g["test5"][1]) and np.allclose(x, [0.25, 0.5, 0.75, 1.0], atol=1e-9)
    assert abs(res - 2.87596e-11) < 1e-10                      # "Residual5"
    B = gpu.Matrix(1, 2, [0, 2], [0, 1], [1., 1.])
    assert np.array_equal(B.spmv_transpose(np.array([3.])), [3., 3.])   # transp_mul( B, b2): 3 3
    b = np.array([1., 2., 3.])
    Sb = np.concatenate([A2.spmv(b[:2]) + B.spmv_transpose(b[2:]), B.spmv(b[:2])])
    assert np.array_equal(Sb, g["block_Sb"]) and np.array_equal(Sb, [5., 4., 3.])


# ---------------------------------------------------------------------------------------------
# vector kernels
# ---------------------------------------------------------------------------------------------
def test_vector_kernels_against_reference(gpu):
    g = load_golden("vectors.npz")
    x, y, y2 = gpu.Vector.from_host(g["x"]), gpu.Vector.from_host(g["y"]), gpu.Vector.from_host(g["y2"])
    z = gpu.Vector(x.n)
    # elementwise updates: bit-identical to the reference's valarray expressions
    gpu.z_xpay(z, x, -1.25, y)
    assert np.array_equal(z.download(), g["z_xpay"])
    gpu.z_xpaypby2(z, x, 0.5, y, -3.0, y2)
    assert np.array_equal(z.download(), g["z_xpaypby2"])
    z.copy_from(y)
    gpu.axpy(0.37, x, z)
    assert np.array_equal(z.download(), g["axpy"])
    # reductions: double-double tree, compared with the exactly rounded value and with the
    # reference's Kahan sum; the tolerance is a few ulp of sum|x_i y_i|
    import math
    exact = math.fsum((g["x"] * g["y"]).tolist())
    scale = float(np.abs(g["x"] * g["y"]).sum())
    d = gpu.dot(x, y)
    assert abs(d - exact) <= 4e-16 * scale
    assert abs(d - float(g["kahan_dot"])) <= 4e-16 * scale
    assert abs(gpu.norm(x) - float(g["kahan_norm"])) <= 4e-16 * float(g["kahan_norm"])
    assert abs(gpu.norm_sq(x) - float(g["norm_sq"])) <= 1e-13 * float(g["norm_sq"])   # reference: plain sum
    assert gpu.supnorm(x) == float(g["supnorm"])
    h = gpu.Vector(1000)
    h.fill_hash()
    assert np.array_equal(h.download(), hash_vector(1000))


@pytest.mark.parametrize("spc,threads,mode", [(1, 32, 1), (2, 64, 1), (16, 128, 1), (1, 64, 2), (16, 128, 2)])
def test_fused_spmv_dot_matches_separate_dot(gpu, spc, threads, mode):
    """The dot product fused into the SpMV epilogue (flat and two-level grid reduction: > 2048
    CTAs with one slice per CTA) equals the standalone reduction of the same products."""
    A = gpu.Matrix.brick(2, 32, 1)
    n = A.rows
    x, y, y2 = gpu.Vector(n), gpu.Vector(n), gpu.Vector(n)
    x.fill_hash()
    try:
        gpu.set_option("spmv.slices_per_cta", spc)
        gpu.set_option("spmv.threads", threads)
        gpu.set_option("spmv.dotmode", mode)
        vals = {A.spmv_dot_dev(x, y, x) for _ in range(3)}
    finally:
        gpu.set_option("spmv.slices_per_cta", 16)
        gpu.set_option("spmv.threads", 128)
        gpu.set_option("spmv.dotmode", 0)
    A.spmv_dev(x, y2)
    assert np.array_equal(y.download(), y2.download())
    assert len(vals) == 1                                  # deterministic
    d = gpu.dot(y2, x)
    scale = float(np.abs(y2.download() * x.download()).sum())
    assert abs(vals.pop() - d) <= 4e-16 * scale            # two correctly-rounded dd sums


def test_reductions_are_deterministic(gpu):
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(1 << 20), rng.standard_normal(1 << 20)
    x, y = gpu.Vector.from_host(a), gpu.Vector.from_host(b)
    vals = {gpu.dot(x, y) for _ in range(5)}
    assert len(vals) == 1


# ---------------------------------------------------------------------------------------------
# Krylov solvers against the reference's outputs
# ---------------------------------------------------------------------------------------------
def check_solve(out, g, key, x_rtol=X_RTOL, res_rtol=1e-3):
    conv, it, res = bool(g[key][0]), int(g[key][1]), float(g[key][2])
    assert out[0] == conv, key
    assert abs(out[1] - it) <= 1, (key, out[1], it)
    if out[1] == it:
        # 1e-15: a residual at rounding level (e.g. exact termination on tiny systems) is noise
        assert abs(out[2] - res) <= res_rtol * abs(res) + 1e-15, (key, out[2], res)
        assert rel(out[3], g[key + "_x"]) <= x_rtol, (key, rel(out[3], g[key + "_x"]))


@pytest.mark.parametrize("path", golden_files("brick_*.npz") + golden_files("random_spd*.npz"),
                         ids=lambda p: p.split("/")[-1][:-4])
def test_solvers_against_reference(gpu, path):
    g = dict(np.load(path))
    A = gmat(gpu, g)
    b, x0 = g["b"], np.zeros(int(g["rows"]))
    tag = "nonsym" if "nonsym" in path else "sym"
    if tag == "sym":
        out = A.pcg_jacobi(x0, b, 400, 1e-10, True, 1.0, hist=True)
        check_solve(out, g, "sym_pcgjac")
        it = min(out[1], int(g["sym_pcgjac"][1]))
        ref_hist = g["sym_pcgjac_hist"]
        # atol: a residual at rounding level (exact termination on tiny systems) is noise
        assert np.allclose(out[4][1:it + 1], ref_hist[1:it + 1], rtol=1e-6, atol=1e-15)
        check_solve(A.pcg_jacobi(x0, b, 400, 1e-10, False, 0.8), g, "sym_pcgjac_w08")
        check_solve(A.pcg_dummy(x0, b, 400, 1e-10, True), g, "sym_pcgdummy")
        out = A.cg(x0, b, 400, 1e-10, True, hist=True)
        check_solve(out, g, "sym_cg")
        it = min(out[1], int(g["sym_cg"][1]))
        assert np.allclose(out[4][1:it + 1], g["sym_cg_hist"][1:it + 1], rtol=1e-6, atol=1e-15)
        check_solve(A.minres(x0, b, 400, 1e-9), g, "sym_minres", x_rtol=1e-8)
        check_solve(A.pminres_jacobi(x0, b, 400, 1e-9), g, "sym_pminresjac", x_rtol=1e-8)
        # fixed work: exactly 7 iterations, not converged, max_iter left untouched
        out = A.pcg_jacobi(x0, b, 7, 0.0, True, 1.0)
        assert out[0] is False and out[1] == 7
        assert abs(out[2] - g["sym_pcgjac_fixed7"][2]) <= 1e-9 * g["sym_pcgjac_fixed7"][2] + 1e-15
        assert rel(out[3], g["sym_pcgjac_fixed7_x"]) <= 1e-12
    check_solve(A.gmres_jacobi(x0, b, 20, 400, 1e-10, True, 1.0, True), g, f"{tag}_gmresjac_l1")
    check_solve(A.gmres_jacobi(x0, b, 20, 400, 1e-10, True, 1.0, False), g, f"{tag}_gmresjac_l0")
    check_solve(A.gmres_jacobi(x0, b, 5, 400, 1e-10, True, 1.0, True), g, f"{tag}_gmresjac_m5", x_rtol=1e-8)
    check_solve(A.gmres_dummy(x0, b, 20, 400, 1e-10, False, True), g, f"{tag}_gmresdummy")


def test_bicgstab_and_gcr_against_reference(gpu):
    """BiCGStab (krylovsolver.h:1189-1258) and GCR(m) (:1410-1483) against the compiled reference's
    results (tests/golden/krylov_extra.npz) and the reference's own known-answer systems."""
    E = load_golden("krylov_extra.npz")
    calls = {"bicgstab_jac": lambda A, x0, b: A.bicgstab(x0, b, 400, 1e-10, True, gpu.PC_JACOBI, 1.0),
             "bicgstab_dummy": lambda A, x0, b: A.bicgstab(x0, b, 400, 1e-10, True, gpu.PC_DUMMY),
             "bicgstab_jac_abs_w08": lambda A, x0, b: A.bicgstab(x0, b, 400, 1e-9, False, gpu.PC_JACOBI, 0.8),
             "gcr_jac_m100": lambda A, x0, b: A.gcr(x0, b, 100, 400, 1e-10, True, gpu.PC_JACOBI, 1.0),
             "gcr_dummy_m100": lambda A, x0, b: A.gcr(x0, b, 100, 400, 1e-10, True, gpu.PC_DUMMY)}
    for path in golden_files("brick_*.npz") + golden_files("random_spd*.npz"):
        g = dict(np.load(path))
        name = path.split("/")[-1][:-4]
        A = gmat(gpu, g)
        n = A.rows
        for key, fn in calls.items():
            out = fn(A, np.zeros(n), g["b"])
            exp, xe = E[f"{name}/{key}"], E[f"{name}/{key}_x"]
            assert out[0] == bool(exp[0]) and abs(out[1] - int(exp[1])) <= 1, (name, key, out[:3], exp)
            if out[1] == int(exp[1]):
                # BiCGStab amplifies rounding differences of its dot products (the reference sums them
                # left to right): compare the solutions at the level both reach (tol 1e-10 / 1e-9)
                assert rel(out[3], xe) <= 1e-7, (name, key, rel(out[3], xe))
        # truncated GCR: same (non-)convergence and residual level as the reference, stale vacc[] included
        out = A.gcr(np.zeros(n), g["b"], 4, 60, 1e-10, True, gpu.PC_JACOBI, 1.0)
        exp = E[f"{name}/gcr_jac_m4_trunc"]
        assert out[0] == bool(exp[0])
        if not out[0]:
            assert out[1] == 60 and 0.2 <= out[2] / exp[2] <= 5.0, (name, out[:3], exp)
    A2 = gpu.Matrix(2, 2, [0, 1, 2], [1, 0], [1., 1.])
    M = np.array([[-249., -453, -397, -52], [-453, -731, -601, -78], [-397, -601, -648, -91], [-52, -78, -91, -13]])
    A4 = gpu.Matrix(4, 4, np.arange(0, 17, 4), np.tile(np.arange(4), 4), M.reshape(-1))
    b2, b4 = np.array([1., 2.]), np.array([-1277. / 2, -2015. / 2, -3907. / 4, -533. / 4])
    out = A2.bicgstab(np.zeros(2), b2, 10, 1e-10, True, gpu.PC_DUMMY)       # bicgstab.cpp Test()
    assert out[0] and out[1] == int(E["kat/bicgstab2"][1]) and abs(out[2]) < 1e-10 and np.allclose(out[3], [2., 1.])
    out = A4.bicgstab(np.zeros(4), b4, 10, 1e-10, True, gpu.PC_DUMMY)       # Test2(): "Case 2 Residual" 0 +- 1e-10
    assert out[0] and abs(out[1] - int(E["kat/bicgstab4"][1])) <= 1 and abs(out[2]) < 1e-10
    assert np.allclose(out[3], [0.25, 0.5, 0.75, 1.0], atol=1e-5)
    out = A2.gcr(np.full(2, 3.0), b2, 10, 100, 1e-10, False, gpu.PC_DUMMY)  # gcr.cpp Test()
    assert out[0] and out[1] == int(E["kat/gcr2"][1]) and abs(out[2]) < 1e-10 and np.allclose(out[3], [2., 1.])
    out = A4.gcr(np.zeros(4), b4, 10, 100, 1e-10, False, gpu.PC_DUMMY)      # Test2()
    assert out[0] and abs(out[1] - int(E["kat/gcr4"][1])) <= 1 and abs(out[2]) < 1e-10
    assert np.allclose(out[3], [0.25, 0.5, 0.75, 1.0], atol=1e-8)
    # live against the oracle on a larger nonsymmetric system
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1, 1.1)
    rows = np.repeat(np.arange(n), np.diff(rb.astype(np.int64)))
    va = va.copy()
    va[ci.astype(np.int64) == rows] *= 1.25
    from oracle import port as oport
    P = oport.PortMatrix(n, n, rb, ci, va)
    G = gpu.Matrix(n, n, rb, ci, va)
    b = P.spmv(gpu.brick_solution_host(2, 12, 1))
    r, o = P.bicgstab(np.zeros(n), b, 500, 1e-9, True, 1), G.bicgstab(np.zeros(n), b, 500, 1e-9, True, gpu.PC_JACOBI)
    assert r[0] and o[0] and abs(o[1] - r[1]) <= 1 and rel(o[3], r[3]) <= 1e-7
    r, o = P.gcr(np.zeros(n), b, 100, 500, 1e-9, True, 1), G.gcr(np.zeros(n), b, 100, 500, 1e-9, True, gpu.PC_JACOBI)
    assert r[0] and o[0] and abs(o[1] - r[1]) <= 1 and rel(o[3], r[3]) <= 1e-8


def test_solver_edge_cases(gpu):
    g = load_golden("brick_fe2_N4_num0.npz")
    A = gmat(gpu, g)
    n = A.rows
    b = g["b"]
    # zero right-hand side: normb == 0 -> 1 (krylovsolver.h:132), converged at iteration 0
    conv, it, res, x = A.pcg_jacobi(np.zeros(n), np.zeros(n), 50, 1e-10, True)
    assert conv and it == 0 and res == 0.0 and not x.any()
    # exact initial guess
    conv, it, res, x = A.pcg_jacobi(g["xstar"], b, 50, 1e-8, True)
    assert conv and it == 0
    # max_iter = 0: not converged, max_iter untouched, x untouched
    conv, it, res, x = A.pcg_jacobi(np.zeros(n), b, 0, 1e-10, True)
    assert not conv and it == 0 and not x.any()
    # check_every = 1 and a chunk size that does not divide the iteration count
    for ce in (1, 5):
        gpu.set_option("pcg.check_every", ce)
        out = A.pcg_jacobi(np.zeros(n), b, 400, 1e-10, True)
        assert out[0] and out[1] == int(g["sym_pcgjac"][1])
        assert rel(out[3], g["sym_pcgjac_x"]) <= 1e-12
    gpu.set_option("pcg.check_every", 16)
    # device-resident entry points give the same answer as the host-buffer ones
    xv, bv = gpu.Vector(n), gpu.Vector.from_host(b)
    conv, it, res = A.pcg_dev(xv, bv, 400, 1e-10, True)
    assert conv and it == int(g["sym_pcgjac"][1]) and rel(xv.download(), g["sym_pcgjac_x"]) <= 1e-12
    # 1x1 system
    A1 = gpu.Matrix(1, 1, [0, 1], [0], [4.0])
    conv, it, res, x = A1.pcg_jacobi(np.zeros(1), np.array([2.0]), 5, 1e-12, False)
    assert conv and it == 1 and x[0] == 0.5
    # a matrix with a missing diagonal entry: Jacobi refuses loudly (the reference reads past the row)
    Z = gpu.Matrix(2, 2, [0, 1, 2], [1, 0], [1., 1.])
    with pytest.raises(gpu.DropsB200Error) as ei:
        Z.pcg_jacobi(np.zeros(2), np.ones(2), 5, 1e-10)
    assert ei.value.code == -5
    # dimension checks
    with pytest.raises(gpu.DropsB200Error):
        A.spmv_dev(gpu.Vector(n + 1), gpu.Vector(n))
    # an empty matrix and rows without entries
    E = gpu.Matrix(3, 3, [0, 0, 0, 0], [], [])
    assert np.array_equal(E.spmv(np.ones(3)), np.zeros(3))
    assert np.array_equal(E.spmv_transpose(np.ones(3)), np.zeros(3))


def test_pcg_graph_replay_matches_eager(gpu, port):
    """Chunks of PCG iterations replayed as one CUDA graph (the launch-bound case: cdrdrops' 250 k
    unknowns) give bit-identical results to eager launches, across option changes, new right-hand
    sides, value updates and device / host entry points."""
    fe, N, numbering = 1, 20, 0
    n, rb, ci, va = gpu.brick_csr_host(fe, N, numbering)
    A = gpu.Matrix(n, n, rb, ci, va)
    P = port.PortMatrix(n, n, rb, ci, va)
    xs = gpu.brick_solution_host(fe, N, numbering)
    b = P.spmv(xs)
    ref = P.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
    outs = {}
    try:
        for graph in (1, 0, 1):
            gpu.set_option("pcg.graph", graph)
            for ce in (16, 4):
                gpu.set_option("pcg.check_every", ce)
                out = A.pcg_jacobi(np.zeros(n), b, 500, 1e-10, True)
                assert out[0] and out[1] == ref[1]
                outs.setdefault("x", out[3])
                assert np.array_equal(out[3], outs["x"])            # same kernels, same order: same bits
                out = A.pcg_jacobi(np.zeros(n), 2.0 * b, 37, 0.0, True)   # fixed work, other rhs, same graph
                assert not out[0] and out[1] == 37
                outs.setdefault("x37", out[3])
                assert np.array_equal(out[3], outs["x37"])
        gpu.set_option("pcg.graph", 1)
        gpu.set_option("pcg.check_every", 16)
        xv, bv = gpu.Vector(n), gpu.Vector.from_host(b)
        for _ in range(3):                                           # replay on device-resident vectors
            xv.fill(0.0)
            conv, it, res = A.pcg_dev(xv, bv, 500, 1e-10, True)
            assert conv and it == ref[1] and np.array_equal(xv.download(), outs["x"])
        A.update_values(3.0 * va, 2)                                 # values change behind the cached graph
        out = A.pcg_jacobi(np.zeros(n), 3.0 * b, 500, 1e-10, True)
        assert out[0] and out[1] == ref[1] and rel(out[3], ref[3]) <= 1e-10
        out = A.cg(np.zeros(n), 3.0 * b, 500, 1e-10, True)           # CG shares the handle: graph is re-keyed
        assert out[0] and rel(out[3], ref[3]) <= 1e-9
    finally:
        gpu.set_option("pcg.graph", 1)
        gpu.set_option("pcg.check_every", 16)


def test_lincomb_on_device_matches_reference(gpu):
    """drops_b200_matrix_lincomb against SparseMatBaseCL::LinComb run by the compiled reference
    (tests/golden/lincomb.npz): merged pattern, values and the product bit-identical; the time-step
    system L = E/dt + theta H then solves like the reference's."""
    L = load_golden("lincomb.npz")
    n = int(L["rows"])
    A = gpu.Matrix(n, n, L["A_rowbeg"], L["A_colind"], L["A_val"])
    B = gpu.Matrix(n, n, L["B_rowbeg"], L["B_colind"], L["B_val"])
    for tag in ("dt", "neg", "zero"):
        a, b = L[f"{tag}_coef"]
        Cm = gpu.Matrix.lincomb(a, A, b, B, version=5)
        rb, ci, va = Cm.download()
        assert np.array_equal(rb, L[f"{tag}_rowbeg"]) and np.array_equal(ci, L[f"{tag}_colind"])
        assert np.array_equal(va, L[f"{tag}_val"])
        assert np.array_equal(Cm.spmv(L["x_in"]), L[f"{tag}_y"])
        assert Cm.info()["version"] == 5
    Cm = gpu.Matrix.lincomb(L["dt_coef"][0], A, L["dt_coef"][1], B)
    out = Cm.gmres_jacobi(np.zeros(n), L["b"], 50, 200, 1e-10, True, 1.0, True)
    assert out[0] == bool(L["dt_gmres"][0]) and abs(out[1] - int(L["dt_gmres"][1])) <= 1
    assert rel(out[3], L["dt_gmres_x"]) <= 1e-8
    with pytest.raises(gpu.DropsB200Error):
        gpu.Matrix.lincomb(1.0, A, 1.0, gpu.Matrix(2, 2, [0, 1, 2], [0, 1], [1., 1.]))


def test_matrix_value_update_tracks_version(gpu):
    g = load_golden("brick_fe2_N4_num1.npz")
    A = gmat(gpu, g)
    assert A.info()["version"] == 1
    y1 = A.spmv(g["x_in"])
    A.update_values(2.0 * g["val"], 2)           # SparseMatBaseCL::operator*=  bumps Version()
    assert A.info()["version"] == 2
    assert np.array_equal(A.spmv(g["x_in"]), 2.0 * y1)
    assert np.array_equal(A.spmv_transpose(g["x_in"]), 2.0 * g["y_ATx"])
    assert np.array_equal(A.diag(), 2.0 * g["diag"])


# ---------------------------------------------------------------------------------------------
# oracle-sized brick systems, checked live against oracle/drops_oracle.c
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fe,N,numbering", [(1, 24, 0), (2, 16, 1), (2, 32, 1), (2, 32, 0)])
def test_brick_parity_with_oracle(gpu, port, fe, N, numbering):
    n, rb, ci, va = gpu.brick_csr_host(fe, N, numbering)
    P = port.PortMatrix(n, n, rb, ci, va)
    # device generator == host generator, bit for bit
    G = gpu.Matrix.brick(fe, N, numbering)
    rb2, ci2, va2 = G.download()
    assert np.array_equal(rb2, rb) and np.array_equal(ci2, ci) and np.array_equal(va2, va)
    xs = gpu.brick_solution_host(fe, N, numbering)
    xv = gpu.Vector(n)
    gpu.brick_solution_device(fe, N, numbering, xv)
    assert np.array_equal(xv.download(), xs)
    # SpMV: bit-identical
    xin = hash_vector(n)
    assert np.array_equal(G.spmv(xin), P.spmv(xin))
    assert np.array_equal(G.spmv_transpose(xin), P.spmv_transpose(xin))
    # PCG-Jacobi to convergence: iteration count +-1, solution 1e-10
    b = P.spmv(xs)
    ref = P.pcg_jacobi(np.zeros(n), b, 2000, 1e-10, True, hist=True)
    out = G.pcg_jacobi(np.zeros(n), b, 2000, 1e-10, True, hist=True)
    assert ref[0] and out[0]
    assert abs(out[1] - ref[1]) <= 1, (out[1], ref[1])
    assert rel(out[3], ref[3]) <= X_RTOL
    assert rel(out[3], xs) <= 1e-8
    it = min(out[1], ref[1])
    assert np.allclose(out[4][:it + 1], ref[4][:it + 1], rtol=1e-5, atol=0)
    if N <= 16:
        r = P.gmres_jacobi(np.zeros(n), b, 30, 1000, 1e-9, True)
        o = G.gmres_jacobi(np.zeros(n), b, 30, 1000, 1e-9, True)
        assert r[0] and o[0] and abs(o[1] - r[1]) <= 1 and rel(o[3], r[3]) <= 1e-8
        r = P.minres(np.zeros(n), b, 1000, 1e-8)
        o = G.minres(np.zeros(n), b, 1000, 1e-8)
        assert r[0] and o[0] and abs(o[1] - r[1]) <= 1 and rel(o[3], r[3]) <= 1e-7


def test_nonsymmetric_gmres_parity_with_oracle(gpu, port):
    """C3 stand-in (level-set advection system L = E/dt + theta H, levelset/coupling.tpp:95-97):
    P2 pattern, nonsymmetric values (upper triangle x 1.1) plus a mass-like diagonal shift (x 1.25),
    GMRES(m) + Jacobi with and without restarts, left and right preconditioning."""
    n, rb, ci, va = gpu.brick_csr_host(2, 12, 1, 1.1)
    rows = np.repeat(np.arange(n), np.diff(rb.astype(np.int64)))
    va = va.copy()
    va[ci.astype(np.int64) == rows] *= 1.25
    P = port.PortMatrix(n, n, rb, ci, va)
    G = gpu.Matrix(n, n, rb, ci, va)
    xs = gpu.brick_solution_host(2, 12, 1)
    b = P.spmv(xs)
    assert np.array_equal(G.spmv(xs), b)
    for m in (8, 30):
        for left in (True, False):
            r = P.gmres_jacobi(np.zeros(n), b, m, 2000, 1e-9, True, 1.0, left)
            o = G.gmres_jacobi(np.zeros(n), b, m, 2000, 1e-9, True, 1.0, left)
            assert r[0] and o[0]
            assert abs(o[1] - r[1]) <= 1, (o[1], r[1])
            assert rel(o[3], r[3]) <= 1e-8


def test_irregular_interface_system_parity(gpu, port):
    """C4 stand-in (TraceFEM): keep only the P1 DOFs within 2h of the sphere |p-c| = 0.3; irregular
    row lengths, CG + Jacobi."""
    import scipy.sparse as sp
    N = 28
    n, rb, ci, va = gpu.brick_csr_host(1, N, 0)
    A = sp.csr_matrix((va, ci.astype(np.int64), rb.astype(np.int64)), shape=(n, n))
    M = N - 1
    idx = np.arange(n)
    p = np.stack([idx % M + 1, (idx // M) % M + 1, idx // (M * M) + 1], 1) / N
    keep = np.abs(np.linalg.norm(p - 0.5, axis=1) - 0.3) < 2.0 / N
    S = A[keep][:, keep].tocsr()          # stored zeros of the pattern stay (poisson.tpp:912-917)
    S.sort_indices()
    m = S.shape[0]
    S.data[S.indices == np.repeat(np.arange(m), np.diff(S.indptr))] += 1e-3
    assert len(set(np.diff(S.indptr))) > 5
    P = port.PortMatrix(m, m, S.indptr, S.indices, S.data)
    G = gpu.Matrix(m, m, S.indptr, S.indices, S.data)
    xin = hash_vector(m)
    assert np.array_equal(G.spmv(xin), P.spmv(xin))
    b = P.spmv(np.ones(m))
    r = P.pcg_jacobi(np.zeros(m), b, 1000, 1e-9, True)
    o = G.pcg_jacobi(np.zeros(m), b, 1000, 1e-9, True)
    assert r[0] and o[0] and abs(o[1] - r[1]) <= 1 and rel(o[3], r[3]) <= X_RTOL


# ---------------------------------------------------------------------------------------------
# BASELINE.json full size (C2): properties that need no CPU copy of the 17 GB matrix
# ---------------------------------------------------------------------------------------------
def test_full_size_c2_properties(gpu):
    fe, N, numbering = 2, 185, 1
    n, nnz = gpu.brick_sizes(fe, N)
    assert n == 50243409
    A = gpu.Matrix.brick(fe, N, numbering)
    info = A.info()
    assert info["nnz"] == nnz
    assert info["padded_nnz"] <= 1.02 * nnz                 # sigma-sorted slices: < 2 % padding
    x, y, z, w = (gpu.Vector(n) for _ in range(4))
    # (1) symmetry through the operator: x'Ay == y'Ax, and A^T x == A x bit for bit
    x.fill_hash()
    gpu.brick_solution_device(fe, N, numbering, y)
    A.spmv_dev(x, z)
    A.spmv_dev(y, w)
    yAx, xAy = gpu.dot(y, z), gpu.dot(x, w)
    assert abs(yAx - xAy) <= 1e-13 * gpu.norm(y) * gpu.norm(z)   # x is mean-free: y'Ax itself cancels
    A.spmv_transpose_dev(x, w)
    gpu.axpy(-1.0, z, w)
    assert gpu.supnorm(w) == 0.0
    # (2) linearity: A(x + 3y) == Ax + 3Ay to rounding
    gpu.z_xpay(w, x, 3.0, y)                                # w = x + 3y
    v = gpu.Vector(n)
    A.spmv_dev(w, v)                                        # v = A w
    A.spmv_dev(y, w)                                        # w = A y
    gpu.axpy(3.0, w, z)                                     # z = Ax + 3Ay
    gpu.axpy(-1.0, z, v)
    assert gpu.supnorm(v) <= 1e-12 * max(gpu.supnorm(z), 1.0)
    # (3) stiffness rows annihilate constants: |A 1|_inf is tiny except next to the boundary, and
    #     sum_i (A 1)_i = 1'A1 > 0 equals the boundary flux; check 1'A1 against x*-weighted identity
    x.fill(1.0)
    A.spmv_dev(x, z)
    one_A_one = gpu.dot(x, z)
    assert one_A_one > 0
    # (4) PCG: the recurrence residual equals the true residual, and the error decreases
    gpu.brick_solution_device(fe, N, numbering, y)          # y = x*
    b = gpu.Vector(n)
    A.spmv_dev(y, b)
    normb = gpu.norm(b)
    x.fill(0.0)
    conv, it, res = A.pcg_dev(x, b, 60, 0.0, True)
    assert not conv and it == 60
    A.spmv_dev(x, z)
    gpu.axpy(-1.0, b, z)                                    # z = Ax - b
    true_res = gpu.norm(z) / normb
    assert abs(true_res - res) <= 1e-8 * max(res, 1e-300) + 1e-13
    gpu.axpy(-1.0, y, x)
    assert gpu.norm(x) < gpu.norm(y)


# ---------------------------------------------------------------------------------------------
# production defaults against the compiled reference: P2 N=64 (2 048 383 rows) is the first size
# of BASELINE.md's parity set at which "sell.block_min_rows" (2e6) switches the block order, the
# row-ordered PCG, the U15/384 fused-dot kernel and the diagonal byte index on BY THEMSELVES.
# No option is touched here; the oracle is oracle/_ref (the unmodified reference headers), with
# the plain-C restatement (bit-identical to it, tests/test_oracle.py) as the stand-in on a box
# where _ref was not built.  PCG: num/krylovsolver.h:107-179.
# ---------------------------------------------------------------------------------------------
def test_production_defaults_parity_p2_n64(gpu, port):
    from oracle import ref as oref
    fe, N, numbering = 2, 64, 1
    n, rb, ci, va = gpu.brick_csr_host(fe, N, numbering)
    assert n == 2048383
    if oref.available():
        oref.set_num_threads(max(1, __import__("os").cpu_count() or 1))
        R = oref.RefMatrix(n, n, rb, ci, va)
    else:
        R = port.PortMatrix(n, n, rb, ci, va)
    G = gpu.Matrix(n, n, rb, ci, va)               # host upload path, the reference's own layout
    lay = G.layout()
    assert lay["block_ordered"] and lay["row_ordered_stream"], lay
    vd = G.value_dictionary()
    assert vd["packed"] and 0 < vd["n_distinct"] <= 47, vd
    xin = hash_vector(n)
    assert np.array_equal(G.spmv(xin), R.spmv(xin))
    assert np.array_equal(G.spmv_transpose(xin), R.spmv_transpose(xin))
    xs = gpu.brick_solution_host(fe, N, numbering)
    b = R.spmv(xs)
    ref = R.pcg_jacobi(np.zeros(n), b, 5000, 1e-10, True)
    out = G.pcg_jacobi(np.zeros(n), b, 5000, 1e-10, True)
    assert ref[0] and out[0]
    assert abs(out[1] - ref[1]) <= 1, (out[1], ref[1])
    assert rel(out[3], ref[3]) <= X_RTOL
    assert G.diag_index() > 0                      # the update kernel read d as one byte per row
    # true residual of the GPU solution, evaluated by the reference
    assert np.linalg.norm(b - R.spmv(out[3])) / np.linalg.norm(b) <= 2e-10
    # the same defaults on the device-generated operator (what bench.py times) give the same bits
    D = gpu.Matrix.brick(fe, N, numbering)
    assert D.layout()["block_ordered"]
    out2 = D.pcg_jacobi(np.zeros(n), b, 5000, 1e-10, True)
    assert out2[1] == out[1] and np.array_equal(out2[3], out[3])
    # line segments are the unit of the blocks at this size: 8 DOF classes x ~63^2 lines x 1 complete segment each
    assert lay["line_segments"] >= 6 * 61 ** 2, lay
    assert lay["segment_entries"] >= 0.4 * len(va), lay     # ... and most of them run from segment descriptors
    # footprint ("sell.release_general"): the 12-byte stream of an operator this large is gone once the packed stream
    # exists -- A: 12 (csr) + 2 x 4 (pk, pkp) bytes per entry, A^T (built by the transp_mul above): 2 x 4, instead of 32 + 32
    nnz = len(va)
    assert G.info()["device_bytes"] < 30 * nnz, (G.info()["device_bytes"], nnz)
    try:
        gpu.set_option("spmv.pk", 0)
        with pytest.raises(gpu.DropsB200Error):
            G.spmv(xin)                            # asked for the released stream: fails loudly
    finally:
        gpu.set_option("spmv.pk", 1)
    # a value update rebuilds the streams from the CSR copy (and releases the general one again)
    G.update_values(va * 2.0, 2)
    assert np.array_equal(G.spmv(xin), 2.0 * R.spmv(xin)) and np.array_equal(G.spmv_transpose(xin), 2.0 * R.spmv_transpose(xin))
    # GMRES and MINRES iterate on the same row-ordered vectors ("krylov.reorder"): iteration counts of the reference
    G.update_values(va, 3)
    og = G.gmres_jacobi(np.zeros(n), b, 30, 70, 1e-30, True, 1.0, True)
    orf = R.gmres_jacobi(np.zeros(n), b, 30, 70, 1e-30, True, 1.0, True)
    assert og[1] == orf[1] and rel(og[3], orf[3]) <= 1e-9, (og[1], orf[1], rel(og[3], orf[3]))
    og = G.minres(np.zeros(n), b, 60, 1e-30, True)
    orf = R.minres(np.zeros(n), b, 60, 1e-30, True)
    assert og[1] == orf[1] and rel(og[3], orf[3]) <= 1e-9, (og[1], orf[1], rel(og[3], orf[3]))
