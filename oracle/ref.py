"""ctypes binding to oracle/_ref/libdrops_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED DROPS reference (headers under /root/reference/src) behind the
extern "C" veneer in oracle/ref_shim.cpp; build it with `make -C oracle ref`.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may import this
module; nothing under drops_b200/ does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdrops_ref.so")

_lib = None

_szp = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise FileNotFoundError(
            f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
    L = C.CDLL(LIB_PATH)
    vp, sz, d, i = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    ip, dpp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    sig = {
        "ref_num_threads": (i, []),
        "ref_set_num_threads": (None, [i]),
        "ref_last_log": (C.c_char_p, []),
        "ref_matrix_create": (vp, [sz, sz, sz, _szp, _szp, _dp]),
        "ref_matrix_build_triplets": (vp, [sz, sz, sz, _szp, _szp, _dp]),
        "ref_matrix_destroy": (None, [vp]),
        "ref_matrix_rows": (sz, [vp]),
        "ref_matrix_cols": (sz, [vp]),
        "ref_matrix_nnz": (sz, [vp]),
        "ref_matrix_version": (sz, [vp]),
        "ref_matrix_get_csr": (None, [vp, _szp, _szp, _dp]),
        "ref_spmv": (None, [vp, _dp, _dp]),
        "ref_y_Ax_raw": (None, [vp, _dp, _dp]),
        "ref_spmv_transpose": (None, [vp, _dp, _dp]),
        "ref_get_diag": (None, [vp, _dp]),
        "ref_dot": (d, [sz, _dp, _dp]),
        "ref_norm_sq": (d, [sz, _dp]),
        "ref_norm": (d, [sz, _dp]),
        "ref_supnorm": (d, [sz, _dp]),
        "ref_kahan_dot": (d, [sz, _dp, _dp]),
        "ref_kahan_norm": (d, [sz, _dp]),
        "ref_axpy": (None, [sz, d, _dp, _dp]),
        "ref_z_xpay": (None, [sz, _dp, _dp, d, _dp]),
        "ref_z_xpaypby2": (None, [sz, _dp, _dp, d, _dp, d, _dp]),
        "ref_jacobi_apply": (None, [vp, d, _dp, _dp]),
        "ref_pcg_jacobi": (i, [vp, _dp, _dp, d, ip, dpp, i, vp, i]),
        "ref_pcg_dummy": (i, [vp, _dp, _dp, ip, dpp, i, vp, i]),
        "ref_pcg_jacobi_timed": (i, [vp, _dp, _dp, d, ip, dpp, i]),
        "ref_cg": (i, [vp, _dp, _dp, ip, dpp, i, vp, i]),
        "ref_pcg_diagpc": (i, [vp, _dp, _dp, _dp, ip, dpp, i]),
        "ref_diagpc_apply": (None, [sz, _dp, _dp, _dp]),
        "ref_gmres_jacobi": (i, [vp, _dp, _dp, d, i, ip, dpp, i, i]),
        "ref_gmres_dummy": (i, [vp, _dp, _dp, i, ip, dpp, i, i]),
        "ref_bicgstab": (i, [vp, _dp, _dp, i, d, ip, dpp, i]),
        "ref_matrix_write": (i, [vp, C.c_char_p, i, i]),
        "ref_matrix_read_text": (vp, [C.c_char_p]),
        "ref_matrix_lincomb": (vp, [d, vp, d, vp]),
        "ref_gcr": (i, [vp, _dp, _dp, i, d, i, ip, dpp, i]),
        "ref_minres": (i, [vp, _dp, _dp, ip, dpp, i]),
        "ref_pminres_jacobi": (i, [vp, _dp, _dp, d, ip, dpp, i]),
        "ref_pminres_dummy": (i, [vp, _dp, _dp, ip, dpp, i]),
        "ref_lanczos_basis": (i, [vp, _dp, i, _dp, _dp, _dp]),
        "ref_blockmat_saddle": (None, [vp, vp, _dp, _dp, _dp]),
        "ref_bbt_diag": (None, [vp, _dp]),
        "ref_max_eig_scaled": (d, [vp, i, _dp, i]),
        "ref_chebyshev_apply": (d, [vp, i, d, i, d, _dp, _dp]),
        "ref_pcg_chebyshev": (i, [vp, _dp, _dp, d, i, d, ip, dpp, i]),
        "ref_pcgne": (i, [vp, _dp, _dp, i, d, d, i, d, ip, dpp, i]),
        "ref_idrs": (i, [vp, _dp, _dp, i, d, i, d, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"), sz,
                         ip, dpp, i]),
        "ref_block_dims": (None, [vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "ref_block_mul": (None, [vp, C.POINTER(C.c_int), i, _dp, _dp]),
        "ref_composite_mul": (None, [vp, i, vp, i, i, _dp, _dp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefMatrix:
    """A DROPS::MatrixCL living inside the compiled reference."""

    def __init__(self, rows, cols, rowbeg, colind, val):
        L = lib()
        self.rows, self.cols = int(rows), int(cols)
        rowbeg, colind, val = _u64(rowbeg), _u64(colind), _f64(val)
        self.nnz = int(val.size)
        if colind.size == 0:  # ndpointer rejects nothing, but keep a valid buffer
            colind, val = np.zeros(1, np.uint64), np.zeros(1, np.float64)
        self.h = L.ref_matrix_create(self.rows, self.cols, self.nnz, rowbeg, colind, val)

    @classmethod
    def from_triplets(cls, rows, cols, ti, tj, tv):
        L = lib()
        self = cls.__new__(cls)
        self.rows, self.cols = int(rows), int(cols)
        ti, tj, tv = _u64(ti), _u64(tj), _f64(tv)
        self.h = L.ref_matrix_build_triplets(self.rows, self.cols, int(tv.size), ti, tj, tv)
        self.nnz = int(L.ref_matrix_nnz(self.h))
        return self

    @classmethod
    def lincomb(cls, a, A, b, B):
        """SparseMatBaseCL::LinComb (spmat.h:1457-1551) run by the reference."""
        L = lib()
        self = cls.__new__(cls)
        self.h = L.ref_matrix_lincomb(float(a), A.h, float(b), B.h)
        self.rows, self.cols, self.nnz = A.rows, A.cols, int(L.ref_matrix_nnz(self.h))
        return self

    def write(self, path, fmt="mtx", precision=17):
        """The reference's own writers: fmt "mtx" = operator<< (spmat.h:1213-1224), "text" = out() (:1227-1238)."""
        if lib().ref_matrix_write(self.h, str(path).encode(), 0 if fmt == "mtx" else 1, int(precision)) != 0:
            raise IOError(f"reference could not write {path}")

    @classmethod
    def read_text(cls, path):
        """The reference's in() (spmat.h:1241-1255)."""
        L = lib()
        h = L.ref_matrix_read_text(str(path).encode())
        if not h:
            raise IOError(f"reference could not read {path}")
        self = cls.__new__(cls)
        self.h = h
        self.rows, self.cols = int(L.ref_matrix_rows(h)), int(L.ref_matrix_cols(h))
        self.nnz = int(L.ref_matrix_nnz(h))
        return self

    def csr(self):
        L = lib()
        rb = np.zeros(self.rows + 1, np.uint64)
        ci = np.zeros(max(self.nnz, 1), np.uint64)
        va = np.zeros(max(self.nnz, 1), np.float64)
        L.ref_matrix_get_csr(self.h, rb, ci, va)
        return rb, ci[:self.nnz], va[:self.nnz]

    def close(self):
        if getattr(self, "h", None):
            lib().ref_matrix_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- products -------------------------------------------------------------------
    def spmv(self, x):
        y = np.zeros(self.rows)
        lib().ref_spmv(self.h, _f64(x), y)
        return y

    def spmv_raw(self, x, y):
        lib().ref_y_Ax_raw(self.h, x, y)

    def spmv_transpose(self, x):
        y = np.zeros(self.cols)
        lib().ref_spmv_transpose(self.h, _f64(x), y)
        return y

    def diag(self):
        d = np.zeros(self.rows)
        lib().ref_get_diag(self.h, d)
        return d

    def jacobi_apply(self, b, omega=1.0):
        x = np.zeros(self.rows)
        lib().ref_jacobi_apply(self.h, float(omega), _f64(b), x)
        return x

    # --- solvers: return (converged, iters, resid, x[, hist]) -------------------------
    def _solve(self, fn, x0, b, max_iter, tol, pre=(), post=(), hist=False):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        args = [self.h, x, _f64(b), *pre, C.byref(it), C.byref(res), *post]
        h = None
        if hist is not None:
            if hist:
                h = np.full(int(max_iter) + 1, np.nan)
                args += [h.ctypes.data_as(C.c_void_p), int(h.size)]
            else:
                args += [None, 0]
        conv = fn(*args)
        out = (bool(conv), it.value, res.value, x)
        return out + (h,) if h is not None else out

    def pcg_jacobi(self, x0, b, max_iter, tol, relative=False, omega=1.0, hist=False):
        return self._solve(lib().ref_pcg_jacobi, x0, b, max_iter, tol,
                           pre=(float(omega),), post=(int(relative),), hist=hist)

    def pcg_dummy(self, x0, b, max_iter, tol, relative=False, hist=False):
        return self._solve(lib().ref_pcg_dummy, x0, b, max_iter, tol,
                           post=(int(relative),), hist=hist)

    def pcg_jacobi_timed(self, x0, b, max_iter, tol, relative=True, omega=1.0):
        return self._solve(lib().ref_pcg_jacobi_timed, x0, b, max_iter, tol,
                           pre=(float(omega),), post=(int(relative),), hist=None)

    def pcg_diagpc(self, x0, b, max_iter, tol, relative, D):
        """PCGSolverCL<DiagPcCL> with DiagPcCL(D) (precond.h:618-642)."""
        return self._solve(lib().ref_pcg_diagpc, x0, b, max_iter, tol, pre=(_f64(D),), post=(int(relative),), hist=None)

    def cg(self, x0, b, max_iter, tol, relative=False, hist=False):
        return self._solve(lib().ref_cg, x0, b, max_iter, tol, post=(int(relative),), hist=hist)

    def gmres_jacobi(self, x0, b, restart, max_iter, tol, relative=True, omega=1.0, left=True):
        return self._solve(lib().ref_gmres_jacobi, x0, b, max_iter, tol,
                           pre=(float(omega), int(restart)),
                           post=(int(relative), int(left)), hist=None)

    def gmres_dummy(self, x0, b, restart, max_iter, tol, relative=True, left=True):
        return self._solve(lib().ref_gmres_dummy, x0, b, max_iter, tol,
                           pre=(int(restart),), post=(int(relative), int(left)), hist=None)

    def bicgstab(self, x0, b, max_iter, tol, relative=True, pc=1, omega=1.0):
        """BICGSTAB (krylovsolver.h:1189-1258); pc: 0 DummyPcCL, 1 JACPcCL.  (converged, iters, resid, x)"""
        return self._solve(lib().ref_bicgstab, x0, b, max_iter, tol, pre=(int(pc), float(omega)),
                           post=(int(relative),), hist=None)

    def gcr(self, x0, b, m, max_iter, tol, relative=True, pc=1, omega=1.0):
        """GCR(m) (krylovsolver.h:1410-1483)."""
        return self._solve(lib().ref_gcr, x0, b, max_iter, tol, pre=(int(pc), float(omega), int(m)),
                           post=(int(relative),), hist=None)

    def minres(self, x0, b, max_iter, tol, relative=False):
        return self._solve(lib().ref_minres, x0, b, max_iter, tol, post=(int(relative),), hist=None)

    def pminres_jacobi(self, x0, b, max_iter, tol, relative=False, omega=1.0):
        return self._solve(lib().ref_pminres_jacobi, x0, b, max_iter, tol,
                           pre=(float(omega),), post=(int(relative),), hist=None)

    def pminres_dummy(self, x0, b, max_iter, tol, relative=False):
        return self._solve(lib().ref_pminres_dummy, x0, b, max_iter, tol,
                           post=(int(relative),), hist=None)

    # --- section 8(f) callers -------------------------------------------------------------------
    def bbt_diag(self):
        d = np.zeros(self.rows)
        lib().ref_bbt_diag(self.h, d)
        return d

    def max_eig_scaled(self, D, ne=False, iters=10):
        return lib().ref_max_eig_scaled(self.h, int(ne), _f64(D), int(iters))

    def chebyshev_apply(self, b, ne=False, scale=None, degree=3, ratio=30.0):
        scale = (1.0 if ne else 1.1) if scale is None else scale
        x = np.zeros(self.rows)
        lam_printed = lib().ref_chebyshev_apply(self.h, int(ne), float(scale), int(degree), float(ratio), _f64(b), x)
        return x, lam_printed

    def pcg_chebyshev(self, x0, b, max_iter, tol, relative=False, scale=1.1, degree=3, ratio=30.0):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = lib().ref_pcg_chebyshev(self.h, x, _f64(b), float(scale), int(degree), float(ratio),
                                       C.byref(it), C.byref(res), int(relative))
        return bool(conv), it.value, res.value, x

    def pcgne(self, u0, b, max_iter, tol, relative=False, pc=0, omega=1.0, scale=1.0, degree=3, ratio=30.0):
        u = _f64(u0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = lib().ref_pcgne(self.h, u, _f64(b), int(pc), float(omega), float(scale), int(degree), float(ratio),
                               C.byref(it), C.byref(res), int(relative))
        return bool(conv), it.value, res.value, u

    def idrs(self, x0, b, max_iter, tol, draws, relative=False, s=4, omega_bound=0.7, pc=1, omega=1.0):
        """draws: the s*n ints rand() returns; the shadow vectors are shadow_from_draws(draws)."""
        x = _f64(x0).copy()
        draws = np.ascontiguousarray(draws, dtype=np.int32).reshape(-1)
        assert draws.size == s * self.rows
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        ret = lib().ref_idrs(self.h, x, _f64(b), int(pc), float(omega), int(s), float(omega_bound), draws, draws.size,
                             C.byref(it), C.byref(res), int(relative))
        return ret, it.value, res.value, x

    def lanczos_basis(self, r0, max_vecs=8):
        n = self.rows
        Q = np.zeros((max_vecs, n))
        a, b = np.zeros(max_vecs), np.zeros(max_vecs)
        k = lib().ref_lanczos_basis(self.h, _f64(r0), max_vecs, Q.reshape(-1), a, b)
        return Q[:k], a[:k], b[:k]


def dot(x, y):
    return lib().ref_dot(len(x), _f64(x), _f64(y))


def norm(x):
    return lib().ref_norm(len(x), _f64(x))


def norm_sq(x):
    return lib().ref_norm_sq(len(x), _f64(x))


def supnorm(x):
    return lib().ref_supnorm(len(x), _f64(x))


def kahan_dot(x, y):
    return lib().ref_kahan_dot(len(x), _f64(x), _f64(y))


def kahan_norm(x):
    return lib().ref_kahan_norm(len(x), _f64(x))


def diagpc_apply(D, b):
    x = np.zeros(len(b))
    lib().ref_diagpc_apply(len(b), _f64(D), _f64(b), x)
    return x


def axpy(a, x, y):
    y = _f64(y).copy()
    lib().ref_axpy(len(x), float(a), _f64(x), y)
    return y


def z_xpay(x, a, y):
    z = np.zeros(len(x))
    lib().ref_z_xpay(len(x), z, _f64(x), float(a), _f64(y))
    return z


def z_xpaypby2(x, a, y, b, y2):
    z = np.zeros(len(x))
    lib().ref_z_xpaypby2(len(x), z, _f64(x), float(a), _f64(y), float(b), _f64(y2))
    return z


def blockmat_saddle(A, B, b):
    n = A.rows + B.rows
    Sb, STb = np.zeros(n), np.zeros(n)
    lib().ref_blockmat_saddle(A.h, B.h, _f64(b), Sb, STb)
    return Sb, STb


RAND_MAX = 2147483647


def shadow_from_draws(draws, s):
    """IDRS :1613: P[i][j] = (double)(2.0*rand())/RAND_MAX - 1.0 for the s*n draws in call order."""
    r = np.asarray(draws, dtype=np.float64).reshape(s, -1)
    return (2.0 * r) / float(RAND_MAX) - 1.0


def _handles(blocks):
    return (C.c_void_p * 4)(*[b.h if b is not None else None for b in blocks])


def block_mul(blocks, ops, x, transposed=False):
    """BlockMatrixCL(A, opA, B, opB, C, opC, D, opD) * x or transp_mul, run by the reference."""
    rows, cols = C.c_size_t(), C.c_size_t()
    hb, ob = _handles(blocks), (C.c_int * 4)(*ops)
    lib().ref_block_dims(C.cast(hb, C.c_void_p), ob, C.byref(rows), C.byref(cols))
    y = np.zeros(cols.value if transposed else rows.value)
    x = _f64(x)
    assert x.size == (rows.value if transposed else cols.value)
    lib().ref_block_mul(C.cast(hb, C.c_void_p), ob, int(transposed), x, y)
    return y


def composite_mul(A0, op0, A1, op1, x, transposed=False):
    n_out = (A1.rows if op1 == 0 else A1.cols) if not transposed else (A0.cols if op0 == 0 else A0.rows)
    n_in = (A0.cols if op0 == 0 else A0.rows) if not transposed else (A1.rows if op1 == 0 else A1.cols)
    assert len(x) == n_in, (len(x), n_in)
    y = np.zeros(n_out)
    lib().ref_composite_mul(A0.h, int(op0), A1.h, int(op1), int(transposed), _f64(x), y)
    return y


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(t):
    lib().ref_set_num_threads(int(t))


def last_log():
    return lib().ref_last_log().decode()
