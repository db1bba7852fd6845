"""ctypes binding to oracle/liboracle_port.so (oracle/drops_oracle.c) -- TEST INFRASTRUCTURE ONLY.

The plain-C restatement of the reference hot path.  Same calling surface as oracle/ref.py so
tests can run either checker.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_port.so")

_lib = None
_szp = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("drops_oracle.c", "drops_oracle_ext.c")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    sz, d, i = C.c_size_t, C.c_double, C.c_int
    ip, dpp, vp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p
    csr = [_dp, _szp, _szp]
    sig = {
        "orc_y_Ax": (None, [_dp, sz] + csr + [_dp]),
        "orc_y_ATx": (None, [_dp, sz, sz] + csr + [_dp]),
        "orc_get_diag": (None, [_dp, sz] + csr),
        "orc_dot": (d, [sz, _dp, _dp]),
        "orc_norm_sq": (d, [sz, _dp]),
        "orc_norm": (d, [sz, _dp]),
        "orc_supnorm": (d, [sz, _dp]),
        "orc_kahan_dot": (d, [sz, _dp, _dp]),
        "orc_kahan_norm": (d, [sz, _dp]),
        "orc_axpy": (None, [sz, d, _dp, _dp]),
        "orc_z_xpay": (None, [sz, _dp, _dp, d, _dp]),
        "orc_z_xpaypby2": (None, [sz, _dp, _dp, d, _dp, d, _dp]),
        "orc_jacobi0": (None, [sz] + csr + [d, _dp, _dp]),
        "orc_cg": (i, [sz] + csr + [_dp, _dp, ip, dpp, i, vp, i]),
        "orc_pcg": (i, [sz] + csr + [i, d, _dp, _dp, ip, dpp, i, vp, i]),
        "orc_gmres": (i, [sz] + csr + [i, d, _dp, _dp, i, ip, dpp, i, i]),
        "orc_bicgstab": (i, [sz] + csr + [i, d, _dp, _dp, ip, dpp, i]),
        "orc_lincomb": (None, [sz, d] + csr + [d] + csr + [_szp, C.c_void_p, C.c_void_p]),
        "orc_gcr": (i, [sz] + csr + [i, d, _dp, _dp, i, ip, dpp, i]),
        "orc_minres": (i, [sz] + csr + [i, i, d, _dp, _dp, ip, dpp, i]),
        "orc_lanczos_basis": (i, [sz] + csr + [_dp, i, _dp, _dp, _dp]),
        # drops_oracle_ext.c
        "orc_bbt_diag": (None, [sz, _dp, _szp, _dp]),
        "orc_max_eig_scaled": (d, [sz, sz] + csr + [i, _dp, i]),
        "orc_chebyshev": (None, [sz, sz] + csr + [i, _dp, _dp, _dp, i, d, d]),
        "orc_pcg_ext": (i, [sz, sz] + csr + [i, i, d, d, i, d, _dp, _dp, ip, dpp, i, vp, i, dpp]),
        "orc_pcg_diagpc": (i, [sz] + csr + [_dp, _dp, _dp, ip, dpp, i, vp, i]),
        "orc_idrs": (i, [sz] + csr + [i, d, i, d, _dp, _dp, _dp, ip, dpp, i]),
        "orc_block_mul": (None, [vp, vp, vp, _szp, _szp, C.POINTER(C.c_int), _szp, _szp, _dp, _dp]),
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


class PortMatrix:
    """CSR triple in the reference layout (size_t rowbeg/colind, double val)."""

    def __init__(self, rows, cols, rowbeg, colind, val):
        self.rows, self.cols = int(rows), int(cols)
        self.rowbeg, self.colind, self.val = _u64(rowbeg), _u64(colind), _f64(val)
        self.nnz = int(self.val.size)
        if self.nnz == 0:
            self.colind, self.val = np.zeros(1, np.uint64), np.zeros(1, np.float64)

    @property
    def _csr(self):
        return (self.val, self.rowbeg, self.colind)

    def csr(self):
        return self.rowbeg, self.colind[:self.nnz], self.val[:self.nnz]

    def spmv(self, x):
        y = np.zeros(self.rows)
        lib().orc_y_Ax(y, self.rows, *self._csr, _f64(x))
        return y

    def spmv_raw(self, x, y):
        lib().orc_y_Ax(y, self.rows, *self._csr, x)

    def spmv_transpose(self, x):
        y = np.zeros(self.cols)
        lib().orc_y_ATx(y, self.rows, self.cols, *self._csr, _f64(x))
        return y

    def diag(self):
        d = np.zeros(self.rows)
        lib().orc_get_diag(d, self.rows, *self._csr)
        return d

    def jacobi_apply(self, b, omega=1.0):
        x = np.zeros(self.rows)
        lib().orc_jacobi0(self.rows, *self._csr, float(omega), _f64(b), x)
        return x

    @staticmethod
    def _hist(max_iter, hist):
        if hist:
            h = np.full(int(max_iter) + 1, np.nan)
            return h, [h.ctypes.data_as(C.c_void_p), int(h.size)]
        return None, [None, 0]

    def _pcg(self, pc_kind, omega, x0, b, max_iter, tol, relative, hist):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        h, hargs = self._hist(max_iter, hist)
        conv = lib().orc_pcg(self.rows, *self._csr, pc_kind, float(omega), x, _f64(b),
                             C.byref(it), C.byref(res), int(relative), *hargs)
        out = (bool(conv), it.value, res.value, x)
        return out + (h,) if hist else out

    def pcg_jacobi(self, x0, b, max_iter, tol, relative=False, omega=1.0, hist=False):
        return self._pcg(1, omega, x0, b, max_iter, tol, relative, hist)

    def pcg_dummy(self, x0, b, max_iter, tol, relative=False, hist=False):
        return self._pcg(0, 1.0, x0, b, max_iter, tol, relative, hist)

    def pcg_jacobi_timed(self, x0, b, max_iter, tol, relative=True, omega=1.0):
        return self._pcg(1, omega, x0, b, max_iter, tol, relative, False)

    def cg(self, x0, b, max_iter, tol, relative=False, hist=False):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        h, hargs = self._hist(max_iter, hist)
        conv = lib().orc_cg(self.rows, *self._csr, x, _f64(b), C.byref(it), C.byref(res),
                            int(relative), *hargs)
        out = (bool(conv), it.value, res.value, x)
        return out + (h,) if hist else out

    def _gmres(self, pc_kind, omega, x0, b, restart, max_iter, tol, relative, left):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = lib().orc_gmres(self.rows, *self._csr, pc_kind, float(omega), x, _f64(b),
                               int(restart), C.byref(it), C.byref(res), int(relative), int(left))
        if conv < 0:
            raise ValueError("restart < 2")
        return bool(conv), it.value, res.value, x

    def gmres_jacobi(self, x0, b, restart, max_iter, tol, relative=True, omega=1.0, left=True):
        return self._gmres(1, omega, x0, b, restart, max_iter, tol, relative, left)

    def gmres_dummy(self, x0, b, restart, max_iter, tol, relative=True, left=True):
        return self._gmres(0, 1.0, x0, b, restart, max_iter, tol, relative, left)

    def bicgstab(self, x0, b, max_iter, tol, relative=True, pc=1, omega=1.0):
        """(converged, iters, resid, x); a breakdown reports converged = False like the reference."""
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = lib().orc_bicgstab(self.rows, *self._csr, int(pc), float(omega), x, _f64(b),
                                  C.byref(it), C.byref(res), int(relative))
        return conv == 1, it.value, res.value, x

    def gcr(self, x0, b, m, max_iter, tol, relative=True, pc=1, omega=1.0):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = lib().orc_gcr(self.rows, *self._csr, int(pc), float(omega), x, _f64(b), int(m),
                             C.byref(it), C.byref(res), int(relative))
        return conv == 1, it.value, res.value, x

    def _minres(self, precond, pc_kind, omega, x0, b, max_iter, tol, relative):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = lib().orc_minres(self.rows, *self._csr, precond, pc_kind, float(omega), x, _f64(b),
                                C.byref(it), C.byref(res), int(relative))
        return bool(conv), it.value, res.value, x

    def minres(self, x0, b, max_iter, tol, relative=False):
        return self._minres(0, 0, 1.0, x0, b, max_iter, tol, relative)

    def pminres_jacobi(self, x0, b, max_iter, tol, relative=False, omega=1.0):
        return self._minres(1, 1, omega, x0, b, max_iter, tol, relative)

    def pminres_dummy(self, x0, b, max_iter, tol, relative=False):
        return self._minres(1, 0, 1.0, x0, b, max_iter, tol, relative)

    # --- drops_oracle_ext.c: PCGNE, Chebyshev, IDR(s) ------------------------------------------
    def bbt_diag(self):
        d = np.zeros(self.rows)
        lib().orc_bbt_diag(self.rows, self.val, self.rowbeg, d)
        return d

    def max_eig_scaled(self, D, ne=False, iters=10):
        return lib().orc_max_eig_scaled(self.rows, self.cols, *self._csr, int(ne), _f64(D), int(iters))

    def chebyshev_apply(self, b, ne=False, scale=None, degree=3, ratio=30.0):
        """(x, lambda_max): SetDiag + Apply of ChebyshevPcCL (ne False) / ChebyshevBBTPcCL (ne True)."""
        scale = (1.0 if ne else 1.1) if scale is None else scale
        D = self.bbt_diag() if ne else self.diag()
        lmax = scale * self.max_eig_scaled(D, ne)
        x = np.zeros(self.rows)
        lib().orc_chebyshev(self.rows, self.cols, *self._csr, int(ne), x, _f64(b), D, int(degree), float(lmax), float(ratio))
        return x, lmax

    def _pcg_ext(self, ne, pc, omega, scale, degree, ratio, x0, b, max_iter, tol, relative, hist):
        x = _f64(x0).copy()
        it, res, lam = C.c_int(int(max_iter)), C.c_double(float(tol)), C.c_double(0.0)
        h, hargs = self._hist(max_iter, hist)
        conv = lib().orc_pcg_ext(self.rows, self.cols, *self._csr, int(ne), int(pc), float(omega), float(scale),
                                 int(degree), float(ratio), x, _f64(b), C.byref(it), C.byref(res), int(relative),
                                 *hargs, C.byref(lam))
        out = (bool(conv), it.value, res.value, x)
        return out + (h,) if hist else out

    def pcg_chebyshev(self, x0, b, max_iter, tol, relative=False, scale=1.1, degree=3, ratio=30.0, hist=False):
        return self._pcg_ext(0, 2, 1.0, scale, degree, ratio, x0, b, max_iter, tol, relative, hist)

    def pcg_diagpc(self, x0, b, max_iter, tol, relative, D, hist=False):
        """PCGSolverCL<DiagPcCL> with DiagPcCL(D) (precond.h:618-642)."""
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        h, hargs = self._hist(max_iter, hist)
        conv = lib().orc_pcg_diagpc(self.rows, *self._csr, _f64(D), x, _f64(b), C.byref(it), C.byref(res), int(relative), *hargs)
        out = (bool(conv), it.value, res.value, x)
        return out + (h,) if hist else out

    def pcgne(self, u0, b, max_iter, tol, relative=False, pc=0, omega=1.0, scale=1.0, degree=3, ratio=30.0, hist=False):
        """pc 0 identity, 1 Jacobi on BBTDiag, 2 ChebyshevBBTPcCL."""
        return self._pcg_ext(1, pc, omega, scale, degree, ratio, u0, b, max_iter, tol, relative, hist)

    def idrs(self, x0, b, max_iter, tol, P, relative=False, s=4, omega_bound=0.7, pc=1, omega=1.0):
        """(ret, iters, resid, x); ret is the reference's return value (see krylovsolver.h:1689-1696), -1 = it threw."""
        x = _f64(x0).copy()
        P = _f64(P).reshape(-1)
        assert P.size == s * self.rows
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        ret = lib().orc_idrs(self.rows, *self._csr, int(pc), float(omega), int(s), float(omega_bound), P, x, _f64(b),
                             C.byref(it), C.byref(res), int(relative))
        return ret, it.value, res.value, x

    def lanczos_basis(self, r0, max_vecs=8):
        n = self.rows
        Q = np.zeros((max_vecs, n))
        a, b = np.zeros(max_vecs), np.zeros(max_vecs)
        k = lib().orc_lanczos_basis(n, *self._csr, _f64(r0), max_vecs, Q.reshape(-1), a, b)
        return Q[:k], a[:k], b[:k]


def dot(x, y):
    return lib().orc_dot(len(x), _f64(x), _f64(y))


def norm(x):
    return lib().orc_norm(len(x), _f64(x))


def norm_sq(x):
    return lib().orc_norm_sq(len(x), _f64(x))


def supnorm(x):
    return lib().orc_supnorm(len(x), _f64(x))


def kahan_dot(x, y):
    return lib().orc_kahan_dot(len(x), _f64(x), _f64(y))


def kahan_norm(x):
    return lib().orc_kahan_norm(len(x), _f64(x))


def axpy(a, x, y):
    y = _f64(y).copy()
    lib().orc_axpy(len(x), float(a), _f64(x), y)
    return y


def z_xpay(x, a, y):
    z = np.zeros(len(x))
    lib().orc_z_xpay(len(x), z, _f64(x), float(a), _f64(y))
    return z


def z_xpaypby2(x, a, y, b, y2):
    z = np.zeros(len(x))
    lib().orc_z_xpaypby2(len(x), z, _f64(x), float(a), _f64(y), float(b), _f64(y2))
    return z


def lincomb(rows, a, A_csr, b, B_csr):
    """a*A + b*B (SparseMatBaseCL::LinComb, spmat.h:1457-1551); *_csr = (rowbeg, colind, val) -> same triple."""
    (ra, ca, va), (rb_, cb, vb) = [( _u64(t[0]), _u64(t[1]), _f64(t[2])) for t in (A_csr, B_csr)]
    L = lib()
    cr = np.zeros(rows + 1, np.uint64)
    L.orc_lincomb(rows, float(a), va, ra, ca, float(b), vb, rb_, cb, cr, None, None)
    nnz = int(cr[-1])
    cc, cv = np.zeros(max(nnz, 1), np.uint64), np.zeros(max(nnz, 1))
    L.orc_lincomb(rows, float(a), va, ra, ca, float(b), vb, rb_, cb, cr, cc.ctypes.data_as(C.c_void_p),
                  cv.ctypes.data_as(C.c_void_p))
    return cr, cc[:nnz], cv[:nnz]


def block_dims(blocks, ops):
    """num_rows()/num_cols() bookkeeping of BlockMatrixBaseCL (spblockmat.h:112-190): per block row / column."""
    def orows(b, o):
        return b.rows if o == 0 else b.cols

    def ocols(b, o):
        return b.cols if o == 0 else b.rows
    nr = [None, None]
    nc = [None, None]
    for br in range(2):
        for b in (2 * br, 2 * br + 1):
            if blocks[b] is not None and nr[br] is None:
                nr[br] = orows(blocks[b], ops[b])
    for bc in range(2):
        for b in (bc, bc + 2):
            if blocks[b] is not None and nc[bc] is None:
                nc[bc] = ocols(blocks[b], ops[b])
    return nr, nc


def block_mul(blocks, ops, x, transposed=False):
    """BlockMatrixBaseCL * x (spblockmat.h:225-267); blocks = [A, B, C, D] PortMatrix or None, ops 0 MUL / 1 TRANSP_MUL.
    transposed: transp_mul = GetTranspose()*x (:194-201, :269-273)."""
    if transposed:
        blocks = [blocks[0], blocks[2], blocks[1], blocks[3]]
        ops = [1 - ops[0], 1 - ops[2], 1 - ops[1], 1 - ops[3]]
    nr, nc = block_dims(blocks, ops)
    PD, PS = C.POINTER(C.c_double), C.POINTER(C.c_size_t)
    val, row, col = (PD * 4)(), (PS * 4)(), (PS * 4)()
    brows, bcols = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
    for k, b in enumerate(blocks):
        if b is None:
            continue
        val[k] = b.val.ctypes.data_as(PD)
        row[k] = b.rowbeg.ctypes.data_as(PS)
        col[k] = b.colind.ctypes.data_as(PS)
        brows[k], bcols[k] = b.rows, b.cols
    y = np.zeros(sum(nr))
    lib().orc_block_mul(C.cast(val, C.c_void_p), C.cast(row, C.c_void_p), C.cast(col, C.c_void_p), brows, bcols,
                        (C.c_int * 4)(*ops), np.array(nr, np.uint64), np.array(nc, np.uint64), _f64(x), y)
    return y


def composite_mul(A0, op0, A1, op1, x, transposed=False):
    """CompositeMatrixBaseCL * x = op1(A1) * (op0(A0) * x) (spblockmat.h:376-397); transposed: GetTranspose()*x."""
    if transposed:
        A0, op0, A1, op1 = A1, 1 - op1, A0, 1 - op0
    t = A0.spmv(x) if op0 == 0 else A0.spmv_transpose(x)
    return A1.spmv(t) if op1 == 0 else A1.spmv_transpose(t)
