"""ctypes view of include/drops_b200.h -- the only way Python reaches the CUDA path.

Used by tests/, bench.py and __graft_entry__.py.  Host vectors are numpy float64 arrays, host
matrices are CSR triples in the reference layout (uint64 rowbeg/colind, float64 val).  There is
no CPU fallback: a missing library or a missing GPU raises DropsB200Error.
"""
import ctypes as C
import os

import numpy as np

from . import _build

PC_DUMMY, PC_JACOBI = 0, 1

_lib = None


class DropsB200Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libdrops_b200 error {code}: {text}")
        self.code = code


_szp = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

_vp, _sz, _d, _i = C.c_void_p, C.c_size_t, C.c_double, C.c_int
_ip, _dpp, _szpp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_size_t)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/drops_b200.h one to one
SIGNATURES = {
    "drops_b200_init": (_i, [_i]),
    "drops_b200_shutdown": (_i, []),
    "drops_b200_last_error": (C.c_char_p, []),
    "drops_b200_abi_version": (_i, []),
    "drops_b200_compiled_arch": (_i, []),
    "drops_b200_device_info": (_i, [_ip, _ip, _szpp, _szpp]),
    "drops_b200_set_option": (_i, [C.c_char_p, _d]),
    "drops_b200_get_option": (_i, [C.c_char_p, _dpp]),
    "drops_b200_timing_get": (_i, [C.c_char_p, _dpp, C.POINTER(C.c_longlong)]),
    "drops_b200_timing_reset": (_i, []),
    "drops_b200_timer_start": (_i, []),
    "drops_b200_timer_stop": (_i, [_dpp]),
    "drops_b200_launch_count": (C.c_longlong, []),
    "drops_b200_synchronize": (_i, []),
    "drops_b200_matrix_create": (_i, [_sz, _sz, _sz, _szp, _szp, _dp, C.c_uint64, _vpp]),
    "drops_b200_matrix_create_device": (_i, [_sz, _sz, _sz, _vp, _vp, _vp, C.c_uint64, _vpp]),
    "drops_b200_matrix_update_values": (_i, [_vp, _dp, C.c_uint64]),
    "drops_b200_matrix_destroy": (_i, [_vp]),
    "drops_b200_matrix_info": (_i, [_vp, _szpp, _szpp, _szpp, C.POINTER(C.c_uint64), _szpp, _szpp]),
    "drops_b200_matrix_download": (_i, [_vp, _szp, _szp, _dp]),
    "drops_b200_spmv": (_i, [_vp, _dp, _dp]),
    "drops_b200_spmv_transpose": (_i, [_vp, _dp, _dp]),
    "drops_b200_matrix_diag": (_i, [_vp, _dp]),
    "drops_b200_vector_create": (_i, [_sz, _vpp]),
    "drops_b200_vector_destroy": (_i, [_vp]),
    "drops_b200_vector_size": (_i, [_vp, _szpp]),
    "drops_b200_vector_upload": (_i, [_vp, _dp]),
    "drops_b200_vector_download": (_i, [_vp, _dp]),
    "drops_b200_vector_fill": (_i, [_vp, _d]),
    "drops_b200_vector_copy": (_i, [_vp, _vp]),
    "drops_b200_vector_fill_hash": (_i, [_vp]),
    "drops_b200_vector_ptr": (_i, [_vp, _vpp]),
    "drops_b200_spmv_dev": (_i, [_vp, _vp, _vp]),
    "drops_b200_spmv_transpose_dev": (_i, [_vp, _vp, _vp]),
    "drops_b200_spmv_dot_dev": (_i, [_vp, _vp, _vp, _vp, _dpp]),
    "drops_b200_dot": (_i, [_vp, _vp, _dpp]),
    "drops_b200_norm_sq": (_i, [_vp, _dpp]),
    "drops_b200_norm": (_i, [_vp, _dpp]),
    "drops_b200_supnorm": (_i, [_vp, _dpp]),
    "drops_b200_axpy": (_i, [_d, _vp, _vp]),
    "drops_b200_z_xpay": (_i, [_vp, _vp, _d, _vp]),
    "drops_b200_z_xpaypby2": (_i, [_vp, _vp, _d, _vp, _d, _vp]),
    "drops_b200_jacobi_apply": (_i, [_vp, _d, _vp, _vp]),
    "drops_b200_pcg": (_i, [_vp, _dp, _dp, _i, _d, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_pcg_dev": (_i, [_vp, _vp, _vp, _i, _d, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_cg": (_i, [_vp, _dp, _dp, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_cg_dev": (_i, [_vp, _vp, _vp, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_gmres": (_i, [_vp, _dp, _dp, _i, _d, _i, _ip, _dpp, _i, _i]),
    "drops_b200_gmres_dev": (_i, [_vp, _vp, _vp, _i, _d, _i, _ip, _dpp, _i, _i]),
    "drops_b200_minres": (_i, [_vp, _dp, _dp, _i, _i, _d, _ip, _dpp, _i]),
    "drops_b200_minres_dev": (_i, [_vp, _vp, _vp, _i, _i, _d, _ip, _dpp, _i]),
    "drops_b200_brick_sizes": (_i, [_i, _i, _szpp, _szpp]),
    "drops_b200_brick_csr_host": (_i, [_i, _i, _i, _d, _szp, _szp, _dp]),
    "drops_b200_brick_matrix_device": (_i, [_i, _i, _i, _d, _vpp]),
    "drops_b200_brick_solution_host": (_i, [_i, _i, _i, _dp]),
    "drops_b200_brick_solution_device": (_i, [_i, _i, _i, _vp]),
}


def lib_path():
    return _build.LIB_PATH


def lib():
    """Load libdrops_b200.so (must have been built: __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB_PATH):
        raise DropsB200Error(-100, f"{_build.LIB_PATH} is missing: run `python -m drops_b200._build` "
                                   "(the product has no CPU fallback)")
    L = C.CDLL(_build.LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(code):
    if code < 0:
        raise DropsB200Error(code, lib().drops_b200_last_error().decode(errors="replace"))
    return code


_inited_device = None


def init(device=0):
    global _inited_device
    check(lib().drops_b200_init(int(device)))
    _inited_device = int(device)


def shutdown():
    global _inited_device
    check(lib().drops_b200_shutdown())
    _inited_device = None


def set_option(key, value):
    check(lib().drops_b200_set_option(key.encode(), float(value)))


def timing(key):
    ms, cnt = C.c_double(), C.c_longlong()
    check(lib().drops_b200_timing_get(key.encode(), C.byref(ms), C.byref(cnt)))
    return ms.value, cnt.value


def timing_reset():
    check(lib().drops_b200_timing_reset())


def timer_start():
    check(lib().drops_b200_timer_start())


def timer_stop():
    ms = C.c_double()
    check(lib().drops_b200_timer_stop(C.byref(ms)))
    return ms.value


def launch_count():
    return int(lib().drops_b200_launch_count())


def synchronize():
    check(lib().drops_b200_synchronize())


def device_info():
    sm, cc = C.c_int(), C.c_int()
    fr, to = C.c_size_t(), C.c_size_t()
    check(lib().drops_b200_device_info(C.byref(sm), C.byref(cc), C.byref(fr), C.byref(to)))
    return {"sm_count": sm.value, "cc": cc.value, "free_bytes": fr.value, "total_bytes": to.value}


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _nonempty(a, dtype):
    return a if a.size else np.zeros(1, dtype)


class Vector:
    """Device-resident VectorCL."""

    def __init__(self, n, data=None):
        self.n = int(n)
        h = C.c_void_p()
        check(lib().drops_b200_vector_create(self.n, C.byref(h)))
        self.h = h
        if data is not None:
            self.upload(data)

    @classmethod
    def from_host(cls, a):
        a = _f64(a)
        return cls(a.size, a)

    def upload(self, a):
        a = _f64(a)
        assert a.size == self.n
        check(lib().drops_b200_vector_upload(self.h, _nonempty(a, np.float64)))

    def download(self):
        out = np.zeros(max(self.n, 1))
        check(lib().drops_b200_vector_download(self.h, out))
        return out[:self.n]

    def fill(self, v):
        check(lib().drops_b200_vector_fill(self.h, float(v)))

    def fill_hash(self):
        check(lib().drops_b200_vector_fill_hash(self.h))

    def copy_from(self, other):
        check(lib().drops_b200_vector_copy(self.h, other.h))

    def close(self):
        if getattr(self, "h", None):
            lib().drops_b200_vector_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dot(x, y):
    r = C.c_double()
    check(lib().drops_b200_dot(x.h, y.h, C.byref(r)))
    return r.value


def norm(x):
    r = C.c_double()
    check(lib().drops_b200_norm(x.h, C.byref(r)))
    return r.value


def norm_sq(x):
    r = C.c_double()
    check(lib().drops_b200_norm_sq(x.h, C.byref(r)))
    return r.value


def supnorm(x):
    r = C.c_double()
    check(lib().drops_b200_supnorm(x.h, C.byref(r)))
    return r.value


def axpy(a, x, y):
    check(lib().drops_b200_axpy(float(a), x.h, y.h))


def z_xpay(z, x, a, y):
    check(lib().drops_b200_z_xpay(z.h, x.h, float(a), y.h))


def z_xpaypby2(z, x, a, y, b, y2):
    check(lib().drops_b200_z_xpaypby2(z.h, x.h, float(a), y.h, float(b), y2.h))


class Matrix:
    """Device-resident MatrixCL (CSR + sliced-ELL operator + diagonal)."""

    def __init__(self, rows, cols, rowbeg, colind, val, version=1):
        rowbeg, colind, val = _u64(rowbeg), _u64(colind), _f64(val)
        h = C.c_void_p()
        check(lib().drops_b200_matrix_create(int(rows), int(cols), int(val.size), rowbeg,
                                             _nonempty(colind, np.uint64), _nonempty(val, np.float64),
                                             int(version), C.byref(h)))
        self.h = h
        self.rows, self.cols, self.nnz = int(rows), int(cols), int(val.size)

    @classmethod
    def _wrap(cls, h):
        self = cls.__new__(cls)
        self.h = h
        info = self.info()
        self.rows, self.cols, self.nnz = info["rows"], info["cols"], info["nnz"]
        return self

    @classmethod
    def brick(cls, fe, N, numbering=0, upper_scale=1.0):
        """Generate the brick system directly on the device."""
        h = C.c_void_p()
        check(lib().drops_b200_brick_matrix_device(int(fe), int(N), int(numbering), float(upper_scale), C.byref(h)))
        return cls._wrap(h)

    def info(self):
        r, c, z, db, pn = (C.c_size_t() for _ in range(5))
        v = C.c_uint64()
        check(lib().drops_b200_matrix_info(self.h, C.byref(r), C.byref(c), C.byref(z), C.byref(v),
                                           C.byref(db), C.byref(pn)))
        return {"rows": r.value, "cols": c.value, "nnz": z.value, "version": v.value,
                "device_bytes": db.value, "padded_nnz": pn.value}

    def download(self):
        rb = np.zeros(self.rows + 1, np.uint64)
        ci = np.zeros(max(self.nnz, 1), np.uint64)
        va = np.zeros(max(self.nnz, 1))
        check(lib().drops_b200_matrix_download(self.h, rb, ci, va))
        return rb, ci[:self.nnz], va[:self.nnz]

    def update_values(self, val, version):
        val = _f64(val)
        assert val.size == self.nnz
        check(lib().drops_b200_matrix_update_values(self.h, _nonempty(val, np.float64), int(version)))

    def close(self):
        if getattr(self, "h", None):
            lib().drops_b200_matrix_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- host-buffer products (copies inside the call) ---------------------------------
    def spmv(self, x):
        x = _f64(x)
        assert x.size == self.cols
        y = np.zeros(max(self.rows, 1))
        check(lib().drops_b200_spmv(self.h, _nonempty(x, np.float64), y))
        return y[:self.rows]

    def spmv_transpose(self, x):
        x = _f64(x)
        assert x.size == self.rows
        y = np.zeros(max(self.cols, 1))
        check(lib().drops_b200_spmv_transpose(self.h, _nonempty(x, np.float64), y))
        return y[:self.cols]

    def diag(self):
        d = np.zeros(max(self.rows, 1))
        check(lib().drops_b200_matrix_diag(self.h, d))
        return d[:self.rows]

    def jacobi_apply(self, b, omega=1.0):
        bv, xv = Vector.from_host(b), Vector(self.rows)
        check(lib().drops_b200_jacobi_apply(self.h, float(omega), bv.h, xv.h))
        return xv.download()

    # --- device-vector products ----------------------------------------------------------
    def spmv_dev(self, x, y):
        check(lib().drops_b200_spmv_dev(self.h, x.h, y.h))

    def spmv_dot_dev(self, x, y, w):
        """y = A x fused with dot(y, w); returns the dot product."""
        r = C.c_double()
        check(lib().drops_b200_spmv_dot_dev(self.h, x.h, y.h, w.h, C.byref(r)))
        return r.value

    def spmv_transpose_dev(self, x, y):
        check(lib().drops_b200_spmv_transpose_dev(self.h, x.h, y.h))

    # --- solvers: (converged, iters, resid, x[, hist]) -----------------------------------
    @staticmethod
    def _hist(max_iter, hist):
        if hist:
            h = np.full(int(max_iter) + 1, np.nan)
            return h, [h.ctypes.data_as(C.c_void_p), int(h.size)]
        return None, [None, 0]

    def _finish(self, conv, it, res, x, h, hist):
        check(conv)
        out = (bool(conv), it.value, res.value, x)
        return out + (h,) if hist else out

    def pcg(self, x0, b, max_iter, tol, relative=False, pc=PC_JACOBI, omega=1.0, hist=False):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        h, hargs = self._hist(max_iter, hist)
        conv = lib().drops_b200_pcg(self.h, x, _f64(b), int(pc), float(omega), C.byref(it), C.byref(res),
                                    int(relative), *hargs)
        return self._finish(conv, it, res, x, h, hist)

    def pcg_jacobi(self, x0, b, max_iter, tol, relative=False, omega=1.0, hist=False):
        return self.pcg(x0, b, max_iter, tol, relative, PC_JACOBI, omega, hist)

    def pcg_dummy(self, x0, b, max_iter, tol, relative=False, hist=False):
        return self.pcg(x0, b, max_iter, tol, relative, PC_DUMMY, 1.0, hist)

    def pcg_dev(self, x, b, max_iter, tol, relative=False, pc=PC_JACOBI, omega=1.0):
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = check(lib().drops_b200_pcg_dev(self.h, x.h, b.h, int(pc), float(omega), C.byref(it),
                                              C.byref(res), int(relative), None, 0))
        return bool(conv), it.value, res.value

    def cg(self, x0, b, max_iter, tol, relative=False, hist=False):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        h, hargs = self._hist(max_iter, hist)
        conv = lib().drops_b200_cg(self.h, x, _f64(b), C.byref(it), C.byref(res), int(relative), *hargs)
        return self._finish(conv, it, res, x, h, hist)

    def gmres(self, x0, b, restart, max_iter, tol, relative=True, pc=PC_JACOBI, omega=1.0, left=True):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = check(lib().drops_b200_gmres(self.h, x, _f64(b), int(pc), float(omega), int(restart),
                                            C.byref(it), C.byref(res), int(relative), int(left)))
        return bool(conv), it.value, res.value, x

    def gmres_jacobi(self, x0, b, restart, max_iter, tol, relative=True, omega=1.0, left=True):
        return self.gmres(x0, b, restart, max_iter, tol, relative, PC_JACOBI, omega, left)

    def gmres_dummy(self, x0, b, restart, max_iter, tol, relative=True, left=True):
        return self.gmres(x0, b, restart, max_iter, tol, relative, PC_DUMMY, 1.0, left)

    def gmres_dev(self, x, b, restart, max_iter, tol, relative=True, pc=PC_JACOBI, omega=1.0, left=True):
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = check(lib().drops_b200_gmres_dev(self.h, x.h, b.h, int(pc), float(omega), int(restart),
                                                C.byref(it), C.byref(res), int(relative), int(left)))
        return bool(conv), it.value, res.value

    def _minres(self, precond, pc, omega, x0, b, max_iter, tol, relative):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = check(lib().drops_b200_minres(self.h, x, _f64(b), int(precond), int(pc), float(omega),
                                             C.byref(it), C.byref(res), int(relative)))
        return bool(conv), it.value, res.value, x

    def minres(self, x0, b, max_iter, tol, relative=False):
        return self._minres(0, PC_DUMMY, 1.0, x0, b, max_iter, tol, relative)

    def pminres_jacobi(self, x0, b, max_iter, tol, relative=False, omega=1.0):
        return self._minres(1, PC_JACOBI, omega, x0, b, max_iter, tol, relative)

    def pminres_dummy(self, x0, b, max_iter, tol, relative=False):
        return self._minres(1, PC_DUMMY, 1.0, x0, b, max_iter, tol, relative)

    def minres_dev(self, x, b, max_iter, tol, relative=False, precond=0, pc=PC_DUMMY, omega=1.0):
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = check(lib().drops_b200_minres_dev(self.h, x.h, b.h, int(precond), int(pc), float(omega),
                                                 C.byref(it), C.byref(res), int(relative)))
        return bool(conv), it.value, res.value


# --- synthetic brick systems (host side: no GPU needed) -----------------------------------------
def brick_sizes(fe, N):
    n, nnz = C.c_size_t(), C.c_size_t()
    check(lib().drops_b200_brick_sizes(int(fe), int(N), C.byref(n), C.byref(nnz)))
    return n.value, nnz.value


def brick_csr_host(fe, N, numbering=0, upper_scale=1.0):
    """(n, rowbeg, colind, val) in the reference layout, generated on the host."""
    n, nnz = brick_sizes(fe, N)
    rb = np.zeros(n + 1, np.uint64)
    ci = np.zeros(nnz, np.uint64)
    va = np.zeros(nnz)
    check(lib().drops_b200_brick_csr_host(int(fe), int(N), int(numbering), float(upper_scale), rb, ci, va))
    assert rb[-1] == nnz, (rb[-1], nnz)
    return n, rb, ci, va


def brick_solution_host(fe, N, numbering=0):
    n, _ = brick_sizes(fe, N)
    x = np.zeros(n)
    check(lib().drops_b200_brick_solution_host(int(fe), int(N), int(numbering), x))
    return x


def brick_solution_device(fe, N, numbering, vec):
    check(lib().drops_b200_brick_solution_device(int(fe), int(N), int(numbering), vec.h))
