"""ctypes view of include/drops_b200.h -- the only way Python reaches the CUDA path.

Used by tests/, bench.py and __graft_entry__.py.  Host vectors are numpy float64 arrays, host
matrices are CSR triples in the reference layout (uint64 rowbeg/colind, float64 val).  There is
no CPU fallback: a missing library or a missing GPU raises DropsB200Error.
"""
import ctypes as C
import os

import numpy as np

from . import _build

PC_DUMMY, PC_JACOBI, PC_CHEBYSHEV, PC_DIAG = 0, 1, 2, 3
OP_MUL, OP_TRANSP_MUL = 0, 1

_lib = None


class DropsB200Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libdrops_b200 error {code}: {text}")
        self.code = code


_szp = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
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
    "drops_b200_matrix_lincomb": (_i, [_d, _vp, _d, _vp, C.c_uint64, _vpp]),
    "drops_b200_matrix_destroy": (_i, [_vp]),
    "drops_b200_matrix_info": (_i, [_vp, _szpp, _szpp, _szpp, C.POINTER(C.c_uint64), _szpp, _szpp]),
    "drops_b200_matrix_format": (_i, [_vp, _szpp, _szpp]),
    "drops_b200_matrix_value_dictionary": (_i, [_vp, _i, _ip, _ip, _ip]),
    "drops_b200_matrix_layout": (_i, [_vp, _ip, _ip]),
    "drops_b200_matrix_diag_index": (_i, [_vp, _ip]),
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
    "drops_b200_bicgstab": (_i, [_vp, _dp, _dp, _i, _d, _ip, _dpp, _i, _ip]),
    "drops_b200_bicgstab_dev": (_i, [_vp, _vp, _vp, _i, _d, _ip, _dpp, _i, _ip]),
    "drops_b200_gcr": (_i, [_vp, _dp, _dp, _i, _d, _i, _ip, _dpp, _i]),
    "drops_b200_gcr_dev": (_i, [_vp, _vp, _vp, _i, _d, _i, _ip, _dpp, _i]),
    "drops_b200_matrix_set_chebyshev": (_i, [_vp, _i, _d, _i, _d]),
    "drops_b200_chebyshev_lambda_max": (_i, [_vp, _i, _dpp]),
    "drops_b200_matrix_line_segments": (_i, [_vp, C.POINTER(C.c_size_t)]),
    "drops_b200_matrix_row_order": (_i, [_vp, C.c_void_p, C.c_size_t]),
    "drops_b200_matrix_segment_stream": (_i, [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "drops_b200_matrix_set_diag_pc": (_i, [_vp, C.c_void_p]),
    "drops_b200_diagpc_apply": (_i, [_vp, _vp, _vp]),
    "drops_b200_matrix_bbt_diag": (_i, [_vp, _dp]),
    "drops_b200_precond_apply_dev": (_i, [_vp, _i, _d, _i, _vp, _vp]),
    "drops_b200_pcg_general_dev": (_i, [_vp, _vp, _vp, _i, _d, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_pcgne": (_i, [_vp, _dp, _dp, _i, _d, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_pcgne_dev": (_i, [_vp, _vp, _vp, _i, _d, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_idrs": (_i, [_vp, _dp, _dp, _i, _d, _i, _d, _vp, _ip, _dpp, _i]),
    "drops_b200_idrs_dev": (_i, [_vp, _vp, _vp, _i, _d, _i, _d, _vp, _ip, _dpp, _i]),
    "drops_b200_block_dims": (_i, [_vp, _ip, _szpp, _szpp]),
    "drops_b200_block_spmv": (_i, [_vp, _ip, _dp, _dp]),
    "drops_b200_block_spmv_dev": (_i, [_vp, _ip, _vp, _vp]),
    "drops_b200_composite_spmv": (_i, [_vp, _i, _vp, _i, _dp, _dp]),
    "drops_b200_composite_spmv_dev": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "drops_b200_brick_sizes": (_i, [_i, _i, _szpp, _szpp]),
    "drops_b200_brick_csr_host": (_i, [_i, _i, _i, _d, _szp, _szp, _dp]),
    "drops_b200_brick_matrix_device": (_i, [_i, _i, _i, _d, _vpp]),
    "drops_b200_brick_solution_host": (_i, [_i, _i, _i, _dp]),
    "drops_b200_brick_solution_device": (_i, [_i, _i, _i, _vp]),
    # row-partitioned systems
    "drops_b200_exchange_create": (_i, [_i, _i, _sz, _i64p, _sz, _i64p, _i32p, _vpp]),
    "drops_b200_exchange_destroy": (_i, [_vp]),
    "drops_b200_exchange_sizes": (_i, [_vp, _szpp, _szpp, _szpp, _ip]),
    "drops_b200_exchange_request": (_i, [_vp, _i, _szpp, _vp]),
    "drops_b200_exchange_set_send": (_i, [_vp, _i, _sz, _i64p]),
    "drops_b200_exchange_send_list": (_i, [_vp, _i, _szpp, _vp]),
    "drops_b200_exchange_local_index": (_i, [_vp, _sz, _i64p, _i64p]),
    "drops_b200_exchange_ready": (_i, [_vp]),
    "drops_b200_find_ghosts": (_i, [_sz, _i64p, _sz, _szp, _vp, _sz, _szpp]),
    "drops_b200_partition_rows": (_i, [_sz, _szp, _szp, _i, _i32p]),
    "drops_b200_comm_unique_id": (_i, [_vp]),
    "drops_b200_comm_create": (_i, [_i, _i, _vp, _vpp]),
    "drops_b200_comm_create_local": (_i, [_i, _vpp]),
    "drops_b200_comm_destroy": (_i, [_vp]),
    "drops_b200_comm_info": (_i, [_vp, _ip, _ip, _ip]),
    "drops_b200_comm_allreduce_max": (_i, [_vp, _dpp]),
    "drops_b200_dist_matrix_create": (_i, [_vp, _vp, _sz, _szp, _szp, _dp, C.c_uint64, _vpp]),
    "drops_b200_dist_matrix_create_device": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, C.c_uint64, _vpp]),
    "drops_b200_dist_matrix_destroy": (_i, [_vp]),
    "drops_b200_dist_matrix_info": (_i, [_vp, _szpp, _szpp, _szpp, _szpp, _szpp, _szpp, _szpp]),
    "drops_b200_dist_matrix_format": (_i, [_vp, _ip, _ip]),
    "drops_b200_dist_matrix_diag": (_i, [_vp, _dp]),
    "drops_b200_dist_spmv_dev": (_i, [_i, _vpp, _vpp, _vpp]),
    "drops_b200_dist_dot": (_i, [_i, _vpp, _vpp, _vpp, _dpp]),
    "drops_b200_dist_pcg_dev": (_i, [_i, _vpp, _vpp, _vpp, _i, _d, _ip, _dpp, _i, _vp, _i]),
    "drops_b200_dist_matrix_segment_stream": (_i, [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "drops_b200_dist_matrix_wait_stats": (_i, [_vp, _dp, _i]),
    "drops_b200_dist_gmres_dev": (_i, [_i, _vpp, _vpp, _vpp, _i, _d, _i, _ip, _dpp, _i, _i]),
    "drops_b200_dist_minres_dev": (_i, [_i, _vpp, _vpp, _vpp, _i, _i, _d, _ip, _dpp, _i]),
    "drops_b200_brick_slab_generate": (_i, [_i, _i, _i, _d, _i, _i, _vpp]),
    "drops_b200_brick_slab_destroy": (_i, [_vp]),
    "drops_b200_brick_slab_info": (_i, [_vp, _szpp, _szpp, _szpp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "drops_b200_brick_slab_owned": (_i, [_vp, _i64p]),
    "drops_b200_brick_slab_ghosts": (_i, [_vp, _i64p, _i32p]),
    "drops_b200_brick_slab_matrix": (_i, [_vp, _vp, _vp, _vpp]),
    "drops_b200_brick_slab_solution": (_i, [_vp, _vp]),
    "drops_b200_brick_owner": (_i, [_i, _i, _i, _i, _sz, _i64p, _i32p]),
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


def get_option(key, default=None):
    """Value of a library option; `default` when it was never set (the library then uses its own default)."""
    v = C.c_double()
    r = lib().drops_b200_get_option(key.encode(), C.byref(v))
    if r < 0:
        if default is None:
            check(r)
        return default
    return v.value


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

    def download_into(self, out):
        """Copy into a caller-supplied contiguous float64 array (e.g. a pinned buffer): no intermediate array."""
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == self.n
        check(lib().drops_b200_vector_download(self.h, out))
        return out

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


def diagpc_apply(D, b, x):
    """DiagPcCL::Apply: x = D o b."""
    check(lib().drops_b200_diagpc_apply(D.h, b.h, x.h))


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

    @classmethod
    def lincomb(cls, a, A, b, B, version=1):
        """a*A + b*B merged on the device (SparseMatBaseCL::LinComb)."""
        h = C.c_void_p()
        check(lib().drops_b200_matrix_lincomb(float(a), A.h, float(b), B.h, int(version), C.byref(h)))
        return cls._wrap(h)

    def info(self):
        r, c, z, db, pn = (C.c_size_t() for _ in range(5))
        v = C.c_uint64()
        check(lib().drops_b200_matrix_info(self.h, C.byref(r), C.byref(c), C.byref(z), C.byref(v),
                                           C.byref(db), C.byref(pn)))
        return {"rows": r.value, "cols": c.value, "nnz": z.value, "version": v.value,
                "device_bytes": db.value, "padded_nnz": pn.value}

    def format(self):
        ns, pn = C.c_size_t(), C.c_size_t()
        check(lib().drops_b200_matrix_format(self.h, C.byref(ns), C.byref(pn)))
        return {"nslices": ns.value, "padded_nnz": pn.value}

    def row_order(self):
        """perm[32 * slice + lane] = row summed there (-1: padding position)."""
        perm = np.full(self.format()["nslices"] * 32, -1, dtype=np.int32)
        check(lib().drops_b200_matrix_row_order(self.h, perm.ctypes.data_as(C.c_void_p), perm.size))
        return perm

    def layout(self):
        bo, ro = C.c_int(), C.c_int()
        check(lib().drops_b200_matrix_layout(self.h, C.byref(bo), C.byref(ro)))
        ls = C.c_size_t()
        check(lib().drops_b200_matrix_line_segments(self.h, C.byref(ls)))
        ss, se = C.c_size_t(), C.c_size_t()
        check(lib().drops_b200_matrix_segment_stream(self.h, C.byref(ss), C.byref(se)))
        return {"block_ordered": bool(bo.value), "row_ordered_stream": bool(ro.value), "row_ordered_kind": ro.value,
                "line_segments": ls.value, "segment_slices": ss.value, "segment_entries": se.value}

    def diag_index(self):
        """Length of the diagonal's value list when PCG reads the Jacobi diagonal as one byte per row (0: fp64 vector)."""
        nv = C.c_int()
        check(lib().drops_b200_matrix_diag_index(self.h, C.byref(nv)))
        return nv.value

    def value_dictionary(self, transposed=False):
        nd, pk, cb = C.c_int(), C.c_int(), C.c_int()
        check(lib().drops_b200_matrix_value_dictionary(self.h, int(transposed), C.byref(nd), C.byref(pk), C.byref(cb)))
        return {"n_distinct": nd.value, "packed": pk.value == 1, "six_byte": pk.value == 2, "column_bits": cb.value}

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

    def bicgstab(self, x0, b, max_iter, tol, relative=True, pc=PC_JACOBI, omega=1.0):
        """(converged, iters, resid, x, breakdown)"""
        x = _f64(x0).copy()
        it, res, bd = C.c_int(int(max_iter)), C.c_double(float(tol)), C.c_int(0)
        conv = check(lib().drops_b200_bicgstab(self.h, x, _f64(b), int(pc), float(omega), C.byref(it), C.byref(res),
                                               int(relative), C.byref(bd)))
        return bool(conv), it.value, res.value, x, bd.value

    def gcr(self, x0, b, truncate, max_iter, tol, relative=True, pc=PC_JACOBI, omega=1.0):
        x = _f64(x0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        conv = check(lib().drops_b200_gcr(self.h, x, _f64(b), int(pc), float(omega), int(truncate), C.byref(it),
                                          C.byref(res), int(relative)))
        return bool(conv), it.value, res.value, x

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


    # --- Chebyshev, PCGNE, IDR(s) ----------------------------------------
    def set_chebyshev(self, normal_eq=False, lambda_scale=None, degree=3, ratio=30.0):
        """Constructor arguments of ChebyshevPcCL (normal_eq False) / ChebyshevBBTPcCL (True)."""
        if lambda_scale is None:
            lambda_scale = 1.0 if normal_eq else 1.1
        check(lib().drops_b200_matrix_set_chebyshev(self.h, int(normal_eq), float(lambda_scale), int(degree),
                                                    float(ratio)))

    def set_diag_pc(self, D):
        """DiagPcCL(D): the vector PC_DIAG multiplies with (None forgets it)."""
        if D is None:
            check(lib().drops_b200_matrix_set_diag_pc(self.h, None))
            return
        D = _f64(D)
        assert D.size == self.rows
        check(lib().drops_b200_matrix_set_diag_pc(self.h, D.ctypes.data_as(C.c_void_p)))

    def chebyshev_lambda_max(self, normal_eq=False):
        r = C.c_double()
        check(lib().drops_b200_chebyshev_lambda_max(self.h, int(normal_eq), C.byref(r)))
        return r.value

    def bbt_diag(self):
        d = np.zeros(max(self.rows, 1))
        check(lib().drops_b200_matrix_bbt_diag(self.h, d))
        return d[:self.rows]

    def precond_apply(self, b, pc, omega=1.0, normal_eq=False):
        bv, xv = Vector.from_host(b), Vector(self.rows)
        check(lib().drops_b200_precond_apply_dev(self.h, int(pc), float(omega), int(normal_eq), bv.h, xv.h))
        return xv.download()

    def pcg_general(self, x0, b, max_iter, tol, relative=False, pc=PC_JACOBI, omega=1.0, hist=False):
        xv, bv = Vector.from_host(x0), Vector.from_host(b)
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        h, hargs = self._hist(max_iter, hist)
        conv = lib().drops_b200_pcg_general_dev(self.h, xv.h, bv.h, int(pc), float(omega), C.byref(it), C.byref(res),
                                                int(relative), *hargs)
        return self._finish(conv, it, res, xv.download(), h, hist)

    def pcgne(self, u0, b, max_iter, tol, relative=False, pc=PC_DUMMY, omega=1.0, hist=False):
        u = _f64(u0).copy()
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        h, hargs = self._hist(max_iter, hist)
        conv = lib().drops_b200_pcgne(self.h, u, _f64(b), int(pc), float(omega), C.byref(it), C.byref(res),
                                      int(relative), *hargs)
        return self._finish(conv, it, res, u, h, hist)

    def idrs(self, x0, b, max_iter, tol, shadow=None, relative=False, s=4, omega_bound=0.7, pc=PC_JACOBI, omega=1.0):
        """(ret, iters, resid, x) with the reference's return convention (include/drops_b200.h)."""
        x = _f64(x0).copy()
        sp_ = None
        if shadow is not None:
            shadow = _f64(shadow).reshape(-1)
            assert shadow.size == s * self.rows
            sp_ = shadow.ctypes.data_as(C.c_void_p)
        it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
        ret = check(lib().drops_b200_idrs(self.h, x, _f64(b), int(pc), float(omega), int(s), float(omega_bound), sp_,
                                          C.byref(it), C.byref(res), int(relative)))
        return ret, it.value, res.value, x


def _block_args(blocks, ops):
    hb = (C.c_void_p * 4)(*[b.h if b is not None else None for b in blocks])
    return C.cast(hb, C.c_void_p), (C.c_int * 4)(*[int(o) for o in ops]), hb


def block_dims(blocks, ops):
    r, c = C.c_size_t(), C.c_size_t()
    hb, ob, _keep = _block_args(blocks, ops)
    check(lib().drops_b200_block_dims(hb, ob, C.byref(r), C.byref(c)))
    return r.value, c.value


def block_spmv(blocks, ops, x, transposed=False):
    """BlockMatrixBaseCL * x; blocks = [A, B, C, D] (Matrix or None), ops OP_MUL / OP_TRANSP_MUL.
    transposed: transp_mul(S, x) = GetTranspose() * x (spblockmat.h:194-201)."""
    if transposed:
        blocks = [blocks[0], blocks[2], blocks[1], blocks[3]]
        ops = [1 - ops[0], 1 - ops[2], 1 - ops[1], 1 - ops[3]]
    rows, cols = block_dims(blocks, ops)
    x = _f64(x)
    assert x.size == cols
    y = np.zeros(max(rows, 1))
    hb, ob, _keep = _block_args(blocks, ops)
    check(lib().drops_b200_block_spmv(hb, ob, _nonempty(x, np.float64), y))
    return y[:rows]


def composite_spmv(A0, op0, A1, op1, x, transposed=False):
    """CompositeMatrixBaseCL * x = op1(A1) * (op0(A0) * x); transposed: GetTranspose() * x (spblockmat.h:364-370)."""
    if transposed:
        A0, op0, A1, op1 = A1, 1 - op1, A0, 1 - op0
    ny = A1.rows if op1 == 0 else A1.cols
    y = np.zeros(max(ny, 1))
    check(lib().drops_b200_composite_spmv(A0.h, int(op0), A1.h, int(op1), _nonempty(_f64(x), np.float64), y))
    return y[:ny]


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


# --- row-partitioned systems ---------------------------------------------------------------------
UNIQUE_ID_BYTES = 128


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def find_ghosts(owned_gid, colind):
    """Sorted unique column ids of the local rows that are not owned (host)."""
    owned_gid, colind = _i64(owned_gid), _u64(colind)
    n = C.c_size_t()
    check(lib().drops_b200_find_ghosts(owned_gid.size, _nonempty(owned_gid, np.int64), colind.size,
                                       _nonempty(colind, np.uint64), None, 0, C.byref(n)))
    out = np.zeros(max(n.value, 1), np.int64)
    check(lib().drops_b200_find_ghosts(owned_gid.size, _nonempty(owned_gid, np.int64), colind.size,
                                       _nonempty(colind, np.uint64), out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
    return out[:n.value]


def partition_rows(rowbeg, colind, nparts):
    """part[row] in [0, nparts): level-structure row partitioner (host)."""
    rowbeg, colind = _u64(rowbeg), _u64(colind)
    part = np.zeros(max(rowbeg.size - 1, 1), np.int32)
    check(lib().drops_b200_partition_rows(rowbeg.size - 1, rowbeg, _nonempty(colind, np.uint64), int(nparts), part))
    return part[:rowbeg.size - 1]


class Exchange:
    """Halo plan of one rank (ExchangeCL stand-in); host only."""

    def __init__(self, rank, world, owned_gid, ghost_gid, ghost_owner):
        owned_gid, ghost_gid, ghost_owner = _i64(owned_gid), _i64(ghost_gid), _i32(ghost_owner)
        h = C.c_void_p()
        check(lib().drops_b200_exchange_create(int(rank), int(world), owned_gid.size, _nonempty(owned_gid, np.int64),
                                               ghost_gid.size, _nonempty(ghost_gid, np.int64),
                                               _nonempty(ghost_owner, np.int32), C.byref(h)))
        self.h, self.rank, self.world = h, int(rank), int(world)
        self.n_owned, self.n_ghost = int(owned_gid.size), int(ghost_gid.size)

    def request(self, peer):
        n = C.c_size_t()
        check(lib().drops_b200_exchange_request(self.h, int(peer), C.byref(n), None))
        out = np.zeros(max(n.value, 1), np.int64)
        check(lib().drops_b200_exchange_request(self.h, int(peer), C.byref(n), out.ctypes.data_as(C.c_void_p)))
        return out[:n.value]

    def set_send(self, peer, gids):
        gids = _i64(gids)
        check(lib().drops_b200_exchange_set_send(self.h, int(peer), gids.size, _nonempty(gids, np.int64)))

    def send_list(self, peer):
        n = C.c_size_t()
        check(lib().drops_b200_exchange_send_list(self.h, int(peer), C.byref(n), None))
        out = np.zeros(max(n.value, 1), np.int64)
        check(lib().drops_b200_exchange_send_list(self.h, int(peer), C.byref(n), out.ctypes.data_as(C.c_void_p)))
        return out[:n.value]

    def local_index(self, gids):
        gids = _i64(gids)
        out = np.zeros(max(gids.size, 1), np.int64)
        check(lib().drops_b200_exchange_local_index(self.h, gids.size, _nonempty(gids, np.int64), out))
        return out[:gids.size]

    def sizes(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        nn = C.c_int()
        check(lib().drops_b200_exchange_sizes(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(nn)))
        return {"n_owned": a.value, "n_ghost": b.value, "n_send": c.value, "neighbours": nn.value}

    def ready(self):
        return bool(check(lib().drops_b200_exchange_ready(self.h)))

    def close(self):
        if getattr(self, "h", None):
            lib().drops_b200_exchange_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    def __init__(self, h):
        self.h = h
        r, w, l = C.c_int(), C.c_int(), C.c_int()
        check(lib().drops_b200_comm_info(h, C.byref(r), C.byref(w), C.byref(l)))
        self.rank, self.world, self.is_local = r.value, w.value, bool(l.value)

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        check(lib().drops_b200_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def nccl(cls, rank, world, unique_id=None):
        h = C.c_void_p()
        check(lib().drops_b200_comm_create(int(rank), int(world), unique_id, C.byref(h)))
        return cls(h)

    @classmethod
    def local(cls, world):
        h = C.c_void_p()
        check(lib().drops_b200_comm_create_local(int(world), C.byref(h)))
        return cls(h)

    def allreduce_max(self, v):
        d = C.c_double(float(v))
        check(lib().drops_b200_comm_allreduce_max(self.h, C.byref(d)))
        return d.value

    def close(self):
        if getattr(self, "h", None):
            lib().drops_b200_comm_destroy(self.h)
            self.h = None


def _harr(objs):
    return (C.c_void_p * len(objs))(*[o.h for o in objs])


class DistMatrix:
    """The rows of one rank of a row-partitioned MatrixCL."""

    def __init__(self, comm, ex, rowbeg, colind_global, val, version=1):
        rowbeg, colind_global, val = _u64(rowbeg), _u64(colind_global), _f64(val)
        h = C.c_void_p()
        check(lib().drops_b200_dist_matrix_create(comm.h, ex.h, val.size, rowbeg, _nonempty(colind_global, np.uint64),
                                                  _nonempty(val, np.float64), int(version), C.byref(h)))
        self.h, self.comm, self.ex = h, comm, ex

    @classmethod
    def _wrap(cls, h, comm, ex):
        self = cls.__new__(cls)
        self.h, self.comm, self.ex = h, comm, ex
        return self

    def info(self):
        v = [C.c_size_t() for _ in range(7)]
        check(lib().drops_b200_dist_matrix_info(self.h, *[C.byref(x) for x in v]))
        keys = ["n_owned", "n_ghost", "nnz", "interior_rows", "boundary_rows", "n_send", "device_bytes"]
        return dict(zip(keys, [x.value for x in v]))

    def format(self):
        pk, nd = C.c_int(), C.c_int()
        check(lib().drops_b200_dist_matrix_format(self.h, C.byref(pk), C.byref(nd)))
        return {"interior_packed": bool(pk.value & 1), "six_byte": bool(pk.value & 2), "row_ordered": bool(pk.value & 4),
                "block_ordered": bool(pk.value & 8), "p2p": bool(pk.value & 16), "p2p_fault": bool(pk.value & 32), "p2p_reduce": bool(pk.value & 64),
                "diag_index": bool(pk.value & 128), "segments": bool(pk.value & 256), "n_distinct": nd.value}

    def segment_stream(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib().drops_b200_dist_matrix_segment_stream(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"segment_slices": a.value, "segment_entries": b.value, "interior_padded_entries": c.value}

    def wait_stats(self, reset=False):
        """Time spent waiting for peers inside the iteration: {point: (microseconds, waits)} (drops_b200_dist_matrix_wait_stats)."""
        o = np.zeros(8)
        check(lib().drops_b200_dist_matrix_wait_stats(self.h, o, 1 if reset else 0))
        names = ["halo", "reduce_pAp", "reduce_rho_resid", "reduce_other"]
        return {nm: (o[2 * k] / 1e3, int(o[2 * k + 1])) for k, nm in enumerate(names)}

    def diag(self):
        d = np.zeros(max(self.info()["n_owned"], 1))
        check(lib().drops_b200_dist_matrix_diag(self.h, d))
        return d[:self.info()["n_owned"]]

    def close(self):
        if getattr(self, "h", None):
            lib().drops_b200_dist_matrix_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dist_spmv(parts, xs, ys):
    check(lib().drops_b200_dist_spmv_dev(len(parts), _harr(parts), _harr(xs), _harr(ys)))


def dist_dot(parts, xs, ys):
    r = C.c_double()
    check(lib().drops_b200_dist_dot(len(parts), _harr(parts), _harr(xs), _harr(ys), C.byref(r)))
    return r.value


def dist_pcg(parts, xs, bs, max_iter, tol, relative=False, pc=PC_JACOBI, omega=1.0, hist=False):
    it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
    h, hargs = Matrix._hist(max_iter, hist)
    conv = check(lib().drops_b200_dist_pcg_dev(len(parts), _harr(parts), _harr(xs), _harr(bs), int(pc), float(omega),
                                               C.byref(it), C.byref(res), int(relative), *hargs))
    out = (bool(conv), it.value, res.value)
    return out + (h,) if hist else out


def dist_gmres(parts, xs, bs, restart, max_iter, tol, relative=True, pc=PC_JACOBI, omega=1.0, left=True):
    it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
    conv = check(lib().drops_b200_dist_gmres_dev(len(parts), _harr(parts), _harr(xs), _harr(bs), int(pc), float(omega),
                                                 int(restart), C.byref(it), C.byref(res), int(relative), int(left)))
    return bool(conv), it.value, res.value


def dist_minres(parts, xs, bs, max_iter, tol, relative=False, precond=0, pc=PC_DUMMY, omega=1.0):
    """MINRES (precond = 0) / PMINRES on the partitioned system; (converged, iterations, residual)."""
    it, res = C.c_int(int(max_iter)), C.c_double(float(tol))
    conv = check(lib().drops_b200_dist_minres_dev(len(parts), _harr(parts), _harr(xs), _harr(bs), int(precond), int(pc),
                                                  float(omega), C.byref(it), C.byref(res), int(relative)))
    return bool(conv), it.value, res.value


class BrickSlab:
    """z-slab of the synthetic brick generated on the device (global column ids)."""

    def __init__(self, fe, N, numbering, world, rank, upper_scale=1.0):
        h = C.c_void_p()
        check(lib().drops_b200_brick_slab_generate(int(fe), int(N), int(numbering), float(upper_scale), int(world),
                                                   int(rank), C.byref(h)))
        self.h = h
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        k0, k1 = C.c_longlong(), C.c_longlong()
        check(lib().drops_b200_brick_slab_info(h, C.byref(a), C.byref(b), C.byref(c), C.byref(k0), C.byref(k1)))
        self.n_owned, self.nnz, self.n_ghost, self.k0, self.k1 = a.value, b.value, c.value, k0.value, k1.value
        self.world, self.rank = int(world), int(rank)

    def owned(self):
        out = np.zeros(max(self.n_owned, 1), np.int64)
        check(lib().drops_b200_brick_slab_owned(self.h, out))
        return out[:self.n_owned]

    def ghosts(self):
        g, o = np.zeros(max(self.n_ghost, 1), np.int64), np.zeros(max(self.n_ghost, 1), np.int32)
        check(lib().drops_b200_brick_slab_ghosts(self.h, g, o))
        return g[:self.n_ghost], o[:self.n_ghost]

    def exchange(self):
        g, o = self.ghosts()
        return Exchange(self.rank, self.world, self.owned(), g, o)

    def matrix(self, comm, ex):
        h = C.c_void_p()
        check(lib().drops_b200_brick_slab_matrix(comm.h, ex.h, self.h, C.byref(h)))
        return DistMatrix._wrap(h, comm, ex)

    def solution(self, vec):
        check(lib().drops_b200_brick_slab_solution(self.h, vec.h))

    def close(self):
        if getattr(self, "h", None):
            lib().drops_b200_brick_slab_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def brick_owner(fe, N, numbering, world, gids):
    gids = _i64(gids)
    out = np.zeros(max(gids.size, 1), np.int32)
    check(lib().drops_b200_brick_owner(int(fe), int(N), int(numbering), int(world), gids.size, _nonempty(gids, np.int64), out))
    return out[:gids.size]
