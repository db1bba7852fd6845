"""CPU tests of the C-ABI library: it builds, loads, exports every symbol include/drops_b200.h
declares, refuses to compute without a GPU (no CPU fallback), and its host-side brick generator
produces the systems SURVEY.md section 8(d) describes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "drops_b200.h")).read()
    return sorted(set(re.findall(r"DROPS_B200_API\s+[\w\s\*]+?\b(drops_b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_built):
    names = declared_symbols()
    assert len(names) >= 50
    L = C.CDLL(lib_built)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_covers_the_header(lib_built):
    from drops_b200 import capi
    assert sorted(capi.SIGNATURES) == declared_symbols()
    assert capi.lib().drops_b200_abi_version() == 1
    assert capi.lib().drops_b200_compiled_arch() == 100


def test_built_for_sm_100a(lib_built):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(lib_built):
    from drops_b200 import capi
    with pytest.raises(capi.DropsB200Error) as ei:
        capi.init(0)
    assert ei.value.code == -3          # DROPS_B200_E_NODEVICE
    rb = np.array([0, 1], np.uint64)
    with pytest.raises(capi.DropsB200Error) as ei:
        capi.Matrix(1, 1, rb, np.array([0], np.uint64), np.array([1.0]))
    assert ei.value.code == -4          # DROPS_B200_E_NOTINIT: nothing computes on the CPU
    with pytest.raises(capi.DropsB200Error):
        capi.Vector(4)


@pytest.mark.parametrize("fe,N,lens", [(1, 8, {15}), (2, 8, {65, 27, 19})])
def test_brick_pattern(lib_built, fe, N, lens):
    import scipy.sparse as sp
    from drops_b200 import capi
    n, nnz = capi.brick_sizes(fe, N)
    assert n == ((2 * N - 1) ** 3 if fe == 2 else (N - 1) ** 3)
    n2, rb, ci, va = capi.brick_csr_host(fe, N, 0)
    assert rb[-1] == nnz == len(va)
    A = sp.csr_matrix((va, ci.astype(np.int64), rb.astype(np.int64)), shape=(n, n))
    assert abs(A - A.T).max() == 0.0
    M = 2 * N - 1 if fe == 2 else N - 1
    # rows whose whole stencil is interior have the DROPS row lengths
    idx = np.arange(n)
    i, j, k = idx % M + 1, (idx // M) % M + 1, idx // (M * M) + 1
    r = 2 if fe == 2 else 1
    inner = (np.minimum.reduce([i, j, k]) > r) & (np.maximum.reduce([i, j, k]) < M + 1 - r)
    assert set(np.diff(rb.astype(np.int64))[inner].tolist()) == lens
    # stiffness rows annihilate constants away from the Dirichlet boundary
    assert np.abs((A @ np.ones(n))[inner]).max() < 1e-15
    # ascending, unique columns (num/spmat.h:664)
    for row in range(0, n, max(1, n // 97)):
        c = ci[rb[row]:rb[row + 1]].astype(np.int64)
        assert np.all(np.diff(c) > 0)
    # SPD
    if n <= 3500:
        assert np.linalg.eigvalsh(A.toarray()).min() > 0


def test_brick_p2_row_length_census(lib_built):
    """65 / 27 / 27 / 19 couplings for vertex / axis-edge+space-diagonal / face-diagonal rows and
    28.75 on average in the interior."""
    from drops_b200 import capi
    N = 10
    n, rb, ci, va = capi.brick_csr_host(2, N, 0)
    M = 2 * N - 1
    idx = np.arange(n)
    i, j, k = idx % M + 1, (idx // M) % M + 1, idx // (M * M) + 1
    inner = (np.minimum.reduce([i, j, k]) > 2) & (np.maximum.reduce([i, j, k]) < M - 1)
    odd = (i & 1) + (j & 1) + (k & 1)
    lens = np.diff(rb.astype(np.int64))
    assert set(lens[inner & (odd == 0)]) == {65}
    assert set(lens[inner & (odd == 1)]) == {27}
    assert set(lens[inner & (odd == 2)]) == {19}
    assert set(lens[inner & (odd == 3)]) == {27}


def test_brick_numberings_are_permutations_of_each_other(lib_built):
    import scipy.sparse as sp
    from drops_b200 import capi
    N = 5
    n, rb0, ci0, va0 = capi.brick_csr_host(2, N, 0)
    _, rb1, ci1, va1 = capi.brick_csr_host(2, N, 1)
    A0 = sp.csr_matrix((va0, ci0.astype(np.int64), rb0.astype(np.int64)), shape=(n, n))
    A1 = sp.csr_matrix((va1, ci1.astype(np.int64), rb1.astype(np.int64)), shape=(n, n))
    x0, x1 = capi.brick_solution_host(2, N, 0), capi.brick_solution_host(2, N, 1)
    # the permutation is defined by the solution samples (distinct DOF positions, distinct keys)
    M = 2 * N - 1
    idx = np.arange(n)
    key0 = idx.astype(np.float64)
    P = sp.csr_matrix((np.ones(n), (np.argsort(np.argsort(x1, kind="stable"), kind="stable"),
                                    np.argsort(np.argsort(x0, kind="stable"), kind="stable"))), shape=(n, n))
    # vertices first: the first (N-1)^3 rows of numbering 1 are the 65/… vertex rows
    nv = (N - 1) ** 3
    assert np.diff(rb1.astype(np.int64))[:nv].min() >= 27
    assert np.allclose(np.sort(A0 @ x0), np.sort(A1 @ x1), atol=1e-13)
    assert A0.nnz == A1.nnz and np.isclose(A0.diagonal().sum(), A1.diagonal().sum())
    # numbering 1 columns ascending too
    for row in range(0, n, 13):
        assert np.all(np.diff(ci1[rb1[row]:rb1[row + 1]].astype(np.int64)) > 0)
    assert key0.size == P.shape[0]


def test_brick_solves_poisson(lib_built, port):
    """A_h u_h = b_h with b_h = A_h I_h(u*) is consistent: PCG recovers I_h(u*) (oracle solve)."""
    from drops_b200 import capi
    N = 6
    n, rb, ci, va = capi.brick_csr_host(2, N, 1)
    xs = capi.brick_solution_host(2, N, 1)
    A = port.PortMatrix(n, n, rb, ci, va)
    b = A.spmv(xs)
    conv, it, res, x = A.pcg_jacobi(np.zeros(n), b, 500, 1e-12, True)
    assert conv and np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-10
    assert 0 < xs.max() <= 1.0


def test_documented_options_exist_in_the_source():
    """Every option DESIGN.md section 7 names is one the library reads (`opt("...")` in csrc/)."""
    import glob
    text = open(os.path.join(ROOT, "DESIGN.md")).read()
    sec = text[text.index("## 7. Options"):text.index("## 8.")]
    names = set(re.findall(r"`([a-z]+\.[a-z0-9_/]+)`", sec))
    assert len(names) >= 40
    src = "".join(open(f).read() for f in glob.glob(os.path.join(ROOT, "drops_b200", "csrc", "*.cu*")))
    missing = []
    for n in sorted(names):
        cands = [n]
        if "/" in n:                                   # `sell.block_tile_a/b/c`
            base, parts = n[:n.rindex("_") + 1], n[n.rindex("_") + 1:].split("/")
            cands = [base + p for p in parts]
        missing += [c for c in cands if '"%s"' % c not in src]
    assert not missing, missing


def test_slab_split_counts_a_cut_face_as_planes():
    """"brick.face_weight": the end slabs of a z-slab partition have one cut face, the others two; the split gives every
    rank the same planes + w * faces.  0 = the even split; two ranks (one face each) are always even.  Host code: no GPU."""
    from drops_b200 import capi as gpu
    fe, N, world = 1, 40, 4
    n, _ = gpu.brick_sizes(fe, N)
    M = round(n ** (1.0 / 3.0))                      # interior lattice planes (Dirichlet boundary removed)
    assert M ** 3 == n
    gids = np.arange(n, dtype=np.int64)
    try:
        planes = {}
        for w in (0.0, 5.0):
            gpu.set_option("brick.face_weight", w)
            own = gpu.brick_owner(fe, N, 0, world, gids)
            assert np.all(np.diff(own) >= 0) and own[0] == 0 and own[-1] == world - 1
            cnt = np.bincount(own, minlength=world)
            assert np.all(cnt % (M * M) == 0)
            planes[w] = (cnt // (M * M)).tolist()
            two = np.bincount(gpu.brick_owner(fe, N, 0, 2, gids), minlength=2) // (M * M)
            assert abs(int(two[0]) - int(two[1])) <= 1
        assert max(planes[0.0]) - min(planes[0.0]) <= 1
        p = planes[5.0]
        assert sum(p) == M and min(p[0], p[-1]) >= max(p[1:-1]) + 3 and max(p[1:-1]) - min(p[1:-1]) <= 1, p
    finally:
        gpu.set_option("brick.face_weight", 6.0)
