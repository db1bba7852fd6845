"""Matrix text formats (src/num/spmat.h:1213-1288): files written by the REFERENCE's own operator<< and
out() -- committed under tests/golden/ (made by this file's `regenerate()` where /root/reference exists) --
are read by drops_b200/matrix_io.py; live round trips against the compiled reference where it is built."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden


def _system():
    g = load_golden("brick_fe2_N4_num1_nonsym.npz")
    return int(g["rows"]), int(g["cols"]), g["rowbeg"], g["colind"], g["val"]


def regenerate():
    from oracle import ref
    n, m, rb, ci, va = _system()
    R = ref.RefMatrix(n, m, rb, ci, va)
    R.write(os.path.join(GOLDEN, "io_ref_operator_out.mtx"), "mtx")
    R.write(os.path.join(GOLDEN, "io_ref_out_p17.txt"), "text", 17)
    R.write(os.path.join(GOLDEN, "io_ref_out_default.txt"), "text", 6)


def test_reads_files_written_by_the_reference():
    from drops_b200 import matrix_io as mio
    n, m, rb, ci, va = _system()
    r = mio.read_matrix(os.path.join(GOLDEN, "io_ref_operator_out.mtx"))      # operator<<: 16 significant digits
    assert r[0] == n and r[1] == m and np.array_equal(r[2], rb) and np.array_equal(r[3], ci)
    assert np.allclose(r[4], va, rtol=2e-16 * 10, atol=0)
    r = mio.read_matrix(os.path.join(GOLDEN, "io_ref_out_p17.txt"))           # out() with precision 17: exact
    assert r[0] == n and np.array_equal(r[2], rb) and np.array_equal(r[3], ci) and np.array_equal(r[4], va)
    r = mio.read_matrix(os.path.join(GOLDEN, "io_ref_out_default.txt"))       # out() at the stream default (6 digits)
    assert np.array_equal(r[2], rb) and np.array_equal(r[3], ci) and np.allclose(r[4], va, rtol=1e-5)


def test_writers_reproduce_the_reference_bytes(tmp_path):
    from drops_b200 import matrix_io as mio
    n, m, rb, ci, va = _system()
    p = tmp_path / "a.mtx"
    mio.write_mtx(p, n, m, rb, ci, va)
    assert open(p).read() == open(os.path.join(GOLDEN, "io_ref_operator_out.mtx")).read()
    q = tmp_path / "a.txt"
    mio.write_drops_text(q, n, m, rb, ci, va)
    r = mio.read_drops_text(q)
    assert np.array_equal(r[2], rb) and np.array_equal(r[3], ci) and np.array_equal(r[4], va)


def test_old_triplet_format_and_errors(tmp_path):
    from drops_b200 import matrix_io as mio
    p = tmp_path / "old.txt"
    p.write_text("% 2 x 3 4 nonzeros\n1 1 1.5\n2 3 -2\n1 2 7\n1 1 9\n")       # repeated (1,1): last value wins
    n, m, rb, ci, va = mio.read_matrix(p)
    assert (n, m) == (2, 3) and list(rb) == [0, 2, 3] and list(ci) == [0, 1, 2] and list(va) == [9.0, 7.0, -2.0]
    p.write_text("% 2 x 2 1 nonzeros\n3 1 1.0\n")
    with pytest.raises(ValueError):
        mio.read_matrix(p)                                                   # "Inconsistent data"
    p.write_text("2 2 3\n0 1 2 \n0 1 \n1 2 \n")
    with pytest.raises(ValueError):
        mio.read_matrix(p)                                                   # nnz mismatch
    s = tmp_path / "sym.mtx"
    s.write_text("%%MatrixMarket matrix coordinate real symmetric\n% comment\n2 2 2\n1 1 4\n2 1 -1\n")
    n, m, rb, ci, va = mio.read_matrix(s)
    assert list(rb) == [0, 2, 3] and list(ci) == [0, 1, 0] and list(va) == [4.0, -1.0, -1.0]


def test_live_round_trip_with_the_compiled_reference(refmod, tmp_path):
    """ours -> reference in(), reference out()/operator<< -> ours, then the same product on both sides."""
    from drops_b200 import matrix_io as mio
    n, m, rb, ci, va = _system()
    q = tmp_path / "ours.txt"
    mio.write_drops_text(q, n, m, rb, ci, va)
    R = refmod.RefMatrix.read_text(q)
    rb2, ci2, va2 = R.csr()
    assert np.array_equal(rb2, rb) and np.array_equal(ci2, ci) and np.array_equal(va2, va)
    R.write(tmp_path / "ref.mtx", "mtx")
    r = mio.read_mtx(tmp_path / "ref.mtx")
    assert np.array_equal(r[3], ci) and np.allclose(r[4], va, rtol=2e-15, atol=0)
    assert open(tmp_path / "ref.mtx").read() == open(os.path.join(GOLDEN, "io_ref_operator_out.mtx")).read()


@pytest.mark.gpu
def test_replay_of_an_exported_matrix_on_the_gpu(gpu, port, tmp_path):
    """A matrix that only exists as a reference-written file goes through the oracle and the CUDA path."""
    from drops_b200 import matrix_io as mio
    n, m, rb, ci, va = mio.read_matrix(os.path.join(GOLDEN, "io_ref_out_p17.txt"))
    g = load_golden("brick_fe2_N4_num1_nonsym.npz")
    A = gpu.Matrix(n, m, rb, ci, va)
    assert np.array_equal(A.spmv(g["x_in"]), g["y_Ax"])
    P = port.PortMatrix(n, m, rb, ci, va)
    r = P.gmres_jacobi(np.zeros(n), g["b"], 20, 400, 1e-10, True)
    o = A.gmres_jacobi(np.zeros(n), g["b"], 20, 400, 1e-10, True)
    assert r[0] and o[0] and abs(o[1] - r[1]) <= 1


if __name__ == "__main__":
    regenerate()
