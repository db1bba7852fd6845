"""Readers / writers for the matrix text formats DROPS itself uses (src/num/spmat.h:1196-1288), so that
a system exported by a real DROPS run -- `A.exportMTX(path)` (:1085-1089), `std::cout << A` (:1213-1224),
`out(os, A)` (:1227-1238; `WriteToFile` in surfactant.cpp:1659-1660) -- can be replayed through the oracle
and through the CUDA path (tools/replay_matrix.py).  Host-side only; CSR comes back in the reference layout
(uint64 rowbeg/colind, float64 val, columns ascending inside a row).
"""
import numpy as np


def _csr_from_triplets(rows, cols, r, c, v):
    """Row-major CSR with ascending columns; a repeated (r, c) keeps the LAST value, like
    SparseMatBuilderCL's `B(r, c) = v` (spmat.h:1283)."""
    r, c, v = np.asarray(r, np.int64), np.asarray(c, np.int64), np.asarray(v, np.float64)
    if r.size and (r.min() < 0 or r.max() >= rows or c.min() < 0 or c.max() >= cols):
        raise ValueError("SparseMatBaseCL operator>>: Inconsistent data.")
    order = np.lexsort((np.arange(r.size), c, r))          # stable in input order inside equal (r, c)
    r, c, v = r[order], c[order], v[order]
    last = np.ones(r.size, bool)
    if r.size > 1:
        last[:-1] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    r, c, v = r[last], c[last], v[last]
    rowbeg = np.zeros(rows + 1, np.uint64)
    np.cumsum(np.bincount(r, minlength=rows), out=rowbeg[1:])
    return rowbeg, c.astype(np.uint64), v


def read_mtx(path):
    """MatrixMarket coordinate file as written by operator<< (`%%MatrixMarket matrix coordinate real
    general`, 1-based).  `symmetric` headers of other writers are expanded.  -> (rows, cols, rowbeg, colind, val)"""
    with open(path) as f:
        header = f.readline()
        if not header.startswith("%%MatrixMarket"):
            raise ValueError(f"{path}: not a MatrixMarket file")
        tokens = header.lower().split()
        if "coordinate" not in tokens or "complex" in tokens or "pattern" in tokens:
            raise ValueError(f"{path}: only real coordinate matrices are on this path")
        symmetric = "symmetric" in tokens
        line = f.readline()
        while line.startswith("%") or not line.strip():
            line = f.readline()
        rows, cols, nnz = (int(t) for t in line.split())
        data = np.loadtxt(f, ndmin=2) if nnz else np.zeros((0, 3))
    if data.shape[0] != nnz:
        raise ValueError(f"{path}: {data.shape[0]} entries, header says {nnz}")
    r, c, v = data[:, 0].astype(np.int64) - 1, data[:, 1].astype(np.int64) - 1, data[:, 2]
    if symmetric:
        off = r != c
        r, c, v = np.r_[r, c[off]], np.r_[c, r[off]], np.r_[v, v[off]]
    rowbeg, colind, val = _csr_from_triplets(rows, cols, r, c, v)
    return rows, cols, rowbeg, colind, val


def write_mtx(path, rows, cols, rowbeg, colind, val):
    """Same bytes as the reference's operator<< (15 digits, scientific, 1-based)."""
    rowbeg = np.asarray(rowbeg, np.int64)
    r = np.repeat(np.arange(rows), np.diff(rowbeg))
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real general\n{rows} {cols} {len(val)}\n")
        for i, j, a in zip(r, np.asarray(colind, np.int64), np.asarray(val, np.float64)):
            f.write(f"{i + 1} {j + 1} {a:.15e}\n")


def read_drops_text(path):
    """The CSR dump of out(os, A): `rows cols nnz`, row_beg[rows+1], col_ind[nnz], val[nnz]."""
    with open(path) as f:
        tok = f.read().split()
    rows, cols, nnz = int(tok[0]), int(tok[1]), int(tok[2])
    p = 3
    rowbeg = np.array(tok[p:p + rows + 1], dtype=np.uint64); p += rows + 1
    colind = np.array(tok[p:p + nnz], dtype=np.uint64); p += nnz
    val = np.array(tok[p:p + nnz], dtype=np.float64); p += nnz
    if len(rowbeg) != rows + 1 or len(colind) != nnz or len(val) != nnz or int(rowbeg[-1]) != nnz:
        raise ValueError(f"{path}: truncated or inconsistent DROPS matrix dump")
    return rows, cols, rowbeg, colind, val


def write_drops_text(path, rows, cols, rowbeg, colind, val, precision=17):
    with open(path, "w") as f:
        f.write(f"{rows} {cols} {len(val)}\n")
        f.write(" ".join(str(int(x)) for x in rowbeg) + " \n")
        f.write(" ".join(str(int(x)) for x in colind) + " \n")
        f.write(" ".join(f"{float(x):.{precision}g}" for x in val) + " \n")


def read_old_drops_triplets(path):
    """What operator>> reads (spmat.h:1258-1288): a comment `% rows x cols nnz nonzeros`, then 1-based
    `r c v` triplets."""
    with open(path) as f:
        text = f.read()
    head, _, rest = text.partition("\n")
    if "%" not in head or "x" not in head:
        raise ValueError('SparseMatBaseCL operator>>: Missing "% rows cols nz" comment.')
    h = head[head.index("%") + 1:]
    rows = int(h[:h.index("x")].split()[0])
    after = h[h.index("x") + 1:].split()
    cols, nnz = int(after[0]), int(after[1])
    tok = rest.split()
    if len(tok) < 3 * nnz:
        raise ValueError("SparseMatBaseCL operator>>: Stream corrupt.")
    data = np.array(tok[:3 * nnz], dtype=np.float64).reshape(nnz, 3)
    rowbeg, colind, val = _csr_from_triplets(rows, cols, data[:, 0].astype(np.int64) - 1,
                                             data[:, 1].astype(np.int64) - 1, data[:, 2])
    return rows, cols, rowbeg, colind, val


def read_matrix(path):
    """Dispatch on the first bytes: MatrixMarket, old triplet format, or CSR dump."""
    with open(path) as f:
        first = f.readline()
    if first.startswith("%%MatrixMarket"):
        return read_mtx(path)
    if first.lstrip().startswith("%"):
        return read_old_drops_triplets(path)
    return read_drops_text(path)
