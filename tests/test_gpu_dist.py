"""GPU parity of the row-partitioned path.  On one GPU all ranks are partitions of one process
(local communicator: device-to-device copies instead of NCCL), which runs the whole partitioned
machinery -- plan, column remap, interior/boundary split, chained double-double partials,
rank-ordered reductions -- against the single-GPU path and the reference goldens.  With >= 2 GPUs
the same checks run over NCCL, one process per GPU (tests/dist_worker_gpu.py).
Bars: SpMV rows bit-identical to the single-GPU product; PCG iteration count identical (+-1
allowed by north_star), solution within 1e-10 relative."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, hash_vector, load_golden

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def build_parts(gpu, rb, ci, va, part, world):
    from drops_b200 import partition
    comm = gpu.Comm.local(world)
    loc = [partition.local_system(rb, ci, va, part, r, world) for r in range(world)]
    partition.connect_in_process([l[4] for l in loc])
    mats = [gpu.DistMatrix(comm, l[4], l[1], l[2], l[3]) for l in loc]
    return comm, loc, mats


def scatter(gpu, loc, v):
    return [gpu.Vector.from_host(v[l[0]]) for l in loc]


def gather(loc, vecs, n):
    out = np.full(n, np.nan)
    for l, v in zip(loc, vecs):
        out[l[0]] = v.download()
    return out


@pytest.mark.parametrize("world", [1, 2, 3])
def test_partitioned_matches_reference_golden(gpu, world):
    g = load_golden("brick_fe2_N6_num1.npz")
    n = int(g["rows"])
    rb, ci, va = g["rowbeg"], g["colind"], g["val"]
    part = gpu.partition_rows(rb, ci, world)
    comm, loc, mats = build_parts(gpu, rb, ci, va, part, world)
    infos = [m.info() for m in mats]
    assert sum(i["n_owned"] for i in infos) == n and sum(i["nnz"] for i in infos) == len(va)
    for i in infos:
        assert i["interior_rows"] + i["boundary_rows"] == i["n_owned"]
        assert (i["boundary_rows"] > 0) == (world > 1)
    # y = A x: every row bit-identical to the compiled reference
    xs = scatter(gpu, loc, g["x_in"])
    ys = [gpu.Vector(l[0].size) for l in loc]
    gpu.dist_spmv(mats, xs, ys)
    assert np.array_equal(gather(loc, ys, n), g["y_Ax"])
    assert np.array_equal(np.concatenate([m.diag() for m in mats])[np.argsort(np.concatenate([l[0] for l in loc]))], g["diag"])
    # dot: rank-ordered double-double sum == single-GPU reduction to the last bits
    d = gpu.dist_dot(mats, xs, ys)
    xv, yv = gpu.Vector.from_host(g["x_in"]), gpu.Vector.from_host(g["y_Ax"])
    assert abs(d - gpu.dot(xv, yv)) <= 4e-16 * float(np.abs(g["x_in"] * g["y_Ax"]).sum())
    # the brick qualifies for the packed stream, so every partition with interior rows has its row-ordered twins
    for m, i in zip(mats, infos):
        f = m.format()
        assert f["interior_packed"] == (i["interior_rows"] > 0) and f["row_ordered"] == (i["interior_rows"] > 0)
    # PCG + Jacobi to convergence: same iteration count as the reference run, same solution
    bs = scatter(gpu, loc, g["b"])
    x0 = [gpu.Vector(l[0].size) for l in loc]
    conv, it, res, hist = gpu.dist_pcg(mats, x0, bs, 400, 1e-10, True, hist=True)
    assert conv and abs(it - int(g["sym_pcgjac"][1])) <= 1
    assert rel(gather(loc, x0, n), g["sym_pcgjac_x"]) <= 1e-10
    k = min(it, int(g["sym_pcgjac"][1]))
    assert np.allclose(hist[1:k + 1], g["sym_pcgjac_hist"][1:k + 1], rtol=1e-6, atol=1e-15)
    # fixed work (tol = 0): exactly 7 iterations, not converged
    for v in x0:
        v.fill(0.0)
    conv, it, res = gpu.dist_pcg(mats, x0, bs, 7, 0.0, True)
    assert not conv and it == 7
    assert abs(res - g["sym_pcgjac_fixed7"][2]) <= 1e-9 * g["sym_pcgjac_fixed7"][2] + 1e-15
    assert rel(gather(loc, x0, n), g["sym_pcgjac_fixed7_x"]) <= 1e-12
    # the same solve in the owned numbering ("dist.reorder" = 0 at solve time): same count, same solution
    try:
        gpu.set_option("dist.reorder", 0)
        x1 = [gpu.Vector(l[0].size) for l in loc]
        conv1, it1, res1 = gpu.dist_pcg(mats, x1, bs, 7, 0.0, True)
        assert not conv1 and it1 == 7 and abs(res1 - res) <= 1e-9 * res and rel(gather(loc, x1, n), gather(loc, x0, n)) <= 1e-12
    finally:
        gpu.set_option("dist.reorder", 1)
    # halo through peer stores (default inside the solver loops) vs copies: the same bits
    if world > 1:
        assert all(m.format()["p2p"] and m.format()["p2p_reduce"] and not m.format()["p2p_fault"] for m in mats)
        try:
            gpu.set_option("dist.p2p", 0)
            gpu.set_option("dist.p2p_reduce", 0)
            x2 = [gpu.Vector(l[0].size) for l in loc]
            conv2, it2, res2 = gpu.dist_pcg(mats, x2, bs, 7, 0.0, True)
            assert it2 == 7 and res2 == res and np.array_equal(gather(loc, x2, n), gather(loc, x0, n))
        finally:
            gpu.set_option("dist.p2p", 1)
            gpu.set_option("dist.p2p_reduce", 1)
    # unpreconditioned (DummyPcCL) and omega != 1
    for v in x0:
        v.fill(0.0)
    conv, it, res = gpu.dist_pcg(mats, x0, bs, 400, 1e-10, True, pc=gpu.PC_DUMMY)
    assert conv and abs(it - int(g["sym_pcgdummy"][1])) <= 1 and rel(gather(loc, x0, n), g["sym_pcgdummy_x"]) <= 1e-10
    for v in x0:
        v.fill(0.0)
    conv, it, res = gpu.dist_pcg(mats, x0, bs, 400, 1e-10, False, omega=0.8)
    assert conv and abs(it - int(g["sym_pcgjac_w08"][1])) <= 1 and rel(gather(loc, x0, n), g["sym_pcgjac_w08_x"]) <= 1e-10


@pytest.mark.parametrize("world,tile", [(2, 2), (3, 8)])
def test_partition_rows_in_graph_blocks(gpu, world, tile):
    """Partitions whose rows are cut in blocks of their own graph (reorder.cu on the owned x owned part) and
    iterate on row-ordered vectors: products keep their bits, the solve its count and solution."""
    g = load_golden("brick_fe2_N6_num1.npz")
    n = int(g["rows"])
    rb, ci, va = g["rowbeg"], g["colind"], g["val"]
    part = gpu.partition_rows(rb, ci, world)
    try:
        gpu.set_option("sell.block_min_rows", 0)
        gpu.set_option("sell.block_tile", tile)
        comm, loc, mats = build_parts(gpu, rb, ci, va, part, world)
        assert all(m.format()["block_ordered"] and m.format()["row_ordered"] for m in mats)
        xs = scatter(gpu, loc, g["x_in"])
        ys = [gpu.Vector(l[0].size) for l in loc]
        gpu.dist_spmv(mats, xs, ys)
        assert np.array_equal(gather(loc, ys, n), g["y_Ax"])
        bs = scatter(gpu, loc, g["b"])
        x0 = [gpu.Vector(l[0].size) for l in loc]
        conv, it, res = gpu.dist_pcg(mats, x0, bs, 400, 1e-10, True)
        assert conv and abs(it - int(g["sym_pcgjac"][1])) <= 1 and rel(gather(loc, x0, n), g["sym_pcgjac_x"]) <= 1e-10
        x1 = [gpu.Vector(l[0].size) for l in loc]
        conv, it, res = gpu.dist_gmres(mats, x1, bs, 20, 400, 1e-10, True)
        assert conv and abs(it - int(g["sym_gmresjac_l1"][1])) <= 1 and rel(gather(loc, x1, n), g["sym_gmresjac_l1_x"]) <= 1e-10
    finally:
        gpu.set_option("sell.block_min_rows", 2.0e6)
        gpu.set_option("sell.block_tile", 8)


@pytest.mark.parametrize("name,tag,world", [("brick_fe2_N6_num1.npz", "sym", 1), ("brick_fe2_N6_num1.npz", "sym", 3),
                                            ("brick_fe2_N4_num1_nonsym.npz", "nonsym", 2), ("brick_fe2_N4_num1_nonsym.npz", "nonsym", 4)])
def test_partitioned_gmres_matches_reference_golden(gpu, name, tag, world):
    """GMRES(m) on the partitioned system (rank-ordered reductions, Givens rotations on the host of every
    rank) against the compiled reference's serial GMRES: iteration counts +-1, solutions 1e-10 / 1e-8."""
    g = load_golden(name)
    n = int(g["rows"])
    rb, ci, va = g["rowbeg"], g["colind"], g["val"]
    part = gpu.partition_rows(rb, ci, world)
    comm, loc, mats = build_parts(gpu, rb, ci, va, part, world)
    bs = scatter(gpu, loc, g["b"])
    single = gpu.Matrix(n, n, rb, ci, va)
    for key, restart, left, x_rtol in ((f"{tag}_gmresjac_l1", 20, True, 1e-10), (f"{tag}_gmresjac_l0", 20, False, 1e-10),
                                       (f"{tag}_gmresjac_m5", 5, True, 1e-8)):
        x0 = [gpu.Vector(l[0].size) for l in loc]
        conv, it, res = gpu.dist_gmres(mats, x0, bs, restart, 400, 1e-10, True, left=left)
        rconv, rit, rres = bool(g[key][0]), int(g[key][1]), float(g[key][2])
        assert conv == rconv and abs(it - rit) <= 1, (key, it, rit)
        xg = gather(loc, x0, n)
        if it == rit:
            assert abs(res - rres) <= 1e-3 * abs(rres) + 1e-15
            assert rel(xg, g[key + "_x"]) <= x_rtol, (key, rel(xg, g[key + "_x"]))
        one = single.gmres_jacobi(np.zeros(n), g["b"], restart, 400, 1e-10, True, 1.0, left)
        assert one[1] == it and rel(xg, one[3]) <= 1e-9
        # "gmres.ortho" = 1: classical Gram-Schmidt, the reference's distributed ModGMRES (krylovsolver.h:465-494, :862-866) --
        # one multi-dot and ONE reduction over the ranks per column; same iteration count, same solution to 1e-8, and the
        # single-GPU driver in the same mode agrees with it
        try:
            gpu.set_option("gmres.ortho", 1)
            x1 = [gpu.Vector(l[0].size) for l in loc]
            conv1, it1, res1 = gpu.dist_gmres(mats, x1, bs, restart, 400, 1e-10, True, left=left)
            assert conv1 == conv and abs(it1 - it) <= 1, (key, it1, it)
            assert rel(gather(loc, x1, n), g[key + "_x"]) <= max(x_rtol, 1e-8), (key, rel(gather(loc, x1, n), g[key + "_x"]))
            one1 = single.gmres_jacobi(np.zeros(n), g["b"], restart, 400, 1e-10, True, 1.0, left)
            assert one1[1] == it1 and rel(gather(loc, x1, n), one1[3]) <= 1e-9
        finally:
            gpu.set_option("gmres.ortho", 0)
    if tag == "sym":
        x0 = [gpu.Vector(l[0].size) for l in loc]
        conv, it, res = gpu.dist_gmres(mats, x0, bs, 20, 400, 1e-10, False, pc=gpu.PC_DUMMY)
        assert conv and abs(it - int(g["sym_gmresdummy"][1])) <= 1
        assert rel(gather(loc, x0, n), g["sym_gmresdummy_x"]) <= 1e-9
    # max_iter reached: not converged, iteration budget untouched
    x0 = [gpu.Vector(l[0].size) for l in loc]
    conv, it, res = gpu.dist_gmres(mats, x0, bs, 4, 6, 1e-30, True)
    assert not conv and it == 6


@pytest.mark.parametrize("world", [1, 2, 4])
def test_partitioned_minres_matches_reference_golden(gpu, world):
    """MINRES and PMINRES<JACPcCL> on the partitioned system (krylovsolver.h:926-1161, the solvers of the reference's
    parallel Stokes drivers) against the compiled reference's serial run: iteration counts +-1, solutions to 1e-8, and
    the single-GPU driver with the same arguments."""
    g = load_golden("brick_fe2_N6_num1.npz")
    n = int(g["rows"])
    rb, ci, va = g["rowbeg"], g["colind"], g["val"]
    part = gpu.partition_rows(rb, ci, world)
    comm, loc, mats = build_parts(gpu, rb, ci, va, part, world)
    bs = scatter(gpu, loc, g["b"])
    single = gpu.Matrix(n, n, rb, ci, va)
    for key, precond, pc in (("sym_minres", 0, gpu.PC_DUMMY), ("sym_pminresjac", 1, gpu.PC_JACOBI)):
        x0 = [gpu.Vector(l[0].size) for l in loc]
        conv, it, res = gpu.dist_minres(mats, x0, bs, 400, 1e-9, False, precond, pc)
        rconv, rit, rres = bool(g[key][0]), int(g[key][1]), float(g[key][2])
        assert conv == rconv and abs(it - rit) <= 1, (key, it, rit)
        xg = gather(loc, x0, n)
        if it == rit:
            assert abs(res - rres) <= 1e-3 * abs(rres) + 1e-15
            assert rel(xg, g[key + "_x"]) <= 1e-8, (key, rel(xg, g[key + "_x"]))
        one = single.minres(np.zeros(n), g["b"], 400, 1e-9) if not precond else single.pminres_jacobi(np.zeros(n), g["b"], 400, 1e-9)
        assert one[1] == it and rel(xg, one[3]) <= 1e-9
    # zero right-hand side: converged before the first step; max_iter reached: not converged
    zs = [gpu.Vector(l[0].size) for l in loc]
    x0 = [gpu.Vector(l[0].size) for l in loc]
    conv, it, res = gpu.dist_minres(mats, x0, zs, 10, 1e-10, False)
    assert conv and it == 0 and res == 0.0
    x0 = [gpu.Vector(l[0].size) for l in loc]
    conv, it, res = gpu.dist_minres(mats, x0, bs, 3, 1e-30, False)
    assert not conv


@pytest.mark.parametrize("fe,N,numbering,world", [(2, 16, 1, 4), (2, 16, 0, 3), (1, 24, 0, 2), (2, 12, 1, 8)])
def test_brick_slabs_match_single_gpu(gpu, fe, N, numbering, world):
    """Device-generated z-slabs (the multi-GPU bench path) against the single-GPU operator."""
    from drops_b200 import partition
    n, nnz = gpu.brick_sizes(fe, N)
    G = gpu.Matrix.brick(fe, N, numbering)
    comm = gpu.Comm.local(world)
    slabs, exs, mats = partition.brick_partition_in_process(fe, N, numbering, world, comm)
    owned = [s.owned() for s in slabs]
    assert sum(o.size for o in owned) == n and sum(m.info()["nnz"] for m in mats) == nnz
    assert np.array_equal(np.sort(np.concatenate(owned)), np.arange(n))
    for r, (s, o) in enumerate(zip(slabs, owned)):
        assert np.all(np.diff(o) > 0)
        assert np.all(gpu.brick_owner(fe, N, numbering, world, o) == r)
        gg, go = s.ghosts()
        assert np.array_equal(gpu.brick_owner(fe, N, numbering, world, gg), go)
        assert set(np.abs(go - r)) <= {1}                       # slabs only talk to their neighbours
    xin = hash_vector(n)
    xs = [gpu.Vector.from_host(xin[o]) for o in owned]
    ys = [gpu.Vector(o.size) for o in owned]
    gpu.dist_spmv(mats, xs, ys)
    y = np.full(n, np.nan)
    for o, v in zip(owned, ys):
        y[o] = v.download()
    assert np.array_equal(y, G.spmv(xin))
    # x* and b = A x*
    xstar = gpu.brick_solution_host(fe, N, numbering)
    xsv = [gpu.Vector(o.size) for o in owned]
    for s, v, o in zip(slabs, xsv, owned):
        s.solution(v)
        assert np.array_equal(v.download(), xstar[o])
    bs = [gpu.Vector(o.size) for o in owned]
    gpu.dist_spmv(mats, xsv, bs)
    b = G.spmv(xstar)
    ref = G.pcg_jacobi(np.zeros(n), b, 2000, 1e-10, True, hist=True)
    x0 = [gpu.Vector(o.size) for o in owned]
    conv, it, res, hist = gpu.dist_pcg(mats, x0, bs, 2000, 1e-10, True, hist=True)
    assert conv and ref[0] and abs(it - ref[1]) <= 1, (it, ref[1])
    x = np.full(n, np.nan)
    for o, v in zip(owned, x0):
        x[o] = v.download()
    assert rel(x, ref[3]) <= 1e-10
    k = min(it, ref[1])
    assert np.allclose(hist[:k + 1], ref[4][:k + 1], rtol=1e-6, atol=1e-15)
    # peer waits are only counted on the peer-to-peer path; the call itself works on every partition
    for m in mats:
        ws = m.wait_stats()
        assert set(ws) == {"halo", "reduce_pAp", "reduce_rho_resid", "reduce_other"}
        assert all(v[0] >= 0.0 and v[1] >= 0 for v in ws.values())
        if m.format()["p2p_reduce"] and world > 1:
            assert ws["reduce_pAp"][1] >= it


@pytest.mark.parametrize("world", [2, 3])
def test_slab_interiors_run_from_segment_descriptors(gpu, world):
    """z-slabs of a P2 brick with lines of 37-39 same-shape rows: the interior operator of every partition is cut in line
    segments (kept whole across the interior / boundary split) and runs from segment descriptors; PCG, GMRES and MINRES
    give the bits of the packed kernel ("spmv.segments" = 0) and the single-GPU iteration counts."""
    from drops_b200 import partition
    fe, N, numbering = 2, 40, 1
    n, nnz = gpu.brick_sizes(fe, N)
    try:
        gpu.set_option("sell.block_min_rows", 0)
        G = gpu.Matrix.brick(fe, N, numbering)
        comm = gpu.Comm.local(world)
        slabs, exs, mats = partition.brick_partition_in_process(fe, N, numbering, world, comm)
        assert all(m.format()["block_ordered"] and m.format()["row_ordered"] and m.format()["segments"] for m in mats)
        owned = [s.owned() for s in slabs]
        xstar = gpu.brick_solution_host(fe, N, numbering)
        b = G.spmv(xstar)
        bs = [gpu.Vector.from_host(b[o]) for o in owned]
        ref = G.pcg_jacobi(np.zeros(n), b, 2000, 1e-10, True)

        def run_all():
            out = []
            x0 = [gpu.Vector(o.size) for o in owned]
            out.append(gpu.dist_pcg(mats, x0, bs, 2000, 1e-10, True) + (np.concatenate([v.download() for v in x0]),))
            x1 = [gpu.Vector(o.size) for o in owned]
            out.append(gpu.dist_gmres(mats, x1, bs, 10, 30, 1e-30, True) + (np.concatenate([v.download() for v in x1]),))
            x2 = [gpu.Vector(o.size) for o in owned]
            out.append(gpu.dist_minres(mats, x2, bs, 40, 1e-30, True) + (np.concatenate([v.download() for v in x2]),))
            return out
        seg = run_all()
        gpu.set_option("spmv.segments", 0)
        pk = run_all()
        gpu.set_option("spmv.segments", 1)
        for a, c in zip(seg, pk):
            assert a[1] == c[1] and a[2] == c[2] and np.array_equal(a[3], c[3])
        assert seg[0][0] and abs(seg[0][1] - ref[1]) <= 1
        x = np.full(n, np.nan)
        x[np.concatenate(owned)] = seg[0][3]
        assert rel(x, ref[3]) <= 1e-10
    finally:
        gpu.set_option("spmv.segments", 1)
        gpu.set_option("sell.block_min_rows", 2.0e6)


def test_partition_with_empty_rank_and_nccl_world1(gpu):
    g = load_golden("brick_fe1_N6_num0.npz")
    n = int(g["rows"])
    rb, ci, va = g["rowbeg"], g["colind"], g["val"]
    part = np.zeros(n, np.int32)
    part[n // 3:] = 2                                          # rank 1 owns nothing
    comm, loc, mats = build_parts(gpu, rb, ci, va, part, 3)
    xs = scatter(gpu, loc, g["x_in"])
    ys = [gpu.Vector(l[0].size) for l in loc]
    gpu.dist_spmv(mats, xs, ys)
    assert np.array_equal(gather(loc, ys, n), g["y_Ax"])
    bs = scatter(gpu, loc, g["b"])
    x0 = [gpu.Vector(l[0].size) for l in loc]
    conv, it, res = gpu.dist_pcg(mats, x0, bs, 400, 1e-10, True)
    assert conv and abs(it - int(g["sym_pcgjac"][1])) <= 1 and rel(gather(loc, x0, n), g["sym_pcgjac_x"]) <= 1e-10
    # an NCCL communicator of one rank needs no NCCL: same entry points, one partition
    from drops_b200 import partition
    c1 = gpu.Comm.nccl(0, 1)
    owned, lrb, lci, lva, ex = partition.local_system(rb, ci, va, np.zeros(n, np.int32), 0, 1)
    A = gpu.DistMatrix(c1, ex, lrb, lci, lva)
    xv, yv = gpu.Vector.from_host(g["x_in"]), gpu.Vector(n)
    gpu.dist_spmv([A], [xv], [yv])
    assert np.array_equal(yv.download(), g["y_Ax"])
    # wrong number of partitions for the communicator / missing diagonal
    with pytest.raises(gpu.DropsB200Error):
        gpu.dist_spmv(mats[:2], xs[:2], ys[:2])
    Z = gpu.DistMatrix(c1, gpu.Exchange(0, 1, [0, 1], [], []), [0, 1, 2], [1, 0], [1., 1.])
    with pytest.raises(gpu.DropsB200Error) as ei:
        gpu.dist_pcg([Z], [gpu.Vector(2)], [gpu.Vector(2)], 5, 1e-10)
    assert ei.value.code == -5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def test_two_ranks_over_nccl(gpu):
    """One process per GPU, NCCL halo + all-gather; needs >= 2 GPUs (skipped on the 1-GPU box)."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("WORKER_OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]
