"""Host-side logic of the row-partitioned path (no GPU): the halo plan (ExchangeCL stand-in), the
row partitioner, the z-slab owner map -- in process and across two real processes over gloo.
The partitioned product is emulated with scipy on the plan's local numbering and must reproduce
the global product bit for bit (same per-row summation order)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, hash_vector


@pytest.fixture(scope="module")
def capi(lib_built):
    from drops_b200 import capi
    return capi


def global_system(capi, fe=2, N=5, numbering=1):
    n, rb, ci, va = capi.brick_csr_host(fe, N, numbering, 1.1)
    A = sp.csr_matrix((va, ci.astype(np.int64), rb.astype(np.int64)), shape=(n, n))
    return n, rb, ci, va, A


def emulate_product(capi, partition, n, rb, ci, va, part, world, x):
    """y = A x through the plans of all ranks; returns the assembled global y."""
    loc = [partition.local_system(rb, ci, va, part, r, world) for r in range(world)]
    exs = [l[4] for l in loc]
    partition.connect_in_process(exs)
    assert all(e.ready() for e in exs)
    y = np.full(n, np.nan)
    for r, (owned, lrb, lci, lva, ex) in enumerate(loc):
        ghost = np.zeros(ex.n_ghost)
        off = 0
        for q in range(world):
            req = ex.request(q)
            if q == r:
                assert req.size == 0
                continue
            oq, exq = loc[q][0], exs[q]
            sent = x[oq][exq.send_list(r)]                 # what q packs for r
            assert np.array_equal(oq[exq.send_list(r)], req)
            ghost[off:off + req.size] = sent
            off += req.size
        assert off == ex.n_ghost
        xl = np.concatenate([x[owned], ghost])
        cols = ex.local_index(lci.astype(np.int64))
        assert cols.min(initial=0) >= 0
        Al = sp.csr_matrix((lva, cols, lrb.astype(np.int64)), shape=(owned.size, xl.size))
        Al.has_canonical_format = True       # keep the stored entry order: ascending GLOBAL column
        y[owned] = Al @ xl
    return y, loc


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_plan_reproduces_global_product(capi, world):
    from drops_b200 import partition
    n, rb, ci, va, A = global_system(capi)
    part = capi.partition_rows(rb, ci, world)
    assert part.min() == 0 and part.max() == world - 1
    counts = np.bincount(part, minlength=world)
    assert counts.min() > 0.5 * n / world                    # balanced
    x = hash_vector(n)
    y, loc = emulate_product(capi, partition, n, rb, ci, va, part, world, x)
    assert np.array_equal(y, A @ x)
    # the level-structure partition has thin interfaces (9 lattice planes only in this tiny brick)
    if world == 2:
        assert sum(l[4].n_ghost for l in loc) < 0.6 * n


def test_brick_slab_owner_matches_geometry(capi):
    from drops_b200 import partition
    for fe, N, numbering in [(1, 7, 0), (2, 5, 0), (2, 5, 1)]:
        n, rb, ci, va, A = global_system(capi, fe, N, numbering)
        for world in (2, 3):
            part = capi.brick_owner(fe, N, numbering, world, np.arange(n))
            M = 2 * N - 1 if fe == 2 else N - 1
            # owner by construction: plane k of every lattice point, numbering applied
            xs = capi.brick_solution_host(fe, N, numbering)   # x*(p) identifies the position
            x0 = capi.brick_solution_host(fe, N, 0)
            k_of_nat = np.arange(n) // (M * M) + 1
            begin = [1 + M * r // world for r in range(world)] + [M + 1]
            owner_nat = np.searchsorted(begin, k_of_nat, side="right") - 1
            assert np.array_equal(np.bincount(part, minlength=world), np.bincount(owner_nat, minlength=world))
            x = hash_vector(n)
            y, _ = emulate_product(capi, partition, n, rb, ci, va, part, world, x)
            assert np.array_equal(y, A @ x)


def test_plan_rejects_inconsistent_input(capi):
    with pytest.raises(capi.DropsB200Error):
        capi.Exchange(0, 2, [0, 1, 1], [], [])                # owned not strictly ascending
    with pytest.raises(capi.DropsB200Error):
        capi.Exchange(0, 2, [0, 1], [1], [1])                 # ghost that is owned
    with pytest.raises(capi.DropsB200Error):
        capi.Exchange(0, 2, [0, 1], [5], [0])                 # ghost owned by myself
    with pytest.raises(capi.DropsB200Error):
        capi.Exchange(0, 2, [0, 1], [5, 5], [1, 1])           # duplicate ghost
    ex = capi.Exchange(0, 2, [0, 1, 7], [5, 3], [1, 1])
    assert not ex.ready()
    assert np.array_equal(ex.request(1), [3, 5])              # ascending inside an owner group
    assert np.array_equal(ex.local_index([0, 7, 3, 5, 4]), [0, 2, 3, 4, -1])
    with pytest.raises(capi.DropsB200Error):
        ex.set_send(1, [2])                                   # peer asks for something we do not own
    ex.set_send(1, [7, 0])
    assert ex.ready() and np.array_equal(ex.send_list(1), [2, 0])
    assert ex.sizes() == {"n_owned": 3, "n_ghost": 2, "n_send": 2, "neighbours": 1}


def test_find_ghosts_and_empty_rank(capi):
    from drops_b200 import partition
    n, rb, ci, va, A = global_system(capi, 1, 6, 0)
    part = np.zeros(n, np.int32)
    part[n // 2:] = 2                                         # rank 1 owns nothing
    x = hash_vector(n)
    y, loc = emulate_product(capi, partition, n, rb, ci, va, part, 3, x)
    assert np.array_equal(y, A @ x)
    assert loc[1][0].size == 0 and loc[1][4].n_ghost == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_process_plan_over_gloo(capi):
    """world_size 2, gloo: requests travel through torch.distributed, the halo and the dot product
    through gloo send/recv/all_reduce; results equal the global ones."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker_cpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("WORKER_OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]
