"""Worker of tests/test_partition_cpu.py::test_two_process_plan_over_gloo (one process per rank)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from drops_b200 import capi, partition
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    fe, N, numbering = 2, 6, 1
    n, rb, ci, va = capi.brick_csr_host(fe, N, numbering)
    A = sp.csr_matrix((va, ci.astype(np.int64), rb.astype(np.int64)), shape=(n, n))
    for kind in ("slab", "levels"):
        part = capi.brick_owner(fe, N, numbering, world, np.arange(n)) if kind == "slab" \
            else capi.partition_rows(rb, ci, world)
        owned, lrb, lci, lva, ex = partition.local_system(rb, ci, va, part, rank, world)
        assert not ex.ready()
        partition.connect(ex, dist)                      # requests through gloo
        assert ex.ready()
        i = np.arange(n, dtype=np.uint64)
        x = ((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)).astype(np.float64) / 4294967296.0 - 0.5
        xl = x[owned]
        # halo: pack what the peers requested, receive my ghosts (order = request order)
        ghost = np.zeros(ex.n_ghost)
        ops, bufs, off = [], [], 0
        for q in range(world):
            if q == rank:
                continue
            s = torch.from_numpy(xl[ex.send_list(q)].copy())
            nreq = ex.request(q).size
            rcv = torch.zeros(nreq, dtype=torch.float64)
            bufs.append((off, nreq, rcv))
            off += nreq
            if s.numel():
                ops.append(dist.P2POp(dist.isend, s, q))
            if nreq:
                ops.append(dist.P2POp(dist.irecv, rcv, q))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for o, m, rcv in bufs:
            ghost[o:o + m] = rcv.numpy()
        cols = ex.local_index(lci.astype(np.int64))
        Al = sp.csr_matrix((lva, cols, lrb.astype(np.int64)), shape=(owned.size, owned.size + ex.n_ghost))
        Al.has_canonical_format = True
        yl = Al @ np.concatenate([xl, ghost])
        assert np.array_equal(yl, (A @ x)[owned]), kind
        # dot product: local partial + all_reduce == global to rounding
        t = torch.tensor([float(np.dot(yl, xl))], dtype=torch.float64)
        dist.all_reduce(t)
        ref = float(np.dot(A @ x, x))
        assert abs(t.item() - ref) <= 1e-12 * np.abs((A @ x) * x).sum(), kind
    dist.barrier()
    print("WORKER_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
