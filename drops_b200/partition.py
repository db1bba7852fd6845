"""Plumbing around the row-partitioned path: who owns which rows, and the transport of the halo
requests between the ranks at SETUP time (torch.distributed, any backend).  Nothing here is on the
data path: per iteration the ranks talk through NCCL inside libdrops_b200 only.

What it stands in for: DROPS derives its ExchangeCL from the ParMETIS-partitioned mesh
(parallel/partitioner.cpp:251-257, misc/problem.cpp:757-760).  Here a partition is just `part[row]`
(from capi.partition_rows, from z-slabs of the synthetic brick, or from the caller) and the halo
plan follows from the matrix pattern.
"""
import numpy as np

from . import capi


def extract_rows(rowbeg, colind, val, rows):
    """CSR of the listed rows (global column ids kept)."""
    rowbeg = np.asarray(rowbeg, dtype=np.int64)
    rows = np.asarray(rows, dtype=np.int64)
    lens = rowbeg[rows + 1] - rowbeg[rows]
    rb = np.zeros(rows.size + 1, np.int64)
    np.cumsum(lens, out=rb[1:])
    idx = np.repeat(rowbeg[rows] - rb[:-1], lens) + np.arange(rb[-1], dtype=np.int64)
    return rb.astype(np.uint64), np.asarray(colind)[idx].astype(np.uint64), np.asarray(val, dtype=np.float64)[idx]


def local_system(rowbeg, colind, val, part, rank, world):
    """(owned_gid, local rowbeg, colind (global ids), val, Exchange with requests) of `rank`."""
    part = np.asarray(part, dtype=np.int32)
    owned = np.flatnonzero(part == rank).astype(np.int64)
    rb, ci, va = extract_rows(rowbeg, colind, val, owned)
    ghosts = capi.find_ghosts(owned, ci)
    ex = capi.Exchange(rank, world, owned, ghosts, part[ghosts])
    return owned, rb, ci, va, ex


def connect_in_process(exchanges):
    """All ranks' plans live in this process (local communicator): hand the requests over directly."""
    for r, er in enumerate(exchanges):
        for q, eq in enumerate(exchanges):
            if q != r:
                eq.set_send(r, er.request(q))       # what r wants from q is what q sends to r


def connect(ex, dist):
    """Transports every rank's requests to their owners through torch.distributed (gloo or nccl)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    req = [ex.request(q) if q != rank else np.zeros(0, np.int64) for q in range(world)]
    counts = torch.tensor([len(a) for a in req], dtype=torch.int64, device=dev)
    allc = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allc, counts)
    allc = torch.stack(allc).cpu().numpy()          # allc[r][q] = what r requests from q
    sends = [torch.from_numpy(req[q].copy()).to(dev) for q in range(world)]
    recvs = [torch.zeros(int(allc[q][rank]), dtype=torch.int64, device=dev) for q in range(world)]
    ops = []
    for q in range(world):
        if q == rank:
            continue
        if sends[q].numel():
            ops.append(dist.P2POp(dist.isend, sends[q], q))
        if recvs[q].numel():
            ops.append(dist.P2POp(dist.irecv, recvs[q], q))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    for q in range(world):
        if q != rank:
            ex.set_send(q, recvs[q].cpu().numpy())


def make_comm(dist, rank, world):
    """NCCL communicator of libdrops_b200; the unique id travels through torch.distributed."""
    uid = [capi.Comm.unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    return capi.Comm.nccl(rank, world, uid[0])


def brick_partition(fe, N, numbering, rank, world, comm, dist=None, upper_scale=1.0):
    """z-slab `rank` of the synthetic brick, generated on the device: (slab, exchange, DistMatrix)."""
    slab = capi.BrickSlab(fe, N, numbering, world, rank, upper_scale)
    ex = slab.exchange()
    if world > 1:
        connect(ex, dist)
    A = slab.matrix(comm, ex)
    return slab, ex, A


def brick_partition_in_process(fe, N, numbering, world, comm, upper_scale=1.0):
    slabs = [capi.BrickSlab(fe, N, numbering, world, r, upper_scale) for r in range(world)]
    exs = [s.exchange() for s in slabs]
    connect_in_process(exs)
    mats = [s.matrix(comm, e) for s, e in zip(slabs, exs)]
    return slabs, exs, mats
