"""One rank of the NCCL parity check (tests/test_gpu_dist.py::test_two_ranks_over_nccl, also run
by tools/gpu_*.sh under torchrun with 2/4/8 ranks): z-slab partition of a P2 brick over NCCL against
the single-GPU operator that every rank builds for itself."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from drops_b200 import capi, partition
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    capi.init(local)
    comm = partition.make_comm(dist, rank, world)
    for fe, N, numbering in [(2, 24, 1), (2, 20, 0), (1, 40, 0)]:
        n, nnz = capi.brick_sizes(fe, N)
        slab, ex, A = partition.brick_partition(fe, N, numbering, rank, world, comm, dist)
        owned = slab.owned()
        G = capi.Matrix.brick(fe, N, numbering)                 # single-GPU operator, same device
        i = np.arange(n, dtype=np.uint64)
        xin = ((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)).astype(np.float64) / 4294967296.0 - 0.5
        x, y = capi.Vector.from_host(xin[owned]), capi.Vector(owned.size)
        capi.dist_spmv([A], [x], [y])
        yref = G.spmv(xin)
        assert np.array_equal(y.download(), yref[owned]), "partitioned SpMV differs from the single-GPU rows"
        d = capi.dist_dot([A], [x], [y])
        dref = capi.dot(capi.Vector.from_host(xin), capi.Vector.from_host(yref))
        assert abs(d - dref) <= 4e-16 * float(np.abs(xin * yref).sum()), (d, dref)
        xs = capi.Vector(owned.size)
        slab.solution(xs)
        b = capi.Vector(owned.size)
        capi.dist_spmv([A], [xs], [b])
        xstar = capi.brick_solution_host(fe, N, numbering)
        bref = G.spmv(xstar)
        assert np.array_equal(b.download(), bref[owned])
        ref = G.pcg_jacobi(np.zeros(n), bref, 3000, 1e-10, True, hist=True)
        x0 = capi.Vector(owned.size)
        conv, it, res, hist = capi.dist_pcg([A], [x0], [b], 3000, 1e-10, True, hist=True)
        assert conv and ref[0] and abs(it - ref[1]) <= 1, (it, ref[1])
        err = np.linalg.norm(x0.download() - ref[3][owned]) / np.linalg.norm(ref[3][owned])
        assert err <= 1e-10, err
        k = min(it, ref[1])
        assert np.allclose(hist[:k + 1], ref[4][:k + 1], rtol=1e-6, atol=1e-15)
        # every rank saw the same iteration count and residual
        t = torch.tensor([float(it), res], dtype=torch.float64, device="cuda")
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        assert all(torch.equal(a, t) for a in tl)
        # GMRES(20) + Jacobi over NCCL against the single-GPU driver
        gref = G.gmres_jacobi(np.zeros(n), bref, 20, 3000, 1e-10, True, 1.0, True)
        xg = capi.Vector(owned.size)
        gconv, git, gres = capi.dist_gmres([A], [xg], [b], 20, 3000, 1e-10, True)
        assert gconv and gref[0] and abs(git - gref[1]) <= 1, (git, gref[1])
        gerr = np.linalg.norm(xg.download() - gref[3][owned]) / np.linalg.norm(gref[3][owned])
        assert gerr <= 1e-8, gerr
        t = torch.tensor([float(git), gres], dtype=torch.float64, device="cuda")
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        assert all(torch.equal(a, t) for a in tl)
        # the same with classical Gram-Schmidt (ModGMRES, krylovsolver.h:465-494): one wide reduction per column
        capi.set_option("gmres.ortho", 1)
        xc = capi.Vector(owned.size)
        cconv, cit, cres = capi.dist_gmres([A], [xc], [b], 20, 3000, 1e-10, True)
        capi.set_option("gmres.ortho", 0)
        assert cconv and abs(cit - git) <= 1, (cit, git)
        cerr = np.linalg.norm(xc.download() - gref[3][owned]) / np.linalg.norm(gref[3][owned])
        assert cerr <= 1e-8, cerr
        # MINRES and PMINRES<JACPcCL> over the ranks against the single-GPU driver
        merr = []
        for precond, pc in ((0, capi.PC_DUMMY), (1, capi.PC_JACOBI)):
            mref = G.minres(np.zeros(n), bref, 3000, 1e-9, True) if not precond else G.pminres_jacobi(np.zeros(n), bref, 3000, 1e-9, True)
            xm = capi.Vector(owned.size)
            mconv, mit, mres = capi.dist_minres([A], [xm], [b], 3000, 1e-9, True, precond, pc)
            assert mconv and mref[0] and abs(mit - mref[1]) <= 1, (mit, mref[1])
            merr.append(np.linalg.norm(xm.download() - mref[3][owned]) / np.linalg.norm(mref[3][owned]))
            assert merr[-1] <= 1e-8, merr
            t = torch.tensor([float(mit), mres], dtype=torch.float64, device="cuda")
            tl = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(tl, t)
            assert all(torch.equal(a, t) for a in tl)
        fmt = A.format()
        assert not fmt["p2p_fault"]
        ws = A.wait_stats()
        if fmt["p2p"] and world > 1:            # every halo exchange and every reduction of the solves above was counted
            assert ws["halo"][1] > 0 and ws["halo"][0] >= 0.0, ws
        if fmt["p2p_reduce"] and world > 1:
            assert ws["reduce_pAp"][1] >= it and ws["reduce_rho_resid"][1] >= it and ws["reduce_other"][1] > 0, ws
            A.wait_stats(reset=True)
            assert A.wait_stats()["reduce_pAp"] == (0.0, 0)
        if fmt["p2p"]:                          # the same solve with the halo over ncclSend/ncclRecv: same bits
            capi.set_option("dist.p2p", 0); capi.set_option("dist.p2p_reduce", 0)
            x1 = capi.Vector(owned.size)
            c1, it1, res1 = capi.dist_pcg([A], [x1], [b], 3000, 1e-10, True)
            capi.set_option("dist.p2p", 1); capi.set_option("dist.p2p_reduce", 1)
            assert it1 == it and res1 == res and np.array_equal(x1.download(), x0.download())
        if rank == 0:
            print(f"P{fe} N={N} numbering={numbering} world={world} p2p={fmt['p2p']} p2p_reduce={fmt['p2p_reduce']}: iters {it} (single GPU {ref[1]}), "
                  f"rel.diff {err:.2e}, gmres iters {git} (single GPU {gref[1]}), rel.diff {gerr:.2e}, cgs iters {cit}, rel.diff {cerr:.2e}, "
                  f"minres/pminres rel.diff {merr[0]:.2e}/{merr[1]:.2e}, "
                  f"ghosts {ex.n_ghost}, info {A.info()}", flush=True)
    dist.barrier()
    print("WORKER_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
