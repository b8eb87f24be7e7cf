"""N > 1 host logic on CPU: two processes on the `gloo` backend (world_size 2, 127.0.0.1).  Each rank reads its own share of
the mesh, all-gathers the process-boundary node ids, builds its halo lists with the product's own host routine
(prfdd_halo_build_lists, what Domain::setup_halo calls) and performs the exchange the GPU path performs -- pack, send/recv,
add in ascending-rank order -- with gloo point-to-point.  Checked against the oracle's gs_add on simulated ranks: the summed
node vector, the 1/multiplicity weights, and a full dssum of a random vector."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, mesh_dir, N, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr
        from oracle import domain as od
        L = pr.lib()
        R = od.DomainRank(mesh_dir, N, rank)                 # this rank's share: reader, boundary-first node order, Q / Qt
        gathered = [None] * world
        dist.all_gather_object(gathered, R.boundary_nodes)
        offsets = np.zeros(world + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([g.size for g in gathered])
        ids = np.concatenate(gathered).astype(np.int64)
        peers = np.zeros(world, np.int32); pcount = np.zeros(world, np.int32); poff = np.zeros(world, np.int32)
        idx = np.zeros(max(R.boundary_nodes.size, 1) * world, np.int32)
        total = C.c_longlong(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        npeers = L.prfdd_halo_build_lists(C.c_int(rank), C.c_int(world), vp(ids), vp(offsets), vp(peers), vp(pcount), vp(poff), vp(idx), C.c_longlong(idx.size), C.byref(total))
        assert npeers >= 0

        def halo_add(nodes):
            """the device exchange of Domain::halo_exchange, with gloo instead of NCCL"""
            recv = {}
            reqs = []
            for k in range(npeers):
                p = int(peers[k]); sl = idx[poff[k]:poff[k] + pcount[k]]
                send = torch.from_numpy(np.ascontiguousarray(nodes[sl]))
                buf = torch.zeros(int(pcount[k]), dtype=torch.float64)
                recv[p] = (buf, sl)
                reqs.append(dist.isend(send, dst=p)); reqs.append(dist.irecv(buf, src=p))
            for r in reqs:
                r.wait()
            nb = R.num_bdary_nodes
            acc = np.zeros(nb)
            for p in sorted(recv):
                if p < rank:
                    acc[recv[p][1]] += recv[p][0].numpy()
            nodes[:nb] = acc + nodes[:nb]
            for p in sorted(recv):
                if p > rank:
                    nodes[recv[p][1]] += recv[p][0].numpy()

        rng = np.random.default_rng(100 + rank)
        u = rng.standard_normal(R.num_local_points)
        nodes = np.zeros(R.num_local_nodes); R.Qt.multiply(nodes, u)
        mult = np.zeros(R.num_local_nodes); R.Qt.multiply(mult, np.ones(R.num_local_points))
        halo_add(nodes); halo_add(mult)
        out.put((rank, u, nodes, mult, int(npeers), peers[:npeers].tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,nel,N", [(2, 6, 3), (3, 4, 2)])
def test_halo_exchange_world_size_2(prfdd, tmp_path, dim, nel, N):
    import torch.multiprocessing as mp
    from oracle import domain as od
    d = str(tmp_path)
    prfdd.mesh_generate_box(d, dim, nel, N, 2, 0.03)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, d, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W = od.DomainWorld(d, N, 2)
    us = [r[1] for r in res]
    for Rk, u in zip(W.ranks, us):
        Rk.Qt.multiply(Rk.work[0], u)
    W._gs_add([Rk.work[0] for Rk in W.ranks])
    for (rank, u, nodes, mult, npeers, peers), Rk in zip(res, W.ranks):
        assert npeers == 1 and peers == [1 - rank]
        assert np.array_equal(nodes, Rk.work[0][:Rk.num_local_nodes])          # same sums, same (ascending-rank) order
        assert np.array_equal(1.0 / mult, Rk.assembled_weight)                  # 1/multiplicity (domain.tpp:296-302)
    # both holders of a shared node end up with bit-identical values
    shared = {}
    for (rank, u, nodes, mult, _, _), Rk in zip(res, W.ranks):
        for s, gid in enumerate(Rk.boundary_nodes.tolist()):
            shared.setdefault(gid, []).append(nodes[s])
    assert all(len(v) == 2 and v[0] == v[1] for v in shared.values())


def test_halo_lists_three_ranks_consistency(prfdd):
    """pure-host check with more ranks than GPUs: lists of a pair are mirror images (same ids, same order)."""
    L = prfdd.lib()
    rng = np.random.default_rng(4)
    world = 5
    pool = rng.choice(10_000, 400, replace=False).astype(np.int64)
    per = [np.unique(rng.choice(pool, 150)) for _ in range(world)]
    per = [rng.permutation(p) for p in per]
    offsets = np.zeros(world + 1, np.int64); offsets[1:] = np.cumsum([p.size for p in per])
    ids = np.concatenate(per)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lists = []
    for r in range(world):
        peers = np.zeros(world, np.int32); pc = np.zeros(world, np.int32); po = np.zeros(world, np.int32)
        idx = np.zeros(per[r].size * world, np.int32); total = C.c_longlong(0)
        n = L.prfdd_halo_build_lists(C.c_int(r), C.c_int(world), vp(ids), vp(offsets), vp(peers), vp(pc), vp(po), vp(idx), C.c_longlong(idx.size), C.byref(total))
        lists.append({int(peers[k]): per[r][idx[po[k]:po[k] + pc[k]]] for k in range(n)})
        assert list(peers[:n]) == sorted(peers[:n])
    for a in range(world):
        for b, g in lists[a].items():
            assert np.array_equal(g, lists[b][a]) and np.all(np.diff(g) > 0)
            assert set(g.tolist()) == set(per[a].tolist()) & set(per[b].tolist())
