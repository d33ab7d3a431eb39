"""world_size-2 (and 3) gloo runs of the multi-rank host logic: slab partition, halo
exchange protocol and result gather (dune_functions_b200/slab.py) with CPU tensors.  The
per-slab coefficient vectors come from the oracle, so no GPU is needed."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tree_name, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import orc
        import dune_functions_b200 as dfx
        from dune_functions_b200 import slab
        from cases import trees
        ids = None
        if tree_name.endswith("_ids"):
            # permuted face DOFs (dfx_basis_set_vertex_ids): every rank takes its share of the global id table
            tree = dfx.lagrange(int(tree_name[len("lagrange"):-len("_ids")]))
            ids = np.random.default_rng(5).permutation(4 * 3 * 7).astype(np.uint64)
        else:
            tree = trees(3)[tree_name]
        whole = dfx.YaspGrid([3, 2, 6], upper=[1.0, 0.5, 2.0], vertex_ids=ids)
        og = orc.OracleBasis(whole, tree)
        gx = np.random.default_rng(7).uniform(-1, 1, og.dimension())
        nz = 6 // world
        gflat = og.flatIndices().reshape(6, 6, -1).astype(np.int64)
        mine = whole.slab(rank, world)
        b = dfx.makeBasis(mine, tree)                 # host-side handle only: no compute call
        ol = orc.OracleBasis(dfx.YaspGrid(list(mine.local_n), vertex_ids=mine.vertex_ids), tree)
        lflat = ol.flatIndices().reshape(nz, 6, -1).astype(np.int64)
        # local vector: owned entries from the global vector, top plane unknown (NaN)
        x = np.full(b.dimension(), np.nan)
        x[lflat] = gx[gflat[nz * rank:nz * (rank + 1)]]
        full = x.copy()
        if rank < world - 1:
            for begin, count in slab.halo_ranges(b, 1):
                x[begin:begin + count] = np.nan
        xt = torch.from_numpy(x)
        ex = slab.HaloExchange(b, rank, world)
        ex(xt)
        ok_halo = bool(np.array_equal(xt.numpy(), full))
        # every rank rebuilds the global vector from the owned ranges
        g = slab.gather_global(b, xt, og.dimension())
        ok_gather = bool(np.array_equal(g.numpy(), gx))
        # the same onto one root only: the others send and get nothing back
        root = world - 1
        gr = slab.gather_global(b, xt, og.dimension(), root=root)
        ok_gather = ok_gather and ((gr is None) if rank != root else bool(np.array_equal(gr.numpy(), gx)))
        results[rank] = (ok_halo, ok_gather, slab.halo_size(b))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("tree_name", ["lagrange2", "lagrange3", "TH_FL(FL)", "lagrangeDG1", "lagrange4_ids"])
def test_slab_halo_and_gather_gloo(world, tree_name):
    port = 29500 + (os.getpid() + hash(tree_name) + world) % 2000
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, port, tree_name, results), nprocs=world, join=True)
        assert len(results) == world
        for r in range(world):
            ok_halo, ok_gather, n = results[r]
            assert ok_halo, f"rank {r}: halo exchange did not restore the interface plane"
            assert ok_gather, f"rank {r}: gathered vector differs from the global one"
        if tree_name == "lagrange2":
            assert results[0][2] == 7 * 5       # (2*3+1) x (2*2+1) lattice points in a z-plane
        if tree_name == "lagrangeDG1":
            assert results[0][2] == 0
