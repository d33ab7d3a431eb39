"""Two ranks on two GPUs of one box (NCCL): the slab-partitioned path end to end — interpolate per
slab, interface plane over both transports (NVLink peer-memory link with CUDA IPC, NCCL
send/recv), evaluation on the slab, gather of the distributed vector — against the serial oracle.
Skipped on a single-GPU box (the world_size-2 logic itself is covered on CPU/gloo by
tests/test_dist_cpu.py and the peer link by a single-process test in tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import orc
        import dune_functions_b200 as dfx
        from dune_functions_b200 import slab
        whole = dfx.YaspGrid([3, 2, 6], upper=[1.0, 0.5, 2.0])
        tree = dfx.lagrange(2)
        og = orc.OracleBasis(whole, tree)
        f = dfx.quadratic(0.5, [1.0, -2.0, 3.0], [1.0, 0.5, -1.0, 2.0, 0.25, -0.5])
        serial = og.interpolate(f)
        mine = whole.slab(rank, world)
        b = dfx.makeBasis(mine, tree, device=rank)
        x = b.newVector(0.0)
        dfx.interpolate(b, x, f)
        ok = {}
        # gather of the distributed vector = the serial interpolation, bit for bit
        g = slab.gather_global(b, x, og.dimension())
        ok["gather"] = bool(np.array_equal(g.cpu().numpy(), serial))
        # both transports restore a destroyed interface plane
        full = x.clone()
        for name, ex in (("peer", slab.PeerHaloExchange(b, rank, world)), ("nccl", slab.HaloExchange(b, rank, world, device=x.device))):
            for step in range(3):                       # several steps: both receive buffers of the link
                y = full.clone()
                if rank < world - 1:
                    for begin, count in slab.halo_ranges(b, 1):
                        y[begin:begin + count] = float("nan")
                ex(y)
                torch.cuda.synchronize()
                ok[f"{name}{step}"] = bool(torch.equal(y, full))
            if name == "peer":
                ok["peer_error"] = ex.error() == 0
                dist.barrier()
                ex.close()
        # evaluation on the slab = the slab's elements of the serial evaluation
        pts, _ = dfx.tensor_gauss(3, 3)
        v = dfx.makeDiscreteGlobalBasisFunction(b, x).evaluate(pts).cpu().numpy()
        vs = og.evaluate(serial, pts)
        nz = 6 // world
        per_layer = 3 * 2
        ref = vs[rank * nz * per_layer:(rank + 1) * nz * per_layer]
        ok["evaluate"] = bool(np.abs(v - ref).max() <= 1e-12)
        results[rank] = ok
    finally:
        dist.destroy_process_group()


def test_two_gpu_slab_run_matches_the_serial_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 1000
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
        assert len(results) == 2
        for r in range(2):
            bad = [k for k, v in results[r].items() if not v]
            assert not bad, f"rank {r}: {bad}"
