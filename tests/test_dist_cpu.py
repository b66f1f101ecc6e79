"""CPU (gloo, world_size 2) tests of the host-side slab logic used by the multi-GPU path: slab bookkeeping, the
out-of-band channel that ships the NCCL unique id, slab-wise image generation and the MAX-over-ranks error norm."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import fans_oracle as fo
from fans_b200 import dist as fdist
from fans_b200 import simple


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the 128-byte id reaches every rank unchanged
        payload = bytes(range(128)) if rank == 0 else None
        got = fdist.broadcast_bytes(payload)
        assert got == bytes(range(128))
        # 2. slab-wise generated image == slab of the full image
        dims = [16, 8, 8]
        x0, n0 = fdist.slab(dims[0], world, rank)
        ms = simple.ellipsoid_microstructure(dims, x0, n0)
        full = simple.ellipsoid_microstructure(dims)
        assert np.array_equal(ms, full[x0:x0 + n0])
        # 3. error norm: local norm + all-reduce MAX (solver.h:430) == the oracle's n_ranks emulation
        r = np.random.default_rng(3).standard_normal((16, 8, 8, 3))
        loc = r[x0:x0 + n0]
        t = torch.tensor([np.abs(loc).sum(), (loc * loc).sum(), np.abs(loc).max()], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, t.tolist()))
    finally:
        dist.destroy_process_group()


def test_slab_bookkeeping():
    assert fdist.slab(512, 8, 3) == (192, 64)
    with pytest.raises(ValueError):
        fdist.slab(10, 4, 0)
    fdist.check_decomposition([32, 32, 32], 8)
    with pytest.raises(ValueError, match="Number of processes too large"):
        fdist.check_decomposition([16, 32, 32], 8)   # reader.cpp:306: n_x/4 < world_size
    with pytest.raises(ValueError):
        fdist.check_decomposition([32, 32, 32], 3)


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r = np.random.default_rng(3).standard_normal((16, 8, 8, 3))
    mats = [{"phases": [0, 1], "matmodel": "LinearElasticIsotropic", "material_properties": {"bulk_modulus": [62.5, 222.222], "shear_modulus": [28.8462, 166.6667]}}]
    for k, measure in enumerate(["L1", "L2", "Linfinity"]):
        sol = fo.OracleSolver(simple.ellipsoid_microstructure([16, 8, 8]), [1, 1, 1], "mechanical", mats, "HEX8", "cg", "small",
                              {"measure": measure, "type": "absolute", "tolerance": 1e-10}, 1, n_ranks=world)
        want = sol.compute_error(r)
        for rank in range(world):
            got = res[rank][k] if measure != "L2" else np.sqrt(res[rank][k])
            assert abs(got - want) <= 1e-12 * abs(want)
