"""The communication plan of the slab-decomposed coarse solve (cubep3m_b200/coarse_slab.py) on the CPU: simulated ranks on the
(2,1,1), (2,2,1) and (2,2,2) grids against the global solve, message sizes / contiguity as the plan promises, and a world-2 gloo run in which
the four exchanges really cross process boundaries."""
import os

import numpy as np
import pytest

from cubep3m_b200 import coarse_slab as cs
from cubep3m_b200 import topology as topo


def _inputs(plan, seed):
    rng = np.random.default_rng(seed)
    cubes = [rng.random((plan.nc,) * 3).astype(np.float32) for _ in range(plan.W)]
    Nx, Ny, Nz = plan.N
    kern = rng.standard_normal((3, Nz, Ny, plan.hc)).astype(np.float32)
    return cubes, kern


@pytest.mark.parametrize("grid,nc", [((2, 1, 1), 8), ((2, 2, 1), 8), ((2, 2, 2), 8), ((2, 2, 2), 16)])
def test_simulated_ranks_match_global_solve(grid, nc):
    plan = cs.SlabPlan(grid, nc)
    cubes, kern = _inputs(plan, 11)
    got = cs.solve_simulated(plan, cubes, kern)
    ref = cs.solve_global(plan, cubes, kern)
    scale = max(np.abs(r).max() for r in ref)
    for g, r in zip(got, ref):
        assert np.abs(g - r).max() < 2e-6 * scale


def test_plan_shapes_and_contiguity():
    plan = cs.SlabPlan((2, 2, 2), 128)                     # 8 x 256^3 particles: global coarse mesh 256^3
    assert (plan.zs, plan.ys, plan.hc) == (32, 32, 129)
    for r in range(plan.W):
        sends = plan.cube_sends(r)
        assert len(sends) == 4 and [z0 for _, z0 in sends] == [0, 32, 64, 96]          # 4 contiguous chunks of 32 planes
        assert all((d * plan.zs) // plan.nc == topo.rank_coords(r, plan.grid)[2] for d, _ in sends)
        assert len(plan.cube_sources(r)) == 4                                           # Dx*Dy cubes share a slab
        # every destination's nc+2 halo planes are served exactly once
        for d in range(plan.W):
            hz = sorted(h for q in range(plan.W) for _, h in plan.halo_planes(q, d))
            assert hz == list(range(plan.nc + 2))
    with pytest.raises(ValueError):
        cs.SlabPlan((2, 2, 2), 6)                          # Nz = 12 is not a multiple of the 8 ranks

def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = cs.SlabPlan((2, 1, 1), 8)
    cubes, kern = _inputs(plan, 23)

    def exchange(tag, msgs):
        # shapes are known to both sides from the plan; here they travel as objects to keep the test short
        box = [None] * world
        dist.all_gather_object(box, {d: np.ascontiguousarray(a) for d, a in msgs.items()})
        return {s: box[s][rank] for s in range(world) if rank in box[s]}

    f = cs.solve_rank(plan, rank, cubes[rank], kern[:, :, rank * plan.ys:(rank + 1) * plan.ys, :], exchange)
    np.save(os.path.join(tmp, f"f{rank}.npy"), f)
    dist.destroy_process_group()


def test_world2_gloo_exchanges(tmp_path):
    import torch.multiprocessing as mp
    port = 29650 + os.getpid() % 200
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    plan = cs.SlabPlan((2, 1, 1), 8)
    cubes, kern = _inputs(plan, 23)
    ref = cs.solve_global(plan, cubes, kern)
    for r in range(2):
        got = np.load(tmp_path / f"f{r}.npy")
        assert np.abs(got - ref[r]).max() < 2e-6 * np.abs(ref[r]).max()
