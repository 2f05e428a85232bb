"""N > 1 logic on the CPU: (1) the multi-rank oracle is decomposition-invariant (8 ranks of a 2x2x2 grid give the same global
result as one rank on the same box), (2) non-cubic rank grids conserve mass / particles, (3) a world_size-2 gloo run of the
particle-pass exchange protocol (plus-going pairs with from-minus even when both neighbours are the same peer) reproduces the
oracle's post-pass particle sets bit-exactly."""
import os

import numpy as np
import pytest

from cubep3m_b200 import default_config, ic
from cubep3m_b200 import topology as topo
from oracle import Oracle
from tests.conftest import sort_records


def _global_ics(nc, seed=5):
    xv = ic.zeldovich_ics(nc, box=50.0, z_i=20.0, seed=seed)
    xv[:, 3:] *= 2.0
    return xv


def test_topology_matches_reference_cart():
    """rank = x + D*y + D^2*z with cart_coords(1) = z (mpi_initialization.f90:55-76)."""
    g = (2, 2, 2)
    assert topo.rank_coords(5, g) == (1, 0, 1)
    assert topo.neighbours(0, g) == (1, 1, 2, 2, 4, 4)
    assert topo.neighbours(3, (2, 2, 1)) == (2, 2, 1, 1, 3, 3)
    assert topo.grid_for_world(4) == (2, 2, 1)


def test_decomposition_invariance():
    """8 ranks (nf_tile=80, T=2 -> 64^3 cells each) vs 1 rank (nf_tile=112, T=2 -> 128^3 cells) on the same global particles."""
    xv = _global_ics(128)
    dt, dt_old, a_mid, mass_p, off = 0.4, 0.2, 0.05, 8.0, (1.25, -0.5, 2.0)
    c1 = default_config(nf_tile=112, tiles_node_dim=2)
    o1 = Oracle(c1)
    o1.set_particles(xv)
    out1 = o1.particle_mesh(dt, dt_old, a_mid, mass_p, off)
    r1 = o1.get_particles()
    o1.close()
    c8 = default_config(nf_tile=80, tiles_node_dim=2, nodes_dim=2)
    o8 = Oracle(c8)
    # ranks own particles by their position AFTER the drift (the pass moves leavers to the neighbour anyway, so assign by start)
    parts = topo.split_global(xv, c8.mT, c8.grid)
    for r, p in enumerate(parts):
        o8.set_particles(p, rank=r)
    out8 = o8.particle_mesh(dt, dt_old, a_mid, mass_p, off)
    res = []
    for r in range(8):
        p = o8.get_particles(rank=r)
        cc = topo.rank_coords(r, c8.grid)
        for a in range(3):
            p[:, a] += np.float32(cc[a] * c8.mT)
        res.append(p)
    r8 = np.concatenate(res)
    o8.close()
    assert out8.np_total == out1.np_total == len(xv) == len(r8)
    assert out8.sum_rho_f == pytest.approx(out1.sum_rho_f) and out8.sum_rho_c == pytest.approx(out1.sum_rho_c, rel=1e-6)
    a, b = r1[np.lexsort((r1[:, 2], r1[:, 1], r1[:, 0]))], r8[np.lexsort((r8[:, 2], r8[:, 1], r8[:, 0]))]
    # positions: local vs global coordinates round differently in fp32 -> tolerance; a handful of particles may swap sort order
    dpos = np.abs(np.sort(a[:, 0]) - np.sort(b[:, 0])).max()
    assert dpos < 1.1e-3
    # velocities: match particles through a KD-free trick — sort by the (unchanged) initial velocity is impossible after the kick,
    # so compare moments and limiters instead
    assert np.allclose(a[:, 3:].mean(0), b[:, 3:].mean(0), atol=2e-5)
    assert np.sqrt((a[:, 3:] ** 2).sum(1)).mean() == pytest.approx(np.sqrt((b[:, 3:] ** 2).sum(1)).mean(), rel=1e-4)
    for f in ("dt_f_acc", "dt_c_acc"):
        assert getattr(out8, f) == pytest.approx(getattr(out1, f), rel=2e-3), f


@pytest.mark.parametrize("grid,nf_tile,lrck", [((2, 1, 1), 112, 1), ((2, 2, 1), 80, 0)])
def test_noncubic_grid_invariants(grid, nf_tile, lrck):
    # LRCKCORR needs every coarse dimension > 16 (kernel_initialization.f90:573-581), hence nf_tile=112 when it is on
    cfg = default_config(nf_tile=nf_tile, tiles_node_dim=2, nodes_dim_xyz=grid, lrckcorr=lrck)
    o = Oracle(cfg)
    one = _global_ics(cfg.mT, seed=9)           # every rank holds the same periodic box: the global field is its replication
    for r in range(cfg.nodes):
        o.set_particles(one, rank=r)
    out = o.particle_mesh(0.4, 0.2, 0.05, 8.0, (3.0, -1.5, 0.25))
    assert out.np_total == cfg.nodes * len(one)
    assert out.sum_rho_f == pytest.approx(cfg.nodes * float(cfg.mT) ** 3)
    assert out.sum_rho_c == pytest.approx(cfg.nodes * float(cfg.mT) ** 3, rel=1e-6)
    # replication symmetry: all ranks end with the same particle set
    ref = sort_records(o.get_particles(rank=0))
    for r in range(1, cfg.nodes):
        got = sort_records(o.get_particles(rank=r))
        assert np.array_equal(got[:, :3], ref[:, :3])
        assert np.abs(got[:, 3:] - ref[:, 3:]).max() < 1e-4 * np.abs(ref[:, 3:]).max()
    o.close()


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = default_config(nf_tile=80, tiles_node_dim=2, nodes_dim_xyz=(2, 1, 1), lrckcorr=0)
    grid = cfg.grid
    nb = topo.neighbours(rank, grid)
    xv = np.load(os.path.join(tmp, f"in{rank}.npy"))

    def make_exchange(axis):
        minus, plus = nb[2 * axis], nb[2 * axis + 1]

        def ex(plus_going, minus_going):
            if grid[axis] == 1:
                return plus_going, minus_going
            # same order as exchange_axis in lib.cu: sends [plus-going -> plus, minus-going -> minus], recvs [from minus, from plus]
            cnt = torch.tensor([len(plus_going), len(minus_going)], dtype=torch.int64)
            rc = torch.zeros(2, dtype=torch.int64)
            ops = [dist.P2POp(dist.isend, cnt[0:1], plus), dist.P2POp(dist.isend, cnt[1:2], minus),
                   dist.P2POp(dist.irecv, rc[0:1], minus), dist.P2POp(dist.irecv, rc[1:2], plus)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            sp, sm = torch.from_numpy(np.ascontiguousarray(plus_going)), torch.from_numpy(np.ascontiguousarray(minus_going))
            rp, rm = torch.empty((int(rc[0]), 6)), torch.empty((int(rc[1]), 6))
            ops = []
            if len(sp): ops.append(dist.P2POp(dist.isend, sp, plus))
            if len(sm): ops.append(dist.P2POp(dist.isend, sm, minus))
            if len(rp): ops.append(dist.P2POp(dist.irecv, rp, minus))
            if len(rm): ops.append(dist.P2POp(dist.irecv, rm, plus))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            return rp.numpy(), rm.numpy()
        return ex

    for axis in range(3):
        xv = topo.pass_axis_host(xv, axis, cfg.mT, cfg.nf_buf, cfg.eps, make_exchange(axis))
    np.save(os.path.join(tmp, f"out{rank}.npy"), xv)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_pass_protocol_world2(tmp_path):
    import torch.multiprocessing as mp
    cfg = default_config(nf_tile=80, tiles_node_dim=2, nodes_dim_xyz=(2, 1, 1), lrckcorr=0)
    rng = np.random.default_rng(4)
    o = Oracle(cfg)
    for r in range(2):
        xv = np.zeros((3000, 6), np.float32)
        xv[:, :3] = rng.random((3000, 3), dtype=np.float32) * np.float32(cfg.mT)
        xv[:, 3:] = rng.standard_normal((3000, 3)).astype(np.float32)
        xv[:5, 0] = [0.0, 0.0005, -0.0003, cfg.mT - 1e-3, 23.9999]       # exercise the eps nudge and the cut boundaries
        np.save(tmp_path / f"in{r}.npy", xv)
        o.set_particles(xv, rank=r)
    o.link_list()
    o.particle_pass()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"out{r}.npy")
        ref = o.get_particles(rank=r)
        assert len(got) == len(ref)
        assert np.array_equal(sort_records(got), sort_records(ref)), f"rank {r}: post-pass particle set"
    o.close()
