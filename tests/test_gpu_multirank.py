"""Multi-GPU parity: one process per GPU over NCCL (particle_pass send/recv, all-gathered coarse solve, all-reduced limiters)
against the multi-rank CPU oracle. Needs >= 2 GPUs; run with `gpurun --gpus 2|4|8 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import os

import numpy as np
import pytest

from cubep3m_b200 import default_config, ic
from cubep3m_b200 import topology as topo
from tests.conftest import sort_records

pytestmark = pytest.mark.gpu


def _worker(rank, world, grid, nf_tile, uid, tmp, lrck):
    import torch
    torch.cuda.set_device(rank)
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=nf_tile, tiles_node_dim=2, nodes_dim_xyz=grid, rank=rank, local_gpu=rank, lrckcorr=lrck, pp_ext=1)
    pm = ParticleMesh(cfg, nccl_id=uid, world_size=world)
    xv = np.load(os.path.join(tmp, f"in{rank}.npy"))
    pm.upload_particles(xv)
    # a halofind step's sequence (cubepm.f90:193-198,228) on the initial particles: the peak pass sees the neighbours' ghosts
    pm.link_list(); pm.particle_pass()
    pk, cft = pm.halofind_peaks(8.0, 4.0, True, True)
    np.save(os.path.join(tmp, f"peaks{rank}.npy"), pk)
    np.save(os.path.join(tmp, f"cft{rank}.npy"), np.array(cft))
    pm.delete_particles()
    outs = []
    for step in range(2):
        out = pm.particle_mesh(0.4, 0.2 * step + 0.1, 0.05, 8.0, (3.0, -1.5, 0.25))
        outs.append([out.np_local, out.np_with_ghosts, out.np_total, out.dt_f_acc, out.dt_pp_acc, out.dt_pp_ext_acc, out.dt_c_acc,
                     out.sum_rho_f, out.sum_rho_c, out.np_buf_max])
        np.save(os.path.join(tmp, f"out{rank}_s{step}.npy"), pm.download_particles())
    np.save(os.path.join(tmp, f"scal{rank}.npy"), np.array(outs, np.float64))
    pm.close()


@pytest.mark.parametrize("case", ["native", "nccl_pass_replicated_coarse"])
def test_multi_gpu_step_matches_oracle(built, tmp_path, case, monkeypatch):
    """native: particle_pass packed into the neighbour's memory + slab-decomposed coarse solve over peer memory (coarse_slab.cuh).
    nccl_pass_replicated_coarse: the fallbacks — ncclSend/Recv particle_pass inside the step (CUBEP3M_B200_P2P=0) and the all-gathered
    replicated coarse solve (CUBEP3M_B200_COARSE=replicated)."""
    if case != "native":
        monkeypatch.setenv("CUBEP3M_B200_P2P", "0")
        monkeypatch.setenv("CUBEP3M_B200_COARSE", "replicated")
    import torch
    import torch.multiprocessing as mp
    from cubep3m_b200.lib import get_unique_id
    from oracle import Oracle
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if ndev >= 8 else (4 if ndev >= 4 else 2)
    grid = topo.grid_for_world(world)
    nf_tile, lrck = (112, 1) if world == 2 else (80, 0)     # LRCKCORR needs every coarse dimension > 16
    if world == 8:
        nf_tile, lrck = 112, 1
    cfg = default_config(nf_tile=nf_tile, tiles_node_dim=2, nodes_dim_xyz=grid, lrckcorr=lrck, pp_ext=1)
    rng = np.random.default_rng(17)
    o = Oracle(cfg)
    base = ic.zeldovich_ics(cfg.mT, box=50.0, z_i=20.0, seed=3)
    for r in range(world):
        xv = base.copy()
        xv[:, 3:] *= np.float32(1.0 + 0.5 * r)           # ranks differ, so a mis-routed exchange cannot cancel out
        xv[:, :3] = np.mod(xv[:, :3] + rng.random(3).astype(np.float32), np.float32(cfg.mT))
        np.save(tmp_path / f"in{r}.npy", xv)
        o.set_particles(xv, rank=r)
    uid = get_unique_id()
    mp.spawn(_worker, args=(world, grid, nf_tile, uid, str(tmp_path), lrck), nprocs=world, join=True)
    o.link_list(); o.particle_pass()
    for r in range(world):
        op, oc = o.find_peaks(8.0, 4.0, True, True, rank=r)
        gp, gc = np.load(tmp_path / f"peaks{r}.npy"), np.load(tmp_path / f"cft{r}.npy")
        key = lambda a: a[np.lexsort((a["i"], a["j"], a["k"], a["tile"]))]
        gp, op = key(gp), key(op)
        assert len(op) > 100 and len(gp) == len(op), (r, len(gp), len(op))
        for f in ("tile", "i", "j", "k", "den"):
            assert np.array_equal(gp[f], op[f]), (r, f)
        for f in ("x", "y", "z"):
            assert np.array_equal(gp[f], op[f], equal_nan=True), (r, f)
        assert gc[0] == pytest.approx(oc[0], rel=1e-12) and gc[1] == pytest.approx(oc[1], rel=1e-12)
    o.delete_particles()
    for step in range(2):
        oo = o.particle_mesh(0.4, 0.2 * step + 0.1, 0.05, 8.0, (3.0, -1.5, 0.25))
        for r in range(world):
            got = np.load(tmp_path / f"out{r}_s{step}.npy")
            ref = o.get_particles(rank=r)
            sc = np.load(tmp_path / f"scal{r}.npy")[step]
            assert len(got) == len(ref) == int(sc[0]), (step, r)
            assert int(sc[2]) == oo.np_total
            g, f = sort_records(got), sort_records(ref)
            if step == 0:
                assert np.array_equal(g[:, :3], f[:, :3]), f"rank {r}: particle positions after the step (bit-exact from identical input)"
            k = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
            g, f = k(got), k(ref)
            if step == 0:
                rel = np.sqrt(((g[:, 3:] - f[:, 3:]) ** 2).sum(1)) / np.maximum(np.sqrt((f[:, 3:] ** 2).sum(1)), 1e-30)
                assert np.sqrt(np.mean(rel ** 2)) < 1e-4, (r, np.sqrt(np.mean(rel ** 2)))
            assert sc[3] == pytest.approx(oo.dt_f_acc, rel=1e-3) and sc[6] == pytest.approx(oo.dt_c_acc, rel=1e-3)
            assert sc[4] == pytest.approx(oo.dt_pp_acc, rel=1e-2)
            if step == 0:
                assert sc[5] == pytest.approx(oo.dt_pp_ext_acc, rel=2e-4), "dt_pp_ext_acc incl. margin particles (particle_mesh_threaded.f90:617,692)"
            assert sc[7] == pytest.approx(oo.sum_rho_f, rel=1e-9) and sc[8] == pytest.approx(oo.sum_rho_c, rel=1e-6)
    o.close()


def _power_worker(rank, world, grid, uid, tmp, box, shake):
    import torch
    torch.cuda.set_device(rank)
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=112, tiles_node_dim=2, nodes_dim_xyz=grid, rank=rank, local_gpu=rank, pp_ext=0)
    pm = ParticleMesh(cfg, nccl_id=uid, world_size=world)
    pm.upload_particles(np.load(os.path.join(tmp, f"pin{rank}.npy")))
    k, d2, sg = pm.cic_power(box, shake=shake, ngp_binning=True)
    np.save(os.path.join(tmp, f"pk{rank}.npy"), np.stack([k, d2, sg]))
    pm.close()


def test_distributed_cic_power_matches_host_twin(built, tmp_path):
    """cic_power over 8 ranks (nodes_dim = 2, the reference's cubic decomposition): CIC deposit added into the owners' z-slabs over NVLink, slab x / y passes,
    transpose into the peers' y-pencils, z pass, shell sums all-reduced — against the float64 host twin on the gathered particles (2e-4: fp32 mesh, atomics)."""
    import torch
    import torch.multiprocessing as mp
    from cubep3m_b200 import power
    from cubep3m_b200.lib import get_unique_id
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs (a cubic rank grid)")
    world, grid = 8, (2, 2, 2)
    cfg = default_config(nf_tile=112, tiles_node_dim=2, nodes_dim_xyz=grid, pp_ext=0)
    nc, box = cfg.mT * 2, 100.0
    shake = np.array([2.5, -1.25, 0.375], np.float32)
    xv = ic.zeldovich_ics(nc, box=box, z_i=10.0, seed=21)
    xs = xv.copy()
    xs[:, :3] = np.mod(xs[:, :3] + shake, np.float32(nc))
    xs[:, :3] = np.minimum(xs[:, :3], np.nextafter(np.float32(nc), np.float32(0)))
    for r, part in enumerate(topo.split_global(xs, cfg.mT, grid)):
        np.save(tmp_path / f"pin{r}.npy", part)
    mp.spawn(_power_worker, args=(world, grid, get_unique_id(), str(tmp_path), box, shake), nprocs=world, join=True)
    kh, dh, sh = power.power_spectrum(np.mod(xs[:, :3] - shake, np.float32(nc)), nc, box)
    for r in range(world):
        k, d2, sg = np.load(tmp_path / f"pk{r}.npy")
        assert np.allclose(k, kh, rtol=1e-12)
        rel = np.abs(d2 - dh) / np.maximum(np.abs(dh), 1e-30)
        assert rel.max() < 2e-4, (r, float(rel.max()))
        assert np.allclose(sg, sh, rtol=5e-3, atol=1e-12)
