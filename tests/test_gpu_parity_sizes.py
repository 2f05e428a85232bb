"""GPU parity at the benchmark tile sizes and on the product branches round 1 left untested (weak #1, #9):
nf_tile = 176 (BASELINE configs[0]) and 304 (configs[1-4]) against the oracle, -DPID_FLAG, move_grid_back, -DCOARSE_NGP, every overflow
status, the slab-decomposed coarse solve on one rank, and invariance of the result under the stream-overlap knobs."""
import numpy as np
import pytest

from cubep3m_b200 import default_config, ic
from tests.conftest import sort_records

pytestmark = pytest.mark.gpu


def _pair(cfg):
    from cubep3m_b200.lib import ParticleMesh
    from oracle import Oracle
    return ParticleMesh(cfg), Oracle(cfg)


def _rms_rel(g, r):
    rel = np.sqrt(((g[:, 3:] - r[:, 3:]) ** 2).sum(1)) / np.maximum(np.sqrt((r[:, 3:] ** 2).sum(1)), 1e-30)
    return float(np.sqrt(np.mean(rel ** 2))), float(rel.max())


def _one_step(cfg, xv, args, vel_tol):
    pm, o = _pair(cfg)
    x = xv.copy()
    x[:, 3:] = 0                       # report_force.f90:33-45: the velocity after the step IS the kick
    pm.upload_particles(x); o.set_particles(x)
    og, oo = pm.particle_mesh(*args), o.particle_mesh(*args)
    g, r = sort_records(pm.download_particles()), sort_records(o.get_particles())
    tg, to = pm.tile_counts(), o.tile_counts()
    pm.close(); o.close()
    assert og.np_local == oo.np_local == len(g) == len(r)
    assert og.np_with_ghosts == oo.np_with_ghosts and og.np_buf_max == oo.np_buf_max
    assert np.array_equal(g[:, :3], r[:, :3]), "positions after the step (bit-exact)"
    assert np.array_equal(tg, to), "per-tile particle counts (bit-exact)"
    rms, mx = _rms_rel(g, r)
    assert rms < vel_tol, (rms, mx)
    assert og.sum_rho_f == pytest.approx(oo.sum_rho_f, rel=1e-9)
    assert og.sum_rho_c == pytest.approx(oo.sum_rho_c, rel=1e-6)
    for f in ("dt_f_acc", "dt_c_acc", "dt_pp_acc", "dt_pp_ext_acc"):
        assert getattr(og, f) == pytest.approx(getattr(oo, f), rel=2e-4), f
    return rms


def test_full_step_config0_tile176(built):
    """BASELINE configs[0]: 128^3 particles, 256^3 fine mesh, tiles_node_dim = 2 (nf_tile = 176 = 16*11), PM only — one step from the
    z = 100 Zel'dovich ICs against the oracle: positions and tile counts bit-exact, rms relative force <= 1e-4 (north_star)."""
    cfg = default_config(nf_tile=176, tiles_node_dim=2, ppint=0, pp_ext=0)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=200.0, z_i=100.0, seed=12345)
    _one_step(cfg, xv, (0.3, 0.0, 0.0101, 8.0, (2.5, -7.0, 0.625)), 1e-4)


def test_full_step_tile304_pp_ext(built):
    """The benchmark tile (nf_tile = 304 = 16*19, the radix-19 kernels) with PPINT + PP_EXT on: one tile (tiles_node_dim = 1) of 128^3
    particles keeps the oracle at a few seconds. Same gates."""
    cfg = default_config(nf_tile=304, tiles_node_dim=1, ppint=1, pp_ext=1)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=100.0, z_i=50.0, seed=4)
    _one_step(cfg, xv, (0.3, 0.1, 0.02, 8.0, (-1.5, 4.25, 0.375)), 1e-4)


def test_pid_follows_particles(built):
    """-DPID_FLAG: ids travel through pack / unpack / scatter / compaction with their records (particle_pass.f90:150-153,
    delete_particles.f90:34): after a step every id sits on the oracle's position, bit-exact."""
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=1, pid=1)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=8)
    pid = (np.arange(len(xv), dtype=np.int64) * 7 + 3)
    pm, o = _pair(cfg)
    pm.upload_particles(xv, pid); o.set_particles(xv, pid=pid)
    args = (0.5, 0.3, 0.05, 8.0, (1.25, -0.5, 2.0))
    pm.particle_mesh(*args); o.particle_mesh(*args)
    g, gp = pm.download_particles(with_pid=True)
    r, rp = o.get_particles(with_pid=True)
    pm.close(); o.close()
    assert len(gp) == len(rp) == len(xv) and np.array_equal(np.sort(gp), np.sort(pid))
    g, r = g[np.argsort(gp)], r[np.argsort(rp)]
    assert np.array_equal(g[:, :3], r[:, :3])
    rms, mx = _rms_rel(g, r)
    assert rms < 1e-4, (rms, mx)


def test_move_grid_back_bit_exact(built):
    """move_grid_back.f90:20-23: x -= shake_offset, unfused, on the resident particles."""
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=9)
    pm, o = _pair(cfg)
    pm.upload_particles(xv); o.set_particles(xv)
    shake = (3.125, -7.3333, 0.0009765)
    pm.move_grid_back(shake); o.move_grid_back(shake)
    g, r = pm.download_particles(), o.get_particles()
    pm.close(); o.close()
    assert np.array_equal(g, r)
    assert not np.array_equal(g[:, :3], xv[:, :3])


def test_coarse_ngp_branch(built):
    """-DCOARSE_NGP (coarse_cic_mass.f90 / coarse_velocity.f90 with dx1 = 0, dx2 = 1)."""
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0, coarse_ngp=1)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=10)
    _one_step(cfg, xv, (0.5, 0.3, 0.05, 8.0, (1.25, -0.5, 2.0)), 1e-4)


def _status(fn):
    from cubep3m_b200.lib import Cubep3mError
    with pytest.raises(Cubep3mError) as e:
        fn()
    return e.value.status


def test_overflow_statuses_match_reference_aborts(built):
    """'not enough buffer space in pass' (particle_pass.f90:96-99), 'exceeded max_np in pass' (:136-139), 'exceeded max_llf'
    (particle_mesh_threaded.f90:280-283): the library returns the distinct status instead of truncating; the oracle aborts the same way."""
    from cubep3m_b200.lib import ParticleMesh
    from oracle import Oracle
    base = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)
    xv = ic.zeldovich_ics(base.nf_physical_dim, box=50.0, z_i=20.0, seed=11)
    args = (0.1, 0.1, 0.05, 8.0, (0.0, 0.0, 0.0))
    cases = [(dict(max_buf=6 * 100), 3), (dict(max_np=len(xv) + 1000), 4), (dict(max_llf=3), 5)]
    for kw, want in cases:
        cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0, **kw)
        x = xv.copy()
        if want == 5:
            x[:8, :3] = np.float32(40.0) + np.linspace(0.05, 0.4, 8, dtype=np.float32)[:, None]     # 8 particles in one fine cell > max_llf = 3
        pm, o = ParticleMesh(cfg), Oracle(cfg)
        pm.upload_particles(x); o.set_particles(x)
        assert _status(lambda: pm.particle_mesh(*args)) == want, kw
        with pytest.raises(RuntimeError):
            o.particle_mesh(*args)
        pm.close(); o.close()


def test_coarse_slab_pipeline_on_one_rank(built, monkeypatch):
    """The slab-decomposed coarse solve (coarse_slab.cuh: cube -> slab scatter, pencil transposes, halo scatter; W = 1 so every 'peer' is this
    GPU) must reproduce the whole-mesh solve: force_c within 2e-6 of its maximum, and the full step against the oracle as usual."""
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=12)
    args = (0.5, 0.3, 0.05, 8.0, (1.25, -0.5, 2.0))
    res = {}
    for mode in ("replicated", "slab"):
        monkeypatch.setenv("CUBEP3M_B200_COARSE", mode)
        pm = ParticleMesh(cfg)
        pm.upload_particles(xv)
        out = pm.particle_mesh(*args)
        res[mode] = (pm.force_c().copy(), sort_records(pm.download_particles()), out)
        pm.close()
    fa, fb = res["replicated"][0], res["slab"][0]
    assert np.abs(fa - fb).max() <= 2e-6 * np.abs(fa).max()
    assert np.array_equal(res["replicated"][1][:, :3], res["slab"][1][:, :3])
    dv = np.abs(res["replicated"][1][:, 3:] - res["slab"][1][:, 3:]).max()
    assert dv <= 1e-5 * np.abs(res["replicated"][1][:, 3:]).max(), ("velocities after the step, slab vs whole-mesh solve", dv)
    assert res["slab"][2].dt_c_acc == pytest.approx(res["replicated"][2].dt_c_acc, rel=1e-5)
    assert res["slab"][2].sum_rho_c == pytest.approx(res["replicated"][2].sum_rho_c, rel=1e-9)
    monkeypatch.setenv("CUBEP3M_B200_COARSE", "slab")
    _one_step(cfg, xv, args, 1e-4)


def test_stream_overlap_does_not_change_positions(built, monkeypatch):
    """Fine tiles in flight (1 or 2 streams) and the concurrent coarse stream read particle records while kick kernels rewrite their
    velocity words (the z word is rewritten with identical bits, 8-byte stores are not torn): positions must be bit-identical and
    velocities equal to fp32 rounding of the per-cell summation order, whatever the overlap."""
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=1)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=13)
    args = (0.5, 0.3, 0.05, 8.0, (1.25, -0.5, 2.0))
    res = []
    for streams in ("1", "2"):
        monkeypatch.setenv("CUBEP3M_B200_TILE_STREAMS", streams)
        pm = ParticleMesh(cfg)
        pm.upload_particles(xv)
        pm.particle_mesh(*args)
        res.append(sort_records(pm.download_particles()))
        pm.close()
    assert np.array_equal(res[0][:, :3], res[1][:, :3])
    rms, mx = _rms_rel(res[0], res[1])
    assert rms < 1e-5, (rms, mx)
