"""cic_power twin (cubep3m_b200/power.py): sanity against the IC generator's input spectrum, and the north-star's 0.1 % P(k) gate
between the GPU path and the oracle after a multi-step run (GPU part marked gpu)."""
import numpy as np
import pytest

from cubep3m_b200 import default_config, ic, power


def test_power_of_zeldovich_ics_matches_input_spectrum():
    nc, box, z_i = 64, 100.0, 50.0
    xv = ic.zeldovich_ics(nc, box=box, z_i=z_i, seed=2)
    k, d2, _ = power.power_spectrum(xv[:, :3], nc, box)
    a = 1.0 / (1.0 + z_i)
    lin = ic.delta2(k, a)
    sel = (k > 4 * 2 * np.pi / box) & (k < 0.25 * np.pi * nc / box)      # away from cosmic variance and from the lattice scale
    ratio = d2[sel] / lin[sel]
    assert 0.6 < np.median(ratio) < 1.4, np.median(ratio)


def test_uniform_lattice_has_no_power():
    nc = 32
    g = (np.arange(0, nc, 2) + 0.5).astype(np.float32)
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    k, d2, _ = power.power_spectrum(pos, nc, 100.0)
    assert np.abs(d2[: nc // 4 - 1]).max() < 1e-10     # below the lattice frequency a perfect lattice has zero power


@pytest.mark.gpu
def test_pk_gpu_vs_oracle_multistep(built):
    """North-star gate: the power spectrum after a multi-step run agrees with the reference path to 0.1 % per bin.
    5 steps from z=20 with the driver twin choosing dt; same shake offsets on both sides."""
    from cubep3m_b200.lib import ParticleMesh, clock_init, timestep, absorb_limiters
    from oracle import Oracle
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)
    box, z_i = 50.0, 20.0
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=box, z_i=z_i, seed=7)
    pm, o = ParticleMesh(cfg), Oracle(cfg)
    pm.upload_particles(xv); o.set_particles(xv)
    ca, cb = clock_init(z_i), clock_init(z_i)
    rng = np.random.default_rng(777)
    shake = np.zeros(3, np.float32)
    for step in range(5):
        timestep(ca); timestep(cb)
        off = ((rng.random(3, dtype=np.float32) - np.float32(0.5)) * np.float32(16.0) - shake).astype(np.float32)
        shake = shake + off
        og = pm.particle_mesh(ca.dt, ca.dt_old, ca.a_mid, 8.0, off)
        oo = o.particle_mesh(cb.dt, cb.dt_old, cb.a_mid, 8.0, off)
        absorb_limiters(ca, og); absorb_limiters(cb, oo)
        assert og.np_total == oo.np_total == len(xv)
        assert ca.dt == pytest.approx(cb.dt, rel=1e-3)
    g, r = pm.download_particles(), o.get_particles()
    pm.close(); o.close()
    nc = cfg.nf_physical_dim
    undo = lambda p: np.mod(p[:, :3] - shake, np.float32(nc))      # checkpoint.f90:92 writes xv - shake_offset
    kg, dg, _ = power.power_spectrum(undo(g), nc, box)
    kr, dr, _ = power.power_spectrum(undo(r), nc, box)
    rel = np.abs(dg - dr) / np.maximum(np.abs(dr), 1e-30)
    assert rel.max() < 1e-3, rel.max()


@pytest.mark.gpu
def test_device_cic_power_matches_host_twin(built):
    """cubep3m_b200_cic_power (deposit + FFT + shell binning on the GPU, fp32 mesh) against the float64 host twin on the same particles,
    for both shell-weight variants, with a non-zero shake offset to undo."""
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)          # nf_physical_dim = 128
    box, z_i = 50.0, 20.0
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=box, z_i=z_i, seed=11)
    nc = cfg.nf_physical_dim
    shake = np.array([3.25, -1.5, 0.75], np.float32)
    xs = xv.copy()
    xs[:, :3] = np.mod(xs[:, :3] + shake, np.float32(nc))
    pm = ParticleMesh(cfg)
    pm.upload_particles(xs)
    for ngp in (True, False):
        kg, dg, sg = pm.cic_power(box, shake=shake, ngp_binning=ngp)
        kh, dh, sh = power.power_spectrum(np.mod(xs[:, :3] - shake, np.float32(nc)), nc, box, ngp_binning=ngp)
        assert np.allclose(kg, kh, rtol=1e-12)
        rel = np.abs(dg - dh) / np.maximum(np.abs(dh), 1e-30)
        assert rel.max() < 2e-4, (ngp, rel.max())
        assert np.allclose(sg, sh, rtol=5e-3, atol=1e-12)
    pm.close()


@pytest.mark.gpu
def test_pk_config0_z100_to_z10(built):
    """BASELINE configs[0] as the reference runs it: 128^3 particles, 256^3 fine mesh, tiles_node_dim = 2 (nf_tile = 176), PM only, dist_init-style
    Zel'dovich ICs at z = 100 evolved to the z = 10 checkpoint by the driver twin (timestep.f90 chooses dt: ~230 steps of da/a <= 1 %, the last
    one cut to land on a_checkpoint, timestep.f90:128-137) — on the GPU path and on the CPU oracle, each with its own clock fed by its own
    limiters, same shake offsets. Gate (north_star): cic_power at the final checkpoint within 0.1 % per bin."""
    from cubep3m_b200.lib import ParticleMesh, clock_init, timestep, absorb_limiters
    from oracle import Oracle
    cfg = default_config(nf_tile=176, tiles_node_dim=2, ppint=0, pp_ext=0)
    box, z_i, z_f = 200.0, 100.0, 10.0
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=box, z_i=z_i, seed=12345)
    nc = cfg.nf_physical_dim
    mass_p = float(np.float32(nc) ** 3 / np.float32(len(xv)))
    pm, o = ParticleMesh(cfg), Oracle(cfg)
    pm.upload_particles(xv); o.set_particles(xv)
    a_t = float(np.float32(1.0) / np.float32(1.0 + z_f))
    ca, cb = clock_init(z_i, ppint=0, a_target=a_t), clock_init(z_i, ppint=0, a_target=a_t)
    rng = np.random.default_rng(777)
    shake = np.zeros(3, np.float32)
    steps = 0
    while not (ca.checkpoint_step and cb.checkpoint_step):
        assert steps < 400, "the clocks never reached the checkpoint"
        assert ca.checkpoint_step == cb.checkpoint_step, "the two sides fell out of step"
        timestep(ca); timestep(cb)
        off = ((rng.random(3, dtype=np.float32) - np.float32(0.5)) * np.float32(16.0) - shake).astype(np.float32)
        shake = shake + off
        og = pm.particle_mesh(ca.dt, ca.dt_old, ca.a_mid, mass_p, off)
        oo = o.particle_mesh(cb.dt, cb.dt_old, cb.a_mid, mass_p, off)
        absorb_limiters(ca, og); absorb_limiters(cb, oo)
        assert og.np_total == oo.np_total == len(xv)
        steps += 1
    assert steps > 150 and ca.nts == cb.nts
    # the last step is cut linearly in dt (timestep.f90:131-133), which lands a within ~2e-5 of a_checkpoint, not on it
    assert ca.a == pytest.approx(a_t, rel=1e-4) and cb.a == pytest.approx(ca.a, rel=2e-6)
    # checkpoint: half drift to the end of the step (cubepm.f90:175-176), positions minus the shake offset (checkpoint.f90:92)
    off = ((rng.random(3, dtype=np.float32) - np.float32(0.5)) * np.float32(16.0) - shake).astype(np.float32)
    shake = shake + off
    pm.update_position(ca.dt, 0.0, off); o.update_position(cb.dt, 0.0, off)
    kd, dd, _ = pm.cic_power(box, shake=shake)                       # device cic_power of the resident particles
    g, r = pm.download_particles(), o.get_particles()
    pm.close(); o.close()
    undo = lambda p: np.mod(p[:, :3] - shake, np.float32(nc))
    kg, dg, _ = power.power_spectrum(undo(g), nc, box)
    kr, dr, _ = power.power_spectrum(undo(r), nc, box)
    rel = np.abs(dg - dr) / np.maximum(np.abs(dr), 1e-30)
    assert rel.max() < 1e-3, (steps, float(rel.max()), int(rel.argmax()))
    reld = np.abs(dd - dr) / np.maximum(np.abs(dr), 1e-30)
    assert reld.max() < 1e-3, ("device cic_power", float(reld.max()))
    # growth sanity: linear growth from z = 100 to z = 10 is (101/11)^2 = 84 in power on large scales (Omega_m ~ 1 at these redshifts)
    k0, d0, _ = power.power_spectrum(xv[:, :3], nc, box)
    big = slice(2, 10)
    assert 60 < np.median(dr[big] / d0[big]) < 110, np.median(dr[big] / d0[big])
