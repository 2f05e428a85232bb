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


@pytest.mark.gpu
def test_cic_power_slab_and_four_step_paths_match_direct(built, monkeypatch):
    """cubep3m_b200_cic_power three ways on the same 512^3 mesh (256^3 particles): everything-on-one-GPU direct transform (default), the z-slab / y-pencil
    pipeline of the multi-rank path with W = 1 (CUBEP3M_B200_POWER=slab), and that pipeline with the four-step transforms of bigfft.cuh that 1024 / 2048
    meshes need (=big; 512 = 16*16*2 on x, 32*16 on y and z, frequencies left digit-transposed). Same fp32 data, different summation orders: 2e-5."""
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=304, tiles_node_dim=2, pp_ext=0)
    nc, box = cfg.nf_physical_dim, 200.0
    xv = ic.zeldovich_ics(nc, box=box, z_i=20.0, seed=3)
    shake = np.array([1.25, -3.5, 0.5], np.float32)
    xv[:, :3] = np.mod(xv[:, :3] + shake, np.float32(nc))
    res = {}
    for mode in ("direct", "slab", "big"):
        monkeypatch.setenv("CUBEP3M_B200_POWER", mode)
        pm = ParticleMesh(cfg)
        pm.upload_particles(xv)
        res[mode] = {ngp: pm.cic_power(box, shake=shake, ngp_binning=ngp) for ngp in (True, False)}
        pm.close()
    for mode in ("slab", "big"):
        for ngp in (True, False):
            k0, d0, s0 = res["direct"][ngp]
            k1, d1, s1 = res[mode][ngp]
            assert np.allclose(k0, k1, rtol=1e-12), mode
            rel = np.abs(d1 - d0) / np.maximum(np.abs(d0), 1e-30)
            assert rel.max() < 2e-5, (mode, ngp, float(rel.max()), int(rel.argmax()))
            assert np.allclose(s0, s1, rtol=1e-3, atol=1e-12), mode


def _expected_power_of_2x_replicated_box(small_pos, nb_small, box_big):
    """Shell-averaged Delta^2 (NGP shells, cic_power.f90:1583-1660) of the 2x2x2 periodic replication of a box, from the small box alone:
    delta_big(2k') = delta_small(k') mode by mode (same CIC cells, 8 replicas over 8x the volume), every mode with an odd component vanishes; the mode counts and
    mean |k| of the big mesh are enumerated exactly, plane by plane."""
    from cubep3m_b200.power import cic_density
    try:
        from scipy import fft as sfft
        kw = {"workers": -1}
    except Exception:
        sfft, kw = np.fft, {}
    nc = 2 * nb_small
    rho = cic_density(small_pos, nb_small)
    dk = sfft.rfftn(rho - 1.0, **kw) / float(nb_small) ** 3
    n2 = nc // 2
    P = np.zeros(n2 + 3); Wn = np.zeros(n2 + 3); K = np.zeros(n2 + 3)
    sinc = lambda kk, n: np.where(kk == 0, 1.0, np.sin(np.pi * kk / n) / np.where(kk == 0, 1.0, np.pi * kk / n))
    kx_s = np.arange(nb_small // 2 + 1, dtype=np.float64)[None, :]
    fs = np.fft.fftfreq(nb_small, 1.0 / nb_small)
    ky_s = fs[:, None]
    # In the big mesh the small box's Nyquist index -nb/2 (fftfreq convention) sits at k = -nb = index nb of a 2nb mesh, i.e. frequency +nb/2 * 2 ... as
    # a SIGNED big-mesh frequency 2k' = -nb which is not the big mesh's Nyquist, so no folding happens: k = 2k' is used as is for every k'.
    for iz, kz in enumerate(fs):
        kr = 2.0 * np.sqrt(kx_s ** 2 + ky_s ** 2 + kz ** 2)
        pw = (dk[iz].real ** 2 + dk[iz].imag ** 2) / (sinc(kx_s, nb_small) * sinc(ky_s, nb_small) * sinc(np.float64(kz), nb_small)) ** 4
        keep = ~((kx_s == 0) & ~((ky_s > 0) | ((ky_s == 0) & (kz > 0)))) & (kr > 0)
        k1 = np.ceil(kr).astype(np.int64)
        sel = keep & (k1 <= n2 + 1)
        np.add.at(P, k1[sel], pw[sel])
    kx_b = np.arange(nc // 2 + 1, dtype=np.float64)[None, :]
    ky_b = np.fft.fftfreq(nc, 1.0 / nc)[:, None]
    for kz in np.fft.fftfreq(nc, 1.0 / nc):
        kr = np.sqrt(kx_b ** 2 + ky_b ** 2 + kz ** 2)
        keep = ~((kx_b == 0) & ~((ky_b > 0) | ((ky_b == 0) & (kz > 0)))) & (kr > 0)
        k1 = np.ceil(kr).astype(np.int64)
        sel = keep & (k1 <= n2 + 1)
        np.add.at(Wn, k1[sel], 1.0); np.add.at(K, k1[sel], kr[sel])
    sh = np.arange(1, n2 + 1)
    kavg = K[sh] / Wn[sh]
    return 2 * np.pi * kavg / box_big, 4 * np.pi * kavg ** 3 * P[sh] / Wn[sh]


def test_replicated_box_expectation_matches_host_twin():
    """The expectation used by the 1024^3 GPU test, checked here on the CPU at 32 -> 64 cells against the host twin run on the replicated particles."""
    nb = 32
    small = ic.zeldovich_ics(nb, box=50.0, z_i=5.0, seed=9)
    small[:, :3] = np.mod(small[:, :3], np.float32(nb))
    big = ic.tile_box(small, nb, 2)
    k, d2, _ = power.power_spectrum(big[:, :3], 2 * nb, 100.0)
    ke, de = _expected_power_of_2x_replicated_box(small[:, :3], nb, 100.0)
    assert np.allclose(k, ke, rtol=1e-9)
    good = de > 1e-12 * de.max()
    assert np.abs(d2[good] - de[good]).max() <= 1e-6 * de.max() and np.allclose(d2[good], de[good], rtol=1e-4)


@pytest.mark.gpu
def test_cic_power_1024_mesh_against_replicated_512(built):
    """BASELINE configs[2]'s mesh (1024^3, a 4.3 GB half spectrum: four-step passes with 64-bit offsets). The particles are a 2x2x2 periodic replication of a
    512-cell box, for which the answer is known from the small box alone (_expected_power_of_2x_replicated_box, float64 on the host); gate 2e-4 (fp32 mesh)."""
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=304, tiles_node_dim=4, ppint=0, pp_ext=0, max_np=40_000_000)
    nc, box, nb_small = cfg.nf_physical_dim, 400.0, 512
    assert nc == 1024
    small = ic.zeldovich_ics(nb_small, box=200.0, z_i=10.0, seed=5)[::8].copy()          # 2.1 M particles of a 512-cell box
    small[:, :3] = np.mod(small[:, :3], np.float32(nb_small))
    xv = ic.tile_box(small, nb_small, 2)
    pm = ParticleMesh(cfg)
    pm.upload_particles(xv)
    k, d2, _ = pm.cic_power(box, ngp_binning=True)
    pm.close()
    ke, want = _expected_power_of_2x_replicated_box(small[:, :3], nb_small, box)
    assert np.allclose(k, ke, rtol=1e-9)
    good = want > 1e-12 * want.max()
    rel = np.abs(d2[good] - want[good]) / want[good]
    assert rel.max() < 2e-4, (float(rel.max()), int(np.argmax(rel)))
