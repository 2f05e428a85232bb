"""dist_init on the device (cubep3m_b200_dist_init, csrc/distinit.cuh) against the host twin of utils/dist_init/dist_init_dm.f90 (cubep3m_b200/ic.py)."""
import numpy as np
import pytest

from cubep3m_b200 import default_config, ic, power

pytestmark = pytest.mark.gpu


def test_device_dist_init_matches_host_twin_on_the_same_noise(built):
    """Same white-noise field on both sides: lattice positions exact, displacements and velocities within 2e-3 of the rms displacement
    (fp32 FFT + the 2048-point log-log table of Delta^2 against the twin's direct evaluation)."""
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)
    nc, box, z_i, seed = cfg.nf_physical_dim, 50.0, 20.0, 7
    ref = ic.zeldovich_ics(nc, box=box, z_i=z_i, seed=seed)
    noise = np.random.default_rng(seed).standard_normal((nc, nc, nc), dtype=np.float32)
    pm = ParticleMesh(cfg)
    n = pm.dist_init(nc, box, z_i, noise=noise)
    got = pm.download_particles()
    assert n == len(ref) == len(got) == (nc // 2) ** 3
    lat = np.round(ref[:, :3] + 0.5) - 0.5                      # the twin's lattice points (displacements are << 1 cell at z = 20)
    dref, dgot = ref[:, :3] - lat, got[:, :3] - lat
    rms = float(np.sqrt(np.mean(dref ** 2)))
    assert rms > 1e-3
    assert np.abs(dgot - dref).max() < 2e-3 * rms + 2e-6 * nc, (np.abs(dgot - dref).max(), rms)
    vr = float(np.sqrt(np.mean(ref[:, 3:] ** 2)))
    assert np.abs(got[:, 3:] - ref[:, 3:]).max() < 2e-3 * vr
    out = pm.particle_mesh(0.1, 0.0, 1.0 / 21.0, 8.0)           # usable as a simulation start
    assert out.np_total == n
    pm.close()


def test_device_noise_gives_the_input_spectrum_and_replicates(built):
    """Philox / Box-Muller noise drawn on the device: the generated particles carry the input spectrum (same gate as the host twin's own test),
    two seeds differ, one seed repeats bit for bit; reps = 2 tiles the periodic box 2^3 times."""
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)
    nc, box, z_i = cfg.nf_physical_dim, 100.0, 50.0
    pm = ParticleMesh(cfg)
    pm.dist_init(nc, box, z_i, seed=11)
    a = pm.download_particles().copy()
    pm.dist_init(nc, box, z_i, seed=11)
    assert np.array_equal(a, pm.download_particles())
    pm.dist_init(nc, box, z_i, seed=12)
    b = pm.download_particles().copy()
    assert not np.array_equal(a, b)
    k, d2, _ = power.power_spectrum(np.mod(a[:, :3], np.float32(nc)), nc, box)
    lin = ic.delta2(k, 1.0 / (1.0 + z_i))
    sel = (k > 4 * 2 * np.pi / box) & (k < 0.25 * np.pi * nc / box)
    assert 0.6 < np.median(d2[sel] / lin[sel]) < 1.4, np.median(d2[sel] / lin[sel])
    n = pm.dist_init(nc // 2, box / 2, z_i, reps=2, seed=3)
    r = pm.download_particles()
    nb = (nc // 4) ** 3
    assert n == 8 * nb
    for rep in range(8):
        off = np.array([rep % 2, (rep // 2) % 2, rep // 4], np.float32) * np.float32(nc // 2)
        assert np.allclose(r[rep * nb:(rep + 1) * nb, :3], r[:nb, :3] + off, atol=1e-4) and np.array_equal(r[rep * nb:(rep + 1) * nb, 3:], r[:nb, 3:])
    pm.close()
