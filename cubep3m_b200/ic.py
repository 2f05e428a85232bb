"""Synthetic LCDM Zel'dovich initial conditions in the reference's `xv<rank>.ic` layout.

Host-side (numpy/scipy) twin of utils/dist_init/dist_init_dm.f90:
  * particles on every 2nd fine cell at (i1 - 0.5), displaced by the centred finite difference of the
    potential, v = displacement * vfactor(a)                       dist_init_dm.f90:1011-1036
  * delta(k) = sqrt(Delta^2(2 pi kr / box) / (4 pi kr^3) * nc^3) * noise(k)          :685-712
  * potential kernel  K(k) = -4 pi / sum_d (2 sin(pi k_d / nc))^2                    :814-832
  * growth factor / vfactor fits                                                     :1303-1337
Differences (documented, both sides of every parity test see the same particles):
  * the Gaussian noise comes from numpy default_rng(seed) — the Fortran RNG is unavailable;
  * the transfer function is the Eisenstein & Hu (1998) no-wiggle fit instead of the reference's
    batch/camb_WMAP5_transfer_z0.dat table (reference data is not copied into this repo);
  * the short-range kernel correction (`correct_kernel`, :850-903) is not applied.
File format: int32 np_local followed by np_local x 6 float32 (x,y,z,vx,vy,vz), stream access.
"""
import numpy as np

try:
    from scipy import fft as _fft
    _KW = {"workers": -1}
except Exception:  # pragma: no cover
    _fft = np.fft
    _KW = {}


def dgrow(a, om, ol):
    """dist_init_dm.f90:1303-1320 (Carroll, Press & Turner fit)."""
    hsq = om / a ** 3 + (1 - om - ol) / a ** 2 + ol
    oma = om / (a ** 3 * hsq)
    ola = ol / hsq
    g = 2.5 * om / (om ** (4. / 7) - ol + (1 + om / 2) * (1 + ol / 70))
    ga = 2.5 * oma / (oma ** (4. / 7) - ola + (1 + oma / 2) * (1 + ola / 70))
    return a * ga / g


def vfactor(a, om, ol):
    """dist_init_dm.f90:1324-1337."""
    lm = ol / om
    km = (1 - om - ol) / om
    H = 2 / (3 * np.sqrt(a ** 3)) * np.sqrt(1 + a * km + a ** 3 * lm)
    return a ** 2 * H


def transfer_nowiggle(k, om, ob, h):
    """Eisenstein & Hu 1998 eq. 26-31 (no-wiggle); k in h/Mpc."""
    theta = 2.728 / 2.7
    omh2, obh2 = om * h * h, ob * h * h
    fb = ob / om
    s = 44.5 * np.log(9.83 / omh2) / np.sqrt(1 + 10 * obh2 ** 0.75)
    alpha = 1 - 0.328 * np.log(431 * omh2) * fb + 0.38 * np.log(22.3 * omh2) * fb ** 2
    kk = k * h  # 1/Mpc
    gamma = om * h * (alpha + (1 - alpha) / (1 + (0.43 * kk * s) ** 4))
    q = k * theta ** 2 / gamma
    L0 = np.log(2 * np.e + 1.8 * q)
    C0 = 14.2 + 731.0 / (1 + 62.5 * q)
    return L0 / (L0 + C0 * q * q)


def delta2(k, a, om=0.24, ol=0.76, ob=0.04, h=0.7, ns=0.96, s8=0.817):
    """Dimensionless Delta^2(k) at scale factor a, sigma_8-normalised at a=1 (dist_init_dm.f90:475-504)."""
    kk = np.logspace(-4, 2, 4000)
    d2 = kk ** (3 + ns) * transfer_nowiggle(kk, om, ob, h) ** 2 / (2 * np.pi ** 2)
    x = kk * 8.0
    W = 3 * (np.sin(x) - x * np.cos(x)) / x ** 3
    v8 = np.trapezoid(d2 * W ** 2 / kk, kk) if hasattr(np, 'trapezoid') else np.trapz(d2 * W ** 2 / kk, kk)
    norm = s8 ** 2 / v8 * dgrow(a, om, ol) ** 2
    return norm * k ** (3 + ns) * transfer_nowiggle(k, om, ob, h) ** 2 / (2 * np.pi ** 2)


def zeldovich_ics(nc, box=200.0, z_i=100.0, om=0.24, ol=0.76, seed=12345, amplitude=1.0):
    """Global IC for an nc^3 fine mesh: returns xv (np^3, 6) float32 with np = nc/2, positions in [0, nc).

    Particle order follows the file order of dist_init (i fastest)."""
    assert nc % 2 == 0
    a = 1.0 / (1.0 + z_i)
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((nc, nc, nc), dtype=np.float32)
    nk = _fft.rfftn(noise, **_KW)
    del noise
    kz = np.fft.fftfreq(nc, 1.0 / nc).astype(np.float32)[:, None, None]
    ky = np.fft.fftfreq(nc, 1.0 / nc).astype(np.float32)[None, :, None]
    kx = np.arange(nc // 2 + 1, dtype=np.float32)[None, None, :]
    kr = np.sqrt(kx * kx + ky * ky + kz * kz)
    kr[0, 0, 0] = 1.0
    d2 = delta2(2 * np.pi * kr / box, a, om, ol).astype(np.float32)
    amp = np.sqrt(d2 / (4 * np.pi * kr ** 3) * float(nc) ** 3).astype(np.float32) * np.float32(amplitude)
    del d2
    den = ((2 * np.sin(np.pi * kx / nc)) ** 2 + (2 * np.sin(np.pi * ky / nc)) ** 2 + (2 * np.sin(np.pi * kz / nc)) ** 2)
    den[0, 0, 0] = 1.0
    kern = -4 * np.pi / den
    del den
    nk *= (amp * kern.astype(np.float32))
    nk[0, 0, 0] = 0
    del amp, kern, kr
    phi = _fft.irfftn(nk, s=(nc, nc, nc), **_KW).astype(np.float32)
    del nk
    npd = nc // 2
    vf = np.float32(vfactor(a, om, ol))
    i1 = 2 * np.arange(npd)  # 0-based index of the lattice cell (Fortran i1 = 2(i-1)+1)
    sub = phi[np.ix_(i1, i1, i1)]
    xv = np.empty((npd, npd, npd, 6), np.float32)
    fourpi = np.float32(4 * np.pi)
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        lo = np.roll(phi, 1, axis=ax)[np.ix_(i1, i1, i1)]
        hi = np.roll(phi, -1, axis=ax)[np.ix_(i1, i1, i1)]
        dis = (lo - hi) / np.float32(2.0) / fourpi
        lat = (i1 + 1).astype(np.float32) - np.float32(0.5)
        shape = [1, 1, 1]
        shape[ax] = npd
        xv[..., d] = dis + lat.reshape(shape)
        xv[..., 3 + d] = dis * vf
    del sub
    return xv.reshape(-1, 6)


def split_ranks(xv, nc, nodes_dim):
    """Distribute a global particle list to the D^3 ranks exactly as dist_init does: by the LATTICE position
    (not the displaced one), local coordinate = global - cart_coords*nf_physical_node_dim.  rank = x + D y + D^2 z."""
    D = nodes_dim
    npd = nc // 2
    per = npd // D
    g = xv.reshape(npd, npd, npd, 6)
    out = []
    L = np.float32(nc // D)
    for rz in range(D):
        for ry in range(D):
            for rx in range(D):
                blk = g[rz * per:(rz + 1) * per, ry * per:(ry + 1) * per, rx * per:(rx + 1) * per].reshape(-1, 6).copy()
                blk[:, 0] -= np.float32(rx) * L
                blk[:, 1] -= np.float32(ry) * L
                blk[:, 2] -= np.float32(rz) * L
                out.append(blk)
    return out


def tile_box(xv, nc, reps):
    """Periodic replication of an nc-cell box reps times per dimension (used for large synthetic inputs)."""
    outs = []
    for rz in range(reps):
        for ry in range(reps):
            for rx in range(reps):
                b = xv.copy()
                b[:, 0] += np.float32(rx * nc)
                b[:, 1] += np.float32(ry * nc)
                b[:, 2] += np.float32(rz * nc)
                outs.append(b)
    return np.concatenate(outs, axis=0)


def write_ic(path, xv):
    """xv<rank>.ic stream format (dist_init_dm.f90:1011-1034; particle_initialization.f90:299-332 with -DBINARY)."""
    xv = np.ascontiguousarray(xv, np.float32).reshape(-1, 6)
    with open(path, "wb") as f:
        np.array([xv.shape[0]], np.int32).tofile(f)
        xv.tofile(f)


def read_ic(path):
    with open(path, "rb") as f:
        n = int(np.fromfile(f, np.int32, 1)[0])
        return np.fromfile(f, np.float32, 6 * n).reshape(n, 6)


def clustered_ics(nc, n_particles, seed=4242, n_halos=64, frac_in_halos=0.5):
    """A seed-fixed clustered input (uniform background + NFW-like clumps) for PP / contention measurements."""
    rng = np.random.default_rng(seed)
    nh = int(n_particles * frac_in_halos)
    nb = n_particles - nh
    pos = [rng.random((nb, 3), dtype=np.float32) * np.float32(nc)]
    centers = rng.random((n_halos, 3)) * nc
    per = np.full(n_halos, nh // n_halos)
    per[: nh - per.sum()] += 1
    for c, k in zip(centers, per):
        rs = 0.5 + 1.5 * rng.random()
        u = rng.random(k)
        r = rs * u / (1.0 - u + 1e-3)      # ~ r^-2 cusp, truncated
        r = np.minimum(r, 12.0)
        d = rng.standard_normal((k, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        pos.append(((c + d * r[:, None]) % nc).astype(np.float32))
    x = np.concatenate(pos, axis=0)
    x = np.minimum(x, np.float32(np.nextafter(np.float32(nc), np.float32(0))))
    xv = np.zeros((n_particles, 6), np.float32)
    xv[:, :3] = x
    return xv
