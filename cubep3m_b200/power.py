"""Host twin of utils/cic_power/cic_power.f90: the 0.1 % P(k) acceptance metric.

  * CIC deposit with the x - 0.5 cell-centre convention and particle mass mp = (nc/np)^3      cic_power.f90:1496-1539
  * delta = rho - 1, forward FFT                                                              :918
  * per mode: skip the redundant half of the kx = 0 plane (:1583-1584); pow = |delta_k / nc^3|^2 / (sinc sinc sinc)^4 (:1590-1615)
  * shells k1 = ceil(kr), k2 = k1 + 1 with weights w1 = k1 - kr, w2 = 1 - w1 (the NGP build, COMPILE_cic_power.csh:20, uses
    w1 = 1, w2 = 0)                                                                           :1586-1589,1611-1614
  * output k = 2 pi kavg / box, Delta^2 = 4 pi kavg^3 P (NGP binning) or 4 pi (kavg-1)^3 P (CIC binning)   :1649-1660
This twin runs on the host with numpy/scipy on the gathered global particle list (single rank; the reference's quirk that kavg comes
from rank 0's slab only does not arise)."""
import numpy as np

try:
    from scipy import fft as _fft
    _KW = {"workers": -1}
except Exception:  # pragma: no cover
    _fft = np.fft
    _KW = {}


def cic_density(pos, nc):
    """rho on an nc^3 periodic mesh, mean 1; pos in [0, nc) fine-cell units, (N,3) float32. cic_power.f90:1496-1539."""
    pos = np.asarray(pos, np.float64)
    npart = len(pos)
    mp = float(nc) ** 3 / npart
    x = pos - 0.5
    i1 = np.floor(x).astype(np.int64)
    d2 = x - i1
    d1 = 1.0 - d2
    rho = np.zeros(nc ** 3, np.float64)
    i1 %= nc
    i2 = (i1 + 1) % nc
    for cz, wz in ((i1[:, 2], d1[:, 2]), (i2[:, 2], d2[:, 2])):
        for cy, wy in ((i1[:, 1], d1[:, 1]), (i2[:, 1], d2[:, 1])):
            for cx, wx in ((i1[:, 0], d1[:, 0]), (i2[:, 0], d2[:, 0])):
                np.add.at(rho, (cz * nc + cy) * nc + cx, mp * wx * wy * wz)
    return rho.reshape(nc, nc, nc)


def power_spectrum(pos, nc, box, ngp_binning=True):
    """Returns (k [h/Mpc], Delta^2(k), sigma) for shells 1..nc/2 as cic_power writes them (columns 1-3 of <z>ngpps_new.dat)."""
    rho = cic_density(pos, nc)
    dk = _fft.rfftn(rho - 1.0, **_KW) / float(nc) ** 3
    kz = np.fft.fftfreq(nc, 1.0 / nc)[:, None, None]
    ky = np.fft.fftfreq(nc, 1.0 / nc)[None, :, None]
    kx = np.arange(nc // 2 + 1, dtype=np.float64)[None, None, :]
    kr = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    sinc = lambda k: np.where(k == 0, 1.0, np.sin(np.pi * k / nc) / np.where(k == 0, 1.0, np.pi * k / nc))
    w = (sinc(kx) * sinc(ky) * sinc(kz)) ** 4
    pw = (dk.real ** 2 + dk.imag ** 2) / w
    # skip kr = 0 and the redundant half of the kx = 0 plane (:1583-1584): keep (ky > 0) or (ky == 0 and kz > 0) there
    keep = np.ones_like(kr, bool)
    keep[0, 0, 0] = False
    plane = np.broadcast_to((ky > 0) | ((ky == 0) & (kz > 0)), kr[:, :, :1].shape)
    keep[:, :, 0] &= plane[:, :, 0]
    kr, pw = kr[keep], pw[keep]
    k1 = np.ceil(kr).astype(np.int64)
    nb = nc // 2 + 2
    if ngp_binning:
        w1, w2 = np.ones_like(kr), np.zeros_like(kr)
    else:
        w1 = k1 - kr
        w2 = 1.0 - w1
    acc = lambda wt, v, idx: np.bincount(idx, weights=wt * v, minlength=nb + 2)[:nb + 2]
    k2 = k1 + 1
    ok1, ok2 = k1 <= nb, k2 <= nb
    P = acc(w1[ok1], pw[ok1], k1[ok1]) + acc(w2[ok2], pw[ok2], k2[ok2])
    P2 = acc(w1[ok1], pw[ok1] ** 2, k1[ok1]) + acc(w2[ok2], pw[ok2] ** 2, k2[ok2])
    W = acc(w1[ok1], np.ones_like(pw[ok1]), k1[ok1]) + acc(w2[ok2], np.ones_like(pw[ok2]), k2[ok2])
    K = acc(w1[ok1], kr[ok1], k1[ok1]) + acc(w2[ok2], kr[ok2], k2[ok2])
    sh = np.arange(1, nc // 2 + 1)
    Wn = np.maximum(W[sh], 1e-300)
    kavg = K[sh] / Wn
    Pm = P[sh] / Wn
    var = np.maximum(P2[sh] / Wn - Pm ** 2, 0.0)
    keff = kavg if ngp_binning else kavg - 1.0
    d2 = 4 * np.pi * keff ** 3 * Pm
    sig = 4 * np.pi * keff ** 3 * np.sqrt(var / np.maximum(W[sh] - 1, 1.0))
    return 2 * np.pi * kavg / box, d2, sig
