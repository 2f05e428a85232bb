#!/usr/bin/env python
"""Verbose stage-by-stage GPU-vs-oracle diagnostic (development aid; the judged tests are tests/test_gpu_*.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cubep3m_b200 import default_config, ic
from cubep3m_b200.lib import ParticleMesh
from oracle import Oracle, oracle_fft3d


def srt(a):
    return a[np.lexsort((a[:, 5], a[:, 4], a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 112
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    pp_ext = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    cfg = default_config(nf_tile=n, tiles_node_dim=T, pp_ext=pp_ext)
    print("cfg n", n, "T", T, "mT", cfg.mT, "nc", cfg.nc_dim, "pp_ext", pp_ext, flush=True)
    t = time.time(); pm = ParticleMesh(cfg); print("gpu init %.2fs" % (time.time() - t), flush=True)
    t = time.time(); o = Oracle(cfg); print("oracle init %.2fs" % (time.time() - t), flush=True)
    rng = np.random.default_rng(0)
    # 1. FFT
    for N in (cfg.nf_tile, cfg.nc_dim):
        x = rng.standard_normal((N, N, N)).astype(np.float32)
        a = np.zeros((N, N, N + 2), np.float32); a[:, :, :N] = x
        f = pm.fft3d(a.copy())
        ref = np.fft.rfftn(x.astype(np.float64))
        got = f.view(np.complex64)
        print("fft fwd N=%d rel err %.3e" % (N, np.abs(got - ref).max() / np.abs(ref).max()))
        b = pm.fft3d(f.copy(), inverse=True)[:, :, :N] / N ** 3
        print("fft roundtrip N=%d max abs err %.3e" % (N, np.abs(b - x).max()), flush=True)
    # 2. kernels
    kf, kfo = pm.kern_f(), o.kern_f()
    print("kern_f rel err %.3e (max %.3e)" % (np.abs(kf - kfo).max() / np.abs(kfo).max(), np.abs(kfo).max()))
    kc, kco = pm.kern_c(), o.kern_c()
    print("kern_c rel err %.3e (max %.3e)" % (np.abs(kc - kco).max() / np.abs(kco).max(), np.abs(kco).max()), flush=True)
    # 3. particles
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=7)
    xv[:, 3:] *= 3.0
    print("np", len(xv), "disp rms", float(np.sqrt(((xv[:, :3] - np.floor(xv[:, :3]) - 0.5) ** 2).mean())))
    off = (1.25, -0.5, 2.0)
    dt, dt_old, a_mid, mass_p = 0.5, 0.3, 0.05, 8.0
    pm.upload_particles(xv); o.set_particles(xv)
    pm.update_position(dt, dt_old, off); o.update_position(dt, dt_old, off)
    g, r = pm.download_particles(), o.get_particles()
    print("drift bit-exact:", np.array_equal(g, r), flush=True)
    ndel = pm.link_list(); o.link_list()
    cg, co = pm.cell_counts(), o.cell_counts()
    print("link: deleted", ndel, "cell counts equal:", np.array_equal(cg, co), "sum", cg.sum(), co.sum(), flush=True)
    npg = pm.particle_pass(); o.particle_pass()
    sg = pm.sorted_particles(); so = o.get_particles()
    print("pass: np gpu %d oracle %d ; sets bit-exact: %s" % (npg, len(so), np.array_equal(srt(sg), srt(so))), flush=True)
    o.link_list_after = None
    cg = pm.cell_counts()
    # oracle cell counts after pass (chains include ghosts)
    co = o.cell_counts()
    print("cell counts after pass equal:", np.array_equal(cg, co), flush=True)
    # 5. fine tile
    tile = 1 if cfg.tiles_node > 1 else 0
    rho_g, frc_g = pm.fine_tile(tile, mass_p)
    # full step on both (fresh)
    pm.upload_particles(xv); o.set_particles(xv)
    o.set_debug_tile(tile + 1)
    t = time.time(); out_o = o.particle_mesh(dt, dt_old, a_mid, mass_p, off); t_o = time.time() - t
    t = time.time(); out_g = pm.particle_mesh(dt, dt_old, a_mid, mass_p, off); t_g = time.time() - t
    rho_o, frc_o = o.fine_tile()
    print("tile rho equal:", np.array_equal(rho_g[:, :, :n], rho_o[:, :, :n]), "sum", rho_g[:, :, :n].sum(), rho_o[:, :, :n].sum())
    print("tile force rel err %.3e (max |F| %.3e)" % (np.abs(frc_g - frc_o).max() / np.abs(frc_o).max(), np.abs(frc_o).max()), flush=True)
    print("tile counts:", pm.tile_counts().tolist(), o.tile_counts().tolist())
    g, r = pm.download_particles(), o.get_particles()
    print("step: np gpu %d oracle %d ghosts %d %d del %d %d bufmax %d %d" % (len(g), len(r), out_g.np_with_ghosts, out_o.np_with_ghosts, out_g.np_deleted_ll, out_o.np_deleted_ll, out_g.np_buf_max, out_o.np_buf_max))
    k = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    g, r = k(g), k(r)
    print("positions bit-exact:", np.array_equal(g[:, :3], r[:, :3]))
    # velocity change relative to drifted-only
    x0 = xv.copy()
    num = np.sqrt(((g[:, 3:] - r[:, 3:]) ** 2).sum(1))
    # dv reference: need initial velocities matched by position -> compare against oracle's own dv via norm of velocity
    den = np.sqrt((r[:, 3:] ** 2).sum(1)) + 1e-30
    print("rms |dv_gpu - dv_ref| / |v_ref| = %.3e ; max %.3e" % (np.sqrt(np.mean((num / den) ** 2)), (num / den).max()))
    for f in ("dt_f_acc", "dt_pp_acc", "dt_pp_ext_acc", "dt_c_acc", "f_force_max", "pp_force_max", "pp_ext_force_max", "c_force_max", "sum_rho_f", "sum_rho_c"):
        print("  %-18s gpu %.7g oracle %.7g" % (f, getattr(out_g, f), getattr(out_o, f)))
    print("oracle step %.2fs (threads %d); gpu wall %.4fs" % (t_o, o.threads, t_g))
    print("gpu stages ms:", {k: round(v, 3) for k, v in out_g.stages().items()})
    print("oracle stages ms:", {k: round(v, 1) for k, v in out_o.stages().items()})
    # second step timing (warm)
    for _ in range(3):
        out_g = pm.particle_mesh(dt, dt, a_mid, mass_p, off)
    print("gpu warm stages ms:", {k: round(v, 3) for k, v in out_g.stages().items()}, "launches", pm.launches, flush=True)


if __name__ == "__main__":
    main()
