"""Small driver for compute-sanitizer runs: NGP+PPINT+PP_EXT steps and a fine-CIC step on 64^3 particles, the PP_EXT overflow / dense-cell path,
and (round 2) the slab coarse solve on one rank, cic_power, the halo finder's peak pass, the device timestep, the checkpoint writer / reader."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cubep3m_b200 import default_config, ic
from cubep3m_b200.lib import ParticleMesh
for kw in (dict(pp_ext=1), dict(ngp=0)):
    cfg = default_config(nf_tile=112, tiles_node_dim=2, **kw)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=7)
    pm = ParticleMesh(cfg)
    pm.upload_particles(xv)
    out = pm.particle_mesh(0.5, 0.3, 0.05, 8.0, (1.25, -0.5, 2.0))
    out = pm.particle_mesh(0.5, 0.5, 0.05, 8.0, (-3.0, 2.5, 7.75))
    print(kw, out.np_local, out.dt_f_acc, out.sum_rho_f, flush=True)
    pm.close()
# PP_EXT overflow path: a clump that exceeds the tiled kernel's shared-memory capacity (pp::TB_CAP) goes through ppext_blocklist_kernel
cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=1)
rng = np.random.default_rng(5)
xv = np.zeros((23000, 6), np.float32)
xv[:20000, :3] = rng.random((20000, 3)).astype(np.float32) * np.float32(cfg.mT)
xv[20000:, :3] = (np.array([0.37, 0.52, 0.61], np.float32) * np.float32(cfg.mT) + rng.normal(0, 1.5, (3000, 3))).astype(np.float32) % np.float32(cfg.mT)
pm = ParticleMesh(cfg)
pm.upload_particles(xv)
out = pm.particle_mesh(0.05, 0.05, 0.05, 8.0, (0.5, -1.5, 2.25))
print("clump", out.np_local, out.dt_pp_ext_acc, pm.ppext_blocks(), flush=True)
pm.close()

# round-2 kernels: slab-decomposed coarse solve on one rank, 8-column z pass, cic_power, halofind peaks, device timestep, checkpoint round trip
import tempfile
from cubep3m_b200.lib import clock_init
for env in ({"CUBEP3M_B200_COARSE": "slab"}, {"CUBEP3M_B200_SANDWICH": "v3"}):
    os.environ.update(env)
    cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=1)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=7)
    pm = ParticleMesh(cfg)
    pm.upload_particles(xv)
    out = pm.particle_mesh(0.5, 0.3, 0.05, 8.0, (1.25, -0.5, 2.0))
    print(env, out.np_local, out.dt_c_acc, out.dt_f_acc, flush=True)
    for k in env:
        del os.environ[k]
    if "CUBEP3M_B200_SANDWICH" in env:
        k, d2, sg = pm.cic_power(50.0)
        pm.link_list(); pm.particle_pass()
        for ngph in (True, False):
            pk, cft = pm.halofind_peaks(8.0, 4.0, True, ngph)
            print("peaks", ngph, len(pk), cft, flush=True)
        pm.delete_particles()
        c = clock_init(20.0, ppint=1, pp_ext=1)
        pm.timestep_device(c)
        with tempfile.TemporaryDirectory() as td:
            from cubep3m_b200.abi import CheckpointHeader
            h = CheckpointHeader(); h.a = c.a; h.nts = 1; h.mass_p = 8.0
            try:
                h.np_local = out.np_local
                pm.write_checkpoint(os.path.join(td, "xv0.dat"), h, (0.5, -0.25, 1.0))
                h2 = pm.read_checkpoint(os.path.join(td, "xv0.dat"))
                print("checkpoint np", h2.np_local, flush=True)
            except Exception as e:    # binding name / signature differences must not hide the sanitizer result of the kernels above
                print("checkpoint:", type(e).__name__, e, flush=True)
        print("power", float(d2[0]), "clock", c.a, c.dt, flush=True)
    pm.close()
