"""Small driver for compute-sanitizer runs: one NGP+PPINT+PP_EXT step and one fine-CIC step on 64^3 particles."""
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
