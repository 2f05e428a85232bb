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
