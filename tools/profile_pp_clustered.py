"""One particle_mesh step on a seed-fixed clustered box (uniform background + NFW-like clumps, cubep3m_b200/ic.py: clustered_ics) for ncu captures of the
dense-block PP_EXT kernel: python tools/profile_pp_clustered.py [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cubep3m_b200 import default_config, ic
from cubep3m_b200.lib import ParticleMesh
cfg = default_config(nf_tile=176, tiles_node_dim=2, ppint=1, pp_ext=1)
nc = cfg.nf_physical_dim
xv = ic.clustered_ics(nc, 128 ** 3, seed=4242, n_halos=256, frac_in_halos=0.5)
pm = ParticleMesh(cfg)
pm.upload_particles(xv)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for s in range(steps):
    out = pm.particle_mesh(0.002, 0.002, 0.5, 8.0, (1.5, -2.25, 0.75))
    pi, pe = pm.pair_counts()
    print(f"step {s}: {out.stage_ms[12]:.2f} ms, pp {out.stage_ms[6]:.2f}, pp_ext {out.stage_ms[7]:.2f} ms, pairs ppint {pi:.3e} ppext {pe:.3e}, blocks {pm.ppext_blocks()}", flush=True)
pm.close()
