import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cubep3m_b200 import default_config, ic
from cubep3m_b200.lib import ParticleMesh
from tests.conftest import sort_records
cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=0)
xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=12)
xv[:, 3:] = 0
args = (0.5, 0.3, 0.05, 8.0, (1.25, -0.5, 2.0))
res = {}
for rep in range(2):
    for mode in ("replicated", "slab"):
        os.environ["CUBEP3M_B200_COARSE"] = mode
        pm = ParticleMesh(cfg)
        pm.upload_particles(xv)
        out = pm.particle_mesh(*args)
        res[(mode, rep)] = (pm.force_c().copy(), sort_records(pm.download_particles()), out.dt_c_acc, out.dt_f_acc)
        pm.close()
a = res[("replicated", 0)]
for k, v in res.items():
    d = np.abs(v[1][:, 3:] - a[1][:, 3:])
    bad = d.max(1) > 1e-4 * np.abs(a[1][:, 3:]).max()
    print(k, "force_c max diff", np.abs(v[0] - a[0]).max(), "vel max diff", d.max(), "bad", bad.sum(), "dt_c", v[2], "dt_f", v[3])
    if bad.sum():
        p = v[1][bad][:, :3]
        print("  bad positions min/max per axis", p.min(0), p.max(0))
        print("  comp-wise max diff", d.max(0))
