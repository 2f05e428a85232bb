#!/usr/bin/env python
"""Times the device cic_power (cubep3m_b200_cic_power) on the bench workload's particles and compares it with the host twin."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cubep3m_b200 import default_config, ic, power
from cubep3m_b200.lib import ParticleMesh

n, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (304, 2)
cfg = default_config(nf_tile=n, tiles_node_dim=T, pp_ext=0)
nc = cfg.nf_physical_dim
xv = ic.zeldovich_ics(nc, box=200.0, z_i=100.0, seed=12345)
pm = ParticleMesh(cfg)
pm.upload_particles(xv)
pm.cic_power(200.0)
t = time.perf_counter(); k, d2, s = pm.cic_power(200.0); tg = time.perf_counter() - t
t = time.perf_counter(); kh, dh, sh = power.power_spectrum(xv[:, :3], nc, 200.0); th = time.perf_counter() - t
rel = np.abs(d2 - dh) / np.maximum(np.abs(dh), 1e-30)
print({"nc": nc, "particles": len(xv), "gpu_ms": round(tg * 1e3, 2), "host_twin_ms": round(th * 1e3, 1), "max_rel_diff": float(rel.max())})
pm.close()
