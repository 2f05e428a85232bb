#!/usr/bin/env python
"""Freeze oracle outputs as golden vectors (run in the build container; commit the .npz).

The reference ships no golden vectors for particle_mesh and cannot be executed here, so these pin the
ORACLE (single-threaded, deterministic) rather than the Fortran; the oracle in turn is pinned by tests/test_oracle.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from cubep3m_b200 import default_config, ic  # noqa: E402
from oracle import Oracle  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
cfg = default_config(nf_tile=112, tiles_node_dim=2, pp_ext=1)
full = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=99)
rng = np.random.default_rng(1)
xv = full[rng.choice(len(full), 6000, replace=False)].copy()
# a few tight pairs / clumps so PPINT and PP_EXT are exercised
c = rng.random((40, 3)).astype(np.float32) * cfg.mT
clump = (c[:, None, :] + rng.normal(0, 0.4, (40, 6, 3)).astype(np.float32)).reshape(-1, 3) % np.float32(cfg.mT)
xc = np.zeros((len(clump), 6), np.float32); xc[:, :3] = clump
xv = np.concatenate([xv, xc]).astype(np.float32)
dt, dt_old, a_mid, mass_p = 0.4, 0.2, 0.05, 8.0
off = np.array([3.5, -2.25, 0.75], np.float32)
o = Oracle(cfg, threads=1)
o.set_particles(xv)
out = o.particle_mesh(dt, dt_old, a_mid, mass_p, off)
np.savez_compressed(os.path.join(here, "step_n112_T2.npz"), xv_in=xv, xv_out=o.get_particles(), dt=dt, dt_old=dt_old, a_mid=a_mid,
                    mass_p=mass_p, offset=off, pp_ext=1, dt_f_acc=out.dt_f_acc, dt_pp_acc=out.dt_pp_acc, dt_pp_ext_acc=out.dt_pp_ext_acc,
                    dt_c_acc=out.dt_c_acc, tile_counts=o.tile_counts(), np_with_ghosts=out.np_with_ghosts)
print("golden written:", len(xv), "particles ->", out.np_total, "ghosts", out.np_with_ghosts, out.dt_f_acc, out.dt_pp_acc, out.dt_pp_ext_acc)
