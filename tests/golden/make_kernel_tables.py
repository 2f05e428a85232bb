#!/usr/bin/env python
"""Regenerate the Green's-function table fixtures from the reference's ASCII tables.

Reads (only in the build container, where /root/reference exists):
  kernels/wfxyzf.3.ascii  (fine mesh, 16^3 rows '3i4,3e16.8'; kernel_initialization.f90:15,25-36)
  kernels/wfxyzc.2.ascii  (coarse mesh, 4^3 rows;             kernel_initialization.f90:344-359)
and writes float32 .npy arrays laid out [k][j][i][component] (row order of the files: i fastest):
  cubep3m_b200/data/wfxyzf3.npy  shape (16,16,16,3)
  cubep3m_b200/data/wfxyzc2.npy  shape (4,4,4,3)
The values are the exact float32 roundings of the e16.8 decimals, i.e. what a Fortran
formatted read into real(4) produces.
"""
import sys, numpy as np, os
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(__file__), "..", "..", "cubep3m_b200", "data")
for src, n, dst in (("wfxyzf.3.ascii", 16, "wfxyzf3.npy"), ("wfxyzc.2.ascii", 4, "wfxyzc2.npy")):
    rows = np.loadtxt(os.path.join(ref, "kernels", src), dtype=np.float64)
    assert rows.shape == (n ** 3, 6), rows.shape
    idx = rows[:, :3].astype(int)
    exp = np.stack(np.meshgrid(np.arange(1, n + 1), np.arange(1, n + 1), np.arange(1, n + 1), indexing="ij"), -1)
    # file order: k outer, j middle, i inner
    exp = exp.transpose(2, 1, 0, 3).reshape(-1, 3)
    assert (idx == exp).all(), "unexpected row order"
    tab = rows[:, 3:].astype(np.float32).reshape(n, n, n, 3)
    np.save(os.path.join(out, dst), tab)
    print(dst, tab.shape, float(np.abs(tab).max()))
