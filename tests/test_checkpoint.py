"""Checkpoint files in the reference's -DBINARY stream format (checkpoint.f90:72-95, particle_initialization.f90:88-189), written from and read
into the device copy. The format oracle is the numpy restatement below of the two Fortran record lists."""
import os

import numpy as np
import pytest

from cubep3m_b200 import CheckpointHeader, checkpoint_name, default_config, ic

HDR_PPINT = np.dtype([("np_local", "<i4"), ("a", "<f4"), ("t", "<f4"), ("tau", "<f4"), ("nts", "<i4"), ("dt_f_acc", "<f4"), ("dt_pp_acc", "<f4"),
                      ("dt_c_acc", "<f4"), ("cur_checkpoint", "<i4"), ("cur_projection", "<i4"), ("cur_halofind", "<i4"), ("mass_p", "<f4")])
HDR_NOPP = np.dtype([(n, t) for n, t in HDR_PPINT.descr if n != "dt_pp_acc"])


def test_checkpoint_name_matches_f7_3_adjustl():
    assert checkpoint_name(10.0, 0) == "10.000xv0.dat"          # write(z_s,'(f7.3)'), adjustl (checkpoint.f90:31-46)
    assert checkpoint_name(0.0, 7, "PID") == "0.000PID7.dat"
    assert checkpoint_name(127.5, 12) == "127.500xv12.dat"
    assert HDR_PPINT.itemsize == 48 and HDR_NOPP.itemsize == 44


def _hdr(np_local):
    h = CheckpointHeader()
    h.np_local, h.a, h.t, h.tau, h.nts = np_local, 0.0909, 1.25, -9.5, 231
    h.dt_f_acc, h.dt_pp_acc, h.dt_c_acc = 3.5, 0.75, 12.0
    h.cur_checkpoint, h.cur_projection, h.cur_halofind, h.mass_p = 2, 1, 1, 8.0
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("ppint,pid", [(1, 0), (0, 1)])
def test_checkpoint_write_read_roundtrip(built, tmp_path, ppint, pid):
    from cubep3m_b200.lib import ParticleMesh
    cfg = default_config(nf_tile=112, tiles_node_dim=2, ppint=ppint, pp_ext=0, pid=pid)
    xv = ic.zeldovich_ics(cfg.nf_physical_dim, box=50.0, z_i=20.0, seed=5)
    ids = np.arange(len(xv), dtype=np.int64) * 11 + 5
    shake = np.array([3.25, -1.5, 0.0078125], np.float32)
    pm = ParticleMesh(cfg)
    pm.upload_particles(xv, ids if pid else None)
    fx = str(tmp_path / checkpoint_name(10.0, 0)); fp = str(tmp_path / checkpoint_name(10.0, 0, "PID"))
    pm.write_checkpoint(fx, _hdr(-1), shake, fp if pid else None)
    # --- the file, read as the reference's reader does
    hd = HDR_PPINT if ppint else HDR_NOPP
    raw = np.fromfile(fx, dtype=np.uint8)
    assert len(raw) == hd.itemsize + 24 * len(xv)
    h = raw[: hd.itemsize].view(hd)[0]
    assert h["np_local"] == len(xv) and h["nts"] == 231 and h["cur_checkpoint"] == 2 and h["mass_p"] == np.float32(8.0)
    assert h["dt_c_acc"] == np.float32(12.0) and (not ppint or h["dt_pp_acc"] == np.float32(0.75))
    body = raw[hd.itemsize:].view(np.float32).reshape(-1, 6)
    want = xv.copy(); want[:, :3] = xv[:, :3] - shake            # checkpoint.f90:92, one fp32 subtraction per coordinate
    assert np.array_equal(body, want)
    if pid:
        rawp = np.fromfile(fp, dtype=np.uint8)
        assert np.array_equal(rawp[: hd.itemsize], raw[: hd.itemsize]) and np.array_equal(rawp[hd.itemsize:].view(np.int64), ids)
    # --- restart: a fresh context reads the files
    pm2 = ParticleMesh(cfg)
    h2 = pm2.read_checkpoint(fx, fp if pid else None)
    assert h2.np_local == len(xv) and h2.nts == 231 and h2.a == pytest.approx(0.0909) and h2.cur_halofind == 1
    got = pm2.download_particles(with_pid=bool(pid))
    if pid:
        assert np.array_equal(got[0], want) and np.array_equal(got[1], ids)
    else:
        assert np.array_equal(got, want)
    out = pm2.particle_mesh(0.1, 0.1, 0.05, 8.0)                # the restarted state is usable
    assert out.np_total == len(xv)
    pm.close(); pm2.close()
    # a file written by the reference's own writer (numpy twin of the record list) is accepted too
    ref = str(tmp_path / "ref.dat")
    hh = np.zeros(1, hd); hh["np_local"] = 1000; hh["a"] = 0.5; hh["mass_p"] = 8.0
    with open(ref, "wb") as f:
        f.write(hh.tobytes()); f.write(xv[:1000].tobytes())
    pm3 = ParticleMesh(default_config(nf_tile=112, tiles_node_dim=2, ppint=ppint, pp_ext=0))
    h3 = pm3.read_checkpoint(ref)
    assert h3.np_local == 1000 and h3.a == 0.5 and np.array_equal(pm3.download_particles(), xv[:1000])
    pm3.close()
