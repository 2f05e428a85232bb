"""TEST INFRASTRUCTURE: ctypes binding of liboracle. See oracle/__init__.py for who may import it."""
import ctypes as C
import os
import subprocess
import numpy as np

from cubep3m_b200.abi import Config, StepOut, ERRORS, max_np
from cubep3m_b200 import tables

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    so = os.path.join(_HERE, "libcubep3m_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cubep3m_oracle.cpp", "fft_ref.h")] + \
           [os.path.join(_HERE, "..", "include", "cubep3m_b200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_oracle())
        _LIB.oracle_create.argtypes = [C.POINTER(Config), _fp, _fp, C.c_int, C.POINTER(C.c_void_p)]
        _LIB.oracle_destroy.argtypes = [C.c_void_p]
        _LIB.oracle_max_np.argtypes = [C.c_void_p]
        _LIB.oracle_set_kernels.argtypes = [C.c_void_p, _fp, _fp]
        _LIB.oracle_set_particles.argtypes = [C.c_void_p, C.c_int, _fp, C.c_void_p, C.c_int]
        _LIB.oracle_get_np.argtypes = [C.c_void_p, C.c_int]
        _LIB.oracle_get_particles.argtypes = [C.c_void_p, C.c_int, _fp, C.c_void_p]
        _LIB.oracle_update_position.argtypes = [C.c_void_p, C.c_float, C.c_float, _fp]
        _LIB.oracle_move_grid_back.argtypes = [C.c_void_p, _fp]
        for f in ("oracle_link_list", "oracle_particle_pass", "oracle_delete_particles"):
            getattr(_LIB, f).argtypes = [C.c_void_p]
        _LIB.oracle_particle_mesh.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, _fp, C.POINTER(StepOut)]
        _LIB.oracle_cell_counts.argtypes = [C.c_void_p, C.c_int, _ip]
        _LIB.oracle_tile_counts.argtypes = [C.c_void_p, C.c_int, _ip]
        _LIB.oracle_kern_f.argtypes = [C.c_void_p, _fp]
        _LIB.oracle_kern_c.argtypes = [C.c_void_p, C.c_int, _fp]
        _LIB.oracle_rho_c.argtypes = [C.c_void_p, C.c_int, _fp]
        _LIB.oracle_force_c.argtypes = [C.c_void_p, C.c_int, _fp]
        _LIB.oracle_set_debug_tile.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _LIB.oracle_fine_tile.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
        _LIB.oracle_fft3d.argtypes = [C.c_int, _fp, C.c_int]
        _LIB.oracle_find_peaks.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        _LIB.oracle_set_num_threads.argtypes = [C.c_int]
    return _LIB


def oracle_fft3d(a: np.ndarray, inverse=False) -> np.ndarray:
    """In-place padded r2c / c2r on a C-ordered array of shape (n, n, n+2) (= Fortran (n+2,n,n))."""
    n = a.shape[0]
    assert a.shape == (n, n, n + 2) and a.dtype == np.float32
    a = np.ascontiguousarray(a)
    _lib().oracle_fft3d(n, a, 1 if inverse else 0)
    return a


def _chk(st):
    if st != 0:
        raise RuntimeError(f"oracle: {ERRORS.get(st, st)}")


class Oracle:
    """All D^3 ranks of a run in one process; mirrors the subroutine names of the reference."""

    def __init__(self, cfg: Config, build_kernels=True, threads=None):
        self.cfg = cfg
        self.lib = _lib()
        if threads:
            self.lib.oracle_set_num_threads(int(threads))
        self.h = C.c_void_p()
        _chk(self.lib.oracle_create(C.byref(cfg), tables.fine_table().ravel(), tables.coarse_table().ravel(),
                                    1 if build_kernels else 0, C.byref(self.h)))
        self.max_np = self.lib.oracle_max_np(self.h)
        assert self.max_np == max_np(cfg), (self.max_np, max_np(cfg))

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self):
        return self.lib.oracle_num_threads()

    def set_particles(self, xv, rank=0, pid=None):
        xv = np.ascontiguousarray(xv, dtype=np.float32).reshape(-1, 6)
        p = None if pid is None else np.ascontiguousarray(pid, dtype=np.int64).ctypes.data_as(C.c_void_p)
        _chk(self.lib.oracle_set_particles(self.h, rank, xv, p, xv.shape[0]))

    def get_particles(self, rank=0, with_pid=False):
        n = self.lib.oracle_get_np(self.h, rank)
        xv = np.empty((n, 6), np.float32)
        pid = np.empty(n, np.int64) if with_pid else None
        self.lib.oracle_get_particles(self.h, rank, xv.reshape(-1), None if pid is None else pid.ctypes.data_as(C.c_void_p))
        return (xv, pid) if with_pid else xv

    def update_position(self, dt, dt_old, offset=(0, 0, 0)):
        self.lib.oracle_update_position(self.h, dt, dt_old, np.asarray(offset, np.float32))

    def move_grid_back(self, shake):
        self.lib.oracle_move_grid_back(self.h, np.asarray(shake, np.float32))

    def link_list(self):
        self.lib.oracle_link_list(self.h)

    def particle_pass(self):
        _chk(self.lib.oracle_particle_pass(self.h))

    def delete_particles(self):
        self.lib.oracle_delete_particles(self.h)

    def particle_mesh(self, dt, dt_old, a_mid, mass_p, offset=(0, 0, 0)) -> StepOut:
        out = StepOut()
        _chk(self.lib.oracle_particle_mesh(self.h, dt, dt_old, a_mid, mass_p, np.asarray(offset, np.float32), C.byref(out)))
        return out

    def cell_counts(self, rank=0):
        H = self.cfg.H
        a = np.empty(H * H * H, np.int32)
        self.lib.oracle_cell_counts(self.h, rank, a)
        return a.reshape(H, H, H)

    def tile_counts(self, rank=0):
        a = np.empty(self.cfg.tiles_node, np.int32)
        self.lib.oracle_tile_counts(self.h, rank, a)
        return a

    def kern_f(self):
        n = self.cfg.nf_tile
        a = np.empty((n, n, n // 2 + 1, 3), np.float32)
        self.lib.oracle_kern_f(self.h, a.reshape(-1))
        return a

    def kern_c(self, rank=0):
        Nx, Ny, Nz = self.cfg.nc_dims
        cubic = not all(v > 0 for v in self.cfg.nodes_dim_xyz) and self.cfg.nc_dim % self.cfg.nodes == 0
        a = np.empty((self.cfg.nc_slab if cubic else Nz, Ny, Nx // 2 + 1, 3), np.float32)
        self.lib.oracle_kern_c(self.h, rank, a.reshape(-1))
        return a

    def rho_c(self, rank=0):
        nc = self.cfg.nc_node
        a = np.empty((nc, nc, nc), np.float32)
        self.lib.oracle_rho_c(self.h, rank, a.reshape(-1))
        return a

    def force_c(self, rank=0):
        nc = self.cfg.nc_node + 2
        a = np.empty((nc, nc, nc, 3), np.float32)
        self.lib.oracle_force_c(self.h, rank, a.reshape(-1))
        return a

    def set_debug_tile(self, tile, rank=0):
        """tile is 1-based cur_tile (particle_mesh_threaded.f90:85)."""
        self.lib.oracle_set_debug_tile(self.h, rank, tile)

    def find_peaks(self, mass_p, den_peak_cutoff=100.0, para_inter_hc=True, ngph=False, rank=0, max_peaks=1 << 20):
        """halofind.f90:564-672: peaks tile by tile in scan order (before the reference's sort), and (cftmass, cftmass2)."""
        dt = [("i", "<i4"), ("j", "<i4"), ("k", "<i4"), ("tile", "<i4"), ("den", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
        pk = np.zeros(max_peaks, dtype=dt)
        n = C.c_int(0)
        cft = (C.c_double * 2)()
        _chk(self.lib.oracle_find_peaks(self.h, rank, mass_p, den_peak_cutoff, int(para_inter_hc), int(ngph), pk.ctypes.data, max_peaks, C.byref(n), cft))
        return pk[:n.value].copy(), (cft[0], cft[1])

    def fine_tile(self, rank=0):
        n, f = self.cfg.nf_tile, self.cfg.m + 3
        rho = np.empty((n, n, n + 2), np.float32)
        frc = np.empty((f, f, f, 3), np.float32)
        _chk(self.lib.oracle_fine_tile(self.h, rank, rho.reshape(-1), frc.reshape(-1)))
        return rho, frc
