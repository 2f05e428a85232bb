"""ctypes binding of libcubep3m_b200.so (the C ABI of include/cubep3m_b200.h).

`ParticleMesh` mirrors the reference's subroutine names (particle_mesh, update_position, link_list,
particle_pass, delete_particles, move_grid_back — particle_mesh_threaded.f90:2, cubepm.f90:98-228) so the
parity tests read like a driver. There is no CPU fallback: if the CUDA library is missing or no GPU is
visible, construction raises."""
import ctypes as C
import os
import subprocess
import numpy as np

from .abi import Config, StepOut, Clock, CheckpointHeader, ERRORS, max_np
from . import tables

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CUBEP3M_B200_SO", os.path.join(_HERE, "libcubep3m_b200.so"))   # override only for A/B builds
_LIB = None
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

# every symbol include/cubep3m_b200.h declares
SYMBOLS = ["cubep3m_b200_version", "cubep3m_b200_strerror", "cubep3m_b200_default_config", "cubep3m_b200_init",
           "cubep3m_b200_get_unique_id", "cubep3m_b200_finalize", "cubep3m_b200_upload_particles",
           "cubep3m_b200_download_particles", "cubep3m_b200_particle_mesh", "cubep3m_b200_update_position",
           "cubep3m_b200_link_list", "cubep3m_b200_particle_pass", "cubep3m_b200_delete_particles",
           "cubep3m_b200_move_grid_back", "cubep3m_b200_debug_cell_counts", "cubep3m_b200_debug_tile_counts",
           "cubep3m_b200_debug_sorted_particles", "cubep3m_b200_debug_kern_f", "cubep3m_b200_debug_kern_c",
           "cubep3m_b200_debug_rho_c", "cubep3m_b200_debug_force_c", "cubep3m_b200_debug_fine_tile",
           "cubep3m_b200_debug_fft3d", "cubep3m_b200_debug_ppext_blocks", "cubep3m_b200_debug_pair_counts", "cubep3m_b200_launch_count", "cubep3m_b200_set_profiling", "cubep3m_b200_set_tile_streams",
           "cubep3m_b200_num_kernel_classes", "cubep3m_b200_kernel_class_name", "cubep3m_b200_get_kernel_times",
           "cubep3m_b200_cic_power", "cubep3m_b200_dist_init", "cubep3m_b200_write_checkpoint", "cubep3m_b200_read_checkpoint", "cubep3m_b200_clock_init",
           "cubep3m_b200_expansion", "cubep3m_b200_timestep", "cubep3m_b200_timestep_device", "cubep3m_b200_halofind_peaks"]


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a (cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "cubep3m_b200.h"))
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", src_dir], capture_output=not verbose, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libcubep3m_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return SO_PATH


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    L.cubep3m_b200_version.restype = C.c_char_p
    L.cubep3m_b200_strerror.restype = C.c_char_p
    L.cubep3m_b200_strerror.argtypes = [C.c_int]
    L.cubep3m_b200_default_config.argtypes = [C.POINTER(Config)]
    L.cubep3m_b200_init.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.POINTER(C.c_void_p)]
    L.cubep3m_b200_get_unique_id.argtypes = [C.c_void_p]
    L.cubep3m_b200_finalize.argtypes = [C.c_void_p]
    L.cubep3m_b200_upload_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.cubep3m_b200_download_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    L.cubep3m_b200_particle_mesh.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, _fp, C.POINTER(StepOut)]
    L.cubep3m_b200_update_position.argtypes = [C.c_void_p, C.c_float, C.c_float, _fp]
    L.cubep3m_b200_link_list.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.cubep3m_b200_particle_pass.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.cubep3m_b200_delete_particles.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.cubep3m_b200_move_grid_back.argtypes = [C.c_void_p, _fp]
    L.cubep3m_b200_debug_cell_counts.argtypes = [C.c_void_p, _ip]
    L.cubep3m_b200_debug_tile_counts.argtypes = [C.c_void_p, _ip]
    L.cubep3m_b200_debug_sorted_particles.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    L.cubep3m_b200_debug_kern_f.argtypes = [C.c_void_p, _fp]
    L.cubep3m_b200_debug_kern_c.argtypes = [C.c_void_p, _fp]
    L.cubep3m_b200_debug_rho_c.argtypes = [C.c_void_p, _fp]
    L.cubep3m_b200_debug_force_c.argtypes = [C.c_void_p, _fp]
    L.cubep3m_b200_debug_fine_tile.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    L.cubep3m_b200_debug_fft3d.argtypes = [C.c_void_p, C.c_int32, _fp, C.c_int32]
    L.cubep3m_b200_debug_pair_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.cubep3m_b200_launch_count.argtypes = [C.c_void_p]
    L.cubep3m_b200_launch_count.restype = C.c_int64
    L.cubep3m_b200_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.cubep3m_b200_set_tile_streams.argtypes = [C.c_void_p, C.c_int]
    L.cubep3m_b200_cic_power.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_double, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.c_int32]
    L.cubep3m_b200_dist_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_uint64, _fp, _fp, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    L.cubep3m_b200_write_checkpoint.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(CheckpointHeader), C.POINTER(C.c_float)]
    L.cubep3m_b200_read_checkpoint.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(CheckpointHeader)]
    L.cubep3m_b200_kernel_class_name.restype = C.c_char_p
    L.cubep3m_b200_kernel_class_name.argtypes = [C.c_int]
    L.cubep3m_b200_get_kernel_times.argtypes = [C.c_void_p, _fp, C.c_void_p]
    L.cubep3m_b200_clock_init.argtypes = [C.POINTER(Clock), C.c_float, C.c_float, C.c_float]
    L.cubep3m_b200_expansion.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.cubep3m_b200_timestep.argtypes = [C.POINTER(Clock)]
    L.cubep3m_b200_timestep_device.argtypes = [C.c_void_p, C.POINTER(Clock)]
    L.cubep3m_b200_halofind_peaks.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    _LIB = L
    return L


def get_unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 creates it, the launcher broadcasts it: INTEGRATION.md)."""
    buf = C.create_string_buffer(128)
    st = load_library().cubep3m_b200_get_unique_id(buf)
    if st != 0:
        raise Cubep3mError(st)
    return buf.raw


class Cubep3mError(RuntimeError):
    def __init__(self, status):
        self.status = status
        super().__init__(f"cubep3m_b200: {ERRORS.get(status, status)} (status {status})")


def _chk(st):
    if st != 0:
        raise Cubep3mError(st)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ParticleMesh:
    """One rank <-> one GPU. All state lives on the device (resident mode); upload/download move xv explicitly."""

    def __init__(self, cfg: Config, kern_f=None, kern_c=None, nccl_id=None, world_size=1):
        self.lib = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        ft, ct = tables.fine_table(), tables.coarse_table()
        kf = None if kern_f is None else np.ascontiguousarray(kern_f, np.float32)
        kc = None if kern_c is None else np.ascontiguousarray(kern_c, np.float32)
        idbuf = None if nccl_id is None else C.create_string_buffer(bytes(nccl_id), 128)
        _chk(self.lib.cubep3m_b200_init(C.byref(cfg), _ptr(ft), _ptr(ct), _ptr(kf), _ptr(kc), idbuf, world_size, C.byref(self.h)))
        self.max_np = max_np(cfg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cubep3m_b200_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- data movement (strict drop-in mode calls these around every step)
    def upload_particles(self, xv, pid=None):
        xv = np.ascontiguousarray(xv, np.float32).reshape(-1, 6)
        p = None if pid is None else np.ascontiguousarray(pid, np.int64)
        _chk(self.lib.cubep3m_b200_upload_particles(self.h, _ptr(xv), _ptr(p), xv.shape[0]))

    def download_particles(self, out=None, with_pid=False):
        n = C.c_int32()
        _chk(self.lib.cubep3m_b200_download_particles(self.h, None, None, C.byref(n)))
        xv = out if out is not None else np.empty((n.value, 6), np.float32)
        pid = np.empty(n.value, np.int64) if with_pid else None
        _chk(self.lib.cubep3m_b200_download_particles(self.h, _ptr(xv), _ptr(pid), C.byref(n)))
        xv = xv[: n.value]
        return (xv, pid) if with_pid else xv

    # --- the reference's subroutines
    def particle_mesh(self, dt, dt_old, a_mid, mass_p, offset=(0, 0, 0)) -> StepOut:
        out = StepOut()
        _chk(self.lib.cubep3m_b200_particle_mesh(self.h, dt, dt_old, a_mid, mass_p, np.asarray(offset, np.float32), C.byref(out)))
        return out

    def update_position(self, dt, dt_old, offset=(0, 0, 0)):
        _chk(self.lib.cubep3m_b200_update_position(self.h, dt, dt_old, np.asarray(offset, np.float32)))

    def link_list(self):
        n = C.c_int32()
        _chk(self.lib.cubep3m_b200_link_list(self.h, C.byref(n)))
        return n.value

    def particle_pass(self):
        n = C.c_int32()
        _chk(self.lib.cubep3m_b200_particle_pass(self.h, C.byref(n)))
        return n.value

    def delete_particles(self):
        n = C.c_int32()
        _chk(self.lib.cubep3m_b200_delete_particles(self.h, C.byref(n)))
        return n.value

    def move_grid_back(self, shake):
        _chk(self.lib.cubep3m_b200_move_grid_back(self.h, np.asarray(shake, np.float32)))

    # --- parity getters
    def cell_counts(self):
        H = self.cfg.H
        a = np.empty(H * H * H, np.int32)
        _chk(self.lib.cubep3m_b200_debug_cell_counts(self.h, a))
        return a.reshape(H, H, H)

    def tile_counts(self):
        a = np.empty(self.cfg.tiles_node, np.int32)
        _chk(self.lib.cubep3m_b200_debug_tile_counts(self.h, a))
        return a

    def ppext_blocks(self):
        """(target blocks of the tiled PP_EXT kernel, blocks that fell back to the direct walk) of the last particle_mesh call."""
        nb, nf = C.c_int32(), C.c_int32()
        _chk(self.lib.cubep3m_b200_debug_ppext_blocks(self.h, C.byref(nb), C.byref(nf)))
        return nb.value, nf.value

    def pair_counts(self):
        """(PPINT, PP_EXT) ordered pair interactions evaluated in the last particle_mesh call."""
        a, b = C.c_int64(), C.c_int64()
        _chk(self.lib.cubep3m_b200_debug_pair_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sorted_particles(self):
        n = C.c_int32()
        _chk(self.lib.cubep3m_b200_debug_sorted_particles(self.h, None, C.byref(n)))
        xv = np.empty((n.value, 6), np.float32)
        _chk(self.lib.cubep3m_b200_debug_sorted_particles(self.h, _ptr(xv), C.byref(n)))
        return xv

    def kern_f(self):
        n = self.cfg.nf_tile
        a = np.empty((n, n, n // 2 + 1, 3), np.float32)
        _chk(self.lib.cubep3m_b200_debug_kern_f(self.h, a.reshape(-1)))
        return a

    def kern_c(self):
        Nx, Ny, Nz = self.cfg.nc_dims
        cubic = not all(v > 0 for v in self.cfg.nodes_dim_xyz) and self.cfg.nc_dim % self.cfg.nodes == 0
        a = np.empty((self.cfg.nc_slab if cubic else Nz, Ny, Nx // 2 + 1, 3), np.float32)
        _chk(self.lib.cubep3m_b200_debug_kern_c(self.h, a.reshape(-1)))
        return a

    def force_c(self):
        nc = self.cfg.nc_node + 2
        a = np.empty((nc, nc, nc, 3), np.float32)
        _chk(self.lib.cubep3m_b200_debug_force_c(self.h, a.reshape(-1)))
        return a

    def fine_tile(self, tile, mass_p, want_force=True):
        """tile is 0-based (cur_tile - 1)."""
        n, f = self.cfg.nf_tile, self.cfg.m + 3
        rho = np.empty((n, n, n + 2), np.float32)
        frc = np.empty((f, f, f, 3), np.float32) if want_force else None
        _chk(self.lib.cubep3m_b200_debug_fine_tile(self.h, tile, mass_p, _ptr(rho), _ptr(frc)))
        return rho, frc

    def halofind_peaks(self, mass_p, den_peak_cutoff=100.0, para_inter_hc=True, ngph=False, max_peaks=1 << 20):
        """find_halos' density + maxima pass (halofind.f90:564-672) over every tile; call after link_list + particle_pass (cubepm.f90:193-198).
        Returns (structured array of peaks, (cftmass, cftmass2))."""
        from .abi import PEAK_DTYPE
        pk = np.zeros(max_peaks, dtype=PEAK_DTYPE)
        n = C.c_int32(0)
        cft = (C.c_double * 2)()
        _chk(self.lib.cubep3m_b200_halofind_peaks(self.h, mass_p, den_peak_cutoff, int(para_inter_hc), int(ngph), pk.ctypes.data, max_peaks, C.byref(n), cft))
        return pk[:n.value].copy(), (cft[0], cft[1])

    def timestep_device(self, c):
        """timestep.f90 evaluated on the device (same state transition as timestep(c))."""
        _chk(self.lib.cubep3m_b200_timestep_device(self.h, C.byref(c)))
        return c

    def fft3d(self, a, inverse=False):
        n = a.shape[0]
        a = np.ascontiguousarray(a, np.float32)
        assert a.shape == (n, n, n + 2)
        _chk(self.lib.cubep3m_b200_debug_fft3d(self.h, n, a.reshape(-1), 1 if inverse else 0))
        return a

    def cic_power(self, box, shake=(0.0, 0.0, 0.0), ngp_binning=True):
        """Device twin of utils/cic_power (cic_power.f90:840-954): returns (k [h/Mpc], Delta^2, sigma) for shells 1..nc/2 of the resident
        physical particles after subtracting the accumulated shake offset (checkpoint.f90:92)."""
        n = self.cfg.mT * self.cfg.grid[0] // 2
        k, d2, sg = (np.empty(n, np.float64) for _ in range(3))
        off = (C.c_float * 3)(*[float(v) for v in shake])
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        _chk(self.lib.cubep3m_b200_cic_power(self.h, off, C.c_double(box), 1 if ngp_binning else 0, dp(k), dp(d2), dp(sg), n))
        return k, d2, sg

    def dist_init(self, nc, box, z_i, reps=1, seed=12345, noise=None, om=0.24, ol=0.76, table=None):
        """Device twin of utils/dist_init (dist_init_dm.f90): Zel'dovich ICs generated in place; returns np_local. `table` = (k, Delta^2) at the
        initial epoch; default: the Eisenstein-Hu no-wiggle spectrum of cubep3m_b200/ic.py (the reference reads a CAMB table)."""
        from . import ic
        a = 1.0 / (1.0 + z_i)
        if table is None:
            k = np.logspace(-4, 2, 2048)
            table = (k, ic.delta2(k, a, om, ol))
        kt = np.ascontiguousarray(table[0], np.float32); dt = np.ascontiguousarray(table[1], np.float32)
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        n = C.c_int32()
        _chk(self.lib.cubep3m_b200_dist_init(self.h, int(nc), int(reps), float(box), float(ic.vfactor(a, om, ol)), int(seed), kt, dt, len(kt), _ptr(nz), C.byref(n)))
        return n.value

    def write_checkpoint(self, path_xv, hdr: CheckpointHeader, shake=(0.0, 0.0, 0.0), path_pid=None):
        """checkpoint.f90:72-95: header + (x - shake_offset, v) records of the resident particles (+ the PID file with -DPID_FLAG)."""
        off = (C.c_float * 3)(*[float(v) for v in shake])
        _chk(self.lib.cubep3m_b200_write_checkpoint(self.h, os.fsencode(path_xv), None if path_pid is None else os.fsencode(path_pid), C.byref(hdr), off))

    def read_checkpoint(self, path_xv, path_pid=None) -> CheckpointHeader:
        """particle_initialization.f90:88-189 (restart_ic): fills the device copy from the checkpoint, returns the header."""
        hdr = CheckpointHeader()
        _chk(self.lib.cubep3m_b200_read_checkpoint(self.h, os.fsencode(path_xv), None if path_pid is None else os.fsencode(path_pid), C.byref(hdr)))
        return hdr

    def set_profiling(self, on=True):
        _chk(self.lib.cubep3m_b200_set_profiling(self.h, 1 if on else 0))

    def set_tile_streams(self, n):
        _chk(self.lib.cubep3m_b200_set_tile_streams(self.h, int(n)))

    def kernel_times(self):
        """{class name: (ms, launches)} of the last profiled particle_mesh call."""
        k = self.lib.cubep3m_b200_num_kernel_classes()
        ms = np.zeros(k, np.float32)
        n = np.zeros(k, np.int64)
        _chk(self.lib.cubep3m_b200_get_kernel_times(self.h, ms, _ptr(n)))
        return {self.lib.cubep3m_b200_kernel_class_name(i).decode(): (float(ms[i]), int(n[i])) for i in range(k)}

    @property
    def launches(self):
        return int(self.lib.cubep3m_b200_launch_count(self.h))


# ---- driver twin (timestep.f90) -----------------------------------------------------------------
def clock_init(z_i, omega_m=0.24, omega_l=0.76, ppint=1, pp_ext=0, a_target=1.0) -> Clock:
    c = Clock()
    load_library().cubep3m_b200_clock_init(C.byref(c), z_i, omega_m, omega_l)
    c.ppint, c.pp_ext, c.a_target = ppint, pp_ext, a_target
    return c


def timestep(c: Clock):
    load_library().cubep3m_b200_timestep(C.byref(c))
    return c


def absorb_limiters(c: Clock, out: StepOut):
    c.dt_f_acc, c.dt_pp_acc, c.dt_pp_ext_acc, c.dt_c_acc = out.dt_f_acc, out.dt_pp_acc, out.dt_pp_ext_acc, out.dt_c_acc
