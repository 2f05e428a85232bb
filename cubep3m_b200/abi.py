"""ctypes mirror of include/cubep3m_b200.h (structs + constants). No compute here."""
import ctypes as C
import math

ST_NAMES = ["drift", "link", "pass", "fine_deposit", "fine_fft", "fine_kick", "pp", "pp_ext",
            "coarse_mass", "coarse_force", "coarse_vel", "delete", "total"]
ST_COUNT = 13
# the reference's -DMPI_TIME tags (timers.f90:68-77) for the stages that have one
REF_TAGS = {"drift": "pos updt", "link": "linklist", "pass": "par pass", "coarse_mass": "cm  mass",
            "coarse_force": "cm force", "coarse_vel": "cm   vel", "delete": "del part"}

ERRORS = {0: "ok", 1: "bad configuration", 2: "CUDA failure", 3: "not enough buffer space in pass",
          4: "exceeded max_np in pass", 5: "exceeded max_llf", 6: "NCCL failure", 7: "call order violated",
          8: "internal work list overflow"}


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("nodes_dim", "tiles_node_dim", "nf_tile", "nf_buf", "nf_cutoff", "mesh_scale", "pp_range",
                 "max_np", "max_buf", "max_llf")] + \
               [(n, C.c_float) for n in ("density_buffer", "rsoft", "pp_bias", "dt_pp_scale", "G", "eps")] + \
               [(n, C.c_int32) for n in
                ("ngp", "ppint", "pp_ext", "coarse_ngp", "pid", "lrckcorr", "move_grid_back",
                 "ngp_fmesh_force", "pp_force_flag", "pp_ext_force_flag", "coarse_vel_update",
                 "rank", "local_gpu")] + \
               [("nodes_dim_xyz", C.c_int32 * 3)]

    # derived sizes, cubepm.par:190-208
    @property
    def m(self): return self.nf_tile - 2 * self.nf_buf
    @property
    def mT(self): return self.m * self.tiles_node_dim
    @property
    def nc_node(self): return self.mT // self.mesh_scale
    @property
    def grid(self):
        g = tuple(self.nodes_dim_xyz)
        return g if all(v > 0 for v in g) else (self.nodes_dim,) * 3
    @property
    def nc_dim(self): return self.nc_node * self.nodes_dim
    @property
    def nc_dims(self): return tuple(self.nc_node * d for d in self.grid)      # global coarse mesh (Nx, Ny, Nz)
    @property
    def nodes(self):
        g = self.grid
        return g[0] * g[1] * g[2]
    def rank_coords(self, rank):
        g = self.grid
        return (rank % g[0], (rank // g[0]) % g[1], rank // (g[0] * g[1]))
    @property
    def nc_slab(self): return self.nc_dim // self.nodes
    @property
    def hoc_l(self): return 1 - self.nf_buf // self.mesh_scale
    @property
    def hoc_h(self): return self.nc_node + self.nf_buf // self.mesh_scale
    @property
    def H(self): return self.hoc_h - self.hoc_l + 1
    @property
    def tiles_node(self): return self.tiles_node_dim ** 3
    @property
    def nf_physical_dim(self): return self.mT * self.nodes_dim


def default_config(**kw) -> Config:
    """parameters.example + cubepm.par + the cpp flags of Make_PP_THREADS:10 (-DNGP -DPPINT -DLRCKCORR; PP_EXT off)."""
    import numpy as np
    c = Config()
    c.nodes_dim, c.tiles_node_dim, c.nf_tile, c.nf_buf, c.nf_cutoff, c.mesh_scale = 1, 2, 176, 24, 16, 4
    c.pp_range, c.max_np, c.max_buf, c.max_llf = 2, 0, 0, 100000
    c.density_buffer, c.rsoft, c.pp_bias, c.dt_pp_scale = 2.0, 0.1, 1.0, 0.05
    pi = np.float32(3.141592654)
    c.G = float(np.float32(1.0) / np.float32(6.0) / pi)   # cubepm.par:149
    c.eps = 1.0e-3
    c.ngp, c.ppint, c.pp_ext, c.coarse_ngp, c.pid, c.lrckcorr, c.move_grid_back = 1, 1, 0, 0, 0, 1, 0
    c.ngp_fmesh_force = c.pp_force_flag = c.pp_ext_force_flag = c.coarse_vel_update = 1
    c.rank, c.local_gpu = 0, 0
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        if k == "nodes_dim_xyz":
            for i in range(3):
                c.nodes_dim_xyz[i] = int(v[i])
        else:
            setattr(c, k, v)
    return c


def copy_config(c: Config, **kw) -> Config:
    d = Config()
    C.memmove(C.byref(d), C.byref(c), C.sizeof(Config))
    for k, v in kw.items():
        setattr(d, k, v)
    return d


class StepOut(C.Structure):
    _fields_ = [("np_local", C.c_int32), ("np_with_ghosts", C.c_int32), ("np_deleted_ll", C.c_int32),
                ("np_buf_max", C.c_int32),
                ("dt_f_acc", C.c_float), ("dt_pp_acc", C.c_float), ("dt_pp_ext_acc", C.c_float),
                ("dt_c_acc", C.c_float), ("f_force_max", C.c_float), ("pp_force_max", C.c_float),
                ("pp_ext_force_max", C.c_float), ("c_force_max", C.c_float),
                ("sum_rho_f", C.c_double), ("sum_rho_c", C.c_double), ("np_total", C.c_int64),
                ("stage_ms", C.c_float * 16)]

    def stages(self):
        return {ST_NAMES[i]: float(self.stage_ms[i]) for i in range(ST_COUNT)}


class Clock(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("a", "a_mid", "t", "tau", "dt", "dt_old", "da",
                                         "dt_f_acc", "dt_pp_acc", "dt_pp_ext_acc", "dt_c_acc",
                                         "omega_m", "omega_l", "wde", "a_target")] + \
               [(n, C.c_int32) for n in ("nts", "ppint", "pp_ext", "cosmo", "checkpoint_step")]


PEAK_DTYPE = [("i", "<i4"), ("j", "<i4"), ("k", "<i4"), ("tile", "<i4"), ("den", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4")]   # cubep3m_b200_peak


class CheckpointHeader(C.Structure):
    """checkpoint.f90:72-78 (the file holds dt_pp_acc only with -DPPINT)."""
    _fields_ = [("np_local", C.c_int32), ("a", C.c_float), ("t", C.c_float), ("tau", C.c_float), ("nts", C.c_int32),
                ("dt_f_acc", C.c_float), ("dt_pp_acc", C.c_float), ("dt_c_acc", C.c_float),
                ("cur_checkpoint", C.c_int32), ("cur_projection", C.c_int32), ("cur_halofind", C.c_int32), ("mass_p", C.c_float)]


def checkpoint_name(z, rank, kind="xv"):
    """checkpoint.f90:31-46: write(z_s,'(f7.3)') z; adjustl; <z>xv<rank>.dat / <z>PID<rank>.dat"""
    return f"{z:7.3f}".strip() + kind + f"{rank:d}" + ".dat"


def max_np(c: Config) -> int:
    """cubepm.par:170-172."""
    if c.max_np > 0:
        return c.max_np
    mT, b = c.mT, c.nf_buf
    half = float(mT // 2)
    buf = (8.0 * b ** 3 + 6.0 * b * float(mT) ** 2 + 12.0 * b * b * mT) / 8.0
    return int(c.density_buffer * (half ** 3 + buf))
