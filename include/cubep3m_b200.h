/*
 * cubep3m_b200.h — C ABI of the B200-native replacement for CUBEP3M's `call particle_mesh`.
 *
 * Every entry point is `extern "C"`, takes plain pointers / scalars, and returns an int status
 * (0 = ok, see CUBEP3M_B200_E*).  The Fortran side binds them with ISO_C_BINDING (see
 * fortran/particle_mesh_b200.f90 and INTEGRATION.md); the tests bind them with ctypes.
 *
 * Each function names the reference interface it replaces (paths relative to the reference root).
 * All reals are IEEE-754 binary32, exactly as the reference's real(4).
 */
#ifndef CUBEP3M_B200_H
#define CUBEP3M_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; the messages mirror the reference's abort texts */
#define CUBEP3M_B200_OK              0
#define CUBEP3M_B200_EINVAL          1  /* bad configuration (mpi_initialization.f90:18-29 style aborts)      */
#define CUBEP3M_B200_ECUDA           2  /* CUDA runtime failure                                                */
#define CUBEP3M_B200_EPASSBUF        3  /* 'not enough buffer space in pass'   particle_pass.f90:96-99          */
#define CUBEP3M_B200_EMAXNP          4  /* 'exceeded max_np in pass'           particle_pass.f90:136-139        */
#define CUBEP3M_B200_EMAXLLF         5  /* 'exceeded max_llf'                  particle_mesh_threaded.f90:280   */
#define CUBEP3M_B200_ENCCL           6  /* NCCL failure (replaces an MPI error)                                */
#define CUBEP3M_B200_ENOTREADY       7  /* call order violated (e.g. particle_mesh before upload)              */
#define CUBEP3M_B200_ECAPACITY       8  /* an internal work list overflowed (boundary candidates / per-tile moves); nothing is truncated silently */

/*
 * Compile-time `parameter`s of the reference turned into runtime values
 * (parameters.example:14-56, cubepm.par:76-92,148-215) plus the cpp flags that change the numerics
 * (Make_PP_THREADS:10).  Derived sizes (nc_tile_dim, nc_node_dim, hoc_nc_l/h ...) are recomputed
 * inside the library exactly as cubepm.par:190-208 does.
 */
typedef struct cubep3m_b200_config {
  int32_t nodes_dim;        /* parameters: nodes_dim                                   */
  int32_t tiles_node_dim;   /* parameters: tiles_node_dim                              */
  int32_t nf_tile;          /* parameters: nf_tile (incl. 2*nf_buf)                    */
  int32_t nf_buf;           /* parameters: nf_buf = nf_cutoff + 8 = 24                 */
  int32_t nf_cutoff;        /* parameters: nf_cutoff = 16                              */
  int32_t mesh_scale;       /* cubepm.par:157  (must be 4)                             */
  int32_t pp_range;         /* cubepm.par:92                                           */
  int32_t max_np;           /* cubepm.par:170-172; 0 => computed from density_buffer   */
  int32_t max_buf;          /* cubepm.par:175 (floats); 0 => 2.2*max_np                */
  int32_t max_llf;          /* cubepm.par:183; the library also stops (EMAXLLF) at 65534 particles in ONE fine cell (16-bit cell histogram) */
  float   density_buffer;   /* parameters: density_buffer                              */
  float   rsoft;            /* cubepm.par:76                                           */
  float   pp_bias;          /* cubepm.par:80                                           */
  float   dt_pp_scale;      /* cubepm.par:78                                           */
  float   G;                /* cubepm.par:149 = 1/6/pi with pi = 3.141592654           */
  float   eps;              /* cubepm.par:150                                          */
  /* cpp flags / logical parameters (0 or 1) */
  int32_t ngp;              /* -DNGP       fine mesh nearest-grid-point (else CIC)     */
  int32_t ppint;            /* -DPPINT     intra-fine-cell pairs                       */
  int32_t pp_ext;           /* -DPP_EXT    extended pp over pp_range fine cells        */
  int32_t coarse_ngp;       /* -DCOARSE_NGP                                            */
  int32_t pid;              /* -DPID_FLAG  carry int64 particle ids                    */
  int32_t lrckcorr;         /* -DLRCKCORR  (only used when the library builds kern_c)  */
  int32_t move_grid_back;   /* -DMOVE_GRID_BACK                                        */
  int32_t ngp_fmesh_force;  /* cubepm.par:84                                           */
  int32_t pp_force_flag;    /* cubepm.par:85                                           */
  int32_t pp_ext_force_flag;/* cubepm.par:86                                           */
  int32_t coarse_vel_update;/* cubepm.par:87                                           */
  /* topology (mpi_initialization.f90:42-76): rank = x + D*y + D*D*z, cart_coords(1)=z   */
  int32_t rank;
  int32_t local_gpu;        /* CUDA device ordinal for this rank                       */
  /* Rank grid (Dx,Dy,Dz) of cubic nodes; all zero => nodes_dim^3 (the reference's only option, mpi_initialization.f90:55-64).
   * 2- and 4-GPU runs use (2,1,1) / (2,2,1): the tiles of one (non-cubic) box split block-wise over the GPUs.
   * rank = x + Dx*(y + Dy*z), as the reference's row-major mpi_cart_create with cart_coords(1) = z.                       */
  int32_t nodes_dim_xyz[3];
} cubep3m_b200_config;

/* Outputs that particle_mesh leaves in COMMON (cubep3m.fh:19-21) or prints with -DDIAG. */
typedef struct cubep3m_b200_step_out {
  int32_t np_local;          /* after delete_particles                                  */
  int32_t np_with_ghosts;    /* np_local after particle_pass (particle_pass.f90:757)    */
  int32_t np_deleted_ll;     /* 'PARTICLE DELETED' count of link_list.f90:29-47         */
  int32_t np_buf_max;        /* largest single-direction pass buffer (particle_pass.f90:103) */
  float   dt_f_acc;          /* particle_mesh_threaded.f90:652                          */
  float   dt_pp_acc;         /* :668                                                    */
  float   dt_pp_ext_acc;     /* :692                                                    */
  float   dt_c_acc;          /* coarse_max_dt.f90:36                                    */
  float   f_force_max;       /* sqrt(max |force_f|^2) before the limiter formula        */
  float   pp_force_max;
  float   pp_ext_force_max;
  float   c_force_max;
  double  sum_rho_f;         /* 'sum of rho_f'  particle_mesh_threaded.f90:703-705      */
  double  sum_rho_c;         /* 'sum of rho_c'  coarse_mesh.f90:31-43                   */
  int64_t np_total;          /* 'total number of particles' delete_particles.f90:61-64  */
  float   stage_ms[16];      /* device time per stage, indices CUBEP3M_B200_ST_*        */
} cubep3m_b200_step_out;

enum {
  CUBEP3M_B200_ST_DRIFT = 0,   /* 'pos updt' */
  CUBEP3M_B200_ST_LINK,        /* 'linklist' : keys + counting sort                    */
  CUBEP3M_B200_ST_PASS,        /* 'par pass' */
  CUBEP3M_B200_ST_FINE_DEPOSIT,
  CUBEP3M_B200_ST_FINE_FFT,    /* forward + kernel multiply + 3 inverse                */
  CUBEP3M_B200_ST_FINE_KICK,
  CUBEP3M_B200_ST_PP,
  CUBEP3M_B200_ST_PP_EXT,
  CUBEP3M_B200_ST_COARSE_MASS, /* 'cm  mass' */
  CUBEP3M_B200_ST_COARSE_FORCE,/* 'cm force' + 'cf  buff' + 'c max dt'                 */
  CUBEP3M_B200_ST_COARSE_VEL,  /* 'cm   vel' */
  CUBEP3M_B200_ST_DELETE,      /* 'del part' */
  CUBEP3M_B200_ST_TOTAL,
  CUBEP3M_B200_ST_COUNT
};

typedef struct cubep3m_b200_ctx cubep3m_b200_ctx; /* opaque */

/* Library / build info: "sm_100a" etc. */
const char* cubep3m_b200_version(void);
const char* cubep3m_b200_strerror(int status);

/* Fills every field with the values of parameters.example + cubepm.par + Make_PP_THREADS:10. */
void cubep3m_b200_default_config(cubep3m_b200_config* cfg);

/*
 * Allocates all device state once (the reference allocates nothing on the hot path: cubep3m.fh COMMON).
 * fine_table  : 16*16*16*3 floats [k][j][i][c]  = kernels/wfxyzf.3.ascii   (kernel_initialization.f90:25-36)
 * coarse_table:  4* 4* 4*3 floats [k][j][i][c]  = kernels/wfxyzc.2.ascii   (kernel_initialization.f90:344-359)
 * kern_f / kern_c: optional host arrays in the reference's layout (cubep3m.fh:35,56), i.e. what the
 * Fortran driver already computed in fine_kernel / coarse_kernel; when NULL the library builds them on
 * the device from the tables (replaces kernel_initialization.f90:2-267, 272-732).
 * nccl_unique_id: 128-byte ncclUniqueId shared by all ranks, or NULL for a single-process run.
 */
int cubep3m_b200_init(const cubep3m_b200_config* cfg,
                      const float* fine_table, const float* coarse_table,
                      const float* kern_f, const float* kern_c,
                      const void* nccl_unique_id, int world_size,
                      cubep3m_b200_ctx** ctx);
int cubep3m_b200_get_unique_id(void* id128);      /* ncclGetUniqueId wrapper for the shim            */
int cubep3m_b200_finalize(cubep3m_b200_ctx* ctx);

/* xv(6,np_local) AoS float32 as in cubep3m.fh:75; pid may be NULL. Strict mode calls these around every step. */
int cubep3m_b200_upload_particles(cubep3m_b200_ctx* ctx, const float* xv, const int64_t* pid, int32_t np_local);
int cubep3m_b200_download_particles(cubep3m_b200_ctx* ctx, float* xv, int64_t* pid, int32_t* np_local);

/*
 * Replaces `call particle_mesh` (particle_mesh_threaded.f90:2; callers cubepm.f90:143, report_force.f90:41,100).
 * offset[3] is the shake offset the driver drew in update_position.f90:56-58 (all zeros without -DDISP_MESH).
 */
int cubep3m_b200_particle_mesh(cubep3m_b200_ctx* ctx, float dt, float dt_old, float a_mid, float mass_p,
                               const float offset[3], cubep3m_b200_step_out* out);

/* Sub-steps the driver also calls on their own (cubepm.f90:98,176,179,193-194,228). */
int cubep3m_b200_update_position(cubep3m_b200_ctx* ctx, float dt, float dt_old, const float offset[3]);
int cubep3m_b200_link_list(cubep3m_b200_ctx* ctx, int32_t* np_deleted);
int cubep3m_b200_particle_pass(cubep3m_b200_ctx* ctx, int32_t* np_with_ghosts);
int cubep3m_b200_delete_particles(cubep3m_b200_ctx* ctx, int32_t* np_local);
int cubep3m_b200_move_grid_back(cubep3m_b200_ctx* ctx, const float shake_offset[3]);

/* Parity / debug getters (host pointers). */
/* per coarse cell of the hoc range [hoc_nc_l,hoc_nc_h]^3 (x fastest): number of chained particles */
int cubep3m_b200_debug_cell_counts(cubep3m_b200_ctx* ctx, int32_t* counts);
/* particles deposited into each tile's padded fine mesh (cic_l..cic_h of particle_mesh_threaded.f90:120-121) */
int cubep3m_b200_debug_tile_counts(cubep3m_b200_ctx* ctx, int32_t* counts);
/* all particles incl. ghosts, in cell-sorted order, after link_list+particle_pass */
int cubep3m_b200_debug_sorted_particles(cubep3m_b200_ctx* ctx, float* xv, int32_t* np);
/* kern_f (3,n/2+1,n,n) and kern_c (3,nc_dim/2+1,nc_dim,nc_slab) as used by the step */
int cubep3m_b200_debug_kern_f(cubep3m_b200_ctx* ctx, float* kern_f);
int cubep3m_b200_debug_kern_c(cubep3m_b200_ctx* ctx, float* kern_c);
/* rho_c (nc_node^3) and force_c (3,0:nc_node+1,...) of the last step */
int cubep3m_b200_debug_rho_c(cubep3m_b200_ctx* ctx, float* rho_c);
int cubep3m_b200_debug_force_c(cubep3m_b200_ctx* ctx, float* force_c);
/* one tile's rho_f after deposit (n+2,n,n) and force_f (3,m+3,m+3,m+3) as the reference lays them out */
int cubep3m_b200_debug_fine_tile(cubep3m_b200_ctx* ctx, int32_t tile, float mass_p, float* rho_f, float* force_f);
/* in-place 3-D r2c / c2r of a (n+2,n,n) padded array with the library's own FFT (parity vs. the FFTW call sites) */
int cubep3m_b200_debug_fft3d(cubep3m_b200_ctx* ctx, int32_t n, float* data, int32_t inverse);
/* PP_EXT of the last particle_mesh call (particle_mesh_threaded.f90:378-624): target blocks launched by the tiled shared-memory kernel and
 * how many of them exceeded the shared-memory source capacity and were walked through the global cell table instead */
int cubep3m_b200_debug_ppext_blocks(cubep3m_b200_ctx* ctx, int32_t* blocks, int32_t* fallback);
/* ordered particle-pair interactions evaluated by PPINT (:324-361) and PP_EXT (:496-590) in the last particle_mesh call: every unordered pair
 * the reference visits once is evaluated twice here (once per partner), so these are 2 x the reference's pair counts */
int cubep3m_b200_debug_pair_counts(cubep3m_b200_ctx* ctx, int64_t* ppint, int64_t* ppext);
/* number of kernels this context launched so far (bench.py's gpu_launches) */
int64_t cubep3m_b200_launch_count(cubep3m_b200_ctx* ctx);
/* Per-kernel-class device time of the last particle_mesh call: when profiling is on every launch is bracketed by
 * CUDA events on the launching stream (the stand-in for the reference's -DMPI_TIME stopwatches, timers.f90:68-77). */
int cubep3m_b200_set_profiling(cubep3m_b200_ctx* ctx, int on);
/* Number of fine tiles kept in flight on separate CUDA streams (1..init-time maximum, default 2). Per-kernel timings are only
 * unambiguous with 1 (no overlap); the reference analogue is the number of OpenMP threads working on tiles (cores, parameters:20). */
int cubep3m_b200_set_tile_streams(cubep3m_b200_ctx* ctx, int n);
int cubep3m_b200_num_kernel_classes(void);
const char* cubep3m_b200_kernel_class_name(int k);
int cubep3m_b200_get_kernel_times(cubep3m_b200_ctx* ctx, float* ms, int64_t* launches);

/*
 * cic_power on the device (utils/cic_power/cic_power.f90:840-954 driver, :1496-1539 CIC deposit, :1583-1615 mode weights and shells,
 * :1649-1660 output columns): power spectrum of the resident physical particles on the global nf_physical_dim^3 mesh. One rank with a mesh the
 * library transforms directly (<= 560): everything on the one GPU. Several ranks (cubic rank grid; collective call, every rank gets the result) or
 * meshes of 1024 / 2048 cells: z-slabs / y-pencils as the reference's cube -> slab -> distributed r2c (:840-954), the deposit added straight into the
 * owners' slabs over NVLink, the transpose stored into the peers' pencils, shell sums all-reduced; 1024 and 2048 go through four-step passes.
 * shake_offset is subtracted first, as checkpoint.f90:92 does before the particles reach cic_power. nshells must be nf_physical_dim/2;
 * k [h/Mpc], Delta^2(k) and its standard error (may be NULL) are written for shells 1..nshells. ngp_binning = 1 is the build
 * COMPILE_cic_power.csh:20 uses (w1 = 1, w2 = 0), 0 the CIC shell weights.
 */
int cubep3m_b200_cic_power(cubep3m_b200_ctx* ctx, const float shake_offset[3], double box, int32_t ngp_binning,
                           double* k, double* delta2, double* sigma, int32_t nshells);

/*
 * dist_init on the device (utils/dist_init/dist_init_dm.f90:448-1046, single rank): Zel'dovich initial conditions generated straight into the
 * resident particle array, in dist_init's file order. nc = mesh cells per dimension of the generated box (nc/2 particles per dimension), reps >= 1
 * replicates the periodic box reps^3 times (nc * reps must equal the node's nf_physical_dim; nc must be a transform length the library supports).
 * (k_table, delta2_table)[n_table]: the dimensionless power spectrum Delta^2(k) at the initial scale factor, k in h/Mpc ascending — what
 * dist_init builds from batch/camb_WMAP5_transfer_z0.dat, sigma_8 and Dgrow (:448-531); interpolated log-log as `power` does (:1270-1299).
 * vfactor = a^2 H(a) (:1324-1337). noise: optional host white-noise field nc^3 (z slowest); NULL draws it on the device (Philox4x32-10, Box-Muller
 * as :617-629). The short-range kernel correction (:850-903) is not applied.
 */
int cubep3m_b200_dist_init(cubep3m_b200_ctx* ctx, int32_t nc, int32_t reps, float box, float vfactor, uint64_t seed, const float* k_table,
                           const float* delta2_table, int32_t n_table, const float* noise, int32_t* np_local);

/*
 * Checkpoint files in the reference's -DBINARY stream format (checkpoint.f90:72-95 writer, particle_initialization.f90:88-189 reader):
 *   <z>xv<rank>.dat  = header, then np_local records of 6 float32 (x - shake_offset, v)      (checkpoint.f90:92)
 *   <z>PID<rank>.dat = the same header, then np_local int64 ids                                (-DPID_FLAG)
 * header = np_local, a, t, tau, nts, dt_f_acc, [dt_pp_acc only with -DPPINT], dt_c_acc, cur_checkpoint, cur_projection, cur_halofind, mass_p
 * (4-byte fields, 48 bytes with PPINT, 44 without: the library follows cfg.ppint). The writer streams the RESIDENT particles from the device in
 * 32 MB blocks (the reference's blocksize) with the shake offset subtracted on the fly, so resident mode never round-trips xv through the driver;
 * the reader fills the device copy (np_local comes from the file) and returns the header. File names are the caller's (the shim builds them as
 * checkpoint.f90:31-46 does). path_pid may be NULL.
 */
typedef struct cubep3m_b200_checkpoint_header {
  int32_t np_local;
  float   a, t, tau;
  int32_t nts;
  float   dt_f_acc, dt_pp_acc, dt_c_acc;
  int32_t cur_checkpoint, cur_projection, cur_halofind;
  float   mass_p;
} cubep3m_b200_checkpoint_header;
int cubep3m_b200_write_checkpoint(cubep3m_b200_ctx* ctx, const char* path_xv, const char* path_pid, const cubep3m_b200_checkpoint_header* hdr,
                                  const float shake_offset[3]);
int cubep3m_b200_read_checkpoint(cubep3m_b200_ctx* ctx, const char* path_xv, const char* path_pid, cubep3m_b200_checkpoint_header* hdr);

/*
 * Driver twin (host C++): restatement of timestep / expansion (timestep.f90:2-293) so the harness can
 * run multi-step parity without the Fortran driver. Not needed when the Fortran driver is present.
 */
typedef struct cubep3m_b200_clock {
  float a, a_mid, t, tau, dt, dt_old, da;
  float dt_f_acc, dt_pp_acc, dt_pp_ext_acc, dt_c_acc;
  float omega_m, omega_l, wde;
  float a_target;            /* next checkpoint scale factor (a_checkpoint(cur_checkpoint)) */
  int32_t nts, ppint, pp_ext, cosmo, checkpoint_step;
} cubep3m_b200_clock;
void cubep3m_b200_clock_init(cubep3m_b200_clock* c, float z_i, float omega_m, float omega_l);
void cubep3m_b200_expansion(float a0, float dt0, float omega_m, float omega_l, float wde, float* da1, float* da2);
void cubep3m_b200_timestep(cubep3m_b200_clock* c);
/* The same timestep evaluated ON THE DEVICE: the clock is copied to device memory, one thread runs the identical
 * real(8) expansion / limiter logic (timestep.f90:54-293), the result is read back. Same results as cubep3m_b200_timestep to the last
 * bit of the float state except where the device's pow()/sqrt() differ from the host libm by an ulp of the real(8) intermediate. */
int cubep3m_b200_timestep_device(cubep3m_b200_ctx* ctx, cubep3m_b200_clock* c);

/*
 * Halo finder, density + maxima pass: halofind.f90:564-672 (find_halos up to the peak sort), called per tile from halofind.f90:48-54 on a
 * halofind step, i.e. after link_list and particle_pass (cubepm.f90:193-198; ENOTREADY otherwise). Per tile the fine density is deposited
 * (ngph != 0: fine_ngp_mass as with -DNGPH, else fine_cic_mass, :597-616), every physical cell that is the maximum of its 3^3 neighbourhood
 * and exceeds den_peak_cutoff (cubepm.par:124) becomes a peak (:620-632), its position refined by per-axis parabolic interpolation when
 * para_inter_hc (cubepm.par:133, :634-655, para_inter :770-778). Output: i, j, k = ipeak (1-based tile-local cell), tile = 0-based tile index
 * (x fastest), den = den_peak, x/y/z = peak_pos + offset (node-local fine-cell units, as halo_pos :723). Order: tile by tile, ascending density
 * inside a tile (what indexedsort leaves, :676-679). cftmass[2] = sums of rho_f and rho_f**2 over the physical cells (:621-622).
 * More than max_peaks maxima ('too many halos', :626-629) returns ECAPACITY with *n_peaks = the number found.
 * The spherical-overdensity mass growth that follows in the reference (:683-745) consumes these peaks sequentially and stays with the driver.
 */
typedef struct cubep3m_b200_peak { int32_t i, j, k, tile; float den, x, y, z; } cubep3m_b200_peak;
int cubep3m_b200_halofind_peaks(cubep3m_b200_ctx* ctx, float mass_p, float den_peak_cutoff, int32_t para_inter_hc, int32_t ngph,
                                cubep3m_b200_peak* peaks, int32_t max_peaks, int32_t* n_peaks, double* cftmass);

#ifdef __cplusplus
}
#endif
#endif
