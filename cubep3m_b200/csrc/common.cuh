// Shared declarations for the single-TU CUDA library (lib.cu). sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../../include/cubep3m_b200.h"

#ifdef CUBEP3M_WITH_NCCL
#include <nccl.h>
#endif

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      fprintf(stderr, "cubep3m_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e__), __FILE__, __LINE__, \
              cudaGetErrorString(e__));                                                            \
      return CUBEP3M_B200_ECUDA;                                                                   \
    }                                                                                              \
  } while (0)

// kernel classes for the per-class device-time breakdown (bench.py's roofline section)
enum KernelClass {
  KC_DRIFT = 0, KC_PASS_PACK, KC_PASS_UNPACK, KC_KEY_HIST, KC_SCAN, KC_SCATTER, KC_DENSITY,
  KC_FFT_X_R2C, KC_FFT_FWD_STRIDED, KC_FFT_INV_Z_MUL, KC_FFT_INV_Y, KC_FFT_X_C2R, KC_FORCE_MAX, KC_NGP_KICK,
  KC_PPINT, KC_PPEXT, KC_PPEXT_DENSE, KC_PPEXT_MARGIN, KC_CIC_MASS, KC_COARSE_FFT, KC_COARSE_MISC, KC_COARSE_XCHG, KC_CIC_KICK, KC_COMPACT, KC_MISC, KC_COUNT
};
static const char* const kKernelClassNames[KC_COUNT] = {
  "drift", "pass_pack", "pass_unpack", "key_hist", "scan", "scatter", "ngp_density",
  "fft_x_r2c", "fft_fwd_strided", "fft_inv_z_mul", "fft_inv_y", "fft_x_c2r", "force_max", "ngp_kick",
  "ppint", "ppext", "ppext_dense", "ppext_margin", "cic_mass", "coarse_fft", "coarse_misc", "coarse_xchg", "cic_kick", "compact", "misc"};

constexpr int PROF_MAX = 8192;   // profiled launches per step

// every kernel launch goes through this: counts launches and, when profiling is on, brackets the launch with
// CUDA events on the launching stream so that per-class device time can be summed after the step.
#define LAUNCH(ctx, kclass, kernel, grid, block, smem, ...)                                    \
  do {                                                                                         \
    const bool prof__ = (ctx)->profiling && (ctx)->prof_n < PROF_MAX;                          \
    if (prof__) cudaEventRecord((ctx)->prof_ev[2 * (ctx)->prof_n], (ctx)->stream);             \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                           \
    if (prof__) {                                                                              \
      cudaEventRecord((ctx)->prof_ev[2 * (ctx)->prof_n + 1], (ctx)->stream);                   \
      (ctx)->prof_class[(ctx)->prof_n++] = (kclass);                                           \
    }                                                                                          \
    (ctx)->launches++;                                                                         \
    (ctx)->class_launches[(kclass)]++;                                                         \
  } while (0)

constexpr int NUM_SMS = 148;   // B200

// derived sizes, cubepm.par:190-208
struct Dims {
  int D, T, n, b, s, m, mT, nc_tile, nc_node, nc_dim, nc_slab, nc_buf, hoc_l, hoc_h, H, nodes, tiles_node;
  int max_np, max_buf;
  int Dg[3];       // rank grid (Dx,Dy,Dz) of cubic nodes
  int coord[3];    // this rank's (x,y,z) in the grid; rank = x + Dx*(y + Dy*z)
  int nbr[6];      // neighbour ranks: -x,+x,-y,+y,-z,+z (periodic)
  int Nc[3];       // global coarse mesh (Nx,Ny,Nz) = nc_node * Dg
  int world;       // Dx*Dy*Dz
  int hc;          // n/2+1
  int fdim;        // m+3: force_f spans nf_buf-1 .. nf_tile-nf_buf+1 (cubep3m.fh:36-37)
  long long NF;    // fine cells of the hoc range: H^3 * 64
};

// every rank's exchange allocation as mapped into THIS process (coarse_slab.cuh); own pointer for the own rank
struct PeerTable { float* base[8]; };

// device counters written by kernels, mirrored to the host once per step
struct DevCounters {
  int np_deleted;        // out-of-range particles dropped by link_list
  int n_send[2];         // pack counts of the current axis: [0] = "+" direction, [1] = "-"
  int n_multi;           // fine cells with >= 2 particles in the physical region (PPINT work list)
  int n_occ;             // occupied physical fine cells (PP_EXT work list)
  int overflow;          // bit 0: pass buffer, bit 1: max_np, bit 2: max_llf
  unsigned int f_force_max2_bits;   // max |force_f|^2 as ordered uint
  unsigned int pp_force_max_bits;
  unsigned int pp_ext_force_max_bits;
  unsigned int c_force_max_bits;
  int np_phys;           // after delete_particles
  int n_cand;            // particles within half an ulp below a fine-cell boundary (see fine::ngp_fixup_kernel)
  int n_blist;           // particles near a y or z face, listed by the first particle_pass kernel
  int n_ppext_items;     // (fine cell, 32-target chunk) items of the dense PP_EXT blocks (pp::ppext_items_kernel)
  int ppext_ticket;      // work ticket of pp::ppext_cell_kernel (must follow n_ppext_items: both are cleared together)
  int n_ppint_items[2];  // (fine cell, chunk) items of PPINT (pp::ppint_items_kernel): heavy (front of the list), light (back); and the heavy items' work ticket (cleared together)
  int ppint_ticket;
  int n_margin_roles;    // (particle, tile) margin roles listed for the PP_EXT limiter (pp::ppext_margin_list_kernel)
  int xchg_timeout;      // a coarse-mesh exchange wait (coarse_slab.cuh) gave up on a peer
  int n_ppext_fallback;  // PP_EXT blocks whose source region exceeded the shared-memory capacity (walked directly instead)
  double sum_rho_f;
  double sum_rho_c;
  unsigned long long pairs_ppint;   // ordered pair interactions evaluated by PPINT / PP_EXT in this step (margin-only limiter sums not counted)
  unsigned long long pairs_ppext;
};

struct cubep3m_b200_ctx {
  cubep3m_b200_config cfg;
  Dims d;
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  long long class_launches[KC_COUNT] = {0};
  bool profiling = false;
  int prof_n = 0;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_class;
  float class_ms[KC_COUNT] = {0};        // of the last profiled step
  long long class_n[KC_COUNT] = {0};
  int fft_class_base = 0;                // KC_COARSE_FFT while the coarse solve runs, else 0
  cubep3m_b200_clock* dclock = nullptr;  // device copy of the driver clock (cubep3m_b200_timestep_device)
  int world = 1;
  // particles (AoS 24-byte records as the reference's xv(6,:)), double buffered
  float* xv[2] = {nullptr, nullptr};
  int64_t* pid[2] = {nullptr, nullptr};
  int cur = 0;
  int np_local = 0;      // physical particles (valid after upload / delete_particles)
  int np_all = 0;        // incl. ghosts (valid after particle_pass)
  bool sorted = false;   // xv[cur] is cell-sorted and fstart is valid
  bool passed = false;
  unsigned int* key = nullptr;
  unsigned short* rank = nullptr;   // rank of every particle inside its fine cell (the count its histogram atomic returned): scatter slot = fstart[key] + rank
  int* blist = nullptr;       // particles near a y / z face, listed by the first pack kernel of particle_pass
  bool keys_fused = false;    // the pass kernels of this step already produced key[] and the histogram (do_sort skips key_hist_kernel)
  int* fstart = nullptr; // exclusive scan of fine-cell counts, NF+1 entries
  unsigned int* fcur = nullptr;   // fine-cell histogram, two 16-bit counters per word (NF/2 words); counts itself back to zero in the scatter
  int* blocksum = nullptr;
  unsigned long long* scan_status = nullptr;   // tile status words of the single-pass scan (+ ticket counter)
  bool scan_onepass = false;
  int nblocksum = 0;
  int* multi_list = nullptr;  // keys of physical fine cells with >= 2 particles
  int* occ_list = nullptr;    // keys of occupied physical fine cells
  int list_cap = 0;
  float* sendbuf[2] = {nullptr, nullptr};
  float* recvbuf[2] = {nullptr, nullptr};
  float* recvbuf_own[2] = {nullptr, nullptr};   // only allocated for multi-rank runs
  int64_t* recvpid_own[2] = {nullptr, nullptr};
  int64_t* sendpid[2] = {nullptr, nullptr};
  int64_t* recvpid[2] = {nullptr, nullptr};
  int* rowoff = nullptr;      // compaction offsets per physical (cy,cz) row
  float* cand = nullptr;      // positions of the boundary-candidate particles (3 floats each)
  int2* deltas = nullptr;     // per-tile (from-cell, to-cell) moves for the fused density pass, tiles * DELTA_CAP
  int* ndelta = nullptr;      // per-tile move counts
  int* tile_counts = nullptr; // per-tile deposited-particle counts (parity getter)
  int cand_cap = 0;
  // fine mesh
  float* kern_f = nullptr;    // [comp][z][y][kf_pitch]: the reference's kern_f(3,hc,n,n) (cubep3m.fh:35) de-interleaved, rows padded to 16 floats
  int kf_pitch = 0;           // row pitch of kern_f (floats), multiple of 16
  long long kf_stride = 0;    // component stride of kern_f (floats)
  static constexpr int MAX_TILE_STREAMS = 4;
  int tile_streams_max = 1;
  int tile_streams = 1;       // S > 1: consecutive tiles rotate over S streams / buffer sets (hides launch bubbles, tails, latency)
  cudaStream_t stream_coarse = nullptr;   // coarse-mesh solve runs concurrently with the fine-tile loop
  int ppext_mode = 1;          // 1: tiled shared-memory kernel (pp::ppext_tiled_kernel), 0: direct one-thread-per-target kernel (CUBEP3M_B200_PPEXT=direct)
  int* ppext_ovf = nullptr;    // ids of the PP_EXT target blocks that exceeded the tiled kernel's shared-memory capacity
  bool ppext_margin_max = true; // also evaluate the margin particles' partial sums for pp_ext_force_max (particle_mesh_threaded.f90:617); CUBEP3M_B200_PPEXT_MARGIN=0 skips it
  int2* ppext_items = nullptr; int ppext_item_cap = 0; bool ppext_cell_mode = true, ppext_dense_tma = false;   // dense-block PP_EXT work items
  int2* ppint_items = nullptr; int ppint_item_cap = 0;
  int2* margin_roles = nullptr; int margin_cap = 0;
  bool want_roles = false, roles_listed = false;      // particle_mesh asks the scatter to list the margin roles; the limiter kernel then skips its own listing   // (particle index, tile) list of the PP_EXT margin roles
  int ppext_blocks = 0, ppext_fallback = 0;
  long long pairs_ppint = 0, pairs_ppext = 0;   // of the last step   // of the last step (debug getter)
  int hist_mode = 0;                  // CUBEP3M_B200_HISTZERO: 0 = scatter by stored rank + memset after the scan (default), 1 = "scatter": + zero store per particle, 2 = "countdown": round 1's second atomic
  bool defer_hist_zero = false;       // set by particle_mesh around its sort when PP_EXT is on: the histogram is cleared under the PP_EXT kernels instead of right after the scan
  bool hist_clean = false;     // fcur (the fine-cell histogram) is all zeros
  cudaStream_t stream_main = nullptr, stream_aux[MAX_TILE_STREAMS] = {nullptr, nullptr, nullptr, nullptr};   // [0] unused
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_TILE_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
  float* tile_rho_s[MAX_TILE_STREAMS] = {nullptr, nullptr, nullptr, nullptr};   // buffer sets 1..S-1 (set 0 = tile_rho / tile_g / force_f)
  float* tile_g_s[MAX_TILE_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
  float* force_f_s[MAX_TILE_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
  float* tile_rho = nullptr;  // (n+2,n,n) real / (hc,n,n) complex, in place
  float* tile_g = nullptr;    // work array for one force component
  float* force_f[3] = {nullptr, nullptr, nullptr};  // (fdim^3) each, SoA
  float2* tw_f = nullptr;     // twiddles exp(-2 pi i t/n)
  // coarse mesh
  float* kern_c = nullptr;    // [comp][z][y][kx] over the global coarse mesh (reference: kern_c(3,hc,nc_dim,nc_slab) per rank)
  float* rho_c = nullptr;     // nc_node^3
  float* slab = nullptr;      // (Nx+2, Ny, Nz): the WHOLE global coarse mesh, replicated on every rank
  float* slab_g = nullptr;
  float* creal = nullptr;     // (Nx, Ny, Nz) real-space force component
  float* gather = nullptr;    // all ranks' rho_c cubes (ncclAllGather target), world * nc_node^3
  float* force_c = nullptr;   // (3, nc_node+2, nc_node+2, nc_node+2) components innermost as cubep3m.fh:59
  float2* tw_c[3] = {nullptr, nullptr, nullptr};   // twiddles for Nx, Ny, Nz
  // slab-decomposed coarse solve over peer memory (coarse_slab.cuh, lib.cu: cs_init / do_coarse_force_slab)
  int coarse_mode = 0;         // 0: the whole coarse mesh on this GPU (one rank; or the all-gather + replicated-solve fallback), 1: slab-decomposed
  int cs_zs = 0, cs_ys = 0;    // z planes per slab (the reference's nc_slab), y rows per pencil block
  float* cs_xchg = nullptr;    // ONE allocation the peers store into: [slab | pencils T | back[3] | force_c | mailbox], offsets in floats
  size_t cs_off_slab = 0, cs_off_T = 0, cs_off_back[3] = {0, 0, 0}, cs_off_force = 0, cs_off_mail = 0, cs_floats = 0;
  PeerTable cs_peers = {{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
  float* cs_G = nullptr;       // (Nz, ys, hc) complex: inverse-z result of one component before it is sent back
  float* cs_real3 = nullptr;   // [comp][zs][Ny][Nx] real-space force of this rank's slab
  float* cs_kern_rows = nullptr;   // [comp][Nz][ys][hc]: this rank's rows of kern_c
  unsigned int cs_epoch = 0;
  std::vector<void*> cs_ipc_opened;
  float* redbuf = nullptr;    // small device scratch for cross-rank reductions
  int* cntbuf = nullptr;      // received pass counts
  // particle_pass over NVLink peer memory (multi-rank, inside particle_mesh): the pack kernel stores straight into the neighbour's receive
  // buffer; counts and completion flags travel through small mailboxes in peer memory (lib.cu: p2p_init / do_pass)
  bool p2p = false;
  int p2p_cap = 0;                 // particles per (axis, direction) region of a receive buffer
  unsigned int p2p_epoch = 0;      // one per particle_mesh call; flags carry it, so they never need resetting
  int* mailbox = nullptr;          // own: [3 axes][2 directions][count, flag]
  float* peer_recv[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // where MY (axis, direction)-going particles land
  int* peer_box[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  std::vector<void*> ipc_opened;
  int* hbox = nullptr;             // pinned: received counts + time-out flag
  DevCounters* dcnt = nullptr;
  DevCounters* hcnt = nullptr; // pinned
  cudaEvent_t ev[CUBEP3M_B200_ST_COUNT + 4];
  bool ev_ok = false;
  std::vector<float> fine_table, coarse_table;
  int last_tile_counts_valid = 0;
#ifdef CUBEP3M_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
};

// ordered-uint encoding of non-negative floats for atomicMax
__device__ __forceinline__ void atomic_max_float_nonneg(unsigned int* addr, float v) { atomicMax(addr, __float_as_uint(v)); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
