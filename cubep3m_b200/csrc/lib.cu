// libcubep3m_b200.so — the C ABI of include/cubep3m_b200.h on top of the hand-written sm_100a kernels.
// Single translation unit: all kernels are in the .cuh parts included below.
// There is NO CPU fallback: every entry point that computes needs a CUDA device and fails with ECUDA otherwise.
#include "common.cuh"
#include "fft_smem.cuh"
#include "fft3d.cuh"
#include "particles.cuh"
#include "fine.cuh"
#include "pp.cuh"
#include "coarse.cuh"
#include "coarse_slab.cuh"
#include "power.cuh"
#include "distinit.cuh"
#include "bigfft.cuh"
#include "halo.cuh"

namespace {

int derive(const cubep3m_b200_config& c, Dims& d) {
  d.D = c.nodes_dim; d.T = c.tiles_node_dim; d.n = c.nf_tile; d.b = c.nf_buf; d.s = c.mesh_scale;
  if (d.D < 1 || d.T < 1 || d.s != 4 || d.b <= 0 || d.b % d.s != 0) return CUBEP3M_B200_EINVAL;
  d.m = d.n - 2 * d.b;                    // nf_physical_tile_dim  cubepm.par:197
  if (d.m <= 0 || d.m % d.s != 0) return CUBEP3M_B200_EINVAL;
  d.mT = d.m * d.T;                       // nf_physical_node_dim  cubepm.par:201
  d.nc_buf = d.b / d.s;                   // cubepm.par:190
  d.nc_tile = d.m / d.s;                  // cubepm.par:191
  d.nc_node = d.nc_tile * d.T;            // cubepm.par:192
  d.nc_dim = d.nc_node * d.D;             // cubepm.par:193
  const bool custom_grid = c.nodes_dim_xyz[0] > 0 && c.nodes_dim_xyz[1] > 0 && c.nodes_dim_xyz[2] > 0;
  for (int a = 0; a < 3; ++a) d.Dg[a] = custom_grid ? c.nodes_dim_xyz[a] : d.D;
  d.world = d.Dg[0] * d.Dg[1] * d.Dg[2];
  d.nodes = d.world;
  if (c.rank < 0 || c.rank >= d.world) return CUBEP3M_B200_EINVAL;
  d.coord[0] = c.rank % d.Dg[0]; d.coord[1] = (c.rank / d.Dg[0]) % d.Dg[1]; d.coord[2] = c.rank / (d.Dg[0] * d.Dg[1]);
  for (int a = 0; a < 3; ++a) {           // mpi_cart_shift of mpi_initialization.f90:73-76 (periodic)
    int cm[3] = {d.coord[0], d.coord[1], d.coord[2]}, cp[3] = {d.coord[0], d.coord[1], d.coord[2]};
    cm[a] = (cm[a] - 1 + d.Dg[a]) % d.Dg[a]; cp[a] = (cp[a] + 1) % d.Dg[a];
    d.nbr[2 * a] = cm[0] + d.Dg[0] * (cm[1] + d.Dg[1] * cm[2]);
    d.nbr[2 * a + 1] = cp[0] + d.Dg[0] * (cp[1] + d.Dg[1] * cp[2]);
    d.Nc[a] = d.nc_node * d.Dg[a];
  }
  // the reference's slab decomposition needs mod(nc_dim, nodes) == 0 (mpi_initialization.f90:25-29); the replicated coarse
  // solve used here does not, but nc_slab is kept for the kern_c getter when the grid is the reference's cubic one
  d.nc_slab = (!custom_grid && d.nc_dim % d.nodes == 0) ? d.nc_dim / d.nodes : 0;
  d.hoc_l = 1 - d.nc_buf; d.hoc_h = d.nc_node + d.nc_buf; d.H = d.hoc_h - d.hoc_l + 1;   // cubepm.par:204-205
  d.tiles_node = d.T * d.T * d.T;
  if (c.max_np > 0) d.max_np = c.max_np;
  else {                                  // cubepm.par:170-172
    const double half = (double)(d.mT / 2);
    const double buf = (8.0 * d.b * d.b * d.b + 6.0 * d.b * (double)d.mT * d.mT + 12.0 * (double)d.b * d.b * d.mT) / 8.0;
    d.max_np = (int)(c.density_buffer * (half * half * half + buf));
  }
  d.max_buf = c.max_buf > 0 ? c.max_buf : (int)(2.2 * (double)d.max_np);   // cubepm.par:175
  d.hc = d.n / 2 + 1;
  d.fdim = d.m + 3;
  d.NF = (long long)d.H * d.H * d.H * 64;
  if (d.NF >= (1LL << 31)) return CUBEP3M_B200_EINVAL;
  if (!fftk::supported(d.n)) return CUBEP3M_B200_EINVAL;
  for (int a = 0; a < 3; ++a) if (!fftk::supported(d.Nc[a])) return CUBEP3M_B200_EINVAL;
  if (c.pp_range < 0 || c.pp_range > 2) return CUBEP3M_B200_EINVAL;
  // LRCKCORR divides by Im(kernel) for |k| <= 8 (kernel_initialization.f90:573-581): Nyquist must lie beyond
  for (int a = 0; a < 3; ++a) if (c.lrckcorr && d.Nc[a] / 2 <= 8) return CUBEP3M_B200_EINVAL;
  return 0;
}

template <typename T> int dmalloc(T** p, size_t count) {
  CK(cudaMalloc((void**)p, count * sizeof(T)));
  return 0;
}

// kern_f(comp) = Im(spectrum), stored with row pitch kp >= hc (rows = (z, y))
__global__ void extract_imag_kernel(const float2* __restrict__ spec, long long n, int hc, int kp, float* __restrict__ kern) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / hc;
    kern[row * kp + (i - row * hc)] = spec[i].y;
  }
}

// host (3, n) interleaved -> device [comp][stride] planes, elements [off, off+count); rows of `width` elements are stored with row
// pitch `pitch` on the device (pitch == width: dense)
int upload_interleaved(float* dev, const float* host, size_t plane, size_t off, size_t count, size_t width = 0, size_t pitch = 0) {
  std::vector<float> tmp(count);
  for (int comp = 0; comp < 3; ++comp) {
    for (size_t i = 0; i < count; ++i) tmp[i] = host[3 * i + comp];
    if (width && pitch != width) CK(cudaMemcpy2D(dev + (size_t)comp * plane, pitch * sizeof(float), tmp.data(), width * sizeof(float), width * sizeof(float), count / width, cudaMemcpyHostToDevice));
    else CK(cudaMemcpy(dev + (size_t)comp * plane + off, tmp.data(), count * sizeof(float), cudaMemcpyHostToDevice));
  }
  return 0;
}
int download_interleaved(float* host, const float* dev, size_t plane, size_t off, size_t count, size_t width = 0, size_t pitch = 0) {
  std::vector<float> tmp(count);
  for (int comp = 0; comp < 3; ++comp) {
    if (width && pitch != width) CK(cudaMemcpy2D(tmp.data(), width * sizeof(float), dev + (size_t)comp * plane, pitch * sizeof(float), width * sizeof(float), count / width, cudaMemcpyDeviceToHost));
    else CK(cudaMemcpy(tmp.data(), dev + (size_t)comp * plane + off, count * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < count; ++i) host[3 * i + comp] = tmp[i];
  }
  return 0;
}

fftk::Mesh3 fine_mesh(const cubep3m_b200_ctx* ctx) { return fftk::Mesh3{ctx->d.n, ctx->d.n, ctx->d.n, ctx->tw_f, ctx->tw_f, ctx->tw_f}; }
fftk::Mesh3 coarse_mesh3(const cubep3m_b200_ctx* ctx) { return fftk::Mesh3{ctx->d.Nc[0], ctx->d.Nc[1], ctx->d.Nc[2], ctx->tw_c[0], ctx->tw_c[1], ctx->tw_c[2]}; }

int grid_for(long long n, int tpb, int cap = NUM_SMS * 16) {
  long long g = (n + tpb - 1) / tpb;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// ---- kernel_initialization.f90:2-267 fine_kernel, with the library's own FFT
int build_kern_f(cubep3m_b200_ctx* ctx) {
  const Dims& d = ctx->d;
  const cubep3m_b200_config& c = ctx->cfg;
  const int n = d.n, n2 = n + 2, nfc = c.nf_cutoff;
  if (nfc != 16 || n < 2 * nfc) return CUBEP3M_B200_EINVAL;
  std::vector<float> rho((size_t)n2 * n * n);
  auto R = [&](int i, int j, int k) -> float& { return rho[(size_t)(i - 1) + (size_t)n2 * ((j - 1) + (size_t)n * (k - 1))]; };
  for (int comp = 0; comp < 3; ++comp) {
    std::fill(rho.begin(), rho.end(), 0.f);
    for (int k = 1; k <= nfc; ++k)
      for (int j = 1; j <= nfc; ++j)
        for (int i = 1; i <= nfc; ++i) R(i, j, k) = ctx->fine_table[(((size_t)(k - 1) * 16 + (j - 1)) * 16 + (i - 1)) * 3 + comp];   // :25-36
    if (c.pp_ext && c.pp_ext_force_flag)                                                                                           // :38-54
      for (int k = 1; k <= c.pp_range + 1; ++k)
        for (int j = 1; j <= c.pp_range + 1; ++j)
          for (int i = 1; i <= c.pp_range + 1; ++i) R(i, j, k) = 0.f;
    const float sy = comp == 1 ? -1.f : 1.f, sx = comp == 0 ? -1.f : 1.f, sz = comp == 2 ? -1.f : 1.f;
    for (int j = 2; j <= nfc; ++j)                     // :71-73
      for (int k = 1; k <= nfc; ++k)
        for (int i = 1; i <= nfc; ++i) R(i, n - j + 2, k) = sy * R(i, j, k);
    for (int i = 2; i <= nfc; ++i)                     // :76-79
      for (int k = 1; k <= nfc; ++k)
        for (int j = 1; j <= n; ++j) R(n - i + 2, j, k) = sx * R(i, j, k);
    for (int k = 2; k <= nfc; ++k)                     // :82-85
      for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) R(i, j, n - k + 2) = sz * R(i, j, k);
    CK(cudaMemcpyAsync(ctx->tile_rho, rho.data(), rho.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (int st = fftk::forward3d(ctx, fine_mesh(ctx), ctx->tile_rho)) return st;   // :89
    const long long ns = (long long)d.hc * n * n;
    LAUNCH(ctx, KC_MISC, extract_imag_kernel, grid_for(ns, 256), 256, 0, reinterpret_cast<const float2*>(ctx->tile_rho), ns, d.hc, ctx->kf_pitch, ctx->kern_f + (size_t)comp * ctx->kf_stride);   // :93-99
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// ---- kernel_initialization.f90:272-732 coarse_kernel on the global mesh (Nx,Ny,Nz); every rank builds the same table.
// For the reference's cubic grids Nx = Ny = Nz = nc_dim and this is line-for-line the reference's construction.
int build_kern_c(cubep3m_b200_ctx* ctx) {
  const Dims& d = ctx->d;
  const cubep3m_b200_config& c = ctx->cfg;
  const int Nx = d.Nc[0], Ny = d.Nc[1], Nz = d.Nc[2], N2 = Nx + 2, hc = Nx / 2 + 1;
  const float pi = 3.141592654f;
  const size_t ncell = (size_t)Nx * Ny * Nz;
  std::vector<float> ck(3 * ncell), ckc;
  auto CKA = [&](std::vector<float>& a, int comp, int i, int j, int k) -> float& {
    return a[(size_t)comp + 3 * ((size_t)(i - 1) + (size_t)Nx * ((j - 1) + (size_t)Ny * (k - 1)))];
  };
  auto fill_plain = [&](std::vector<float>& a) {       // :302-336
    for (int k = 1; k <= Nz; ++k) {
      float z = (k < Nz / 2 + 2) ? (float)(k - 1) : (float)(k - 1 - Nz); z = d.s * z;
      for (int j = 1; j <= Ny; ++j) {
        float y = (j < Ny / 2 + 2) ? (float)(j - 1) : (float)(j - 1 - Ny); y = d.s * y;
        for (int i = 1; i <= Nx; ++i) {
          float x = (i < Nx / 2 + 2) ? (float)(i - 1) : (float)(i - 1 - Nx); x = d.s * x;
          const float r = sqrtf(x * x + y * y + z * z);
          if (r == 0.0f) { CKA(a, 0, i, j, k) = 0.f; CKA(a, 1, i, j, k) = 0.f; CKA(a, 2, i, j, k) = 0.f; }
          else { const float r3 = r * r * r; CKA(a, 0, i, j, k) = -x / r3; CKA(a, 1, i, j, k) = -y / r3; CKA(a, 2, i, j, k) = -z / r3; }
        }
      }
    }
  };
  fill_plain(ck);
  for (int oz = -3; oz <= 3; ++oz)                     // :344-457 near-field table in all octants
    for (int oy = -3; oy <= 3; ++oy)
      for (int ox = -3; ox <= 3; ++ox) {
        const int i = ox >= 0 ? ox + 1 : Nx + ox + 1, j = oy >= 0 ? oy + 1 : Ny + oy + 1, k = oz >= 0 ? oz + 1 : Nz + oz + 1;
        if (i < 1 || j < 1 || k < 1 || i > Nx || j > Ny || k > Nz) continue;
        const int o[3] = {ox, oy, oz};
        for (int comp = 0; comp < 3; ++comp) {
          const float v = ctx->coarse_table[(((size_t)abs(oz) * 4 + abs(oy)) * 4 + abs(ox)) * 3 + comp];
          CKA(ck, comp, i, j, k) = (o[comp] < 0) ? -v : v;
        }
      }
  std::vector<float> slab((size_t)N2 * Ny * Nz), tmp;
  const size_t ncs = (size_t)hc * Ny * Nz;
  std::vector<float> kc(3 * ncs);   // [comp][z][y][kx]
  auto transform = [&](std::vector<float>& a, int comp) -> int {
    for (int k = 1; k <= Nz; ++k)
      for (int j = 1; j <= Ny; ++j) {
        float* row = &slab[(size_t)N2 * ((j - 1) + (size_t)Ny * (k - 1))];
        for (int i = 1; i <= Nx; ++i) row[i - 1] = CKA(a, comp, i, j, k);
        row[Nx] = row[Nx + 1] = 0.f;
      }
    CK(cudaMemcpyAsync(ctx->slab, slab.data(), slab.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (int st = fftk::forward3d(ctx, coarse_mesh3(ctx), ctx->slab)) return st;
    CK(cudaMemcpyAsync(slab.data(), ctx->slab, slab.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  };
  if (c.lrckcorr) {
    ckc = ck;                                          // :469-475
    fill_plain(ck);                                    // :479-513
  }
  for (int comp = 0; comp < 3; ++comp) {
    if (c.lrckcorr) {
      if (int st = transform(ck, comp)) return st;     // :519-551
      tmp = slab;
      if (int st = transform(ckc, comp)) return st;
      for (int k = 1; k <= Nz; ++k) {                  // :558-591 / :602-635 / :646-679
        const int kz = (k < Nz / 2 + 2) ? k - 1 : k - 1 - Nz;
        for (int j = 1; j <= Ny; ++j) {
          const int ky = (j < Ny / 2 + 2) ? j - 1 : j - 1 - Ny;
          for (int i = 1; i <= Nx + 2; i += 2) {
            const int kx = (i - 1) / 2;
            const float kr = sqrtf((float)(kx * kx + ky * ky + kz * kz));
            if (kr <= 8.f) {
              const float ka = 2 * sinf(pi * kx / (float)Nx), kb = 2 * sinf(pi * ky / (float)Ny), kcc = 2 * sinf(pi * kz / (float)Nz);
              const int kd = comp == 0 ? kx : (comp == 1 ? ky : kz);
              const float kk = comp == 0 ? ka : (comp == 1 ? kb : kcc);
              if (kd != 0) {
                const size_t o = (size_t)i + (size_t)N2 * ((j - 1) + (size_t)Ny * (k - 1));
                const float wa = slab[o], wb = tmp[o];
                const float wc = 4.f * pi * kk / (ka * ka + kb * kb + kcc * kcc) / 16.f;
                slab[o] = wa * (wc / wb);
              }
            }
          }
        }
      }
    } else {
      if (int st = transform(ck, comp)) return st;     // :695-723
    }
    for (int k = 1; k <= Nz; ++k)                      // :593-599
      for (int j = 1; j <= Ny; ++j)
        for (int i = 1; i <= hc; ++i)
          kc[(size_t)comp * ncs + ((size_t)(i - 1) + (size_t)hc * ((j - 1) + (size_t)Ny * (k - 1)))] =
              slab[(size_t)(2 * i - 1) + (size_t)N2 * ((j - 1) + (size_t)Ny * (k - 1))];
  }
  CK(cudaMemcpy(ctx->kern_c, kc.data(), kc.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

#ifdef CUBEP3M_WITH_NCCL
#define NCK(call)                                                                                   \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess) {                                                                       \
      fprintf(stderr, "cubep3m_b200: NCCL error %s at %s:%d\n", ncclGetErrorString(r__), __FILE__, __LINE__); \
      return CUBEP3M_B200_ENCCL;                                                                    \
    }                                                                                               \
  } while (0)
#endif

// Peer-memory set-up for particle_pass: every rank exports its two receive buffers and its mailbox with cudaIpc, the handles are
// all-gathered over NCCL, and each rank maps the buffers of its (up to six) neighbours. Falls back to the NCCL send/recv path (p2p stays
// false) if any step is unavailable; CUBEP3M_B200_P2P=0 forces the fallback.
int p2p_init(cubep3m_b200_ctx* ctx) {
#ifdef CUBEP3M_WITH_NCCL
  const Dims& d = ctx->d;
  const char* e = getenv("CUBEP3M_B200_P2P");
  if (e && atoi(e) == 0) return 0;
  if (ctx->cfg.pid) return 0;
  CK(cudaMalloc((void**)&ctx->mailbox, 3 * 2 * 2 * sizeof(int)));
  CK(cudaMemset(ctx->mailbox, 0, 3 * 2 * 2 * sizeof(int)));
  CK(cudaMallocHost((void**)&ctx->hbox, 4 * sizeof(int)));
  struct Handles { cudaIpcMemHandle_t h[3]; int ok; int pad[3]; };
  Handles mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = cudaIpcGetMemHandle(&mine.h[0], ctx->recvbuf_own[0]) == cudaSuccess && cudaIpcGetMemHandle(&mine.h[1], ctx->recvbuf_own[1]) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.h[2], ctx->mailbox) == cudaSuccess;
  cudaGetLastError();
  Handles* dall = nullptr;
  CK(cudaMalloc((void**)&dall, sizeof(Handles) * d.world));
  CK(cudaMemcpy(dall + ctx->cfg.rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice));
  NCK(ncclAllGather(dall + ctx->cfg.rank, dall, sizeof(Handles), ncclChar, ctx->comm, ctx->stream));
  std::vector<Handles> all(d.world);
  CK(cudaMemcpyAsync(all.data(), dall, sizeof(Handles) * d.world, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dall);
  bool ok = true;
  for (int r = 0; r < d.world; ++r) ok = ok && all[r].ok;
  std::vector<float*> rb0(d.world, nullptr), rb1(d.world, nullptr);
  std::vector<int*> mb(d.world, nullptr);
  ctx->p2p_cap = d.max_buf / 6 / 3;
  for (int axis = 0; axis < 3 && ok; ++axis) {
    if (d.Dg[axis] == 1) continue;
    for (int dir = 0; dir < 2 && ok; ++dir) {
      const int peer = dir == 0 ? d.nbr[2 * axis + 1] : d.nbr[2 * axis];    // "+"-going particles land at the + neighbour
      if (peer == ctx->cfg.rank) { rb0[peer] = ctx->recvbuf_own[0]; rb1[peer] = ctx->recvbuf_own[1]; mb[peer] = ctx->mailbox; }
      if (!mb[peer]) {
        void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
        ok = cudaIpcOpenMemHandle(&p0, all[peer].h[0], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
             cudaIpcOpenMemHandle(&p1, all[peer].h[1], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
             cudaIpcOpenMemHandle(&p2, all[peer].h[2], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        if (p0) ctx->ipc_opened.push_back(p0);
        if (p1) ctx->ipc_opened.push_back(p1);
        if (p2) ctx->ipc_opened.push_back(p2);
        if (!ok) { cudaGetLastError(); break; }
        rb0[peer] = (float*)p0; rb1[peer] = (float*)p1; mb[peer] = (int*)p2;
      }
      // my "+"-going particles are what the peer receives "from its - neighbour": its buffer 0, mailbox entry (axis, 0); "-"-going: 1
      ctx->peer_recv[axis][dir] = (dir == 0 ? rb0[peer] : rb1[peer]) + (size_t)axis * ctx->p2p_cap * 6;
      ctx->peer_box[axis][dir] = mb[peer] + (axis * 2 + dir) * 2;
    }
  }
  // every rank must take the same path: agree on it
  int* dflag = ctx->cntbuf;
  const int mine_ok = ok ? 1 : 0;
  CK(cudaMemcpy(dflag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice));
  NCK(ncclAllReduce(dflag, dflag, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  int all_ok = 0;
  CK(cudaMemcpyAsync(&all_ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->p2p = all_ok == 1;
  if (!ctx->p2p && ctx->cfg.rank == 0) fprintf(stderr, "cubep3m_b200: peer-memory particle_pass unavailable, using NCCL send/recv\n");
#endif
  return 0;
}


// ---- slab-decomposed coarse solve (coarse_slab.cuh): allocation, peer mapping, the rank's rows of kern_c
bool cs_supported(const Dims& d) {
  const int W = d.world;
  if (W > cslab::MAXW) return false;
  if (d.Nc[2] % W || d.Nc[1] % W) return false;
  const int zs = d.Nc[2] / W;
  return zs >= 1 && d.nc_node % zs == 0;
}
int cs_alloc(cubep3m_b200_ctx* ctx) {
  const Dims& d = ctx->d;
  const int W = d.world, Nx = d.Nc[0], Ny = d.Nc[1], Nz = d.Nc[2], hc = Nx / 2 + 1;
  ctx->cs_zs = Nz / W; ctx->cs_ys = Ny / W;
  auto al = [](size_t n) { return (n + 63) / 64 * 64; };
  const size_t slab_f = al((size_t)2 * hc * Ny * ctx->cs_zs), pen_f = al((size_t)2 * hc * ctx->cs_ys * Nz);
  const size_t force_f = al((size_t)3 * (d.nc_node + 2) * (d.nc_node + 2) * (d.nc_node + 2));
  size_t o = 0;
  ctx->cs_off_slab = o; o += slab_f;
  ctx->cs_off_T = o; o += pen_f;
  for (int c = 0; c < 3; ++c) { ctx->cs_off_back[c] = o; o += slab_f; }
  ctx->cs_off_force = o; o += force_f;
  ctx->cs_off_mail = o; o += al((size_t)cslab::PH_COUNT * cslab::MAXW);
  ctx->cs_floats = o;
  CK(cudaMalloc((void**)&ctx->cs_xchg, o * sizeof(float)));
  CK(cudaMemset(ctx->cs_xchg, 0, o * sizeof(float)));
  CK(cudaMalloc((void**)&ctx->cs_G, pen_f * sizeof(float)));
  CK(cudaMalloc((void**)&ctx->cs_real3, (size_t)3 * ctx->cs_zs * Ny * Nx * sizeof(float)));
  CK(cudaMalloc((void**)&ctx->cs_kern_rows, (size_t)3 * Nz * ctx->cs_ys * hc * sizeof(float)));
  ctx->force_c = ctx->cs_xchg + ctx->cs_off_force;
  for (int r = 0; r < cslab::MAXW; ++r) ctx->cs_peers.base[r] = nullptr;
  ctx->cs_peers.base[ctx->cfg.rank] = ctx->cs_xchg;
  return 0;
}
// maps every other rank's copy of an exchange allocation (cudaIpc handles all-gathered over NCCL) into P->base[]; ok = false if any rank could not.
// Collective: every rank must call it with its own allocation.
int map_peers(cubep3m_b200_ctx* ctx, float* mine_ptr, PeerTable* P, std::vector<void*>* opened, bool* ok_out) {
  *ok_out = true;
  for (int r = 0; r < 8; ++r) P->base[r] = nullptr;
  P->base[ctx->cfg.rank] = mine_ptr;
  if (ctx->d.world == 1) return 0;
#ifdef CUBEP3M_WITH_NCCL
  const Dims& d = ctx->d;
  struct Handle { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  Handle mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = cudaIpcGetMemHandle(&mine.h, mine_ptr) == cudaSuccess;
  cudaGetLastError();
  Handle* dall = nullptr;
  CK(cudaMalloc((void**)&dall, sizeof(Handle) * d.world));
  CK(cudaMemcpy(dall + ctx->cfg.rank, &mine, sizeof(Handle), cudaMemcpyHostToDevice));
  NCK(ncclAllGather(dall + ctx->cfg.rank, dall, sizeof(Handle), ncclChar, ctx->comm, ctx->stream));
  std::vector<Handle> all(d.world);
  CK(cudaMemcpyAsync(all.data(), dall, sizeof(Handle) * d.world, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dall);
  bool ok = true;
  for (int r = 0; r < d.world; ++r) ok = ok && all[r].ok;
  for (int r = 0; r < d.world && ok; ++r) {
    if (r == ctx->cfg.rank) continue;
    void* p = nullptr;
    ok = cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (p) opened->push_back(p);
    if (!ok) { cudaGetLastError(); break; }
    P->base[r] = (float*)p;
  }
  int* dflag = ctx->cntbuf;
  const int mine_ok = ok ? 1 : 0;
  CK(cudaMemcpy(dflag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice));
  NCK(ncclAllReduce(dflag, dflag, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  int all_ok = 0;
  CK(cudaMemcpyAsync(&all_ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *ok_out = all_ok == 1;
  return 0;
#else
  *ok_out = false;
  return 0;
#endif
}
int cs_map_peers(cubep3m_b200_ctx* ctx, bool* ok_out) { return map_peers(ctx, ctx->cs_xchg, &ctx->cs_peers, &ctx->cs_ipc_opened, ok_out); }
// a stream-ordered barrier over all ranks (a one-int all-reduce)
int rank_barrier(cubep3m_b200_ctx* ctx) {
  if (ctx->d.world == 1) return 0;
#ifdef CUBEP3M_WITH_NCCL
  NCK(ncclAllReduce(ctx->cntbuf + 4, ctx->cntbuf + 4, 1, ncclInt32, ncclSum, ctx->comm, ctx->stream));
  return 0;
#else
  return CUBEP3M_B200_ENCCL;
#endif
}
void cs_free(cubep3m_b200_ctx* ctx) {
  for (void* q : ctx->cs_ipc_opened) cudaIpcCloseMemHandle(q);
  ctx->cs_ipc_opened.clear();
  if (ctx->cs_xchg) { cudaFree(ctx->cs_xchg); if (ctx->force_c == ctx->cs_xchg + ctx->cs_off_force) ctx->force_c = nullptr; ctx->cs_xchg = nullptr; }
  if (ctx->cs_G) cudaFree(ctx->cs_G);
  if (ctx->cs_real3) cudaFree(ctx->cs_real3);
  if (ctx->cs_kern_rows) cudaFree(ctx->cs_kern_rows);
  ctx->cs_G = ctx->cs_real3 = ctx->cs_kern_rows = nullptr;
}

int fetch_counters(cubep3m_b200_ctx* ctx) {
  CK(cudaMemcpyAsync(ctx->hcnt, ctx->dcnt, sizeof(DevCounters), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int overflow_status(const DevCounters* h) {
  if (h->overflow & 1) return CUBEP3M_B200_EPASSBUF;
  if (h->overflow & 2) return CUBEP3M_B200_EMAXNP;
  if (h->overflow & 4) return CUBEP3M_B200_EMAXLLF;
  if (h->overflow & 8) return CUBEP3M_B200_ECAPACITY;
  return 0;
}

// ---------------------------------------------------------------- stages
int do_drift(cubep3m_b200_ctx* ctx, float dt, float dt_old, const float off[3]) {
  if (ctx->np_local > 0)
    LAUNCH(ctx, KC_DRIFT, part::drift_kernel, (ctx->np_local + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->xv[ctx->cur], ctx->np_local, dt + dt_old, off[0], off[1], off[2]);
  CK(cudaGetLastError());
  ctx->sorted = false; ctx->passed = false;
  ctx->np_all = ctx->np_local;
  return 0;
}

// particle_pass.f90:69-722 for nodes_dim = 1 (every neighbour is this rank) or over NCCL
int exchange_axis(cubep3m_b200_ctx* ctx, int axis, int n_plus_out, int n_minus_out, int* n_from_minus, int* n_from_plus);

int do_pass(cubep3m_b200_ctx* ctx, int* np_buf_max, const float* drift = nullptr, bool in_step = false) {   // drift = {dt+dt_old, ox, oy, oz}: fuse update_position into the first pack
  const Dims& d = ctx->d;
  const float lo = -(float)d.b, hi = (float)d.mT + (float)d.b;
  const float fmT = (float)d.mT, rnf = (float)d.b;
  const float cut_hi = fmT - rnf, cut_lo = rnf;
  const float hi_clamp = (fmT + rnf) - ctx->cfg.eps;
  const int cap = d.max_buf / 6;
  int np = ctx->np_all;
  const int np_first = np;                          // particles present before the first exchange
  int nlist = 0;
  int* blist = ctx->blist;
  const Dims& dd = ctx->d;
  // CUBEP3M_B200_FUSE_KEYS=1: key + histogram ride on the pack / unpack kernels (do_sort then skips key_hist_kernel). Measured at 512^3: pass 1.8 -> 5.0 ms
  // against 1.4 ms saved — the table atomic's DRAM round trip lands in front of the pack kernel's warp-synchronous slot allocation and the kernel turns
  // latency-bound; the stand-alone key_hist_kernel (load -> atomic -> exit) runs at the DRAM limit. Off by default.
  static const bool fuse_env = [] { const char* e = getenv("CUBEP3M_B200_FUSE_KEYS"); return e && atoi(e) == 1; }();
  const bool fuse_keys = fuse_env && in_step && drift != nullptr;
  part::KeyArgs KA{lo, hi, dd.b, dd.H, fuse_keys ? ctx->key : nullptr, ctx->fcur, ctx->cand, ctx->cand_cap, ctx->rank};
  if (fuse_keys) {
    if (!ctx->hist_clean) CK(cudaMemsetAsync(ctx->fcur, 0, sizeof(unsigned int) * (dd.NF / 2), ctx->stream));
    ctx->hist_clean = false;
    CK(cudaMemsetAsync(&ctx->dcnt->np_deleted, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(&ctx->dcnt->n_cand, 0, sizeof(int), ctx->stream));
  }
  ctx->keys_fused = fuse_keys;
  *np_buf_max = 0;
  CK(cudaMemsetAsync(&ctx->dcnt->n_blist, 0, sizeof(int), ctx->stream));
  for (int axis = 0; axis < 3; ++axis) {
    CK(cudaMemsetAsync(&ctx->dcnt->n_send[0], 0, 2 * sizeof(int), ctx->stream));
    // peer-memory path: only inside particle_mesh (its end-of-step all-reduce is what keeps a fast rank from overwriting a region the
    // neighbour has not consumed yet) and without PIDs
    const bool p2p = ctx->p2p && in_step && d.Dg[axis] > 1 && !ctx->cfg.pid;
    float* dst_plus = p2p ? ctx->peer_recv[axis][0] : ctx->sendbuf[0];
    float* dst_minus = p2p ? ctx->peer_recv[axis][1] : ctx->sendbuf[1];
    const int cap_axis = p2p ? ctx->p2p_cap : cap;
    if (np > 0) {
      const float z4[4] = {0.f, 0.f, 0.f, 0.f};
      const float* dr = drift ? drift : z4;
      const long long nvis = axis == 0 ? np : (long long)nlist + (np - np_first);
      const int grid = (int)((nvis + part::TPB - 1) / part::TPB);
#define PACK_ARGS ctx->xv[ctx->cur], ctx->pid[ctx->cur], np, axis, lo, hi, cut_hi, cut_lo, dst_plus, dst_minus, ctx->sendpid[0], ctx->sendpid[1], cap_axis, \
                  ctx->dcnt, dr[0], dr[1], dr[2], dr[3], blist, nlist, np_first, KA
      if (axis == 0 && drift && fuse_keys) LAUNCH(ctx, KC_PASS_PACK, (part::pass_pack_kernel<true, false, true>), grid, part::TPB, 0, PACK_ARGS);
      else if (axis == 0 && drift) LAUNCH(ctx, KC_PASS_PACK, (part::pass_pack_kernel<true, false>), grid, part::TPB, 0, PACK_ARGS);
      else if (axis == 0) LAUNCH(ctx, KC_PASS_PACK, (part::pass_pack_kernel<false, false>), grid, part::TPB, 0, PACK_ARGS);
      else if (grid > 0) LAUNCH(ctx, KC_PASS_PACK, (part::pass_pack_kernel<false, true>), grid, part::TPB, 0, PACK_ARGS);
#undef PACK_ARGS
    }
    CK(cudaGetLastError());
    if (p2p) {
      // one kernel publishes counts + flags into the neighbours' mailboxes, one waits for both of ours: no NCCL call and a single
      // host synchronisation per axis (the NCCL path needs two groups and two synchronisations)
      ctx->hbox[0] = ctx->hbox[1] = ctx->hbox[2] = 0;
      CK(cudaMemsetAsync(ctx->cntbuf, 0, 4 * sizeof(int), ctx->stream));
      LAUNCH(ctx, KC_PASS_PACK, part::pass_publish_kernel, 1, 32, 0, ctx->dcnt, ctx->peer_box[axis][0], ctx->peer_box[axis][1], (int)ctx->p2p_epoch);
      LAUNCH(ctx, KC_PASS_UNPACK, part::pass_wait_kernel, 1, 32, 0, ctx->mailbox + axis * 4, (int)ctx->p2p_epoch, ctx->cntbuf, 60000000000LL);   // ~30 s: ranks may enter a step seconds apart
      CK(cudaMemcpyAsync(ctx->hbox, ctx->cntbuf, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (int st = fetch_counters(ctx)) return st;
    if (int st = overflow_status(ctx->hcnt)) return st;
    const int n_plus = ctx->hcnt->n_send[0], n_minus = ctx->hcnt->n_send[1];
    if (axis == 0) nlist = ctx->hcnt->n_blist;
    if (n_plus * 6 > d.max_buf || n_minus * 6 > d.max_buf) return CUBEP3M_B200_EPASSBUF;      // particle_pass.f90:96-99
    *np_buf_max = std::max(*np_buf_max, std::max(n_plus, n_minus));
    int r_plus = 0, r_minus = 0;   // r_plus: particles that travelled in + direction (arrive from the - neighbour)
    if (p2p) {
      if (ctx->hbox[2]) { fprintf(stderr, "cubep3m_b200: particle_pass timed out waiting for a neighbour (axis %d)\n", axis); return CUBEP3M_B200_ENCCL; }
      r_plus = ctx->hbox[0]; r_minus = ctx->hbox[1];
      if (r_plus > ctx->p2p_cap || r_minus > ctx->p2p_cap) return CUBEP3M_B200_EPASSBUF;       // the sender flagged the overflow too
      ctx->recvbuf[0] = ctx->recvbuf_own[0] + (size_t)axis * ctx->p2p_cap * 6;
      ctx->recvbuf[1] = ctx->recvbuf_own[1] + (size_t)axis * ctx->p2p_cap * 6;
    } else if (int st = exchange_axis(ctx, axis, n_plus, n_minus, &r_plus, &r_minus)) return st;
    if ((long long)np + r_plus + r_minus > d.max_np) return CUBEP3M_B200_EMAXNP;              // particle_pass.f90:136-139
    const int nr = r_plus + r_minus;
    if (nr > 0)
      LAUNCH(ctx, KC_PASS_UNPACK, part::pass_unpack_kernel, (nr + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->xv[ctx->cur], ctx->pid[ctx->cur], np, axis,
             ctx->recvbuf[0], r_plus, ctx->recvbuf[1], r_minus, ctx->recvpid[0], ctx->recvpid[1], fmT, rnf, ctx->cfg.eps, hi_clamp, KA, ctx->dcnt);
    CK(cudaGetLastError());
    np += nr;
  }
  ctx->np_all = np;
  ctx->passed = true;
  ctx->sorted = false;
  return 0;
}

int exchange_axis(cubep3m_b200_ctx* ctx, int axis, int n_plus_out, int n_minus_out, int* r_plus, int* r_minus) {
  const Dims& d = ctx->d;
  if (d.Dg[axis] == 1) {
    // the + and - neighbours are this rank: what was sent in + direction is received "from the - neighbour"
    ctx->recvbuf[0] = ctx->sendbuf[0]; ctx->recvbuf[1] = ctx->sendbuf[1];
    ctx->recvpid[0] = ctx->sendpid[0]; ctx->recvpid[1] = ctx->sendpid[1];
    *r_plus = n_plus_out; *r_minus = n_minus_out;
    return 0;
  }
#ifdef CUBEP3M_WITH_NCCL
  // particle_pass.f90:125,141-144 (+ pass) and :217,233-236 (- pass): count exchange, then the particle payloads.
  // With Dg == 2 both neighbours are the same peer; NCCL matches sends and receives to one peer in issue order, and both sides
  // issue [plus-going, minus-going] / [from-minus, from-plus], which pairs plus-going with from-minus as required.
  const int minus = d.nbr[2 * axis], plus = d.nbr[2 * axis + 1];
  NCK(ncclGroupStart());
  NCK(ncclSend(&ctx->dcnt->n_send[0], 1, ncclInt32, plus, ctx->comm, ctx->stream));
  NCK(ncclSend(&ctx->dcnt->n_send[1], 1, ncclInt32, minus, ctx->comm, ctx->stream));
  NCK(ncclRecv(&ctx->cntbuf[0], 1, ncclInt32, minus, ctx->comm, ctx->stream));
  NCK(ncclRecv(&ctx->cntbuf[1], 1, ncclInt32, plus, ctx->comm, ctx->stream));
  NCK(ncclGroupEnd());
  int hc[2];
  CK(cudaMemcpyAsync(hc, ctx->cntbuf, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *r_plus = hc[0]; *r_minus = hc[1];
  if ((long long)hc[0] * 6 > d.max_buf || (long long)hc[1] * 6 > d.max_buf) return CUBEP3M_B200_EPASSBUF;
  ctx->recvbuf[0] = ctx->recvbuf_own[0]; ctx->recvbuf[1] = ctx->recvbuf_own[1];
  ctx->recvpid[0] = ctx->recvpid_own[0]; ctx->recvpid[1] = ctx->recvpid_own[1];
  NCK(ncclGroupStart());
  if (n_plus_out > 0) NCK(ncclSend(ctx->sendbuf[0], (size_t)6 * n_plus_out, ncclFloat, plus, ctx->comm, ctx->stream));
  if (n_minus_out > 0) NCK(ncclSend(ctx->sendbuf[1], (size_t)6 * n_minus_out, ncclFloat, minus, ctx->comm, ctx->stream));
  if (hc[0] > 0) NCK(ncclRecv(ctx->recvbuf[0], (size_t)6 * hc[0], ncclFloat, minus, ctx->comm, ctx->stream));
  if (hc[1] > 0) NCK(ncclRecv(ctx->recvbuf[1], (size_t)6 * hc[1], ncclFloat, plus, ctx->comm, ctx->stream));
  if (ctx->cfg.pid) {                                   // particle_pass.f90:150-153
    if (n_plus_out > 0) NCK(ncclSend(ctx->sendpid[0], (size_t)n_plus_out, ncclInt64, plus, ctx->comm, ctx->stream));
    if (n_minus_out > 0) NCK(ncclSend(ctx->sendpid[1], (size_t)n_minus_out, ncclInt64, minus, ctx->comm, ctx->stream));
    if (hc[0] > 0) NCK(ncclRecv(ctx->recvpid[0], (size_t)hc[0], ncclInt64, minus, ctx->comm, ctx->stream));
    if (hc[1] > 0) NCK(ncclRecv(ctx->recvpid[1], (size_t)hc[1], ncclInt64, plus, ctx->comm, ctx->stream));
  }
  NCK(ncclGroupEnd());
  return 0;
#else
  (void)n_plus_out; (void)n_minus_out; (void)r_plus; (void)r_minus;
  return CUBEP3M_B200_ENCCL;
#endif
}

// cross-rank reductions of the limiters and DIAG sums (the mpi_reduce/mpi_bcast pairs of particle_mesh_threaded.f90:646-703,
// coarse_max_dt.f90:34-37, delete_particles.f90:63)
int reduce_scalars(cubep3m_b200_ctx* ctx, float* maxv, int nmax, double* sumv, int nsum) {
  if (ctx->d.world == 1) return 0;
#ifdef CUBEP3M_WITH_NCCL
  float* dmax = ctx->redbuf;
  double* dsum = reinterpret_cast<double*>(ctx->redbuf + 16);
  CK(cudaMemcpyAsync(dmax, maxv, nmax * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dsum, sumv, nsum * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NCK(ncclGroupStart());
  NCK(ncclAllReduce(dmax, dmax, nmax, ncclFloat, ncclMax, ctx->comm, ctx->stream));
  NCK(ncclAllReduce(dsum, dsum, nsum, ncclDouble, ncclSum, ctx->comm, ctx->stream));
  NCK(ncclGroupEnd());
  CK(cudaMemcpyAsync(maxv, dmax, nmax * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(sumv, dsum, nsum * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
#else
  (void)maxv; (void)nmax; (void)sumv; (void)nsum;
  return CUBEP3M_B200_ENCCL;
#endif
}

// cell sort of xv[cur][0:np_all) -> xv[cur^1], builds fstart and the PP work lists
int do_sort(cubep3m_b200_ctx* ctx, int* np_deleted) {
  const Dims& d = ctx->d;
  const int np = ctx->np_all;
  const float lo = -(float)d.b, hi = (float)d.mT + (float)d.b;
  CK(cudaMemsetAsync(&ctx->dcnt->n_multi, 0, 2 * sizeof(int), ctx->stream));
  if (!ctx->keys_fused) {
    if (!ctx->hist_clean) CK(cudaMemsetAsync(ctx->fcur, 0, sizeof(unsigned int) * (d.NF / 2), ctx->stream));   // two 16-bit counters per word; otherwise particle_mesh cleared it behind the last step's PP stage
    ctx->hist_clean = false;
    CK(cudaMemsetAsync(&ctx->dcnt->np_deleted, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(&ctx->dcnt->n_cand, 0, sizeof(int), ctx->stream));
    if (np > 0) {
      part::KeyArgs KA{lo, hi, d.b, d.H, ctx->key, ctx->fcur, ctx->cand, ctx->cand_cap, ctx->rank};
      LAUNCH(ctx, KC_KEY_HIST, part::key_hist_kernel, (np + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->xv[ctx->cur], np, KA, ctx->dcnt);
    }
  }
  ctx->keys_fused = false;
  const int nb = (int)((d.NF + part::SCAN_BLOCK - 1) / part::SCAN_BLOCK);
  if (ctx->scan_onepass) {
    CK(cudaMemsetAsync(ctx->scan_status, 0, sizeof(unsigned long long) * ((size_t)nb + 1), ctx->stream));
    LAUNCH(ctx, KC_SCAN, part::scan_apply_kernel<true>, nb, part::TPB, 0, ctx->fcur, d.NF, nullptr, ctx->fstart, d.H, d.nc_buf, d.nc_node, ctx->multi_list,
           ctx->occ_list, ctx->list_cap, ctx->cfg.ppint ? 1 : 0, 0, ctx->dcnt, ctx->scan_status, nb);
  } else {
    LAUNCH(ctx, KC_SCAN, part::scan_reduce_kernel, (nb + part::SCAN_RB - 1) / part::SCAN_RB, part::TPB, 0, ctx->fcur, d.NF, ctx->blocksum, nb);
    LAUNCH(ctx, KC_SCAN, part::scan_blocksums_kernel, 1, 1024, 0, ctx->blocksum, nb);
    LAUNCH(ctx, KC_SCAN, part::scan_apply_kernel<false>, nb, part::TPB, 0, ctx->fcur, d.NF, ctx->blocksum, ctx->fstart, d.H, d.nc_buf, d.nc_node, ctx->multi_list,
           ctx->occ_list, ctx->list_cap, ctx->cfg.ppint ? 1 : 0, 0, ctx->dcnt, nullptr, nb);
  }
  // The histogram has had its last reader: clear it for the next sort (0.06 ms at 256^3 particles, 0.4 ms at 512^3; hiding it under the PP_EXT kernels
  // — defer_hist_zero, an A/B knob — did not pay).
  if (!ctx->defer_hist_zero && ctx->hist_mode == 0) {
    CK(cudaMemsetAsync(ctx->fcur, 0, sizeof(unsigned int) * (d.NF / 2), ctx->stream));
    ctx->hist_clean = true;
  }
  // inside particle_mesh with PP_EXT on, the scatter also lists the margin roles for the PP_EXT limiter (do_pp_ext_margin)
  const bool roles = ctx->want_roles && ctx->cfg.pp_ext && ctx->cfg.pp_range > 0 && ctx->ppext_margin_max && ctx->margin_roles;
  ctx->roles_listed = false;
  if (np > 0) {
    const part::MarginGeom G{d.H, d.b, d.m, d.T, ctx->cfg.pp_range};
    if (roles) {
      CK(cudaMemsetAsync(&ctx->dcnt->n_margin_roles, 0, sizeof(int), ctx->stream));
      LAUNCH(ctx, KC_SCATTER, part::scatter_kernel<true>, (np + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->xv[ctx->cur], ctx->pid[ctx->cur], ctx->key, np, ctx->rank, ctx->fcur, ctx->fstart,
             ctx->xv[ctx->cur ^ 1], ctx->pid[ctx->cur ^ 1], d.max_np, G, ctx->margin_roles, ctx->margin_cap, &ctx->dcnt->n_margin_roles, ctx->hist_mode);
      ctx->roles_listed = true;
    } else {
      LAUNCH(ctx, KC_SCATTER, part::scatter_kernel<false>, (np + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->xv[ctx->cur], ctx->pid[ctx->cur], ctx->key, np, ctx->rank, ctx->fcur, ctx->fstart,
             ctx->xv[ctx->cur ^ 1], ctx->pid[ctx->cur ^ 1], d.max_np, G, nullptr, 0, nullptr, ctx->hist_mode);
    }
  }
  CK(cudaGetLastError());
  if (int st = fetch_counters(ctx)) return st;
  // the histogram stays as counted (the scatter no longer counts it down): the next sort clears it first, unless particle_mesh has meanwhile
  // cleared it behind its PP stage (hist_clean)
  if (ctx->hcnt->overflow & (4 | 8)) { ctx->hist_clean = false; return overflow_status(ctx->hcnt); }   // a wrapped counter: clear everything before the next sort
  if (ctx->hist_mode != 0) ctx->hist_clean = true;      // every occupied cell's word was stored / counted back to zero by the scatter
  ctx->cur ^= 1;
  ctx->np_all = np - ctx->hcnt->np_deleted;
  if (!ctx->passed) ctx->np_local = ctx->np_all;
  if (np_deleted) *np_deleted = ctx->hcnt->np_deleted;
  ctx->sorted = true;
  return 0;
}

// delete_particles.f90:14-50 on the sorted array
// kick != nullptr: {a_mid, dt}: coarse_velocity (coarse_velocity.f90:137-179) is applied on the way (fused kernel, see coarse.cuh)
int do_delete(cubep3m_b200_ctx* ctx, const float* kick = nullptr) {
  const Dims& d = ctx->d;
  const int rows = d.nc_node * d.nc_node;
  CK(cudaMemsetAsync(ctx->rowoff + rows, 0, sizeof(int), ctx->stream));
  LAUNCH(ctx, KC_COMPACT, part::row_count_kernel, (rows + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->fstart, d.H, d.nc_buf, d.nc_node, ctx->rowoff);
  LAUNCH(ctx, KC_SCAN, part::scan_blocksums_kernel, 1, 1024, 0, ctx->rowoff, rows + 1);   // entry [rows] (initialised to 0) becomes the total
  if (kick)
    LAUNCH(ctx, KC_CIC_KICK, coarse::cic_kick_compact_kernel, rows, coarse::TPB, 0, ctx->xv[ctx->cur], ctx->pid[ctx->cur], ctx->fstart, ctx->rowoff, ctx->force_c, d.H,
           d.nc_buf, d.nc_node, kick[0], ctx->cfg.G, kick[1], ctx->cfg.coarse_ngp, 1, ctx->xv[ctx->cur ^ 1], ctx->pid[ctx->cur ^ 1]);
  else
    LAUNCH(ctx, KC_COMPACT, part::compact_rows_kernel, rows, part::TPB, 0, ctx->xv[ctx->cur], ctx->pid[ctx->cur], ctx->fstart, ctx->rowoff, d.H, d.nc_buf, d.nc_node,
           ctx->xv[ctx->cur ^ 1], ctx->pid[ctx->cur ^ 1]);
  CK(cudaGetLastError());
  int total = 0;
  CK(cudaMemcpyAsync(&ctx->hcnt->np_phys, ctx->rowoff + rows, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  total = ctx->hcnt->np_phys;
  ctx->cur ^= 1;
  ctx->np_local = total; ctx->np_all = total;
  ctx->sorted = false; ctx->passed = false;
  return 0;
}

// fine CIC deposit of one tile into t_rho ((n + 2) x n x n): shared-memory-staged scatter (default) or the gather kernel (CUBEP3M_B200_FINECIC=gather)
int launch_cic_density(cubep3m_b200_ctx* ctx, float* t_rho, int tx, int ty, int tz, float mass_p, double* dsum, int* count) {
  const Dims& d = ctx->d;
  const int n = d.n;
  static const bool gather = [] { const char* e = getenv("CUBEP3M_B200_FINECIC"); return e && !strcmp(e, "gather"); }();
  const size_t smem = (size_t)25 * (n + 2) * sizeof(float);
  if (gather || smem > 200 * 1024) {
    LAUNCH(ctx, KC_DENSITY, fine::cic_density_kernel, NUM_SMS * 16, fine::TPB, 0, ctx->xv[ctx->cur], ctx->fstart, t_rho, n, d.b, d.m, d.H, tx, ty, tz, mass_p, dsum, count);
    return 0;
  }
  static bool attr_dev[64] = {false};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (!attr_dev[dev & 63]) { CK(cudaFuncSetAttribute(fine::cic_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024))); attr_dev[dev & 63] = true; }
  CK(cudaMemsetAsync(t_rho, 0, sizeof(float) * (size_t)(n + 2) * n * n, ctx->stream));
  LAUNCH(ctx, KC_DENSITY, fine::cic_scatter_kernel, (n / 4) * (n / 4), fine::TPB, smem, ctx->xv[ctx->cur], ctx->fstart, t_rho, n, d.b, d.m, d.H, tx, ty, tz, mass_p, dsum, count);
  return 0;
}

// fine-mesh solve of one tile: (density fused into) forward FFT -> fused z pass with the Green's functions -> inverse y, x + crop.
// materialise = true writes rho_f to tile_rho first with the stand-alone deposit kernels (debug getter / reference ordering).
int fine_tile_solve(cubep3m_b200_ctx* ctx, int tile, float mass_p, bool materialise, int* scratch_count, int set = 0) {
  const Dims& d = ctx->d;
  float* t_rho = set ? ctx->tile_rho_s[set] : ctx->tile_rho;
  float* t_g = set ? ctx->tile_g_s[set] : ctx->tile_g;
  float* t_force = set ? ctx->force_f_s[set] : ctx->force_f[0];
  const int T = d.T, n = d.n;
  const int tz = tile / (T * T), ty = (tile / T) % T, tx = tile % T;   // particle_mesh_threaded.f90:86-90 (cur_tile-1, x fastest)
  const float scale = 1.0f / (((float)n * (float)n) * (float)n);       // fft_fine.f90:51
  if (!ctx->cfg.ngp) {   // fine CIC: materialised gather deposit, then the generic solve
    if (int st = launch_cic_density(ctx, t_rho, tx, ty, tz, mass_p, &ctx->dcnt->sum_rho_f, scratch_count)) return st;
    return fftk::fine_solve(ctx, fine_mesh(ctx), t_rho, t_g, d.hc, ctx->kern_f, ctx->kf_stride, ctx->kf_pitch, t_force, d.b - 2, d.fdim, scale, &ctx->dcnt->f_force_max2_bits);
  }
  if (materialise) {
    LAUNCH(ctx, KC_DENSITY, fine::ngp_density_kernel, NUM_SMS * 8, fine::TPB, 0, ctx->fstart, ctx->tile_rho, n, d.b, d.m, d.H, tx, ty, tz, mass_p,
           &ctx->dcnt->sum_rho_f, scratch_count);
    if (ctx->hcnt->n_cand > 0)
      LAUNCH(ctx, KC_DENSITY, fine::ngp_fixup_kernel, std::min(NUM_SMS, (std::min(ctx->hcnt->n_cand, ctx->cand_cap) + fine::TPB - 1) / fine::TPB), fine::TPB, 0,
             ctx->cand, &ctx->dcnt->n_cand, ctx->cand_cap, ctx->tile_rho, n, d.b, d.m, tx, ty, tz, mass_p, &ctx->dcnt->sum_rho_f);
    return fftk::fine_solve(ctx, fine_mesh(ctx), ctx->tile_rho, ctx->tile_g, d.hc, ctx->kern_f, ctx->kf_stride, ctx->kf_pitch, ctx->force_f[0], d.b - 2, d.fdim, scale, &ctx->dcnt->f_force_max2_bits);
  }
  fftk::NgpSource src{ctx->fstart, d.H, d.b, tx * d.m, ty * d.m, tz * d.m, mass_p, ctx->deltas + (size_t)tile * fine::DELTA_CAP, ctx->ndelta + tile,
                      fine::DELTA_CAP, &ctx->dcnt->sum_rho_f};
  return fftk::fine_solve(ctx, fine_mesh(ctx), t_rho, t_g, ctx->kf_pitch, ctx->kern_f, ctx->kf_stride, ctx->kf_pitch, t_force, d.b - 2, d.fdim, scale, &ctx->dcnt->f_force_max2_bits, &src);
}

int do_fine(cubep3m_b200_ctx* ctx, float a_mid, float dt, float mass_p, float* ms_dep_fft, float* ms_kick) {
  const Dims& d = ctx->d;
  (void)ms_dep_fft; (void)ms_kick;
  // once per step: per-tile particle counts (parity getter) and the per-tile lists of ulp-boundary mass moves
  CK(cudaMemsetAsync(ctx->ndelta, 0, sizeof(int) * d.tiles_node, ctx->stream));
  CK(cudaMemsetAsync(ctx->tile_counts, 0, sizeof(int) * d.tiles_node, ctx->stream));
  if (ctx->cfg.ngp) LAUNCH(ctx, KC_DENSITY, fine::tile_counts_kernel, d.tiles_node, fine::TPB, 0, ctx->fstart, d.H, d.nc_buf, d.nc_tile, d.T, ctx->tile_counts);
  if (ctx->cfg.ngp && ctx->hcnt->n_cand > 0)
    LAUNCH(ctx, KC_DENSITY, fine::build_tile_deltas_kernel, std::min(NUM_SMS, (std::min(ctx->hcnt->n_cand, ctx->cand_cap) + fine::TPB - 1) / fine::TPB), fine::TPB,
           0, ctx->cand, &ctx->dcnt->n_cand, ctx->cand_cap, d.n, d.b, d.m, d.T, mass_p, ctx->deltas, ctx->ndelta, &ctx->dcnt->sum_rho_f, &ctx->dcnt->overflow);
  const int S = std::min(ctx->tile_streams, d.tiles_node);
  const size_t fstride = (size_t)d.fdim * d.fdim * d.fdim;
  if (S > 1) {
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream_main));
    for (int q = 1; q < S; ++q) CK(cudaStreamWaitEvent(ctx->stream_aux[q], ctx->ev_fork, 0));
  }
  int status = 0;
  for (int tile = 0; tile < d.tiles_node && !status; ++tile) {
    const int set = tile % S;
    ctx->stream = set ? ctx->stream_aux[set] : ctx->stream_main;      // LAUNCH() targets ctx->stream
    status = fine_tile_solve(ctx, tile, mass_p, false, ctx->cfg.ngp ? nullptr : ctx->tile_counts + tile, set);
    if (!status && !ctx->cfg.ngp) {
      const int T = d.T;
      const int tz = tile / (T * T), ty = (tile / T) % T, tx = tile % T;
      float* ff = set ? ctx->force_f_s[set] : ctx->force_f[0];
      LAUNCH(ctx, KC_NGP_KICK, fine::cic_fine_kick_kernel, d.nc_tile * d.nc_tile, fine::TPB, 0, ctx->xv[ctx->cur], ctx->fstart, ff, ff + fstride, ff + 2 * fstride, d.H,
             d.nc_buf, d.nc_tile, d.b, d.m, d.fdim, tx, ty, tz, a_mid, ctx->cfg.G, dt);
    } else if (!status && ctx->cfg.ngp_fmesh_force) {
      const int T = d.T;
      const int tz = tile / (T * T), ty = (tile / T) % T, tx = tile % T;
      float* ff = set ? ctx->force_f_s[set] : ctx->force_f[0];
      LAUNCH(ctx, KC_NGP_KICK, fine::ngp_kick_kernel, d.nc_tile * d.nc_tile, fine::TPB, 0, ctx->xv[ctx->cur], ctx->fstart, ff, ff + fstride, ff + 2 * fstride, d.H, d.nc_buf,
             d.nc_tile, d.b, d.m, d.fdim, tx, ty, tz, a_mid, ctx->cfg.G, dt);
    }
  }
  ctx->stream = ctx->stream_main;
  for (int q = 1; q < S; ++q) { CK(cudaEventRecord(ctx->ev_join[q], ctx->stream_aux[q])); CK(cudaStreamWaitEvent(ctx->stream_main, ctx->ev_join[q], 0)); }
  if (status) return status;
  CK(cudaGetLastError());
  return 0;
}

int do_pp(cubep3m_b200_ctx* ctx, float a_mid, float dt, float mass_p) {
  pp::PPParams P;
  P.mass_p = mass_p; P.rsoft = ctx->cfg.rsoft; P.pp_bias = ctx->cfg.pp_bias; P.a_mid = a_mid; P.G = ctx->cfg.G; P.dt = dt;
  P.cutoff = (float)ctx->cfg.nf_cutoff;
  if (ctx->cfg.ppint && ctx->cfg.ngp) {      // the llf binning of particle_mesh_threaded.f90:274-285 only exists inside #ifdef NGP
    P.apply = ctx->cfg.pp_force_flag;
    const int n_multi = std::min(ctx->hcnt->n_multi, ctx->list_cap);
    if (n_multi > 0) {
      // (cell, 32-target chunk) items, one warp each; the one-warp-per-cell kernel only if the item list overflowed (decided on the device)
      const int icap = ctx->ppext_cell_mode ? ctx->ppint_item_cap : 0;
      CK(cudaMemsetAsync(&ctx->dcnt->n_ppint_items[0], 0, 3 * sizeof(int), ctx->stream));
      if (icap > 0) {
        LAUNCH(ctx, KC_PPINT, pp::ppint_items_kernel, std::min((n_multi + pp::TB_NT - 1) / pp::TB_NT, NUM_SMS * 8), pp::TB_NT, 0, ctx->fstart, ctx->multi_list, &ctx->dcnt->n_multi,
               ctx->list_cap, ctx->cfg.max_llf, ctx->ppint_items, icap, ctx->dcnt->n_ppint_items, ctx->dcnt);
        LAUNCH(ctx, KC_PPINT, pp::ppint_cell_kernel, NUM_SMS * 16, pp::TB_NT, 0, ctx->xv[ctx->cur], ctx->fstart, P, ctx->dcnt, ctx->ppint_items, icap, ctx->dcnt->n_ppint_items,
               &ctx->dcnt->ppint_ticket);
      }
      LAUNCH(ctx, KC_PPINT, pp::ppint_kernel, std::min((n_multi + 3) / 4, NUM_SMS * 16), pp::TPB, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->multi_list,
             &ctx->dcnt->n_multi, ctx->list_cap, P, ctx->cfg.max_llf, ctx->dcnt, ctx->dcnt->n_ppint_items, icap);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
int do_pp_ext(cubep3m_b200_ctx* ctx, float a_mid, float dt, float mass_p) {
  pp::PPParams P;
  P.mass_p = mass_p; P.rsoft = ctx->cfg.rsoft; P.pp_bias = ctx->cfg.pp_bias; P.a_mid = a_mid; P.G = ctx->cfg.G; P.dt = dt;
  P.cutoff = (float)ctx->cfg.nf_cutoff;
  if (ctx->cfg.pp_ext && ctx->cfg.pp_range > 0) {
    P.apply = ctx->cfg.pp_ext_force_flag;
    ctx->ppext_blocks = 0;
    if (ctx->np_all > 0 && ctx->ppext_mode == 1 && ctx->cfg.pp_range <= pp::TB_HALO) {
      const int nc = ctx->d.nc_node;
      const int nbx = (nc + pp::TB_X - 1) / pp::TB_X, nby = (nc + pp::TB_Y - 1) / pp::TB_Y, nbz = (nc + pp::TB_Z - 1) / pp::TB_Z;
      ctx->ppext_blocks = nbx * nby * nbz;
      if (ctx->cfg.pp_range == 2)
        LAUNCH(ctx, KC_PPEXT, pp::ppext_tiled_kernel<2>, ctx->ppext_blocks, pp::TB_NT, pp::TB_SMEM, ctx->xv[ctx->cur], ctx->fstart, ctx->d.H, ctx->d.b, ctx->d.nc_buf, nc, nbx, nby,
               2, P, ctx->dcnt, &ctx->dcnt->n_ppext_fallback, ctx->ppext_ovf);
      else
        LAUNCH(ctx, KC_PPEXT, pp::ppext_tiled_kernel<-1>, ctx->ppext_blocks, pp::TB_NT, pp::TB_SMEM, ctx->xv[ctx->cur], ctx->fstart, ctx->d.H, ctx->d.b, ctx->d.nc_buf, nc, nbx, nby,
               ctx->cfg.pp_range, P, ctx->dcnt, &ctx->dcnt->n_ppext_fallback, ctx->ppext_ovf);
      // dense blocks (overflow list): (cell, chunk) items -> one warp per item; the direct walk only if the item list overflowed
      const int icap = ctx->ppext_cell_mode ? ctx->ppext_item_cap : 0;
      CK(cudaMemsetAsync(&ctx->dcnt->n_ppext_items, 0, 2 * sizeof(int), ctx->stream));
      if (icap > 0) {
        LAUNCH(ctx, KC_PPEXT, pp::ppext_items_kernel, std::min(ctx->ppext_blocks, NUM_SMS * 8), pp::TB_NT, 0, ctx->fstart, ctx->d.H, ctx->d.nc_buf, nc, nbx, nby,
               &ctx->dcnt->n_ppext_fallback, ctx->ppext_ovf, ctx->ppext_items, icap, &ctx->dcnt->n_ppext_items);
        if (ctx->ppext_dense_tma)
          LAUNCH(ctx, KC_PPEXT_DENSE, pp::ppext_cell_tma_kernel, NUM_SMS * 8, pp::TB_NT, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->d.H, ctx->cfg.pp_range, P, ctx->dcnt, ctx->ppext_items,
                 icap, &ctx->dcnt->n_ppext_items, &ctx->dcnt->ppext_ticket);
        else
          LAUNCH(ctx, KC_PPEXT_DENSE, pp::ppext_cell_kernel, NUM_SMS * 16, pp::TB_NT, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->d.H, ctx->cfg.pp_range, P, ctx->dcnt, ctx->ppext_items, icap,
                 &ctx->dcnt->n_ppext_items, &ctx->dcnt->ppext_ticket);
      }
      LAUNCH(ctx, KC_PPEXT_DENSE, pp::ppext_blocklist_kernel, std::min(ctx->ppext_blocks, NUM_SMS * 8), pp::TB_NT, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->d.H, ctx->d.b, ctx->d.nc_buf, nc,
             nbx, nby, ctx->cfg.pp_range, P, ctx->dcnt, &ctx->dcnt->n_ppext_fallback, ctx->ppext_ovf, &ctx->dcnt->n_ppext_items, icap);
    } else if (ctx->np_all > 0) {
      LAUNCH(ctx, KC_PPEXT, pp::ppext_kernel, (ctx->np_all + pp::EXT_TPB - 1) / pp::EXT_TPB, pp::EXT_TPB, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->np_all, ctx->d.H,
             ctx->d.b, ctx->d.nc_buf, ctx->d.nc_node, ctx->cfg.pp_range, P, ctx->dcnt);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

// :617 takes the maximum over the margin particles' partial sums as well (limiter only, no kick). Needs only the sorted positions and the cell
// table: particle_mesh runs it on the coarse stream behind the coarse solve, where it fills the SM slots the bandwidth-bound fine-tile kernels
// leave free instead of adding 2.6 ms (512^3) behind the PP_EXT kick kernels.
int do_pp_ext_margin(cubep3m_b200_ctx* ctx, float a_mid, float dt, float mass_p) {
  if (!(ctx->cfg.pp_ext && ctx->cfg.pp_range > 0 && ctx->np_all > 0 && ctx->ppext_margin_max)) return 0;
  pp::PPParams P;
  P.mass_p = mass_p; P.rsoft = ctx->cfg.rsoft; P.pp_bias = ctx->cfg.pp_bias; P.a_mid = a_mid; P.G = ctx->cfg.G; P.dt = dt;
  P.cutoff = (float)ctx->cfg.nf_cutoff; P.apply = 0;
  const pp::MarginGeom G{ctx->d.H, ctx->d.b, ctx->d.m, ctx->d.T, ctx->cfg.pp_range};
  if (!ctx->roles_listed) {       // normally the scatter of this step's cell sort has listed the roles already
    CK(cudaMemsetAsync(&ctx->dcnt->n_margin_roles, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, KC_PPEXT_MARGIN, pp::ppext_margin_list_kernel, (ctx->np_all + pp::EXT_TPB - 1) / pp::EXT_TPB, pp::EXT_TPB, 0, ctx->xv[ctx->cur], ctx->np_all, G, ctx->margin_roles,
           ctx->margin_cap, &ctx->dcnt->n_margin_roles);
  }
  LAUNCH(ctx, KC_PPEXT_MARGIN, pp::ppext_margin_roles_kernel, NUM_SMS * 16, pp::EXT_TPB, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->np_all, G, ctx->margin_roles, ctx->margin_cap,
         &ctx->dcnt->n_margin_roles, P, ctx->dcnt);
  CK(cudaGetLastError());
  return 0;
}

// coarse_mesh.f90 for nodes_dim = 1 (whole coarse mesh on this GPU)
int do_coarse_mass(cubep3m_b200_ctx* ctx, float mass_p) {
  const Dims& d = ctx->d;
  const size_t nrc = (size_t)d.nc_node * d.nc_node * d.nc_node;
  CK(cudaMemsetAsync(ctx->rho_c, 0, nrc * sizeof(float), ctx->stream));
  const size_t win = (size_t)9 * (d.nc_node + 4) * sizeof(float);
  static const bool plain = [] { const char* e = getenv("CUBEP3M_B200_CICMASS"); return e && !strcmp(e, "global"); }();   // A/B knob: one global atomic per contribution
  if (!plain && win <= 48 * 1024)
    LAUNCH(ctx, KC_CIC_MASS, coarse::cic_mass_smem_kernel, (d.nc_node + 2) * (d.nc_node + 2), coarse::TPB, win, ctx->xv[ctx->cur], ctx->fstart, ctx->rho_c, d.H, d.nc_buf,
           d.nc_node, mass_p, ctx->cfg.coarse_ngp);
  else
    LAUNCH(ctx, KC_CIC_MASS, coarse::cic_mass_kernel, (d.nc_node + 2) * (d.nc_node + 2), coarse::TPB, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->rho_c, d.H, d.nc_buf,
           d.nc_node, mass_p, ctx->cfg.coarse_ngp);
  CK(cudaGetLastError());
  return 0;
}
int do_coarse_force_slab(cubep3m_b200_ctx* ctx);
int do_coarse_force(cubep3m_b200_ctx* ctx) {
  // coarse_force.f90:18-90 with the cube<->slab repack + distributed FFT of fft_coarse.f90 replaced by: all-gather the ranks' rho_c
  // cubes over NVLink, solve the WHOLE (Nx,Ny,Nz) coarse mesh on every GPU (<= 0.5 GB even for 8 x 512^3 particles), and gather
  // this rank's cube plus its one-cell halo from the periodic result (which is what coarse_force_buffer.f90:23-63 exchanges).
  if (ctx->coarse_mode == 1) return do_coarse_force_slab(ctx);
  const Dims& d = ctx->d;
  const int Nx = d.Nc[0], Ny = d.Nc[1], Nz = d.Nc[2], nc = d.nc_node;
  const long long nrc = (long long)nc * nc * nc;
  const float* cubes = ctx->rho_c;
  if (d.world > 1) {
#ifdef CUBEP3M_WITH_NCCL
    NCK(ncclAllGather(ctx->rho_c, ctx->gather, (size_t)nrc, ncclFloat, ctx->comm, ctx->stream));
    cubes = ctx->gather;
#else
    return CUBEP3M_B200_ENCCL;
#endif
  }
  for (int r = 0; r < d.world; ++r) {
    const int rx = r % d.Dg[0], ry = (r / d.Dg[0]) % d.Dg[1], rz = r / (d.Dg[0] * d.Dg[1]);
    const float* cube = (d.world > 1) ? cubes + (size_t)r * nrc : cubes;
    LAUNCH(ctx, KC_COARSE_MISC, coarse::cube_to_slab_kernel, grid_for(nrc, coarse::TPB), coarse::TPB, 0, cube, ctx->slab, nc, Nx, Ny, rx * nc, ry * nc,
           rz * nc, (r == ctx->cfg.rank) ? &ctx->dcnt->sum_rho_c : nullptr);
  }
  ctx->fft_class_base = KC_COARSE_FFT;
  const fftk::Mesh3 g = coarse_mesh3(ctx);
  if (int st = fftk::forward3d(ctx, g, ctx->slab)) return st;                     // coarse_force.f90:18
  const float scale = 1.0f / (((float)Nx * (float)Ny) * (float)Nz);                 // fft_coarse.f90:186
  const int lo[3] = {0, 0, 0}, cnt[3] = {Nx, Ny, Nz};
  const size_t ncs = (size_t)(Nx / 2 + 1) * Ny * Nz;
  const long long nfc = (long long)(nc + 2) * (nc + 2) * (nc + 2);
  for (int comp = 0; comp < 3; ++comp) {                                            // coarse_force.f90:37-90
    if (int st = fftk::backward3d(ctx, g, ctx->slab, ctx->slab_g, ctx->kern_c + (size_t)comp * ncs, ctx->creal, lo, cnt, Nx, Ny, scale)) return st;
    LAUNCH(ctx, KC_COARSE_MISC, coarse::extract_force_kernel, grid_for(nfc, coarse::TPB), coarse::TPB, 0, ctx->creal, Nx, Ny, Nz, nc, d.coord[0], d.coord[1],
           d.coord[2], ctx->force_c, comp);
  }
  ctx->fft_class_base = 0;
  LAUNCH(ctx, KC_COARSE_MISC, coarse::force_max_kernel, grid_for(nrc, coarse::TPB), coarse::TPB, 0, ctx->force_c, nc, &ctx->dcnt->c_force_max_bits);
  CK(cudaGetLastError());
  return 0;
}

// coarse_force.f90:18-90 on the slab decomposition (coarse_slab.cuh): everything is enqueued on ctx->stream, nothing synchronises the host
int do_coarse_force_slab(cubep3m_b200_ctx* ctx) {
  const Dims& d = ctx->d;
  const int W = d.world, me = ctx->cfg.rank, Nx = d.Nc[0], Ny = d.Nc[1], Nz = d.Nc[2], nc = d.nc_node, hc = Nx / 2 + 1, zs = ctx->cs_zs, ys = ctx->cs_ys;
  const long long nrc = (long long)nc * nc * nc;
  const int ep = (int)++ctx->cs_epoch;
  const cslab::Peers& P = ctx->cs_peers;
  const int* mail = reinterpret_cast<const int*>(ctx->cs_xchg + ctx->cs_off_mail);
  const long long tmo = 60000000000LL;                 // ~30 s of spinning before a wait gives up (ranks may enter a step seconds apart)
  auto sig = [&](int ph) { LAUNCH(ctx, KC_COARSE_MISC, cslab::signal_kernel, 1, 32, 0, P, (long long)ctx->cs_off_mail, W, me, ph, ep); };
  auto wait = [&](int ph) { LAUNCH(ctx, KC_COARSE_XCHG, cslab::wait_kernel, 1, 32, 0, mail, W, ph, ep, &ctx->dcnt->xchg_timeout, tmo); };
  float* slab = ctx->cs_xchg + ctx->cs_off_slab;
  float2* T = reinterpret_cast<float2*>(ctx->cs_xchg + ctx->cs_off_T);
  float2* G = reinterpret_cast<float2*>(ctx->cs_G);
  ctx->fft_class_base = KC_COARSE_FFT;
  // 1. pack_slab (fftw3ds.f90:4-54)
  LAUNCH(ctx, KC_COARSE_XCHG, cslab::scatter_cube_kernel, grid_for(nrc, cslab::TPB), cslab::TPB, 0, ctx->rho_c, P, (long long)ctx->cs_off_slab, nc, zs, Nx, Ny, d.coord[0],
         d.coord[1], d.coord[2], &ctx->dcnt->sum_rho_c);
  sig(cslab::PH_CUBE); wait(cslab::PH_CUBE);
  // 2. slab: r2c along x, forward y
  if (int st = fftk::launch_x_r2c(ctx, KC_COARSE_FFT, Nx, slab, Ny * zs, ctx->tw_c[0])) return st;
  float2* cs = reinterpret_cast<float2*>(slab);
  if (int st = fftk::launch_strided(ctx, KC_COARSE_FFT, Ny, false, cs, cs, hc, (long long)hc, (long long)Ny * hc, 0, zs, nullptr, 0, 0, 0, Ny - 1, ctx->tw_c[1])) return st;
  // 3. transpose to y-pencils
  const dim3 tgrid((unsigned)std::max(1, std::min(64, (ys * hc + cslab::TPB - 1) / cslab::TPB)), (unsigned)(W * zs));
  LAUNCH(ctx, KC_COARSE_XCHG, cslab::transpose_kernel<true>, tgrid, cslab::TPB, 0, cs, P, (long long)ctx->cs_off_T, W, me, zs, ys, Ny, hc);
  sig(cslab::PH_FWD); wait(cslab::PH_FWD);
  // 4. forward z on the pencils; per component: x i*kern_c fused into the inverse z pass; 5. transpose back
  if (int st = fftk::launch_strided(ctx, KC_COARSE_FFT, Nz, false, T, T, hc, (long long)ys * hc, (long long)hc, 0, ys, nullptr, 0, 0, 0, Nz - 1, ctx->tw_c[2])) return st;
  const long long krow = (long long)Nz * ys * hc;
  for (int comp = 0; comp < 3; ++comp) {                                            // coarse_force.f90:37-90
    if (int st = fftk::launch_strided(ctx, KC_COARSE_FFT, Nz, true, T, G, hc, (long long)ys * hc, (long long)hc, 0, ys, ctx->cs_kern_rows + comp * krow, (long long)ys * hc,
                                      (long long)hc, 0, Nz - 1, ctx->tw_c[2])) return st;
    LAUNCH(ctx, KC_COARSE_XCHG, cslab::transpose_kernel<false>, tgrid, cslab::TPB, 0, G, P, (long long)ctx->cs_off_back[comp], W, me, zs, ys, Ny, hc);
    sig(cslab::PH_BWD0 + comp);
  }
  // 6. slab: inverse y, c2r along x, 1/(Nx Ny Nz)  (fftw3ds.f90:161)
  const float scale = 1.0f / (((float)Nx * (float)Ny) * (float)Nz);
  for (int comp = 0; comp < 3; ++comp) {
    wait(cslab::PH_BWD0 + comp);
    float2* b = reinterpret_cast<float2*>(ctx->cs_xchg + ctx->cs_off_back[comp]);
    if (int st = fftk::launch_strided(ctx, KC_COARSE_FFT, Ny, true, b, b, hc, (long long)hc, (long long)Ny * hc, 0, zs, nullptr, 0, 0, 0, Ny - 1, ctx->tw_c[1])) return st;
    if (int st = fftk::launch_x_c2r(ctx, KC_COARSE_FFT, Nx, b, ctx->cs_real3 + (size_t)comp * zs * Ny * Nx, 0, Nx, 0, Ny, 0, zs, Ny, (long long)Nx, (long long)Ny, scale,
                                    ctx->tw_c[0])) return st;
  }
  // 7. unpack_slab (fftw3ds.f90:56-101) + coarse_force_buffer.f90:23-63
  const int fc = nc + 2;
  LAUNCH(ctx, KC_COARSE_XCHG, cslab::scatter_halo_kernel, dim3((unsigned)std::min(16, (fc * fc + cslab::TPB - 1) / cslab::TPB), (unsigned)(W * fc)), cslab::TPB, 0, ctx->cs_real3, P,
         (long long)ctx->cs_off_force, me, nc, zs, Nx, Ny, Nz, d.Dg[0], d.Dg[1]);
  sig(cslab::PH_HALO); wait(cslab::PH_HALO);
  ctx->fft_class_base = 0;
  LAUNCH(ctx, KC_COARSE_MISC, coarse::force_max_kernel, grid_for(nrc, coarse::TPB), coarse::TPB, 0, ctx->force_c, nc, &ctx->dcnt->c_force_max_bits);
  CK(cudaGetLastError());
  return 0;
}
int do_coarse_vel(cubep3m_b200_ctx* ctx, float a_mid, float dt) {
  const Dims& d = ctx->d;
  LAUNCH(ctx, KC_CIC_KICK, coarse::cic_kick_kernel, d.nc_node * d.nc_node, coarse::TPB, 0, ctx->xv[ctx->cur], ctx->fstart, ctx->force_c, d.H, d.nc_buf, d.nc_node,
         a_mid, ctx->cfg.G, dt, ctx->cfg.coarse_ngp);
  CK(cudaGetLastError());
  return 0;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }

}  // namespace

// =====================================================================================================

// ---------------------------------------------------------------- driver twin: timestep.f90 (host and device)
__host__ __device__ inline void expansion_core(float a0, float dt0, float omega_m, float omega_l, float wde, float* da1, float* da2) {
  // timestep.f90:241-293: two half steps of a 3rd-order Taylor expansion, real(8) arithmetic from real(4) arguments
  const float dt_x = dt0 / 2;
  double a_x = a0;
  for (int half = 0; half < 2; ++half) {
    const double omHsq = 4.0 / 9.0;
    const double a3rlm = pow(a_x, (double)(-3 * wde)) * omega_l / omega_m;
    const double arkm = a_x * (1.0 - omega_m - omega_l) / omega_m;
    const double adot = sqrt(omHsq * a_x * a_x * a_x * (1.0 + arkm + a3rlm));
    const double addot = a_x * a_x * omHsq * (1.5 + 2.0 * arkm + 1.5 * (1.0 - wde) * a3rlm);
    const double atdot = a_x * adot * omHsq * (3.0 + 6.0 * arkm + 1.5 * (2.0 - 3.0 * wde) * (1.0 - wde) * a3rlm);
    const float da = (float)(adot * dt_x + (addot * (double)dt_x * dt_x) / 2.0 + (atdot * (double)dt_x * dt_x * dt_x) / 6.0);
    if (half == 0) { *da1 = da; a_x = (double)(a0 + da); } else *da2 = da;
  }
}
__host__ __device__ inline void timestep_core(cubep3m_b200_clock* c) {
  // timestep.f90:20-196 (cosmo branch, dark matter only); dt_max = 1, ra_max = 0.01, dt_scale = 1 (cubepm.par:26-29)
  const float dt_max = 1.0f, ra_max = 0.01f, dt_scale = 1.0f;
  c->nts += 1;
  if (c->nts != 1) c->dt_old = c->dt;
  float da_1 = 0, da_2 = 0;
  if (c->cosmo) {
    float dt_e = dt_max;
    int n = 0;
    for (;;) {
      n++;
      expansion_core(c->a, dt_e, c->omega_m, c->omega_l, c->wde, &da_1, &da_2);
      c->da = da_1 + da_2;
      const float ra = c->da / (c->a + c->da);
      if (ra > ra_max) dt_e = dt_e * (ra_max / ra); else break;
      if (n > 10) break;
    }
    float dt = fminf(dt_e, fminf(c->dt_f_acc, c->dt_c_acc));
    if (c->ppint) dt = fminf(dt, c->dt_pp_acc);
    if (c->ppint && c->pp_ext) dt = fminf(dt, c->dt_pp_ext_acc);
    dt = dt * dt_scale;
    expansion_core(c->a, dt, c->omega_m, c->omega_l, c->wde, &da_1, &da_2);
    c->da = da_1 + da_2;
    c->checkpoint_step = 0;
    if (c->a + c->da > c->a_target) {                 // timestep.f90:128-137
      c->checkpoint_step = 1;
      dt = dt * (c->a_target - c->a) / c->da;
      expansion_core(c->a, dt, c->omega_m, c->omega_l, c->wde, &da_1, &da_2);
    }
    c->da = da_1 + da_2;
    c->a_mid = c->a + (c->da / 2);
    c->dt = dt;
    c->tau += dt; c->t += dt; c->a += c->da;
  } else {                                            // timestep.f90:198-221
    c->a = 1.0f; c->a_mid = 1.0f; c->da = 0.f;
    float dt = fminf(1.0f, fminf(c->dt_f_acc, c->dt_c_acc));
    if (c->ppint) dt = fminf(dt, c->dt_pp_acc);
    if (c->ppint && c->pp_ext) dt = fminf(dt, c->dt_pp_ext_acc);
    c->dt = dt; c->t += dt;
  }
}
__global__ void timestep_kernel(cubep3m_b200_clock* c) { if (threadIdx.x == 0 && blockIdx.x == 0) timestep_core(c); }

extern "C" {

const char* cubep3m_b200_version(void) { return "cubep3m_b200 0.1 (sm_100a, hand-written CUDA, no CPU fallback)"; }

const char* cubep3m_b200_strerror(int st) {
  switch (st) {
    case 0: return "ok";
    case CUBEP3M_B200_EINVAL: return "bad configuration";
    case CUBEP3M_B200_ECUDA: return "CUDA failure";
    case CUBEP3M_B200_EPASSBUF: return "not enough buffer space in pass";
    case CUBEP3M_B200_EMAXNP: return "exceeded max_np in pass";
    case CUBEP3M_B200_EMAXLLF: return "exceeded max_llf";
    case CUBEP3M_B200_ENCCL: return "NCCL failure";
    case CUBEP3M_B200_ENOTREADY: return "call order violated";
    case CUBEP3M_B200_ECAPACITY: return "internal work list overflow";
  }
  return "unknown";
}

void cubep3m_b200_default_config(cubep3m_b200_config* c) {
  memset(c, 0, sizeof(*c));
  c->nodes_dim = 1; c->tiles_node_dim = 2; c->nf_tile = 176; c->nf_buf = 24; c->nf_cutoff = 16; c->mesh_scale = 4;
  c->pp_range = 2; c->max_np = 0; c->max_buf = 0; c->max_llf = 100000;
  c->density_buffer = 2.0f; c->rsoft = 0.1f; c->pp_bias = 1.0f; c->dt_pp_scale = 0.05f;
  const float pi = 3.141592654f;
  c->G = 1.0f / 6.0f / pi;
  c->eps = 1.0e-3f;
  c->ngp = 1; c->ppint = 1; c->pp_ext = 0; c->coarse_ngp = 0; c->pid = 0; c->lrckcorr = 1; c->move_grid_back = 0;
  c->ngp_fmesh_force = c->pp_force_flag = c->pp_ext_force_flag = c->coarse_vel_update = 1;
  c->rank = 0; c->local_gpu = 0;
  c->nodes_dim_xyz[0] = c->nodes_dim_xyz[1] = c->nodes_dim_xyz[2] = 0;
}

int cubep3m_b200_get_unique_id(void* id128) {
#ifdef CUBEP3M_WITH_NCCL
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return CUBEP3M_B200_ENCCL;
  memcpy(id128, &id, sizeof(id));
  return 0;
#else
  (void)id128;
  return CUBEP3M_B200_ENCCL;
#endif
}

int cubep3m_b200_finalize(cubep3m_b200_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  auto F = [](void* p) { if (p) cudaFree(p); };
  for (int i = 0; i < 2; ++i) { F(ctx->xv[i]); F(ctx->pid[i]); F(ctx->sendbuf[i]); F(ctx->sendpid[i]); F(ctx->recvbuf_own[i]); F(ctx->recvpid_own[i]); }
  for (void* q : ctx->ipc_opened) cudaIpcCloseMemHandle(q);
  cs_free(ctx);
  if (ctx->mailbox) cudaFree(ctx->mailbox);
  if (ctx->hbox) cudaFreeHost(ctx->hbox);
#ifdef CUBEP3M_WITH_NCCL
  if (ctx->comm) ncclCommDestroy(ctx->comm);
#endif
  F(ctx->cand); F(ctx->deltas); F(ctx->ndelta); F(ctx->tile_counts); F(ctx->key); F(ctx->rank); F(ctx->blist); F(ctx->fstart); F(ctx->fcur); F(ctx->blocksum); F(ctx->scan_status); F(ctx->multi_list); F(ctx->occ_list); F(ctx->rowoff);
  F(ctx->dclock); F(ctx->kern_f); F(ctx->tile_rho); F(ctx->tile_g); F(ctx->force_f[0]); 
  for (int q = 1; q < cubep3m_b200_ctx::MAX_TILE_STREAMS; ++q) {
    F(ctx->tile_rho_s[q]); F(ctx->tile_g_s[q]); F(ctx->force_f_s[q]);
    if (ctx->stream_aux[q]) cudaStreamDestroy(ctx->stream_aux[q]);
    if (ctx->ev_join[q]) cudaEventDestroy(ctx->ev_join[q]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->stream_coarse) cudaStreamDestroy(ctx->stream_coarse);
  F(ctx->tw_f); F(ctx->kern_c); F(ctx->rho_c); F(ctx->slab); F(ctx->slab_g); F(ctx->creal); F(ctx->gather); F(ctx->force_c); F(ctx->redbuf); F(ctx->cntbuf); F(ctx->dcnt); F(ctx->ppext_ovf); F(ctx->margin_roles); F(ctx->ppext_items); F(ctx->ppint_items);
  for (int a = 0; a < 3; ++a) { bool dup = false; for (int b2 = 0; b2 < a; ++b2) dup |= (ctx->tw_c[b2] == ctx->tw_c[a]); if (!dup) F(ctx->tw_c[a]); }
  if (ctx->hcnt) cudaFreeHost(ctx->hcnt);
  if (ctx->ev_ok) for (auto& e : ctx->ev) cudaEventDestroy(e);
  for (auto& e : ctx->prof_ev) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

int cubep3m_b200_init(const cubep3m_b200_config* cfg, const float* fine_table, const float* coarse_table, const float* kern_f,
                      const float* kern_c, const void* nccl_unique_id, int world_size, cubep3m_b200_ctx** out) {
  if (!cfg || !out) return CUBEP3M_B200_EINVAL;
  Dims d;
  if (int st = derive(*cfg, d)) return st;
  if (d.world > 1 && (!nccl_unique_id || world_size != d.world)) return CUBEP3M_B200_EINVAL;
  if ((!kern_f && !fine_table) || (!kern_c && !coarse_table)) return CUBEP3M_B200_EINVAL;
  if (cfg->move_grid_back) return CUBEP3M_B200_EINVAL;   // the in-step shift of particle_mesh_threaded.f90:714-716 is not supported (out of scope); the
                                                         // driver-level call (cubepm.f90:179) is cubep3m_b200_move_grid_back
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1 || cfg->local_gpu >= ndev) return CUBEP3M_B200_ECUDA;
  CK(cudaSetDevice(cfg->local_gpu));
  cubep3m_b200_ctx* ctx = new cubep3m_b200_ctx();
  ctx->cfg = *cfg; ctx->d = d; ctx->device = cfg->local_gpu; ctx->world = world_size > 0 ? world_size : 1;
  if (fine_table) ctx->fine_table.assign(fine_table, fine_table + 16 * 16 * 16 * 3);
  if (coarse_table) ctx->coarse_table.assign(coarse_table, coarse_table + 4 * 4 * 4 * 3);
#define TRY(x) do { int st__ = (x); if (st__) { cubep3m_b200_finalize(ctx); return st__; } } while (0)
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CUBEP3M_B200_ECUDA; }
  for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  ctx->ev_ok = true;
  TRY(fftk::init_constants());
  for (int i = 0; i < 2; ++i) {
    TRY(dmalloc(&ctx->xv[i], (size_t)6 * d.max_np + 64));   // + slack: bulk copies of the PP kernels round their last chunk up to 16 bytes
    if (cfg->pid) TRY(dmalloc(&ctx->pid[i], (size_t)d.max_np));
    TRY(dmalloc(&ctx->sendbuf[i], (size_t)d.max_buf));
    if (cfg->pid) TRY(dmalloc(&ctx->sendpid[i], (size_t)d.max_buf / 6 + 1));
  }
  TRY(dmalloc(&ctx->key, (size_t)d.max_np));
  TRY(dmalloc(&ctx->rank, (size_t)d.max_np));
  TRY(dmalloc(&ctx->blist, (size_t)d.max_np));
  ctx->cand_cap = std::max(4096, d.max_np / 64);
  TRY(dmalloc(&ctx->cand, (size_t)3 * ctx->cand_cap));
  TRY(dmalloc(&ctx->deltas, (size_t)d.tiles_node * fine::DELTA_CAP));
  TRY(dmalloc(&ctx->ndelta, (size_t)d.tiles_node));
  TRY(dmalloc(&ctx->tile_counts, (size_t)d.tiles_node));
  TRY(dmalloc(&ctx->fstart, (size_t)d.NF + 64));
  TRY(dmalloc(&ctx->fcur, (size_t)d.NF / 2 + 64));
  ctx->nblocksum = (int)((d.NF + part::SCAN_BLOCK - 1) / part::SCAN_BLOCK);
  TRY(dmalloc(&ctx->blocksum, (size_t)ctx->nblocksum + 1));
  TRY(dmalloc(&ctx->scan_status, (size_t)ctx->nblocksum + 1));
  {
    // "1pass": single-pass scan with decoupled look-back. Measured on B200 it is SLOWER than the three-kernel scan (0.59 vs 0.36 ms over
    // 175 M cells, 4.0 vs 2.9 ms over 1.2 G cells): with 4096-cell tiles the 43 k-long prefix chain, not the second read of the 16-bit histogram
    // it saves, sets the pace. Kept as an option (bit-identical results, tests/test_gpu_parity.py::test_scan_variants_agree); default: three kernels.
    const char* e = getenv("CUBEP3M_B200_SCAN");
    ctx->scan_onepass = e && !strcmp(e, "1pass");
  }
  {   // A/B: how the sort's cell histogram returns to zero: a memset after the scan (default), a low-footprint clearing kernel under the PP_EXT kernels
      // (=under), or a plain store per particle in the scatter (=scatter; measured: the partial-sector stores cost the scatter 1.5 ms at 512^3)
    const char* e = getenv("CUBEP3M_B200_HISTZERO");
    ctx->hist_mode = (e && !strcmp(e, "scatter")) ? 1 : (e && !strcmp(e, "countdown")) ? 2 : 0;
  }
  {   // A/B: "direct" = round 1's per-target / per-cell kernels, "tma" = the cell kernel with bulk-copy staging of the long source ranges
    const char* e = getenv("CUBEP3M_B200_PPEXT_DENSE");
    ctx->ppext_cell_mode = !(e && !strcmp(e, "direct"));
    ctx->ppext_dense_tma = e && !strcmp(e, "tma");
  }
  ctx->list_cap = d.max_np / 2 + 1024;
  if (cfg->ppint) {
    TRY(dmalloc(&ctx->multi_list, (size_t)ctx->list_cap));
    ctx->ppint_item_cap = d.max_np / 4 + 4096;
    TRY(dmalloc(&ctx->ppint_items, (size_t)ctx->ppint_item_cap));
  }
  if (cfg->pp_ext) {
    const int nc = ctx->d.nc_node;
    TRY(dmalloc(&ctx->ppext_ovf, (size_t)((nc + pp::TB_X - 1) / pp::TB_X) * ((nc + pp::TB_Y - 1) / pp::TB_Y) * ((nc + pp::TB_Z - 1) / pp::TB_Z)));
    ctx->ppext_item_cap = d.max_np / 8 + 4096;
    TRY(dmalloc(&ctx->ppext_items, (size_t)ctx->ppext_item_cap));

    ctx->margin_cap = d.max_np / 8 + 4096;
    TRY(dmalloc(&ctx->margin_roles, (size_t)ctx->margin_cap));
  }
  const size_t rowoff_n = (size_t)d.nc_node * d.nc_node + 16 + d.tiles_node;
  TRY(dmalloc(&ctx->rowoff, rowoff_n));
  if (cudaMemset(ctx->rowoff, 0, rowoff_n * sizeof(int)) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  // spectra of the fused NGP path use a row pitch of a multiple of 16 complex (128-byte aligned 16-column blocks): size for it
  const size_t tile_elems = (size_t)2 * ((d.hc + 15) / 16 * 16) * d.n * d.n;
  TRY(dmalloc(&ctx->tile_rho, tile_elems));
  TRY(dmalloc(&ctx->tile_g, 3 * tile_elems));
  if (cudaMemset(ctx->tile_rho, 0, tile_elems * sizeof(float)) != cudaSuccess || cudaMemset(ctx->tile_g, 0, 3 * tile_elems * sizeof(float)) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  TRY(dmalloc(&ctx->force_f[0], (size_t)3 * d.fdim * d.fdim * d.fdim));
  ctx->force_f[1] = ctx->force_f[0] + (size_t)d.fdim * d.fdim * d.fdim; ctx->force_f[2] = ctx->force_f[1] + (size_t)d.fdim * d.fdim * d.fdim;
  {
    // tuning knob: fine tiles in flight on separate streams / buffer sets. Two: the second tile's kernels fill the tails of the first tile's
    // persistent grids (measured at 256^3 particles with the session-5 kernels: 6.16 ms/step with 1, 6.04 with 2; the session-4 kernels, which
    // were longer and thrashed the L2 with two 113 MB working sets, lost 8 % with 2). Per-kernel timings are only unambiguous with 1, which is
    // what bench.py switches to for its instrumented region (cubep3m_b200_set_tile_streams).
    const char* e = getenv("CUBEP3M_B200_TILE_STREAMS");
    ctx->tile_streams_max = e ? std::max(1, std::min(atoi(e), (int)cubep3m_b200_ctx::MAX_TILE_STREAMS)) : 2;
    ctx->tile_streams = ctx->tile_streams_max;
  }
  {
    const char* e = getenv("CUBEP3M_B200_PPEXT");       // "direct": the one-thread-per-target kernel (A/B measurements); default: tiled
    ctx->ppext_mode = (e && !strcmp(e, "direct")) ? 0 : 1;
    const char* em = getenv("CUBEP3M_B200_PPEXT_MARGIN");
    ctx->ppext_margin_max = !(em && atoi(em) == 0);
    if (cudaFuncSetAttribute(pp::ppext_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp::TB_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(pp::ppext_tiled_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp::TB_SMEM) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  }
  ctx->stream_main = ctx->stream;
  if (cudaStreamCreateWithFlags(&ctx->stream_coarse, cudaStreamNonBlocking) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  if (ctx->tile_streams > 1) {
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
    for (int q = 1; q < ctx->tile_streams; ++q) {
      if (cudaStreamCreateWithFlags(&ctx->stream_aux[q], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&ctx->ev_join[q], cudaEventDisableTiming) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
      TRY(dmalloc(&ctx->tile_rho_s[q], tile_elems));
      TRY(dmalloc(&ctx->tile_g_s[q], 3 * tile_elems));
      if (cudaMemset(ctx->tile_rho_s[q], 0, tile_elems * sizeof(float)) != cudaSuccess || cudaMemset(ctx->tile_g_s[q], 0, 3 * tile_elems * sizeof(float)) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
      TRY(dmalloc(&ctx->force_f_s[q], (size_t)3 * d.fdim * d.fdim * d.fdim));
    }
  }
  ctx->kf_pitch = (d.hc + 15) / 16 * 16;     // 16-column blocks of kern_f start on 64-byte boundaries (16-byte cp.async in fft_z_sandwich2)
  ctx->kf_stride = (long long)ctx->kf_pitch * d.n * d.n;
  TRY(dmalloc(&ctx->kern_f, (size_t)3 * ctx->kf_stride));
  if (cudaMemset(ctx->kern_f, 0, (size_t)3 * ctx->kf_stride * sizeof(float)) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  TRY(fftk::make_twiddles(d.n, &ctx->tw_f));
  const int Nx = d.Nc[0], Ny = d.Nc[1], Nz = d.Nc[2];
  const size_t ncs = (size_t)(Nx / 2 + 1) * Ny * Nz;
  const size_t nrc = (size_t)d.nc_node * d.nc_node * d.nc_node;
  TRY(dmalloc(&ctx->rho_c, nrc));
  {
    // coarse solve: one rank keeps the whole mesh on its GPU; several ranks use the slab decomposition over peer memory (coarse_slab.cuh).
    // CUBEP3M_B200_COARSE=slab forces the slab pipeline on one rank too (tests), =replicated the all-gather + replicated solve (A/B, and
    // the automatic fallback when the peers' memory cannot be mapped).
    const char* e = getenv("CUBEP3M_B200_COARSE");
    const bool want_slab = e ? !strcmp(e, "slab") : d.world > 1;
    ctx->coarse_mode = (want_slab && cs_supported(d)) ? 1 : 0;
  }
  if (ctx->coarse_mode == 1) TRY(cs_alloc(ctx));
  for (int a = 0; a < 3; ++a) {
    for (int b2 = 0; b2 < a; ++b2) if (d.Nc[b2] == d.Nc[a]) ctx->tw_c[a] = ctx->tw_c[b2];
    if (!ctx->tw_c[a]) TRY(fftk::make_twiddles(d.Nc[a], &ctx->tw_c[a]));
  }
  TRY(dmalloc(&ctx->redbuf, (size_t)64));
  TRY(dmalloc(&ctx->cntbuf, (size_t)8));
  if (d.world > 1) {
#ifdef CUBEP3M_WITH_NCCL
    for (int i = 0; i < 2; ++i) {
      TRY(dmalloc(&ctx->recvbuf_own[i], (size_t)d.max_buf));
      if (cfg->pid) TRY(dmalloc(&ctx->recvpid_own[i], (size_t)d.max_buf / 6 + 1));
    }
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    if (ncclCommInitRank(&ctx->comm, d.world, id, cfg->rank) != ncclSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ENCCL; }
    if (int st = p2p_init(ctx)) { cubep3m_b200_finalize(ctx); return st; }
    if (ctx->coarse_mode == 1) {
      bool ok = false;
      if (int st = cs_map_peers(ctx, &ok)) { cubep3m_b200_finalize(ctx); return st; }
      if (!ok) {
        if (cfg->rank == 0) fprintf(stderr, "cubep3m_b200: peer memory unavailable for the slab-decomposed coarse solve, using the all-gathered replicated solve\n");
        cs_free(ctx);
        ctx->coarse_mode = 0;
      }
    }
#else
    cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ENCCL;
#endif
  }
  TRY(dmalloc(&ctx->dcnt, 1));
  if (cudaMemset(ctx->dcnt, 0, sizeof(DevCounters)) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  if (cudaMallocHost((void**)&ctx->hcnt, sizeof(DevCounters)) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  memset(ctx->hcnt, 0, sizeof(DevCounters));
  if (kern_f) TRY(upload_interleaved(ctx->kern_f, kern_f, (size_t)ctx->kf_stride, 0, (size_t)d.hc * d.n * d.n, (size_t)d.hc, (size_t)ctx->kf_pitch));
  else TRY(build_kern_f(ctx));
  // Coarse Green's function. The whole-mesh solve keeps the global table [comp][z][y][kx]; the slab solve keeps only this rank's rows of it
  // (the global table and the FFT workspace that builds it are init-time temporaries then).
  TRY(dmalloc(&ctx->kern_c, 3 * ncs));
  TRY(dmalloc(&ctx->slab, (size_t)(Nx + 2) * Ny * Nz));
  if (ctx->coarse_mode == 0) {
    TRY(dmalloc(&ctx->slab_g, (size_t)(Nx + 2) * Ny * Nz));
    TRY(dmalloc(&ctx->creal, (size_t)(Nx + 2) * Ny * Nz));
    TRY(dmalloc(&ctx->force_c, (size_t)3 * (d.nc_node + 2) * (d.nc_node + 2) * (d.nc_node + 2)));
    if (d.world > 1) TRY(dmalloc(&ctx->gather, nrc * d.world));
  }
  if (kern_c && d.world == 1) TRY(upload_interleaved(ctx->kern_c, kern_c, ncs, 0, ncs));
  else if (kern_c && d.nc_slab > 0) {
    // the driver's kern_c(3,hc,nc_dim,nc_slab) is this rank's z-slab (cubep3m.fh:56; slabs are rank-major in z): all-gather the slabs per component
#ifdef CUBEP3M_WITH_NCCL
    const size_t per = (size_t)(Nx / 2 + 1) * Ny * d.nc_slab;
    TRY(upload_interleaved(ctx->kern_c, kern_c, ncs, per * cfg->rank, per));
    for (int comp = 0; comp < 3; ++comp)
      if (ncclAllGather(ctx->kern_c + comp * ncs + per * cfg->rank, ctx->kern_c + comp * ncs, per, ncclFloat, ctx->comm, ctx->stream) != ncclSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ENCCL; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
#endif
  } else { if (!coarse_table) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_EINVAL; } TRY(build_kern_c(ctx)); }
  // cudaMemcpy / cudaMemset above run on the legacy default stream, which the library's non-blocking streams do not wait for, and a pageable
  // host-to-device cudaMemcpy may return before its last DMA has landed: settle the device before any of the library's streams reads the tables
  // (an intermittent garbage row table of the slab solve — extracted right below from kern_c — was the symptom)
  if (cudaDeviceSynchronize() != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  if (ctx->coarse_mode == 1) {
    LAUNCH(ctx, KC_MISC, cslab::extract_rows_kernel, grid_for((long long)3 * Nz * ctx->cs_ys * (Nx / 2 + 1), cslab::TPB), cslab::TPB, 0, ctx->kern_c, ctx->cs_kern_rows, Nz, Ny,
           Nx / 2 + 1, ctx->cs_ys, cfg->rank * ctx->cs_ys);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
    if (d.world > 1) { cudaFree(ctx->kern_c); cudaFree(ctx->slab); ctx->kern_c = nullptr; ctx->slab = nullptr; }   // one rank keeps them for the debug getters / cic_power
  }
#undef TRY
  if (cudaDeviceSynchronize() != cudaSuccess) { cubep3m_b200_finalize(ctx); return CUBEP3M_B200_ECUDA; }
  *out = ctx;
  return 0;
}

int cubep3m_b200_upload_particles(cubep3m_b200_ctx* ctx, const float* xv, const int64_t* pid, int32_t np_local) {
  if (!ctx || np_local < 0 || np_local > ctx->d.max_np) return CUBEP3M_B200_EMAXNP;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->xv[ctx->cur], xv, sizeof(float) * 6 * (size_t)np_local, cudaMemcpyHostToDevice, ctx->stream));
  if (pid && ctx->cfg.pid) CK(cudaMemcpyAsync(ctx->pid[ctx->cur], pid, sizeof(int64_t) * (size_t)np_local, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->np_local = np_local; ctx->np_all = np_local; ctx->sorted = false; ctx->passed = false;
  return 0;
}

int cubep3m_b200_download_particles(cubep3m_b200_ctx* ctx, float* xv, int64_t* pid, int32_t* np_local) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const int np = ctx->passed ? ctx->np_all : ctx->np_local;
  if (xv) CK(cudaMemcpyAsync(xv, ctx->xv[ctx->cur], sizeof(float) * 6 * (size_t)np, cudaMemcpyDeviceToHost, ctx->stream));
  if (pid && ctx->cfg.pid) CK(cudaMemcpyAsync(pid, ctx->pid[ctx->cur], sizeof(int64_t) * (size_t)np, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (np_local) *np_local = np;
  return 0;
}

int cubep3m_b200_update_position(cubep3m_b200_ctx* ctx, float dt, float dt_old, const float offset[3]) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const float zero[3] = {0.f, 0.f, 0.f};
  if (int st = do_drift(ctx, dt, dt_old, offset ? offset : zero)) return st;
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int cubep3m_b200_move_grid_back(cubep3m_b200_ctx* ctx, const float s[3]) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  if (ctx->np_local > 0)
    LAUNCH(ctx, KC_DRIFT, part::shift_kernel, (ctx->np_local + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->xv[ctx->cur], ctx->np_local, s[0], s[1], s[2]);
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->sorted = false;
  return 0;
}

int cubep3m_b200_link_list(cubep3m_b200_ctx* ctx, int32_t* np_deleted) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  ctx->np_all = ctx->np_local; ctx->passed = false; ctx->keys_fused = false;
  CK(cudaMemsetAsync(&ctx->dcnt->overflow, 0, sizeof(int), ctx->stream));
  return do_sort(ctx, np_deleted);
}

int cubep3m_b200_particle_pass(cubep3m_b200_ctx* ctx, int32_t* np_with_ghosts) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  if (ctx->passed) return CUBEP3M_B200_ENOTREADY;
  int bufmax = 0;
  ctx->np_all = ctx->np_local;
  CK(cudaMemsetAsync(&ctx->dcnt->overflow, 0, sizeof(int), ctx->stream));
  if (int st = do_pass(ctx, &bufmax)) return st;
  if (int st = do_sort(ctx, nullptr)) return st;
  if (np_with_ghosts) *np_with_ghosts = ctx->np_all;
  // The stand-alone pass receives through NCCL into the same buffers a neighbour's NEXT particle_mesh packs into over peer memory. Inside a step the
  // end-of-step all-reduce keeps a fast rank from overwriting what its neighbour has not consumed; here (the halofind / projection sequence of
  // cubepm.f90:193-198 followed by the next step) the same guarantee needs a fence: nobody leaves particle_pass before everybody has unpacked
  // (the reference's blocking mpi_sendrecv_replace calls give it for free, particle_pass.f90:141-144).
  if (ctx->d.world > 1 && ctx->p2p) {
    float mx[1] = {0.f};
    double sm[1] = {0.0};
    if (int st = reduce_scalars(ctx, mx, 1, sm, 1)) return st;
  }
  return 0;
}

int cubep3m_b200_delete_particles(cubep3m_b200_ctx* ctx, int32_t* np_local) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemsetAsync(&ctx->dcnt->overflow, 0, sizeof(int), ctx->stream));
  if (!ctx->sorted) { ctx->keys_fused = false; if (int st = do_sort(ctx, nullptr)) return st; }
  if (int st = do_delete(ctx)) return st;
  if (np_local) *np_local = ctx->np_local;
  return 0;
}

int cubep3m_b200_particle_mesh(cubep3m_b200_ctx* ctx, float dt, float dt_old, float a_mid, float mass_p, const float offset[3],
                               cubep3m_b200_step_out* out) {
  if (!ctx || !out) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const Dims& d = ctx->d;
  const cubep3m_b200_config& c = ctx->cfg;
  memset(out, 0, sizeof(*out));
  const float zero[3] = {0.f, 0.f, 0.f};
  cudaEvent_t* ev = ctx->ev;
  // reset per-step device accumulators
  CK(cudaMemsetAsync(ctx->dcnt, 0, sizeof(DevCounters), ctx->stream));
  CK(cudaMemsetAsync(ctx->rowoff + d.nc_node * d.nc_node, 0, sizeof(int) * (16 + d.tiles_node), ctx->stream));
  ctx->prof_n = 0;
  CK(cudaEventRecord(ev[0], ctx->stream));
  // update_position (particle_mesh_threaded.f90:56) is fused into the first pack kernel of particle_pass (:63; range check of :61 folded in)
  const float* off3 = offset ? offset : zero;
  const float drift4[4] = {dt + dt_old, off3[0], off3[1], off3[2]};
  ctx->sorted = false; ctx->passed = false; ctx->np_all = ctx->np_local;
  CK(cudaEventRecord(ev[1], ctx->stream));
  int bufmax = 0, ndel = 0;
  ctx->p2p_epoch++;
  if (ctx->np_local > 0) { if (int st = do_pass(ctx, &bufmax, drift4, true)) return st; }
  else { if (int st = do_drift(ctx, dt, dt_old, off3)) return st; if (int st = do_pass(ctx, &bufmax, nullptr, true)) return st; }
  CK(cudaEventRecord(ev[2], ctx->stream));
  // CUBEP3M_B200_ROLES=scatter: the scatter lists the PP_EXT margin roles itself (it knows every particle's cell and sorted index) instead of a separate
  // listing kernel. Measured at 512^3: listing kernel -1.26 ms, scatter +1.94 ms (the role arithmetic lengthens a DRAM-latency-bound kernel). Off.
  static const bool roles_in_scatter = [] { const char* e = getenv("CUBEP3M_B200_ROLES"); return e && !strcmp(e, "scatter"); }();
  ctx->want_roles = roles_in_scatter;
  // A/B knob: CUBEP3M_B200_HISTZERO=under clears the histogram with a one-warp-per-CTA kernel under the PP_EXT kernels instead of a memset right after
  // the scan (measured at 512^3: 60.4-60.5 vs 60.1-60.4 ms/step — the slow clearing kernel costs PP_EXT what the memset costs the sort)
  static const bool zero_under = [] { const char* e = getenv("CUBEP3M_B200_HISTZERO"); return e && !strcmp(e, "under"); }();
  ctx->defer_hist_zero = ctx->cfg.pp_ext && ctx->cfg.pp_range > 0 && zero_under && ctx->hist_mode == 0;
  const int sort_st = do_sort(ctx, &ndel);                                                // :61 link_list as a cell sort
  ctx->defer_hist_zero = false;
  ctx->want_roles = false;
  if (sort_st) return sort_st;
  const int np_ghost = ctx->np_all;
  CK(cudaEventRecord(ev[3], ctx->stream));
  // INVARIANT of the overlapped streams below (coarse stream, two fine tiles in flight, PP after the tiles): kernels on other streams or CTAs
  // may READ a record's position words while a kick kernel stores the record's (z, vx) float2. The stored z is the value that was loaded
  // (bit-identical) and an aligned 8-byte store is not torn, so every reader sees the one valid position; velocities are only read by the
  // kernel that updates them. tests/test_gpu_parity_sizes.py::test_stream_overlap_does_not_change_positions pins this.
  // The coarse-mesh density and force solve only need the sorted positions: run them on their own stream, concurrently with
  // the fine-tile loop (the reference cannot: rho_f/rho_c and force_f/force_c are EQUIVALENCEd, cubep3m.fh:134-135).
  CK(cudaStreamWaitEvent(ctx->stream_coarse, ev[3], 0));
  ctx->stream = ctx->stream_coarse;
  CK(cudaEventRecord(ev[6], ctx->stream));
  int cst = do_coarse_mass(ctx, mass_p);                                                  // coarse_mesh.f90:28
  CK(cudaEventRecord(ev[7], ctx->stream));
  if (!cst) cst = do_coarse_force(ctx);                                                   // coarse_mesh.f90:84-100
  CK(cudaEventRecord(ev[8], ctx->stream));
  CK(cudaEventRecord(ev[13], ctx->stream));
  ctx->stream = ctx->stream_main;
  if (cst) return cst;
  if (int st = do_fine(ctx, a_mid, dt, mass_p, nullptr, nullptr)) return st;              // :84-368 (mesh part)
  CK(cudaEventRecord(ev[4], ctx->stream));
  if (int st = do_pp(ctx, a_mid, dt, mass_p)) return st;                                  // :274-361
  CK(cudaEventRecord(ev[5], ctx->stream));
  // the limiter part of PP_EXT (:617) only reads positions: with CUBEP3M_B200_MARGIN=overlap it runs on the (by now idle) coarse stream next to the
  // PP_EXT kick kernels instead of behind them. (Running it next to the fine tiles cost more than it hid: fine_fft 33.1 -> 37.9 ms at 512^3.)
  static const bool margin_overlap = [] { const char* e = getenv("CUBEP3M_B200_MARGIN"); return e && !strcmp(e, "overlap"); }();
  {
    // the sort's cell histogram is no longer needed (the scan was its last reader): clear it for the next step on the idle coarse stream, under the
    // issue-bound PP_EXT kernels that leave the DRAM idle (2.5 GB at 512^3 particles; the main stream joins ev[13] below)
    if (!ctx->hist_clean) {
      CK(cudaStreamWaitEvent(ctx->stream_coarse, ev[5], 0));
      ctx->stream = ctx->stream_coarse;
      LAUNCH(ctx, KC_MISC, part::zero_words_kernel, NUM_SMS * 2, 32, 0, reinterpret_cast<uint4*>(ctx->fcur), (long long)(d.NF / 8));   // NF / 2 words, NF % 64 == 0
      ctx->stream = ctx->stream_main;
      CK(cudaEventRecord(ev[13], ctx->stream_coarse));
      ctx->hist_clean = true;
    }
  }
  if (margin_overlap) {
    CK(cudaStreamWaitEvent(ctx->stream_coarse, ev[5], 0));
    ctx->stream = ctx->stream_coarse;
    const int mst = do_pp_ext_margin(ctx, a_mid, dt, mass_p);
    CK(cudaEventRecord(ev[13], ctx->stream));
    ctx->stream = ctx->stream_main;
    if (mst) return mst;
  }
  if (int st = do_pp_ext(ctx, a_mid, dt, mass_p)) return st;                              // :378-624
  if (!margin_overlap) { if (int st = do_pp_ext_margin(ctx, a_mid, dt, mass_p)) return st; }
  CK(cudaEventRecord(ev[11], ctx->stream));
  CK(cudaStreamWaitEvent(ctx->stream, ev[13], 0));                                        // join the coarse stream
  CK(cudaEventRecord(ev[12], ctx->stream));
  CK(cudaEventRecord(ev[9], ctx->stream));
  if (int st = fetch_counters(ctx)) return st;
  const DevCounters hc = *ctx->hcnt;
  ctx->ppext_fallback = hc.n_ppext_fallback;
  ctx->pairs_ppint = (long long)hc.pairs_ppint; ctx->pairs_ppext = (long long)hc.pairs_ppext;
  if (int st = overflow_status(&hc)) return st;
  if (hc.xchg_timeout) { fprintf(stderr, "cubep3m_b200: coarse-mesh exchange timed out waiting for a peer\n"); return CUBEP3M_B200_ENCCL; }
  // coarse_velocity (coarse_mesh.f90:106) rides on delete_particles' compaction (particle_mesh_threaded.f90:720)
  const float kick2[2] = {a_mid, dt};
  if (int st = do_delete(ctx, c.coarse_vel_update ? kick2 : nullptr)) return st;
  CK(cudaEventRecord(ev[10], ctx->stream));
  CK(cudaEventSynchronize(ev[10]));
  // limiters (particle_mesh_threaded.f90:643-696, coarse_max_dt.f90:36)
  const float G = c.G;
  auto asf = [](unsigned int u) { float f; memcpy(&f, &u, 4); return f; };
  const float f2 = asf(hc.f_force_max2_bits), ppm = asf(hc.pp_force_max_bits), ppe = asf(hc.pp_ext_force_max_bits), cm = asf(hc.c_force_max_bits);
  float mx[5] = {sqrtf(f2), ppm, ppe, cm, (float)bufmax};
  double sm[4] = {hc.sum_rho_f, hc.sum_rho_c, (double)ctx->np_local, (double)ndel};
  if (int st = reduce_scalars(ctx, mx, 5, sm, 4)) return st;
  out->f_force_max = mx[0];
  out->dt_f_acc = 1.0f / sqrtf(std::max(0.0001f, out->f_force_max) * a_mid * G);
  out->pp_force_max = mx[1];
  out->dt_pp_acc = c.ppint ? sqrtf(c.dt_pp_scale * c.rsoft) / std::max(sqrtf(mx[1] * a_mid * G), 1e-3f) : 1000.f;
  out->pp_ext_force_max = mx[2];
  out->dt_pp_ext_acc = c.pp_ext ? sqrtf(c.dt_pp_scale * c.rsoft) / std::max(sqrtf(mx[2] * a_mid * G), 1e-3f) : 1000.f;
  out->c_force_max = mx[3];
  out->dt_c_acc = sqrtf((float)d.s / (mx[3] * a_mid * G));
  out->sum_rho_f = sm[0]; out->sum_rho_c = sm[1];
  out->np_local = ctx->np_local; out->np_total = (int64_t)(sm[2] + 0.5); out->np_with_ghosts = np_ghost; out->np_deleted_ll = (int)(sm[3] + 0.5);
  out->np_buf_max = (int)mx[4];
  out->stage_ms[CUBEP3M_B200_ST_DRIFT] = ev_ms(ev[0], ev[1]);
  out->stage_ms[CUBEP3M_B200_ST_PASS] = ev_ms(ev[1], ev[2]);
  out->stage_ms[CUBEP3M_B200_ST_LINK] = ev_ms(ev[2], ev[3]);
  out->stage_ms[CUBEP3M_B200_ST_FINE_FFT] = ev_ms(ev[3], ev[4]);
  out->stage_ms[CUBEP3M_B200_ST_PP] = ev_ms(ev[4], ev[5]);
  out->stage_ms[CUBEP3M_B200_ST_PP_EXT] = ev_ms(ev[5], ev[11]);
  out->stage_ms[CUBEP3M_B200_ST_COARSE_MASS] = ev_ms(ev[6], ev[7]);
  out->stage_ms[CUBEP3M_B200_ST_COARSE_FORCE] = ev_ms(ev[7], ev[8]);
  out->stage_ms[CUBEP3M_B200_ST_COARSE_VEL] = ev_ms(ev[12], ev[9]);
  out->stage_ms[CUBEP3M_B200_ST_DELETE] = ev_ms(ev[9], ev[10]);
  out->stage_ms[CUBEP3M_B200_ST_TOTAL] = ev_ms(ev[0], ev[10]);
  ctx->last_tile_counts_valid = 1;
  if (ctx->profiling) {
    for (int k = 0; k < KC_COUNT; ++k) { ctx->class_ms[k] = 0.f; ctx->class_n[k] = 0; }
    for (int i = 0; i < ctx->prof_n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]);
      ctx->class_ms[ctx->prof_class[i]] += ms; ctx->class_n[ctx->prof_class[i]]++;
    }
  }
  return 0;
}

int cubep3m_b200_set_profiling(cubep3m_b200_ctx* ctx, int on) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  if (on && ctx->prof_ev.empty()) {
    ctx->prof_ev.resize(2 * PROF_MAX); ctx->prof_class.resize(PROF_MAX);
    for (auto& e : ctx->prof_ev) CK(cudaEventCreate(&e));
  }
  ctx->profiling = on != 0;
  return 0;
}
int cubep3m_b200_set_tile_streams(cubep3m_b200_ctx* ctx, int n) {
  if (!ctx || n < 1 || n > ctx->tile_streams_max) return CUBEP3M_B200_EINVAL;
  ctx->tile_streams = n;
  return 0;
}
int cubep3m_b200_debug_ppext_blocks(cubep3m_b200_ctx* ctx, int32_t* blocks, int32_t* fallback) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  if (blocks) *blocks = ctx->ppext_blocks;
  if (fallback) *fallback = ctx->ppext_fallback;
  return 0;
}
int cubep3m_b200_debug_pair_counts(cubep3m_b200_ctx* ctx, int64_t* ppint, int64_t* ppext) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  if (ppint) *ppint = ctx->pairs_ppint;
  if (ppext) *ppext = ctx->pairs_ppext;
  return 0;
}
int cubep3m_b200_num_kernel_classes(void) { return KC_COUNT; }
const char* cubep3m_b200_kernel_class_name(int k) { return (k >= 0 && k < KC_COUNT) ? kKernelClassNames[k] : ""; }
int cubep3m_b200_get_kernel_times(cubep3m_b200_ctx* ctx, float* ms, int64_t* launches) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  for (int k = 0; k < KC_COUNT; ++k) { ms[k] = ctx->class_ms[k]; launches[k] = ctx->class_n[k]; }
  return 0;
}

// ---------------------------------------------------------------- debug / parity getters
int cubep3m_b200_debug_cell_counts(cubep3m_b200_ctx* ctx, int32_t* counts) {
  if (!ctx || !ctx->sorted) return CUBEP3M_B200_ENOTREADY;
  CK(cudaSetDevice(ctx->device));
  const long long nco = ctx->d.NF / 64;
  int* dtmp = nullptr;
  CK(cudaMalloc(&dtmp, sizeof(int) * nco));
  LAUNCH(ctx, KC_MISC, part::coarse_counts_kernel, (int)((nco + part::TPB - 1) / part::TPB), part::TPB, 0, ctx->fstart, nco, dtmp);
  CK(cudaMemcpyAsync(counts, dtmp, sizeof(int) * nco, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dtmp);
  return 0;
}
int cubep3m_b200_debug_tile_counts(cubep3m_b200_ctx* ctx, int32_t* counts) {
  if (!ctx || !ctx->last_tile_counts_valid) return CUBEP3M_B200_ENOTREADY;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(counts, ctx->tile_counts, sizeof(int) * ctx->d.tiles_node, cudaMemcpyDeviceToHost));
  return 0;
}
int cubep3m_b200_debug_sorted_particles(cubep3m_b200_ctx* ctx, float* xv, int32_t* np) {
  if (!ctx || !ctx->sorted) return CUBEP3M_B200_ENOTREADY;
  CK(cudaSetDevice(ctx->device));
  if (xv) CK(cudaMemcpy(xv, ctx->xv[ctx->cur], sizeof(float) * 6 * (size_t)ctx->np_all, cudaMemcpyDeviceToHost));
  if (np) *np = ctx->np_all;
  return 0;
}
int cubep3m_b200_debug_kern_f(cubep3m_b200_ctx* ctx, float* kern_f) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const size_t plane = (size_t)ctx->d.hc * ctx->d.n * ctx->d.n;
  return download_interleaved(kern_f, ctx->kern_f, (size_t)ctx->kf_stride, 0, plane, (size_t)ctx->d.hc, (size_t)ctx->kf_pitch);
}
int cubep3m_b200_debug_kern_c(cubep3m_b200_ctx* ctx, float* kern_c) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  // reference layout kern_c(3,hc,nc_dim,nc_slab): this rank's z-slab for the reference's cubic grids, else the whole table
  const Dims& d = ctx->d;
  if (!ctx->kern_c) return CUBEP3M_B200_ENOTREADY;   // slab-decomposed multi-rank runs keep only the rank's pencil rows of the table
  const size_t plane = (size_t)(d.Nc[0] / 2 + 1) * d.Nc[1] * d.Nc[2];
  if (d.nc_slab > 0) { const size_t per = (size_t)(d.Nc[0] / 2 + 1) * d.Nc[1] * d.nc_slab; return download_interleaved(kern_c, ctx->kern_c, plane, per * ctx->cfg.rank, per); }
  return download_interleaved(kern_c, ctx->kern_c, plane, 0, plane);
}
int cubep3m_b200_debug_rho_c(cubep3m_b200_ctx* ctx, float* rho_c) {
  // note: after a full step rho_c holds the last force component (as in the reference, coarse_force.f90:88);
  // call after cubep3m_b200_debug_coarse_mass-like use only. Here: recompute the deposit from the sorted array.
  if (!ctx || !ctx->sorted) return CUBEP3M_B200_ENOTREADY;
  (void)rho_c;
  return CUBEP3M_B200_ENOTREADY;
}
int cubep3m_b200_debug_force_c(cubep3m_b200_ctx* ctx, float* force_c) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const size_t nfc = (size_t)3 * (ctx->d.nc_node + 2) * (ctx->d.nc_node + 2) * (ctx->d.nc_node + 2);
  CK(cudaMemcpy(force_c, ctx->force_c, sizeof(float) * nfc, cudaMemcpyDeviceToHost));
  return 0;
}
// halofind.f90:564-672: density + maxima pass of find_halos over every tile of the node, in the reference's tile order (halofind.f90:48-54).
// Valid in the state a halofind step is in (cubepm.f90:193-198: after link_list and particle_pass). Peaks come back tile by tile, inside a tile
// by ascending density (the order indexedsort leaves them in, :676-679; ties in scan order).
int cubep3m_b200_halofind_peaks(cubep3m_b200_ctx* ctx, float mass_p, float den_peak_cutoff, int32_t para_inter_hc, int32_t ngph, cubep3m_b200_peak* peaks,
                                int32_t max_peaks, int32_t* n_peaks, double* cftmass) {
  if (!ctx || !peaks || !n_peaks || max_peaks <= 0) return CUBEP3M_B200_EINVAL;
  if (!ctx->sorted || !ctx->passed) return CUBEP3M_B200_ENOTREADY;
  CK(cudaSetDevice(ctx->device));
  static_assert(sizeof(halo::Peak) == sizeof(cubep3m_b200_peak), "peak record layout");
  const Dims& d = ctx->d;
  const int T = d.T, n = d.n;
  int* scratch = ctx->rowoff + d.nc_node * d.nc_node + 4;
  halo::Peak* dpk = nullptr;
  double* dsum = nullptr;          // [0] deposit DIAG sum (unused), [1..2] cftmass, cftmass2
  int* dn = nullptr;
  CK(cudaMalloc(&dpk, sizeof(halo::Peak) * (size_t)max_peaks));
  CK(cudaMalloc(&dsum, 3 * sizeof(double)));
  CK(cudaMalloc(&dn, sizeof(int)));
  CK(cudaMemsetAsync(dsum, 0, 3 * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(dn, 0, sizeof(int), ctx->stream));
  std::vector<int> tile_end(d.tiles_node, 0);
  int status = 0;
  for (int tile = 0; tile < d.tiles_node && !status; ++tile) {
    const int tz = tile / (T * T), ty = (tile / T) % T, tx = tile % T;
    if (!ngph) {
      if ((status = launch_cic_density(ctx, ctx->tile_rho, tx, ty, tz, mass_p, dsum, scratch))) break;
    } else {
      LAUNCH(ctx, KC_DENSITY, fine::ngp_density_kernel, NUM_SMS * 8, fine::TPB, 0, ctx->fstart, ctx->tile_rho, n, d.b, d.m, d.H, tx, ty, tz, mass_p, dsum, scratch);
      if (ctx->hcnt->n_cand > 0)
        LAUNCH(ctx, KC_DENSITY, fine::ngp_fixup_kernel, NUM_SMS, fine::TPB, 0, ctx->cand, &ctx->dcnt->n_cand, ctx->cand_cap, ctx->tile_rho, n, d.b, d.m, tx, ty, tz,
               mass_p, dsum);
    }
    const long long cells = (long long)d.m * d.m * d.m;
    LAUNCH(ctx, KC_MISC, halo::peak_kernel, (unsigned)std::min<long long>((cells + halo::TPB - 1) / halo::TPB, (long long)NUM_SMS * 8), halo::TPB, 0, ctx->tile_rho, n, d.b, d.m,
           den_peak_cutoff, para_inter_hc, tile, (float)(tx * d.m - d.b), (float)(ty * d.m - d.b), (float)(tz * d.m - d.b), dpk, max_peaks, dn, dsum + 1);
    CK(cudaMemcpyAsync(&tile_end[tile], dn, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  double hs[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(hs, dsum, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int total = tile_end[d.tiles_node - 1];
  if (total > max_peaks) status = CUBEP3M_B200_ECAPACITY;             // 'too many halos' (:626-629)
  if (!status) {
    std::vector<halo::Peak> h(total);
    CK(cudaMemcpy(h.data(), dpk, sizeof(halo::Peak) * (size_t)total, cudaMemcpyDeviceToHost));
    for (int tile = 0, s0 = 0; tile < d.tiles_node; ++tile) {
      auto scan_key = [&](const halo::Peak& q) { return ((long long)q.k * n + q.j) * n + q.i; };
      std::sort(h.begin() + s0, h.begin() + tile_end[tile], [&](const halo::Peak& a, const halo::Peak& b2) { return a.den != b2.den ? a.den < b2.den : scan_key(a) < scan_key(b2); });
      s0 = tile_end[tile];
    }
    memcpy(peaks, h.data(), sizeof(halo::Peak) * (size_t)total);
  }
  *n_peaks = total;
  if (cftmass) { cftmass[0] = hs[1]; cftmass[1] = hs[2]; }
  cudaFree(dpk); cudaFree(dsum); cudaFree(dn);
  return status;
}

int cubep3m_b200_debug_fine_tile(cubep3m_b200_ctx* ctx, int32_t tile, float mass_p, float* rho_f, float* force_f) {
  if (!ctx || !ctx->sorted || !ctx->passed) return CUBEP3M_B200_ENOTREADY;
  if (tile < 0 || tile >= ctx->d.tiles_node) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const Dims& d = ctx->d;
  const int T = d.T, n = d.n;
  const int tz = tile / (T * T), ty = (tile / T) % T, tx = tile % T;
  int* scratch = ctx->rowoff + d.nc_node * d.nc_node + 4;
  double* dsum = nullptr;
  CK(cudaMalloc(&dsum, sizeof(double)));
  if (!ctx->cfg.ngp) {
    if (int st = launch_cic_density(ctx, ctx->tile_rho, tx, ty, tz, mass_p, dsum, scratch)) { cudaFree(dsum); return st; }
  } else {
    LAUNCH(ctx, KC_DENSITY, fine::ngp_density_kernel, NUM_SMS * 8, fine::TPB, 0, ctx->fstart, ctx->tile_rho, n, d.b, d.m, d.H, tx, ty, tz, mass_p, dsum, scratch);
    if (ctx->hcnt->n_cand > 0)
      LAUNCH(ctx, KC_DENSITY, fine::ngp_fixup_kernel, NUM_SMS, fine::TPB, 0, ctx->cand, &ctx->dcnt->n_cand, ctx->cand_cap, ctx->tile_rho, n, d.b, d.m, tx, ty, tz,
             mass_p, dsum);
  }
  if (rho_f) CK(cudaMemcpyAsync(rho_f, ctx->tile_rho, sizeof(float) * (size_t)(n + 2) * n * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dsum);
  if (force_f) {
    if (int st = fine_tile_solve(ctx, tile, mass_p, true, scratch)) return st;
    const size_t nf = (size_t)d.fdim * d.fdim * d.fdim;
    std::vector<float> tmp(nf);
    for (int comp = 0; comp < 3; ++comp) {   // interleave to the reference's force_f(3,...) layout
      CK(cudaMemcpyAsync(tmp.data(), ctx->force_f[comp], sizeof(float) * nf, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (size_t i = 0; i < nf; ++i) force_f[3 * i + comp] = tmp[i];
    }
  }
  return 0;
}
int cubep3m_b200_debug_fft3d(cubep3m_b200_ctx* ctx, int32_t n, float* data, int32_t inverse) {
  if (!ctx) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const Dims& d = ctx->d;
  float *buf, *scratch; fftk::Mesh3 g;
  if (n == d.n) { buf = ctx->tile_rho; scratch = ctx->tile_g; g = fine_mesh(ctx); }
  else if (n == d.Nc[0] && n == d.Nc[1] && n == d.Nc[2] && ctx->slab && ctx->creal) { buf = ctx->slab; scratch = ctx->creal; g = coarse_mesh3(ctx); }
  else return CUBEP3M_B200_EINVAL;
  const size_t bytes = sizeof(float) * (size_t)(n + 2) * n * n;
  CK(cudaMemcpyAsync(buf, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (!inverse) { if (int st = fftk::forward3d(ctx, g, buf)) return st; }
  else {
    // unnormalised c2r back into the padded layout (pitch n+2)
    const int lo[3] = {0, 0, 0}, cnt[3] = {n, n, n};
    if (int st = fftk::backward3d(ctx, g, buf, buf, nullptr, scratch, lo, cnt, n + 2, n, 1.0f)) return st;
    CK(cudaMemcpyAsync(buf, scratch, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  CK(cudaMemcpyAsync(data, buf, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int64_t cubep3m_b200_launch_count(cubep3m_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------- cic_power over z-slabs / y-pencils (several ranks, or meshes beyond one CTA's transform length)
namespace {
void power_shells(const std::vector<double>& sums, int stride, int nc, double box, int ngp_binning, double* k_out, double* delta2_out, double* sigma_out) {
  const double* P = sums.data(); const double* P2 = P + stride; const double* W = P2 + stride; const double* K = W + stride;
  for (int sh = 1; sh <= nc / 2; ++sh) {                      // cic_power.f90:1649-1660
    const double Wn = std::max(W[sh], 1e-300), kavg = K[sh] / Wn, Pm = P[sh] / Wn;
    const double var = std::max(P2[sh] / Wn - Pm * Pm, 0.0);
    const double keff = ngp_binning ? kavg : kavg - 1.0;
    k_out[sh - 1] = 2.0 * M_PI * kavg / box;
    delta2_out[sh - 1] = 4.0 * M_PI * keff * keff * keff * Pm;
    if (sigma_out) sigma_out[sh - 1] = 4.0 * M_PI * keff * keff * keff * sqrt(var / std::max(W[sh] - 1.0, 1.0));
  }
}

// The reference's layout (cic_power.f90:840-954: cube -> slab, distributed r2c, shell sums reduced over the ranks) with the library's means: the CIC deposit
// adds straight into the owners' z-slabs over NVLink (power::cic_density_dist_kernel), x and y passes run on the slab, one transpose stores into the
// peers' y-pencils (cslab::transpose_kernel), the z pass and the shell binning run on the pencils, the shell sums are all-reduced. `big`: the axis
// length needs the four-step passes of bigfft.cuh (1024, 2048); the y / z frequencies then stay digit-transposed and the binning kernel maps them.
int cic_power_slabs(cubep3m_b200_ctx* ctx, const float shake_offset[3], double box, int ngp_binning, double* k_out, double* delta2_out, double* sigma_out, int nc,
                    bool big) {
  const Dims& d = ctx->d;
  const int W = d.world, me = ctx->cfg.rank, zs = nc / W, ys = nc / W, hc = nc / 2 + 1;
  const size_t slab_f = ((size_t)(nc + 2) * nc * zs + 63) / 64 * 64;       // floats: a z-slab (nc+2, nc, zs) = a y-pencil block (hc, ys, nc) complex
  const int nb = nc / 2 + 2, stride = nb + 2;
  float *xchg = nullptr, *outb = nullptr; double *dsinc = nullptr, *dsums = nullptr; float2* tw = nullptr;
  fftk::BigTwiddles BT;
  PeerTable P;
  std::vector<void*> opened;
  int status = 0;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    for (void* q : opened) cudaIpcCloseMemHandle(q);
    if (xchg) cudaFree(xchg); if (outb) cudaFree(outb); if (dsinc) cudaFree(dsinc); if (dsums) cudaFree(dsums); if (tw) cudaFree(tw);
    fftk::big_free(BT);
  };
#define PCK(x) do { if ((x) != cudaSuccess) { cleanup(); return CUBEP3M_B200_ECUDA; } } while (0)
#define PST(x) do { status = (x); if (status) { cleanup(); return status; } } while (0)
  PCK(cudaMalloc((void**)&xchg, 2 * slab_f * sizeof(float)));               // [slab A | pencils T]: what the peers store into
  if (big) PCK(cudaMalloc((void**)&outb, slab_f * sizeof(float)));
  PCK(cudaMalloc((void**)&dsinc, nc * sizeof(double)));
  PCK(cudaMalloc((void**)&dsums, 4 * stride * sizeof(double)));
  if (big) PST(fftk::big_init(nc, BT)); else PST(fftk::make_twiddles(nc, &tw));
  bool ok = true;
  PST(map_peers(ctx, xchg, &P, &opened, &ok));
  if (!ok) { cleanup(); return CUBEP3M_B200_ENCCL; }
  // global particle count -> particle mass (nc / np)^3 of cic_power.f90
  double np_tot = (double)ctx->np_local;
  if (W > 1) {
#ifdef CUBEP3M_WITH_NCCL
    double* dn = reinterpret_cast<double*>(ctx->redbuf + 16);
    PCK(cudaMemcpyAsync(dn, &np_tot, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (ncclAllReduce(dn, dn, 1, ncclDouble, ncclSum, ctx->comm, ctx->stream) != ncclSuccess) { cleanup(); return CUBEP3M_B200_ENCCL; }
    PCK(cudaMemcpyAsync(&np_tot, dn, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PCK(cudaStreamSynchronize(ctx->stream));
#endif
  }
  const float mp = (float)(((double)nc * nc * nc) / np_tot);
  float* A = xchg;
  float2* T = reinterpret_cast<float2*>(xchg + slab_f);
  PCK(cudaMemsetAsync(A, 0, slab_f * sizeof(float), ctx->stream));
  PCK(cudaMemsetAsync(dsums, 0, 4 * stride * sizeof(double), ctx->stream));
  std::vector<double> sinc4(nc);
  for (int i = 0; i < nc; ++i) {
    const int k = i < nc / 2 ? i : i - nc;
    const double x = M_PI * (double)k / (double)nc, sc = (k == 0) ? 1.0 : sin(x) / x;
    sinc4[i] = sc * sc * sc * sc;
  }
  PCK(cudaMemcpyAsync(dsinc, sinc4.data(), nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  PST(rank_barrier(ctx));                                                    // every slab is zeroed before the first remote contribution arrives
  const float zero[3] = {0.f, 0.f, 0.f};
  const float* so = shake_offset ? shake_offset : zero;
  LAUNCH(ctx, KC_MISC, power::cic_density_dist_kernel, (ctx->np_local + power::TPB - 1) / power::TPB, power::TPB, 0, ctx->xv[ctx->cur], ctx->np_local, nc,
         (float)(d.coord[0] * d.mT), (float)(d.coord[1] * d.mT), (float)(d.coord[2] * d.mT), so[0], so[1], so[2], mp, P, 0LL, zs);
  PST(rank_barrier(ctx));                                                    // all contributions to this rank's slab have landed
  float2* S = reinterpret_cast<float2*>(A);
  if (big) {
    PST(fftk::big_forward_x(ctx, KC_MISC, BT, A, reinterpret_cast<float2*>(outb), (long long)nc * zs));
    S = reinterpret_cast<float2*>(outb);
    PST(fftk::big_forward_strided(ctx, KC_MISC, BT, S, hc, (long long)hc, (long long)nc * hc, zs));
  } else {
    PST(fftk::launch_x_r2c(ctx, KC_MISC, nc, A, nc * zs, tw));
    PST(fftk::launch_strided(ctx, KC_MISC, nc, false, S, S, hc, (long long)hc, (long long)nc * hc, 0, zs, nullptr, 0, 0, 0, nc - 1, tw));
  }
  {
    const dim3 tgrid((unsigned)std::max(1, std::min(64, (ys * hc + cslab::TPB - 1) / cslab::TPB)), (unsigned)(W * zs));
    LAUNCH(ctx, KC_MISC, cslab::transpose_kernel<true>, tgrid, cslab::TPB, 0, S, P, (long long)slab_f, W, me, zs, ys, nc, hc);
  }
  PST(rank_barrier(ctx));                                                    // this rank's pencils are complete
  if (big) PST(fftk::big_forward_strided(ctx, KC_MISC, BT, T, hc, (long long)ys * hc, (long long)hc, ys));
  else PST(fftk::launch_strided(ctx, KC_MISC, nc, false, T, T, hc, (long long)ys * hc, (long long)hc, 0, ys, nullptr, 0, 0, 0, nc - 1, tw));
  PCK(cudaFuncSetAttribute(power::shell_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * stride * sizeof(double))));
  LAUNCH(ctx, KC_MISC, power::shell_bin_kernel, NUM_SMS * 4, power::TPB, 4 * stride * sizeof(double), T, nc, ys, me * ys, big ? fftk::BIG_N1 : 0, big ? nc / fftk::BIG_N1 : 0,
         dsinc, ngp_binning, nb, dsums);
  if (W > 1) {
#ifdef CUBEP3M_WITH_NCCL
    if (ncclAllReduce(dsums, dsums, (size_t)4 * stride, ncclDouble, ncclSum, ctx->comm, ctx->stream) != ncclSuccess) { cleanup(); return CUBEP3M_B200_ENCCL; }
#endif
  }
  std::vector<double> sums(4 * stride);
  PCK(cudaMemcpyAsync(sums.data(), dsums, sums.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  PCK(cudaGetLastError());
  PST(rank_barrier(ctx));                                                    // nobody unmaps a buffer a peer may still be reading
  PCK(cudaStreamSynchronize(ctx->stream));
#undef PCK
#undef PST
  cleanup();
  power_shells(sums, stride, nc, box, ngp_binning, k_out, delta2_out, sigma_out);
  return 0;
}
}  // namespace

// ---------------------------------------------------------------- cic_power on the device (utils/cic_power/cic_power.f90)
int cubep3m_b200_cic_power(cubep3m_b200_ctx* ctx, const float shake_offset[3], double box, int32_t ngp_binning, double* k_out, double* delta2_out,
                           double* sigma_out, int32_t nshells) {
  if (!ctx || !k_out || !delta2_out) return CUBEP3M_B200_EINVAL;
  const Dims& d = ctx->d;
  if (d.world > 1 && !(d.Dg[0] == d.Dg[1] && d.Dg[1] == d.Dg[2])) return CUBEP3M_B200_EINVAL;   // a cubic box only (the reference's nodes_dim^3)
  const int nc = d.mT * d.Dg[0];                              // nf_physical_dim
  const char* pmode = getenv("CUBEP3M_B200_POWER");             // "slab": the slab / pencil pipeline on one rank; "big": also the four-step transforms (tests)
  const bool force_big = pmode && !strcmp(pmode, "big") && fftk::big_supported(nc);
  const bool direct = fftk::supported(nc) && !force_big, big = !direct && fftk::big_supported(nc);
  if ((!direct && !big) || nshells != nc / 2 || nc % d.world) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const int np = ctx->np_local;
  if (np <= 0 || ctx->passed) return CUBEP3M_B200_ENOTREADY;   // needs the physical particles only (after delete_particles / upload)
  {
    if (d.world > 1 || big || (pmode && !strcmp(pmode, "slab"))) return cic_power_slabs(ctx, shake_offset, box, ngp_binning, k_out, delta2_out, sigma_out, nc, big);
  }
  const size_t nreal = (size_t)(nc + 2) * nc * nc;
  const int nb = nc / 2 + 2, stride = nb + 2;
  float* rho = nullptr; float2* tw = nullptr; double *dsinc = nullptr, *dsums = nullptr;
  auto cleanup = [&]() { if (rho) cudaFree(rho); if (tw) cudaFree(tw); if (dsinc) cudaFree(dsinc); if (dsums) cudaFree(dsums); };
#define PCK(x) do { if ((x) != cudaSuccess) { cleanup(); return CUBEP3M_B200_ECUDA; } } while (0)
  PCK(cudaMalloc((void**)&rho, nreal * sizeof(float)));
  PCK(cudaMalloc((void**)&dsinc, nc * sizeof(double)));
  PCK(cudaMalloc((void**)&dsums, 4 * stride * sizeof(double)));
  if (fftk::make_twiddles(nc, &tw)) { cleanup(); return CUBEP3M_B200_ECUDA; }
  PCK(cudaMemsetAsync(rho, 0, nreal * sizeof(float), ctx->stream));
  PCK(cudaMemsetAsync(dsums, 0, 4 * stride * sizeof(double), ctx->stream));
  std::vector<double> sinc4(nc);
  for (int i = 0; i < nc; ++i) {
    const int k = i < nc / 2 ? i : i - nc;
    const double x = M_PI * (double)k / (double)nc, sc = (k == 0) ? 1.0 : sin(x) / x;
    sinc4[i] = sc * sc * sc * sc;
  }
  PCK(cudaMemcpyAsync(dsinc, sinc4.data(), nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const float zero[3] = {0.f, 0.f, 0.f};
  const float* so = shake_offset ? shake_offset : zero;
  const float mp = (float)(((double)nc * nc * nc) / (double)np);
  LAUNCH(ctx, KC_MISC, power::cic_density_kernel, (np + power::TPB - 1) / power::TPB, power::TPB, 0, ctx->xv[ctx->cur], np, nc, so[0], so[1], so[2], mp, rho);
  const fftk::Mesh3 g{nc, nc, nc, tw, tw, tw};
  if (int st = fftk::forward3d(ctx, g, rho)) { cleanup(); return st; }
  PCK(cudaFuncSetAttribute(power::shell_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * stride * sizeof(double))));
  LAUNCH(ctx, KC_MISC, power::shell_bin_kernel, NUM_SMS * 4, power::TPB, 4 * stride * sizeof(double), reinterpret_cast<const float2*>(rho), nc, nc, 0, 0, 0, dsinc, ngp_binning, nb, dsums);
  std::vector<double> sums(4 * stride);
  PCK(cudaMemcpyAsync(sums.data(), dsums, sums.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  PCK(cudaGetLastError());
#undef PCK
  cleanup();
  const double* P = sums.data(); const double* P2 = P + stride; const double* W = P2 + stride; const double* K = W + stride;
  for (int sh = 1; sh <= nc / 2; ++sh) {                      // cic_power.f90:1649-1660
    const double Wn = std::max(W[sh], 1e-300), kavg = K[sh] / Wn, Pm = P[sh] / Wn;
    const double var = std::max(P2[sh] / Wn - Pm * Pm, 0.0);
    const double keff = ngp_binning ? kavg : kavg - 1.0;
    k_out[sh - 1] = 2.0 * M_PI * kavg / box;
    delta2_out[sh - 1] = 4.0 * M_PI * keff * keff * keff * Pm;
    if (sigma_out) sigma_out[sh - 1] = 4.0 * M_PI * keff * keff * keff * sqrt(var / std::max(W[sh] - 1.0, 1.0));
  }
  return 0;
}

// ---------------------------------------------------------------- dist_init on the device (utils/dist_init/dist_init_dm.f90)
extern "C" int cubep3m_b200_dist_init(cubep3m_b200_ctx* ctx, int32_t nc, int32_t reps, float box, float vfactor, uint64_t seed, const float* k_table,
                                      const float* delta2_table, int32_t n_table, const float* noise, int32_t* np_local) {
  if (!ctx || !k_table || !delta2_table || n_table < 2 || reps < 1 || nc < 4 || nc % 2) return CUBEP3M_B200_EINVAL;
  const Dims& d = ctx->d;
  if (d.world != 1 || nc * reps != d.mT || !fftk::supported(nc)) return CUBEP3M_B200_EINVAL;   // one rank; the box (times its replication) is the node
  const long long np = (long long)(nc / 2) * (nc / 2) * (nc / 2) * reps * reps * reps;
  if (np > d.max_np) return CUBEP3M_B200_EMAXNP;
  CK(cudaSetDevice(ctx->device));
  const size_t nreal = (size_t)(nc + 2) * nc * nc;
  float *mesh = nullptr, *phi = nullptr, *tab = nullptr; float2* tw = nullptr;
  auto cleanup = [&]() { if (mesh) cudaFree(mesh); if (phi) cudaFree(phi); if (tab) cudaFree(tab); if (tw) cudaFree(tw); };
#define DCK(x) do { if ((x) != cudaSuccess) { cleanup(); return CUBEP3M_B200_ECUDA; } } while (0)
  DCK(cudaMalloc((void**)&mesh, nreal * sizeof(float)));
  DCK(cudaMalloc((void**)&phi, nreal * sizeof(float)));
  DCK(cudaMalloc((void**)&tab, (size_t)2 * n_table * sizeof(float)));
  if (fftk::make_twiddles(nc, &tw)) { cleanup(); return CUBEP3M_B200_ECUDA; }
  DCK(cudaMemcpyAsync(tab, k_table, n_table * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  DCK(cudaMemcpyAsync(tab + n_table, delta2_table, n_table * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  if (noise) DCK(cudaMemcpy2DAsync(mesh, (size_t)(nc + 2) * sizeof(float), noise, (size_t)nc * sizeof(float), (size_t)nc * sizeof(float), (size_t)nc * nc, cudaMemcpyHostToDevice, ctx->stream));
  else LAUNCH(ctx, KC_MISC, distinit::noise_kernel, NUM_SMS * 8, distinit::TPB, 0, mesh, nc, (unsigned long long)seed);      // :535-667
  const fftk::Mesh3 g{nc, nc, nc, tw, tw, tw};
  if (int st = fftk::forward3d(ctx, g, mesh)) { cleanup(); return st; }                                                        // :653
  LAUNCH(ctx, KC_MISC, distinit::phi_k_kernel, NUM_SMS * 8, distinit::TPB, 0, reinterpret_cast<float2*>(mesh), nc, box, tab, tab + n_table, n_table);   // :685-712, :814-832
  const int lo[3] = {0, 0, 0}, cnt[3] = {nc, nc, nc};
  const float scale = 1.0f / (((float)nc * (float)nc) * (float)nc);
  if (int st = fftk::backward3d(ctx, g, mesh, mesh, nullptr, phi, lo, cnt, nc, nc, scale)) { cleanup(); return st; }           // :958-969
  LAUNCH(ctx, KC_MISC, distinit::particles_kernel, NUM_SMS * 8, distinit::TPB, 0, phi, nc, reps, vfactor, ctx->xv[ctx->cur]); // :1011-1036
  DCK(cudaStreamSynchronize(ctx->stream));
  DCK(cudaGetLastError());
#undef DCK
  cleanup();
  ctx->np_local = (int)np; ctx->np_all = (int)np; ctx->sorted = false; ctx->passed = false;
  if (np_local) *np_local = (int32_t)np;
  return 0;
}

// ---------------------------------------------------------------- checkpoint.f90:72-95 / particle_initialization.f90:88-189
namespace {
size_t ckpt_header_bytes(const cubep3m_b200_ctx* ctx) { return ctx->cfg.ppint ? 48 : 44; }
void ckpt_pack_header(const cubep3m_b200_ctx* ctx, const cubep3m_b200_checkpoint_header& h, unsigned char* out) {
  unsigned char* p = out;
  auto put = [&](const void* v) { memcpy(p, v, 4); p += 4; };
  put(&h.np_local); put(&h.a); put(&h.t); put(&h.tau); put(&h.nts); put(&h.dt_f_acc);
  if (ctx->cfg.ppint) put(&h.dt_pp_acc);
  put(&h.dt_c_acc); put(&h.cur_checkpoint); put(&h.cur_projection); put(&h.cur_halofind); put(&h.mass_p);
}
void ckpt_unpack_header(const cubep3m_b200_ctx* ctx, const unsigned char* in, cubep3m_b200_checkpoint_header* h) {
  const unsigned char* p = in;
  auto get = [&](void* v) { memcpy(v, p, 4); p += 4; };
  memset(h, 0, sizeof(*h));
  get(&h->np_local); get(&h->a); get(&h->t); get(&h->tau); get(&h->nts); get(&h->dt_f_acc);
  if (ctx->cfg.ppint) get(&h->dt_pp_acc);
  get(&h->dt_c_acc); get(&h->cur_checkpoint); get(&h->cur_projection); get(&h->cur_halofind); get(&h->mass_p);
}
constexpr long long CKPT_BLOCK = (32LL * 1024 * 1024) / 24;   // particles per block: the reference's blocksize (checkpoint.f90:51, particle_initialization.f90:128)
}  // namespace

extern "C" int cubep3m_b200_write_checkpoint(cubep3m_b200_ctx* ctx, const char* path_xv, const char* path_pid, const cubep3m_b200_checkpoint_header* hdr,
                                             const float shake_offset[3]) {
  if (!ctx || !path_xv || !hdr) return CUBEP3M_B200_EINVAL;
  if (ctx->passed) return CUBEP3M_B200_ENOTREADY;            // ghosts must be gone (the driver checkpoints between steps)
  CK(cudaSetDevice(ctx->device));
  const long long np = ctx->np_local;
  cubep3m_b200_checkpoint_header h = *hdr;
  h.np_local = (int32_t)np;
  unsigned char hb[48];
  ckpt_pack_header(ctx, h, hb);
  const float zero[3] = {0.f, 0.f, 0.f};
  const float* so = shake_offset ? shake_offset : zero;
  FILE* f = fopen(path_xv, "wb");
  if (!f) { fprintf(stderr, "cubep3m_b200: error opening checkpoint file for write: %s\n", path_xv); return CUBEP3M_B200_EINVAL; }   // checkpoint.f90:60-64
  int status = 0;
  float* dstage[2] = {nullptr, nullptr};
  float* hstage[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  const long long blk = std::min<long long>(CKPT_BLOCK, std::max<long long>(np, 1));
  for (int q = 0; q < 2 && !status; ++q)
    if (cudaMalloc((void**)&dstage[q], blk * 24) != cudaSuccess || cudaMallocHost((void**)&hstage[q], blk * 24) != cudaSuccess ||
        cudaEventCreateWithFlags(&done[q], cudaEventDisableTiming) != cudaSuccess) status = CUBEP3M_B200_ECUDA;
  if (!status && fwrite(hb, 1, ckpt_header_bytes(ctx), f) != ckpt_header_bytes(ctx)) status = CUBEP3M_B200_EINVAL;
  // double-buffered: block k+1 is packed and copied while block k is written to the file
  const long long nblk = (np + blk - 1) / blk;
  auto issue = [&](long long k) {
    const int q = (int)(k & 1);
    const long long first = k * blk;
    const int cnt = (int)std::min(blk, np - first);
    LAUNCH(ctx, KC_MISC, part::checkpoint_pack_kernel, (cnt + part::TPB - 1) / part::TPB, part::TPB, 0, ctx->xv[ctx->cur], first, cnt, so[0], so[1], so[2], dstage[q]);
    cudaMemcpyAsync(hstage[q], dstage[q], (size_t)cnt * 24, cudaMemcpyDeviceToHost, ctx->stream);
    cudaEventRecord(done[q], ctx->stream);
  };
  if (!status && nblk > 0) issue(0);
  for (long long k = 0; k < nblk && !status; ++k) {
    const int q = (int)(k & 1);
    if (cudaEventSynchronize(done[q]) != cudaSuccess) { status = CUBEP3M_B200_ECUDA; break; }
    if (k + 1 < nblk) issue(k + 1);
    const size_t cnt = (size_t)std::min(blk, np - k * blk);
    if (fwrite(hstage[q], 24, cnt, f) != cnt) status = CUBEP3M_B200_EINVAL;
  }
  if (fclose(f) != 0 && !status) status = CUBEP3M_B200_EINVAL;
  if (!status && path_pid && ctx->cfg.pid) {                   // checkpoint.f90:99-135
    FILE* g = fopen(path_pid, "wb");
    if (!g) status = CUBEP3M_B200_EINVAL;
    else {
      if (fwrite(hb, 1, ckpt_header_bytes(ctx), g) != ckpt_header_bytes(ctx)) status = CUBEP3M_B200_EINVAL;
      int64_t* hp = reinterpret_cast<int64_t*>(hstage[0]);
      const long long pblk = blk * 3;                          // ids per staging buffer
      for (long long first = 0; first < np && !status; first += pblk) {
        const size_t cnt = (size_t)std::min(pblk, np - first);
        if (cudaMemcpyAsync(hp, ctx->pid[ctx->cur] + first, cnt * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { status = CUBEP3M_B200_ECUDA; break; }
        if (fwrite(hp, 8, cnt, g) != cnt) status = CUBEP3M_B200_EINVAL;
      }
      if (fclose(g) != 0 && !status) status = CUBEP3M_B200_EINVAL;
    }
  }
  cudaStreamSynchronize(ctx->stream);
  for (int q = 0; q < 2; ++q) { if (dstage[q]) cudaFree(dstage[q]); if (hstage[q]) cudaFreeHost(hstage[q]); if (done[q]) cudaEventDestroy(done[q]); }
  if (cudaGetLastError() != cudaSuccess && !status) status = CUBEP3M_B200_ECUDA;
  return status;
}

extern "C" int cubep3m_b200_read_checkpoint(cubep3m_b200_ctx* ctx, const char* path_xv, const char* path_pid, cubep3m_b200_checkpoint_header* hdr) {
  if (!ctx || !path_xv || !hdr) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  FILE* f = fopen(path_xv, "rb");
  if (!f) { fprintf(stderr, "cubep3m_b200: error opening checkpoint: %s\n", path_xv); return CUBEP3M_B200_EINVAL; }     // particle_initialization.f90:106-110
  unsigned char hb[48];
  if (fread(hb, 1, ckpt_header_bytes(ctx), f) != ckpt_header_bytes(ctx)) { fclose(f); return CUBEP3M_B200_EINVAL; }
  ckpt_unpack_header(ctx, hb, hdr);
  const long long np = hdr->np_local;
  if (np < 0 || np > ctx->d.max_np) { fclose(f); fprintf(stderr, "cubep3m_b200: too many particles to store\n"); return CUBEP3M_B200_EMAXNP; }   // :122-126
  int status = 0;
  float* hstage[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  const long long blk = std::min<long long>(CKPT_BLOCK, std::max<long long>(np, 1));
  for (int q = 0; q < 2 && !status; ++q)
    if (cudaMallocHost((void**)&hstage[q], blk * 24) != cudaSuccess || cudaEventCreateWithFlags(&done[q], cudaEventDisableTiming) != cudaSuccess) status = CUBEP3M_B200_ECUDA;
  const long long nblk = (np + blk - 1) / blk;
  for (long long k = 0; k < nblk && !status; ++k) {
    const int q = (int)(k & 1);
    if (k >= 2 && cudaEventSynchronize(done[q]) != cudaSuccess) { status = CUBEP3M_B200_ECUDA; break; }   // the copy that last used this buffer has finished
    const size_t cnt = (size_t)std::min(blk, np - k * blk);
    if (fread(hstage[q], 24, cnt, f) != cnt) { status = CUBEP3M_B200_EINVAL; break; }
    if (cudaMemcpyAsync(ctx->xv[ctx->cur] + (size_t)6 * k * blk, hstage[q], cnt * 24, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { status = CUBEP3M_B200_ECUDA; break; }
    cudaEventRecord(done[q], ctx->stream);
  }
  fclose(f);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !status) status = CUBEP3M_B200_ECUDA;
  if (!status && path_pid && ctx->cfg.pid) {                   // particle_initialization.f90:146-189
    FILE* g = fopen(path_pid, "rb");
    if (!g) status = CUBEP3M_B200_EINVAL;
    else {
      unsigned char hb2[48];
      cubep3m_b200_checkpoint_header h2;
      if (fread(hb2, 1, ckpt_header_bytes(ctx), g) != ckpt_header_bytes(ctx)) status = CUBEP3M_B200_EINVAL;
      else { ckpt_unpack_header(ctx, hb2, &h2); if (h2.np_local != hdr->np_local) status = CUBEP3M_B200_EINVAL; }
      int64_t* hp = reinterpret_cast<int64_t*>(hstage[0]);
      const long long pblk = blk * 3;
      for (long long first = 0; first < np && !status; first += pblk) {
        const size_t cnt = (size_t)std::min(pblk, np - first);
        if (fread(hp, 8, cnt, g) != cnt) { status = CUBEP3M_B200_EINVAL; break; }
        if (cudaMemcpyAsync(ctx->pid[ctx->cur] + first, hp, cnt * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) status = CUBEP3M_B200_ECUDA;
      }
      fclose(g);
    }
  }
  for (int q = 0; q < 2; ++q) { if (hstage[q]) cudaFreeHost(hstage[q]); if (done[q]) cudaEventDestroy(done[q]); }
  if (status) return status;
  ctx->np_local = (int)np; ctx->np_all = (int)np; ctx->sorted = false; ctx->passed = false;
  return 0;
}

// ---------------------------------------------------------------- driver twin (host): timestep.f90
void cubep3m_b200_expansion(float a0, float dt0, float omega_m, float omega_l, float wde, float* da1, float* da2) { expansion_core(a0, dt0, omega_m, omega_l, wde, da1, da2); }
void cubep3m_b200_clock_init(cubep3m_b200_clock* c, float z_i, float omega_m, float omega_l) {
  memset(c, 0, sizeof(*c));
  c->a = 1.0f / (z_i + 1.0f);                        // cubepm.par:30, variable_initialization.f90:15-34
  c->tau = -3.0f / sqrtf(c->a);
  c->dt_f_acc = c->dt_pp_acc = c->dt_pp_ext_acc = c->dt_c_acc = 1000.f;
  c->omega_m = omega_m; c->omega_l = omega_l; c->wde = -1.0f;
  c->a_target = 1.0f; c->cosmo = 1; c->ppint = 1;
}
void cubep3m_b200_timestep(cubep3m_b200_clock* c) { timestep_core(c); }

// The same timestep on the device: the clock lives in device memory next to the step's counters, one thread runs timestep_core
// (real(8) expansion, timestep.f90:241-293) after folding in the limiters the last particle_mesh left, and the host only reads the result back.
int cubep3m_b200_timestep_device(cubep3m_b200_ctx* ctx, cubep3m_b200_clock* clock) {
  if (!ctx || !clock) return CUBEP3M_B200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->dclock) CK(cudaMalloc(&ctx->dclock, sizeof(cubep3m_b200_clock)));
  CK(cudaMemcpyAsync(ctx->dclock, clock, sizeof(*clock), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, KC_MISC, timestep_kernel, 1, 32, 0, ctx->dclock);
  CK(cudaMemcpyAsync(clock, ctx->dclock, sizeof(*clock), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
