// Slab-decomposed coarse-mesh solve over NVLink peer memory (coarse_force.f90:18-90 with fftw3ds.f90's pack_slab :4-54, the distributed
// r2c / c2r of :103-183, unpack_slab :56-101 and the face exchanges of coarse_force_buffer.f90:23-63).
//
// Global coarse mesh (Nx,Ny,Nz) = nc_node * (Dx,Dy,Dz) on W = Dx*Dy*Dz ranks, rank = x + Dx*(y + Dy*z):
//   1. cube -> z-slabs   rank q owns the zs = Nz/W planes [q*zs, (q+1)*zs) (the reference's nc_slab, rank-major). The pack kernel stores every
//                        chunk of the rank's rho_c cube STRAIGHT INTO THE OWNER'S SLAB at (x0,y0) = (rx,ry)*nc (pack_slab's isend + placement).
//   2. slab              r2c along x, FFT along y (local, the library's own passes).
//   3. transpose         to y-pencils: rank s owns rows [s*ys, (s+1)*ys), ys = Ny/W, for ALL z. Rank q stores slab[:, s*ys:(s+1)*ys, :] into
//                        planes [q*zs, (q+1)*zs) of s's pencil array T[z][yl][kx]: contiguous blocks on both sides, no unpack.
//   4. pencils           FFT along z; for each component: x i*kern_c(comp) (the rank's rows of the table) fused into the inverse z pass.
//   5. transpose back    s stores G[q*zs:(q+1)*zs] into rows [s*ys, (s+1)*ys) of q's slab for that component.
//   6. slab              inverse y, c2r along x, 1/(Nx*Ny*Nz) (fftw3ds.f90:161).
//   7. slab -> cube+halo rank q stores, for every rank d, the planes of its slab that fall in d's periodic z range [rz*nc-1, rz*nc+nc], cut to
//                        d's periodic (nc+2)^2 window, into d's force_c (3 components interleaved): unpack_slab AND the six face exchanges at once.
// Every exchange is ONE kernel on the sending GPU whose stores land in the receiver's memory (cudaIpc-mapped), followed by a flag raised in
// the receiver's mailbox behind __threadfence_system(); the receiver's stream spins on the W flags of that phase. No host synchronisation and
// no collective call inside the solve. Flags carry a per-step epoch and are never reset; a buffer is only rewritten after its owner has
// provably consumed it (the phase order and particle_mesh's end-of-step all-reduce).
#pragma once
#include "common.cuh"

namespace cslab {

constexpr int TPB = 256;
constexpr int MAXW = 8;
enum Phase { PH_CUBE = 0, PH_FWD, PH_BWD0, PH_BWD1, PH_BWD2, PH_HALO, PH_COUNT };

using Peers = PeerTable;   // every rank's exchange allocation as mapped into THIS process (own pointer for the own rank)

// 1. pack_slab: rho_c[z][y][x] of the cube at (rx,ry,rz) -> slab[(z%zs)][ry*nc + y][rx*nc + x] of rank rz*(nc/zs) + z/zs; padded row pitch Nx+2
__global__ void __launch_bounds__(TPB) scatter_cube_kernel(const float* __restrict__ rho_c, Peers P, long long off_slab, int nc, int zs, int Nx, int Ny,
                                                           int rx, int ry, int rz, double* __restrict__ sum) {
  const long long total = (long long)nc * nc * nc;
  const int per = nc / zs;
  double s = 0.0;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int x = (int)(t % nc), y = (int)((t / nc) % nc), z = (int)(t / ((long long)nc * nc));
    const float v = rho_c[t];
    s += (double)v;
    float* slab = P.base[rz * per + z / zs] + off_slab;
    slab[((long long)(z % zs) * Ny + (ry * nc + y)) * (Nx + 2) + (rx * nc + x)] = v;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(sum, s);
}

// 3. / 5. block transposes. Both directions move W*zs blocks of ys*hc contiguous complex elements:
//   forward : src = slab[zl][s*ys ..][.]            -> T of rank s   [(me*zs + zl)][.][.]
//   backward: src = G[(q*zs + zl)][.][.] (own)      -> slab of rank q [zl][me*ys ..][.]
template <bool FWD>
__global__ void __launch_bounds__(TPB) transpose_kernel(const float2* __restrict__ src, Peers P, long long off_dst, int W, int me, int zs, int ys, int Ny, int hc) {
  const int blk = blockIdx.y;                       // (peer, zl)
  const int peer = blk / zs, zl = blk - peer * zs;
  const long long len = (long long)ys * hc;
  const float2* s = FWD ? src + ((long long)zl * Ny + (long long)peer * ys) * hc : src + ((long long)(peer * zs + zl) * ys) * hc;
  float2* d = reinterpret_cast<float2*>(P.base[peer] + off_dst) + (FWD ? ((long long)(me * zs + zl) * ys) * hc : ((long long)zl * Ny + (long long)me * ys) * hc);
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < len; i += (long long)gridDim.x * TPB) d[i] = s[i];
}

// 7. unpack_slab + coarse_force_buffer: real3 = the three real-space components of this rank's slab, [comp][zl][y][x] (pitch Nx).
// blockIdx.y = (dest rank d, halo plane hz); the plane is written only by the rank whose slab holds global z = (rz_d*nc - 1 + hz) mod Nz.
__global__ void __launch_bounds__(TPB) scatter_halo_kernel(const float* __restrict__ real3, Peers P, long long off_force, int me, int nc, int zs, int Nx, int Ny,
                                                           int Nz, int Dx, int Dy) {
  const int fc = nc + 2;
  const int d = blockIdx.y / fc, hz = blockIdx.y - d * fc;
  const int dx = d % Dx, dy = (d / Dx) % Dy, dz = d / (Dx * Dy);
  const int z = (dz * nc - 1 + hz + Nz) % Nz;
  if (z / zs != me) return;
  const int zl = z - me * zs;
  const long long cstride = (long long)zs * Ny * Nx;
  const float* plane = real3 + (long long)zl * Ny * Nx;
  float* out = P.base[d] + off_force + (long long)hz * fc * fc * 3;
  for (int t = blockIdx.x * TPB + threadIdx.x; t < fc * fc; t += gridDim.x * TPB) {
    const int i = t % fc, j = t / fc;
    const int gx = (dx * nc - 1 + i + Nx) % Nx, gy = (dy * nc - 1 + j + Ny) % Ny;
    const long long o = (long long)gy * Nx + gx;
    float* q = out + (long long)t * 3;
    q[0] = plane[o]; q[1] = plane[o + cstride]; q[2] = plane[o + 2 * cstride];
  }
}

// ---- flags. mailbox layout (ints, in every rank's exchange allocation): flag[phase][source rank].
__global__ void signal_kernel(Peers P, long long off_mail, int W, int me, int phase, int epoch) {
  const int d = threadIdx.x;
  if (d < W) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(reinterpret_cast<int*>(P.base[d] + off_mail) + phase * MAXW + me) = epoch;
  }
}
// spins until every rank's flag of `phase` carries `epoch`; err[0] is raised on time-out
__global__ void wait_kernel(const int* mail, int W, int phase, int epoch, int* __restrict__ err, long long timeout_cycles) {
  const int s = threadIdx.x;
  if (s < W) {
    const volatile int* flag = mail + phase * MAXW + s;
    const long long t0 = clock64();
    while (*flag != epoch) {
      if (clock64() - t0 > timeout_cycles) { atomicExch(err, 1); break; }
      __nanosleep(100);
    }
    __threadfence_system();
  }
}

// kern_rows[comp][z][yl][kx] <- table[comp][z][me*ys + yl][kx] (global [comp][z][y][kx] table built at init)
__global__ void __launch_bounds__(TPB) extract_rows_kernel(const float* __restrict__ table, float* __restrict__ rows, int Nz, int Ny, int hc, int ys, int y0) {
  const long long per = (long long)Nz * ys * hc, total = 3 * per;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int comp = (int)(t / per);
    const long long r = t - comp * per;
    const int kx = (int)(r % hc), yl = (int)((r / hc) % ys), z = (int)(r / ((long long)hc * ys));
    rows[t] = table[(((long long)comp * Nz + z) * Ny + (y0 + yl)) * hc + kx];
  }
}
}  // namespace cslab
