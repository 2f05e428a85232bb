// Coarse-mesh stages: CIC mass assignment (coarse_mass.f90:82-99 + coarse_cic_mass(_buffer).f90), force halo
// (coarse_force_buffer.f90:23-63), max force (coarse_max_dt.f90:19-31), CIC force interpolation + kick
// (coarse_velocity.f90:137-179).  The FFT solve between them lives in lib.cu (uses fft3d.cuh).
#pragma once
#include "common.cuh"

namespace coarse {

constexpr int TPB = 256;

// CIC weights of coarse_cic_mass.f90:18-27 / coarse_velocity.f90:143-151: x = xv/4 - 0.5, i1 = floor(x)+1
__device__ __forceinline__ void cic_setup(float p, int coarse_ngp, int& i1, float& d1, float& d2) {
  const float x = __fsub_rn(__fmul_rn(0.25f, p), 0.5f);
  const int f = (int)floorf(x);
  i1 = f + 1;                 // 1-based index of the lower cell
  if (coarse_ngp) { d1 = 0.f; d2 = 1.f; }
  else { d1 = (float)i1 - x; d2 = 1.0f - d1; }
}

// Every particle chained in coarse cells 0..nc_node+1 (coarse_mass.f90:83-86) deposits with a bounds check
// (coarse_cic_mass_buffer.f90:59-113); interior cells never fail the check, so one kernel covers both routines.
// The sorted array holds those cells' particles in (nc_node+2)^2 contiguous x-rows.
__global__ void __launch_bounds__(TPB) cic_mass_kernel(const float* __restrict__ xv, const int* __restrict__ fstart, float* __restrict__ rho_c,
                                                       int H, int nc_buf, int nc_node, float mass_p, int coarse_ngp) {
  const int rows = nc_node + 2;
  const int ry = blockIdx.x % rows, rz = blockIdx.x / rows;
  const int cy = nc_buf - 1 + ry, cz = nc_buf - 1 + rz, cx0 = nc_buf - 1;   // hoc cell 0 is index nc_buf-1 (0-based in the hoc range)
  const long long k0 = ((long long)(cz * H + cy) * H + cx0) * 64;
  const int s0 = fstart[k0], s1 = fstart[k0 + (long long)rows * 64];
  for (int i = s0 + threadIdx.x; i < s1; i += TPB) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    const float z = p[1].x;
    int ix, iy, iz; float dx1, dx2, dy1, dy2, dz1, dz2;
    cic_setup(a.x, coarse_ngp, ix, dx1, dx2);
    cic_setup(a.y, coarse_ngp, iy, dy1, dy2);
    cic_setup(z, coarse_ngp, iz, dz1, dz2);
    dx1 = mass_p * dx1; dx2 = mass_p * dx2;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int jx = ix + (c & 1), jy = iy + ((c >> 1) & 1), jz = iz + (c >> 2);
      if (jx < 1 || jx > nc_node || jy < 1 || jy > nc_node || jz < 1 || jz > nc_node) continue;
      const float w = (((c & 1) ? dx2 : dx1) * (((c >> 1) & 1) ? dy2 : dy1)) * ((c >> 2) ? dz2 : dz1);
      atomicAdd(&rho_c[((long long)(jz - 1) * nc_node + (jy - 1)) * nc_node + (jx - 1)], w);
    }
  }
}

// The same deposit, staged in shared memory: the particles of one coarse x-row (cy, cz) only reach the 3 x 3 rho_c rows (cy-1..cy+1, cz-1..cz+1)
// (x = xv/4 - 0.5 puts i1 in {c-1, c} for a particle of coarse cell c), so the CTA accumulates them in a 9-row shared-memory window with
// shared-memory atomics and flushes each row once with coalesced global atomics: 9*nc_node global atomics per CTA instead of 8 per particle
// (8x fewer at the mean density of 8 particles per coarse cell, and contention-free however strongly the box clusters).
__global__ void __launch_bounds__(TPB) cic_mass_smem_kernel(const float* __restrict__ xv, const int* __restrict__ fstart, float* __restrict__ rho_c,
                                                            int H, int nc_buf, int nc_node, float mass_p, int coarse_ngp) {
  extern __shared__ float win[];                       // [3 z][3 y][W], W = nc_node + 4: x cells -1 .. nc_node + 2
  const int W = nc_node + 4;
  const int rows = nc_node + 2;
  const int ry = blockIdx.x % rows, rz = blockIdx.x / rows;     // hoc cells 0 .. nc_node + 1
  const int cy = nc_buf - 1 + ry, cz = nc_buf - 1 + rz, cx0 = nc_buf - 1;
  const long long k0 = ((long long)(cz * H + cy) * H + cx0) * 64;
  const int s0 = fstart[k0], s1 = fstart[k0 + (long long)rows * 64];
  if (s1 == s0) return;
  for (int t = threadIdx.x; t < 9 * W; t += TPB) win[t] = 0.f;
  __syncthreads();
  for (int i = s0 + threadIdx.x; i < s1; i += TPB) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    const float z = p[1].x;
    int ix, iy, iz; float dx1, dx2, dy1, dy2, dz1, dz2;
    cic_setup(a.x, coarse_ngp, ix, dx1, dx2);
    cic_setup(a.y, coarse_ngp, iy, dy1, dy2);
    cic_setup(z, coarse_ngp, iz, dz1, dz2);
    dx1 = mass_p * dx1; dx2 = mass_p * dx2;
    // window-local indices: y, z relative to (ry - 1, rz - 1) (1-based cells), x relative to cell -1
    const int ly = iy - (ry - 1), lz = iz - (rz - 1), lx = ix + 1;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = (((c & 1) ? dx2 : dx1) * (((c >> 1) & 1) ? dy2 : dy1)) * ((c >> 2) ? dz2 : dz1);
      atomicAdd(&win[((lz + (c >> 2)) * 3 + (ly + ((c >> 1) & 1))) * W + lx + (c & 1)], w);
    }
  }
  __syncthreads();
  // flush: window row (wz, wy) is rho_c row (jy, jz) = (ry - 1 + wy, rz - 1 + wz), kept if inside 1..nc_node (coarse_cic_mass_buffer.f90:59-113)
  for (int r = 0; r < 9; ++r) {
    const int jy = ry - 1 + r % 3, jz = rz - 1 + r / 3;
    if (jy < 1 || jy > nc_node || jz < 1 || jz > nc_node) continue;
    float* out = rho_c + ((long long)(jz - 1) * nc_node + (jy - 1)) * nc_node;
    for (int x = threadIdx.x; x < nc_node; x += TPB) {
      const float v = win[r * W + x + 2];                // cell jx = x + 1 sits at window index jx + 1
      if (v != 0.f) atomicAdd(&out[x], v);
    }
  }
}

// copy one rank's cube into the padded global FFT array (Nx+2, Ny, Nz) at offset (ox,oy,oz) — the pack_slab of
// fft_coarse.f90:4-54 for a replicated global mesh — and (for the rank's own cube) the DIAG sum of coarse_mesh.f90:31-43
__global__ void __launch_bounds__(TPB) cube_to_slab_kernel(const float* __restrict__ rho_c, float* __restrict__ slab, int nc_node, int Nx, int Ny,
                                                           int ox, int oy, int oz, double* __restrict__ sum) {
  const long long total = (long long)nc_node * nc_node * nc_node;
  double s = 0.0;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int x = (int)(t % nc_node), y = (int)((t / nc_node) % nc_node), z = (int)(t / ((long long)nc_node * nc_node));
    const float v = rho_c[t];
    s += (double)v;
    slab[((long long)(z + oz) * Ny + (y + oy)) * (Nx + 2) + (x + ox)] = v;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(sum, s);
}

// force_c(comp, 1:nc, 1:nc, 1:nc) = cube of the real-space result at offset (ox,oy,oz)   (coarse_force.f90:52 + unpack_slab)
__global__ void __launch_bounds__(TPB) slab_to_force_kernel(const float* __restrict__ real, long long pitch_x, long long pitch_y, int nc_node, int ox, int oy,
                                                            int oz, float* __restrict__ force_c, int comp) {
  const long long total = (long long)nc_node * nc_node * nc_node;
  const int fc = nc_node + 2;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int x = (int)(t % nc_node), y = (int)((t / nc_node) % nc_node), z = (int)(t / ((long long)nc_node * nc_node));
    force_c[(((long long)(z + 1) * fc + (y + 1)) * fc + (x + 1)) * 3 + comp] = real[((long long)(z + oz) * pitch_y + (y + oy)) * pitch_x + (x + ox)];
  }
}

// force_c(comp, 0:nc+1, 0:nc+1, 0:nc+1) <- global real-space component with periodic wrap: the rank's cube (coarse_force.f90:52 +
// unpack_slab) AND its one-cell halo (what coarse_force_buffer.f90:23-63 obtains from the six neighbours) in one gather.
__global__ void __launch_bounds__(TPB) extract_force_kernel(const float* __restrict__ real, int Nx, int Ny, int Nz, int nc_node, int cx, int cy, int cz,
                                                            float* __restrict__ force_c, int comp) {
  const int fc = nc_node + 2;
  const long long total = (long long)fc * fc * fc;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int i = (int)(t % fc), j = (int)((t / fc) % fc), k = (int)(t / ((long long)fc * fc));
    const int gx = (cx * nc_node + i - 1 + Nx) % Nx, gy = (cy * nc_node + j - 1 + Ny) % Ny, gz = (cz * nc_node + k - 1 + Nz) % Nz;
    force_c[t * 3 + comp] = real[((long long)gz * Ny + gy) * Nx + gx];
  }
}

// periodic self-halo for nodes_dim = 1: axis by axis so edges and corners propagate (coarse_force_buffer.f90:25-63)
__global__ void __launch_bounds__(TPB) halo_self_kernel(float* __restrict__ force_c, int nc_node, int axis) {
  const int fc = nc_node + 2;
  const long long total = (long long)fc * fc * 3;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int c = (int)(t % 3), u = (int)((t / 3) % fc), v = (int)(t / (3LL * fc));
    int i[3];
    auto idx = [&](int w) {
      if (axis == 0) { i[0] = w; i[1] = u; i[2] = v; } else if (axis == 1) { i[0] = u; i[1] = w; i[2] = v; } else { i[0] = u; i[1] = v; i[2] = w; }
      return (((long long)i[2] * fc + i[1]) * fc + i[0]) * 3 + c;
    };
    force_c[idx(nc_node + 1)] = force_c[idx(1)];
    force_c[idx(0)] = force_c[idx(nc_node)];
  }
}

// face pack / unpack for nodes_dim > 1 (the mpi_sendrecv_replace pairs of coarse_force_buffer.f90)
__global__ void __launch_bounds__(TPB) face_pack_kernel(const float* __restrict__ force_c, int nc_node, int axis, int layer, float* __restrict__ buf) {
  const int fc = nc_node + 2;
  const long long total = (long long)fc * fc * 3;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int c = (int)(t % 3), u = (int)((t / 3) % fc), v = (int)(t / (3LL * fc));
    int i0, i1, i2;
    if (axis == 0) { i0 = layer; i1 = u; i2 = v; } else if (axis == 1) { i0 = u; i1 = layer; i2 = v; } else { i0 = u; i1 = v; i2 = layer; }
    buf[t] = force_c[(((long long)i2 * fc + i1) * fc + i0) * 3 + c];
  }
}
__global__ void __launch_bounds__(TPB) face_unpack_kernel(float* __restrict__ force_c, int nc_node, int axis, int layer, const float* __restrict__ buf) {
  const int fc = nc_node + 2;
  const long long total = (long long)fc * fc * 3;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int c = (int)(t % 3), u = (int)((t / 3) % fc), v = (int)(t / (3LL * fc));
    int i0, i1, i2;
    if (axis == 0) { i0 = layer; i1 = u; i2 = v; } else if (axis == 1) { i0 = u; i1 = layer; i2 = v; } else { i0 = u; i1 = v; i2 = layer; }
    force_c[(((long long)i2 * fc + i1) * fc + i0) * 3 + c] = buf[t];
  }
}

// max |force_c| over 1..nc_node  (coarse_max_dt.f90:19-31)
__global__ void __launch_bounds__(TPB) force_max_kernel(const float* __restrict__ force_c, int nc_node, unsigned int* __restrict__ out_bits) {
  const long long total = (long long)nc_node * nc_node * nc_node;
  const int fc = nc_node + 2;
  float mx = 0.f;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int x = (int)(t % nc_node), y = (int)((t / nc_node) % nc_node), z = (int)(t / ((long long)nc_node * nc_node));
    const float* f = force_c + (((long long)(z + 1) * fc + (y + 1)) * fc + (x + 1)) * 3;
    mx = fmaxf(mx, sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]));
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomic_max_float_nonneg(out_bits, mx);
}

// coarse_velocity.f90:137-179: particles chained in coarse cells 1..nc_node, 8-point CIC gather, sequential adds
__global__ void __launch_bounds__(TPB) cic_kick_kernel(float* __restrict__ xv, const int* __restrict__ fstart, const float* __restrict__ force_c,
                                                       int H, int nc_buf, int nc_node, float a_mid, float G, float dt, int coarse_ngp) {
  const int ry = blockIdx.x % nc_node, rz = blockIdx.x / nc_node;
  const int cy = nc_buf + ry, cz = nc_buf + rz;
  const long long k0 = ((long long)(cz * H + cy) * H + nc_buf) * 64;
  const int s0 = fstart[k0], s1 = fstart[k0 + (long long)nc_node * 64];
  const int fc = nc_node + 2;
  const float agd = (a_mid * G) * dt;
  for (int i = s0 + threadIdx.x; i < s1; i += TPB) {
    float2* p = reinterpret_cast<float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    float2 b = p[1], c = p[2];
    int ix, iy, iz; float dx1, dx2, dy1, dy2, dz1, dz2;
    cic_setup(a.x, coarse_ngp, ix, dx1, dx2);
    cic_setup(a.y, coarse_ngp, iy, dy1, dy2);
    cic_setup(b.x, coarse_ngp, iz, dz1, dz2);
    float vx = b.y, vy = c.x, vz = c.y;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int jx = ix + (q & 1), jy = iy + ((q >> 1) & 1), jz = iz + (q >> 2);
      const float dV = ((agd * ((q & 1) ? dx2 : dx1)) * (((q >> 1) & 1) ? dy2 : dy1)) * ((q >> 2) ? dz2 : dz1);
      const float* f = force_c + (((long long)jz * fc + jy) * fc + jx) * 3;
      vx += f[0] * dV; vy += f[1] * dV; vz += f[2] * dV;
    }
    b.y = vx; c.x = vy; c.y = vz;
    p[1] = b; p[2] = c;
  }
}

// coarse kick fused with delete_particles' compaction (delete_particles.f90:14-50): the physical particles are exactly the coarse-cell rows
// 1..nc_node of the sorted array, so the kick's row loop can write each kicked record straight to its compacted place (rowoff = exclusive scan
// of the row lengths) instead of updating it in place and copying it afterwards: one read and one write of the particle array instead of two.
__global__ void __launch_bounds__(TPB) cic_kick_compact_kernel(const float* __restrict__ xv, const int64_t* __restrict__ pid_in, const int* __restrict__ fstart,
                                                               const int* __restrict__ rowoff, const float* __restrict__ force_c, int H, int nc_buf, int nc_node,
                                                               float a_mid, float G, float dt, int coarse_ngp, int kick, float* __restrict__ xv_out,
                                                               int64_t* __restrict__ pid_out) {
  const int ry = blockIdx.x % nc_node, rz = blockIdx.x / nc_node;
  const int cy = nc_buf + ry, cz = nc_buf + rz;
  const long long k0 = ((long long)(cz * H + cy) * H + nc_buf) * 64;
  const int s0 = fstart[k0], s1 = fstart[k0 + (long long)nc_node * 64];
  const int dst0 = rowoff[blockIdx.x] - s0;
  const int fc = nc_node + 2;
  const float agd = (a_mid * G) * dt;
  for (int i = s0 + threadIdx.x; i < s1; i += TPB) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    float2 b = p[1], c = p[2];
    if (kick) {
      int ix, iy, iz; float dx1, dx2, dy1, dy2, dz1, dz2;
      cic_setup(a.x, coarse_ngp, ix, dx1, dx2);
      cic_setup(a.y, coarse_ngp, iy, dy1, dy2);
      cic_setup(b.x, coarse_ngp, iz, dz1, dz2);
      float vx = b.y, vy = c.x, vz = c.y;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int jx = ix + (q & 1), jy = iy + ((q >> 1) & 1), jz = iz + (q >> 2);
        const float dV = ((agd * ((q & 1) ? dx2 : dx1)) * (((q >> 1) & 1) ? dy2 : dy1)) * ((q >> 2) ? dz2 : dz1);
        const float* f = force_c + (((long long)jz * fc + jy) * fc + jx) * 3;
        vx += f[0] * dV; vy += f[1] * dV; vz += f[2] * dV;
      }
      b.y = vx; c.x = vy; c.y = vz;
    }
    float2* o = reinterpret_cast<float2*>(xv_out) + 3LL * (dst0 + i);
    o[0] = a; o[1] = b; o[2] = c;
    if (pid_in) pid_out[dst0 + i] = pid_in[i];
  }
}

}  // namespace coarse
