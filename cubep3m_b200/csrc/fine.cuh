// Fine-mesh stages of the tile loop (particle_mesh_threaded.f90:84-368), one tile at a time:
//   deposit  : NGP density from the fine-cell occupancy table built by the cell sort (no atomics)
//   kick     : NGP force lookup + velocity update of the particles of the physical tile
//   force max: max |force_f|^2 over the cropped force cube
#pragma once
#include "common.cuh"
#include "particles.cuh"

namespace fine {

constexpr int TPB = 256;

// key of fine cell (gx,gy,gz) in the extended node frame (see part::make_key)
__device__ __forceinline__ long long cell_key(int gx, int gy, int gz, int H) {
  return ((long long)(((gz >> 2) * H + (gy >> 2))) * H + (gx >> 2)) * 64 + (((gz & 3) << 4) | ((gy & 3) << 2) | (gx & 3));
}

// NGP: rho_f(i1) += mass_p for every particle chained in coarse cells cic_l..cic_h (particle_mesh_threaded.f90:120-151).
// Those cells cover tile-local fine cells [4, n-5] (0-based) on every axis; everything else stays 0 (:100).
// One thread produces 4 consecutive x cells (= one coarse cell's x-row: 5 consecutive fstart entries).
// Also accumulates the DIAG mass sum over the physical part (:167-173) and the deposited-particle count.
__global__ void __launch_bounds__(TPB) ngp_density_kernel(const int* __restrict__ fstart, float* __restrict__ rho, int n, int b, int m, int H,
                                                          int tx, int ty, int tz, float mass_p, double* __restrict__ sum_phys,
                                                          int* __restrict__ tile_count) {
  const int nq = (n + 2 + 3) / 4;             // float4 groups per padded row (n+2 floats; n%4==0 -> (n+4)/4 groups, last half-used)
  const long long total = (long long)nq * n * n;
  double msum = 0.0;
  int pcount = 0;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int xq = (int)(t % nq);
    const long long r = t / nq;
    const int y = (int)(r % n), z = (int)(r / n);
    const int x0 = xq * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const bool yz_in = (y >= 4 && y <= n - 5 && z >= 4 && z <= n - 5);
    if (yz_in && x0 >= 4 && x0 <= n - 8) {
      const int gx = x0 + tx * m, gy = y + ty * m, gz = z + tz * m;
      const long long k = cell_key(gx, gy, gz, H);      // gx % 4 == 0 because m % 4 == 0
      const int s0 = fstart[k], s1 = fstart[k + 1], s2 = fstart[k + 2], s3 = fstart[k + 3], s4 = fstart[k + 4];
      const int c0 = s1 - s0, c1 = s2 - s1, c2 = s3 - s2, c3 = s4 - s3;
      v[0] = mass_p * (float)c0; v[1] = mass_p * (float)c1; v[2] = mass_p * (float)c2; v[3] = mass_p * (float)c3;
      pcount += c0 + c1 + c2 + c3;
      if (y >= b && y < n - b && z >= b && z < n - b && x0 >= b && x0 < n - b) msum += (double)v[0] + (double)v[1] + (double)v[2] + (double)v[3];
    }
    float* row = rho + ((long long)z * n + y) * (n + 2);
    if (x0 + 3 < n + 2) {
      *reinterpret_cast<float2*>(row + x0) = make_float2(v[0], v[1]);
      *reinterpret_cast<float2*>(row + x0 + 2) = make_float2(v[2], v[3]);
    } else {
      for (int q = 0; q < 4; ++q) if (x0 + q < n + 2) row[x0 + q] = v[q];
    }
  }
  msum = warp_sum_d(msum);
  pcount = warp_sum_i(pcount);
  if ((threadIdx.x & 31) == 0) {
    if (msum != 0.0) atomicAdd(sum_phys, msum);
    if (pcount) atomicAdd(tile_count, pcount);
  }
}

// Exactness fix-up of the NGP deposit. The reference bins with i1 = floor(fl(x + offset)) + 1 where offset = nf_buf - tile*m
// (particle_mesh_threaded.f90:134-143); the occupancy table bins with floor(x) + nf_buf. The two differ only when the fp32
// sum x + offset rounds up across an integer, i.e. for the few listed candidates; move their mass to the reference's cell.
__global__ void __launch_bounds__(TPB) ngp_fixup_kernel(const float* __restrict__ cand, const int* __restrict__ n_cand_ptr, int cand_cap,
                                                        float* __restrict__ rho, int n, int b, int m, int tx, int ty, int tz, float mass_p,
                                                        double* __restrict__ sum_phys) {
  const int nc = min(*n_cand_ptr, cand_cap);
  const float offx = (float)(b - tx * m), offy = (float)(b - ty * m), offz = (float)(b - tz * m);
  for (int i = blockIdx.x * TPB + threadIdx.x; i < nc; i += gridDim.x * TPB) {
    const float x = cand[3 * i], y = cand[3 * i + 1], z = cand[3 * i + 2];
    const int kx = (int)floorf(x) + b - tx * m, ky = (int)floorf(y) + b - ty * m, kz = (int)floorf(z) + b - tz * m;
    if (kx < 4 || kx > n - 5 || ky < 4 || ky > n - 5 || kz < 4 || kz > n - 5) continue;   // not deposited into this tile
    const int rx = (int)floorf(__fadd_rn(x, offx)), ry = (int)floorf(__fadd_rn(y, offy)), rz = (int)floorf(__fadd_rn(z, offz));
    if (rx == kx && ry == ky && rz == kz) continue;
    atomicAdd(&rho[((long long)kz * n + ky) * (n + 2) + kx], -mass_p);
    atomicAdd(&rho[((long long)rz * n + ry) * (n + 2) + rx], mass_p);
    const bool pk = kx >= b && kx < n - b && ky >= b && ky < n - b && kz >= b && kz < n - b;
    const bool pr = rx >= b && rx < n - b && ry >= b && ry < n - b && rz >= b && rz < n - b;
    if (pk != pr) atomicAdd(sum_phys, pr ? (double)mass_p : -(double)mass_p);
  }
}

// ---- fine CIC (the reference's non -DNGP builds: fine_cic_mass.f90:12-42, fine_cic_mass_buffer.f90, particle_mesh_threaded.f90:123-124,154-160).
// Gather formulation on the cell-sorted array: grid point g receives (g+1-x) from the particles of cell g and (x-(g-1)) from those
// of cell g-1 on each axis, i.e. it visits the 8 cells (g-1..g)^3. No atomics, deterministic for a given array order; the boundary
// routine's bounds checks are implicit (grid points outside [0,n) do not exist). Particles outside the tile's coarse-cell range
// cic_l..cic_h = tile-local fine cells [0, n-1] never reach it.
__global__ void __launch_bounds__(TPB) cic_density_kernel(const float* __restrict__ xv, const int* __restrict__ fstart, float* __restrict__ rho, int n, int b,
                                                          int m, int H, int tx, int ty, int tz, float mass_p, double* __restrict__ sum_phys,
                                                          int* __restrict__ tile_count) {
  const long long total = (long long)n * n * n;
  const float offx = (float)(b - tx * m), offy = (float)(b - ty * m), offz = (float)(b - tz * m);
  double msum = 0.0;
  int pcount = 0;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int x = (int)(t % n), y = (int)((t / n) % n), z = (int)(t / ((long long)n * n));
    float acc = 0.f;
    for (int dz = -1; dz <= 0; ++dz)
      for (int dy = -1; dy <= 0; ++dy)
        for (int dx = -1; dx <= 0; ++dx) {
          const int cx = x + dx, cy = y + dy, cz = z + dz;
          if (cx < 0 || cy < 0 || cz < 0) continue;
          const long long k = cell_key(cx + tx * m, cy + ty * m, cz + tz * m, H);
          const int s0 = fstart[k], s1 = fstart[k + 1];
          if (dx == 0 && dy == 0 && dz == 0) pcount += s1 - s0;
          for (int i = s0; i < s1; ++i) {
            const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
            const float2 a = p[0];
            const float pz = p[1].x;
            const float xt = __fadd_rn(a.x, offx), yt = __fadd_rn(a.y, offy), zt = __fadd_rn(pz, offz);
            // weights as the reference forms them: dx1 = i1 - x (lower point), dx2 = 1 - dx1 (upper point); mass_p folded into the x weight
            const float dx1 = (float)(cx + 1) - xt, dy1 = (float)(cy + 1) - yt, dz1 = (float)(cz + 1) - zt;
            const float wx = mass_p * (dx == 0 ? dx1 : 1.0f - dx1), wy = (dy == 0 ? dy1 : 1.0f - dy1), wz = (dz == 0 ? dz1 : 1.0f - dz1);
            acc += (wx * wy) * wz;
          }
        }
    rho[((long long)z * n + y) * (n + 2) + x] = acc;
    if (x >= b && x < n - b && y >= b && y < n - b && z >= b && z < n - b) msum += (double)acc;
  }
  msum = warp_sum_d(msum);
  pcount = warp_sum_i(pcount);
  if ((threadIdx.x & 31) == 0) {
    if (msum != 0.0) atomicAdd(sum_phys, msum);
    if (pcount && tile_count) atomicAdd(tile_count, pcount);
  }
}

// The same deposit as a SCATTER on the cell-sorted array (the north star's formulation): one CTA per coarse x-row (cy, cz) of the tile — a contiguous range of
// the sorted array whose particles reach only the fine rows y in [4 cy, 4 cy + 4], z in [4 cz, 4 cz + 4] — accumulates the eight weights of every particle
// (formed exactly as fine_cic_mass.f90:17-41 forms them, from i1 = floor(fl(x + offset))) with shared-memory atomics into a 5 x 5-row staging tile and flushes the
// tile with coalesced global atomic adds (rows are shared with the neighbouring CTAs through the +1 spill). Cells outside [0, n) are dropped as
// fine_cic_mass_buffer.f90 drops them. rho must be zero on entry. Summation order differs from the chain order by fp32 rounding only. The gather kernel above
// (deterministic, no atomics) walks 8 cells' particle lists per grid point and is 7x slower at the mean density (0.67 ms per 304^3 tile).
__global__ void __launch_bounds__(TPB) cic_scatter_kernel(const float* __restrict__ xv, const int* __restrict__ fstart, float* __restrict__ rho, int n, int b, int m,
                                                          int H, int tx, int ty, int tz, float mass_p, double* __restrict__ sum_phys, int* __restrict__ tile_count) {
  extern __shared__ float stile[];                       // [5 z][5 y][n + 2]
  const int ncr = n >> 2, n2 = n + 2;
  const int ry = blockIdx.x % ncr, rz = blockIdx.x / ncr;
  for (int t = threadIdx.x; t < 25 * n2; t += TPB) stile[t] = 0.f;
  const long long rk = (long long)((rz + tz * (m >> 2)) * H + (ry + ty * (m >> 2))) * H + tx * (m >> 2);
  const int s0 = fstart[rk * 64], s1 = fstart[(rk + ncr - 1) * 64 + 64];
  const float offx = (float)(b - tx * m), offy = (float)(b - ty * m), offz = (float)(b - tz * m);
  __syncthreads();
  for (int i = s0 + threadIdx.x; i < s1; i += TPB) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    const float xt = __fadd_rn(a.x, offx), yt = __fadd_rn(a.y, offy), zt = __fadd_rn(p[1].x, offz);
    const int ix = (int)floorf(xt), iy = (int)floorf(yt), iz = (int)floorf(zt);             // i1 - 1
    const float dx1 = __fmul_rn(mass_p, (float)(ix + 1) - xt), dx2 = __fmul_rn(mass_p, 1.0f - ((float)(ix + 1) - xt));
    const float dy1 = (float)(iy + 1) - yt, dy2 = 1.0f - dy1, dz1 = (float)(iz + 1) - zt, dz2 = 1.0f - dz1;
    const int ly = iy - 4 * ry, lz = iz - 4 * rz;       // 0..3 (4 only when fl(y + offset) rounded up to the next coarse cell: then the +1 weight is exactly 0)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int jx = ix + (q & 1), jy = ly + ((q >> 1) & 1), jz = lz + (q >> 2);
      if (jx < n && jy < 5 && jz < 5) {
        const float w = __fmul_rn(__fmul_rn((q & 1) ? dx2 : dx1, ((q >> 1) & 1) ? dy2 : dy1), (q >> 2) ? dz2 : dz1);
        atomicAdd(&stile[(jz * 5 + jy) * n2 + jx], w);
      }
    }
  }
  __syncthreads();
  double msum = 0.0;
  for (int t = threadIdx.x; t < 25 * n; t += TPB) {
    const int r = t / n, x = t - r * n;
    const int y = 4 * ry + r % 5, z = 4 * rz + r / 5;
    if (y < n && z < n) {
      const float v = stile[r * n2 + x];
      if (v != 0.f) {
        atomicAdd(&rho[((long long)z * n + y) * n2 + x], v);
        if (x >= b && x < n - b && y >= b && y < n - b && z >= b && z < n - b) msum += (double)v;
      }
    }
  }
  msum = warp_sum_d(msum);
  if ((threadIdx.x & 31) == 0 && msum != 0.0) atomicAdd(sum_phys, msum);
  if (threadIdx.x == 0 && tile_count && s1 > s0) atomicAdd(tile_count, s1 - s0);
}

// CIC force interpolation + kick, particle_mesh_threaded.f90:289-316 (sequential adds in the reference's corner order)
__global__ void __launch_bounds__(TPB) cic_fine_kick_kernel(float* __restrict__ xv, const int* __restrict__ fstart, const float* __restrict__ fx,
                                                            const float* __restrict__ fy, const float* __restrict__ fz, int H, int nc_buf, int nc_tile,
                                                            int b, int m, int fdim, int tx, int ty, int tz, float a_mid, float G, float dt) {
  const int ry = blockIdx.x % nc_tile, rz = blockIdx.x / nc_tile;
  const int cy = nc_buf + ty * nc_tile + ry, cz = nc_buf + tz * nc_tile + rz, cx0 = nc_buf + tx * nc_tile;
  const long long k0 = ((long long)(cz * H + cy) * H + cx0) * 64;
  const int s0 = fstart[k0], s1 = fstart[k0 + (long long)nc_tile * 64];
  const float offx = (float)b - (float)(tx * m), offy = (float)b - (float)(ty * m), offz = (float)b - (float)(tz * m);
  const float agd = (a_mid * G) * dt;
  for (int i = s0 + threadIdx.x; i < s1; i += TPB) {
    float2* p = reinterpret_cast<float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    float2 bb = p[1], c = p[2];
    const float xt = __fadd_rn(a.x, offx), yt = __fadd_rn(a.y, offy), zt = __fadd_rn(bb.x, offz);
    const int i1x = (int)floorf(xt) + 1, i1y = (int)floorf(yt) + 1, i1z = (int)floorf(zt) + 1;     // 1-based lower grid point
    const float dx1 = (float)i1x - xt, dy1 = (float)i1y - yt, dz1 = (float)i1z - zt;
    const float dx2 = 1.0f - dx1, dy2 = 1.0f - dy1, dz2 = 1.0f - dz1;
    float vx = bb.y, vy = c.x, vz = c.y;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int jx = i1x + (q & 1) - (b - 1), jy = i1y + ((q >> 1) & 1) - (b - 1), jz = i1z + (q >> 2) - (b - 1);   // index into the cropped cube
      const float dVc = ((agd * ((q & 1) ? dx2 : dx1)) * (((q >> 1) & 1) ? dy2 : dy1)) * ((q >> 2) ? dz2 : dz1);
      const long long o = ((long long)jz * fdim + jy) * fdim + jx;
      vx += fx[o] * dVc; vy += fy[o] * dVc; vz += fz[o] * dVc;
    }
    bb.y = vx; c.x = vy; c.y = vz;
    p[1] = bb; p[2] = c;
  }
}

// ---- the same fix-up for the fused density + r2c kernel (fft_x_r2c_ngp): per-tile lists of (from-cell, to-cell) moves.
// Built once per step for all tiles; also applies the corresponding +-mass_p to the DIAG mass sum.
constexpr int DELTA_CAP = 2048;   // moves per tile (expected: a few tens)
__global__ void __launch_bounds__(TPB) build_tile_deltas_kernel(const float* __restrict__ cand, const int* __restrict__ n_cand_ptr, int cand_cap, int n, int b,
                                                                int m, int T, float mass_p, int2* __restrict__ deltas, int* __restrict__ ndelta,
                                                                double* __restrict__ sum_phys, int* __restrict__ overflow) {
  const int nc = min(*n_cand_ptr, cand_cap);
  for (int i = blockIdx.x * TPB + threadIdx.x; i < nc; i += gridDim.x * TPB) {
    const float p[3] = {cand[3 * i], cand[3 * i + 1], cand[3 * i + 2]};
    int g[3];
    for (int a = 0; a < 3; ++a) g[a] = (int)floorf(p[a]) + b;
    // tiles whose NGP deposit range [4, n-5] contains the particle: t with 4 <= g - t*m <= n-5
    int tl[3], th[3];
    for (int a = 0; a < 3; ++a) {
      tl[a] = max(0, (g[a] - (n - 5) + m - 1) / m);
      th[a] = min(T - 1, (g[a] - 4) / m);
      if (g[a] - 4 < 0) th[a] = -1;
    }
    for (int tz = tl[2]; tz <= th[2]; ++tz)
      for (int ty = tl[1]; ty <= th[1]; ++ty)
        for (int tx = tl[0]; tx <= th[0]; ++tx) {
          const int t3[3] = {tx, ty, tz};
          int k[3], r[3];
          bool diff = false, ok = true;
          for (int a = 0; a < 3; ++a) {
            k[a] = g[a] - t3[a] * m;
            ok &= (k[a] >= 4 && k[a] <= n - 5);
            r[a] = (int)floorf(__fadd_rn(p[a], (float)(b - t3[a] * m)));
            diff |= (r[a] != k[a]);
          }
          if (!ok || !diff) continue;
          const int tile = (tz * T + ty) * T + tx;
          const int slot = atomicAdd(&ndelta[tile], 1);
          if (slot >= DELTA_CAP) { atomicOr(overflow, 8); continue; }   // never silently truncated: the step returns ECAPACITY
          deltas[(long long)tile * DELTA_CAP + slot] = make_int2((k[2] * n + k[1]) * n + k[0], (r[2] * n + r[1]) * n + r[0]);
          const bool pk = k[0] >= b && k[0] < n - b && k[1] >= b && k[1] < n - b && k[2] >= b && k[2] < n - b;
          const bool pr = r[0] >= b && r[0] < n - b && r[1] >= b && r[1] < n - b && r[2] >= b && r[2] < n - b;
          if (pk != pr) atomicAdd(sum_phys, pr ? (double)mass_p : -(double)mass_p);
        }
  }
}

// particles deposited into each tile's padded mesh = particles chained in coarse cells cic_l..cic_h (:120-121): one CTA per tile
__global__ void __launch_bounds__(TPB) tile_counts_kernel(const int* __restrict__ fstart, int H, int nc_buf, int nc_tile, int T, int* __restrict__ counts) {
  const int tile = blockIdx.x;
  const int tx = tile % T, ty = (tile / T) % T, tz = tile / (T * T);
  const int span = nc_tile + 2 * nc_buf - 2;             // coarse cells cic_l..cic_h
  const int c0x = tx * nc_tile + 1, c0y = ty * nc_tile + 1, c0z = tz * nc_tile + 1;   // 0-based index of cic_l inside the hoc range
  int s = 0;
  for (int r = threadIdx.x; r < span * span; r += TPB) {
    const int cy = c0y + r % span, cz = c0z + r / span;
    const long long k0 = ((long long)(cz * H + cy) * H + c0x) * 64;
    s += fstart[k0 + (long long)span * 64] - fstart[k0];
  }
  s = warp_sum_i(s);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&counts[tile], s);
}

// max over the cropped force cube of fx^2+fy^2+fz^2 (particle_mesh_threaded.f90:208-223)
__global__ void __launch_bounds__(TPB) force_max_kernel(const float* __restrict__ fx, const float* __restrict__ fy, const float* __restrict__ fz,
                                                        long long n, unsigned int* __restrict__ out_bits) {
  float mx = 0.f;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += (long long)gridDim.x * TPB) {
    const float a = fx[i], b = fy[i], c = fz[i];
    mx = fmaxf(mx, a * a + b * b + c * c);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomic_max_float_nonneg(out_bits, mx);
}

// NGP kick (particle_mesh_threaded.f90:227-266): for every particle chained in the coarse cells of the physical tile
//   x_t = x + (nf_buf - tile*m); i1 = floor(x_t)+1; v += force_f(:, i1) * a_mid * G * dt
// One CTA per (cy,cz) coarse row of the tile = one contiguous range of the sorted array.
__global__ void __launch_bounds__(TPB) ngp_kick_kernel(float* __restrict__ xv, const int* __restrict__ fstart, const float* __restrict__ fx,
                                                       const float* __restrict__ fy, const float* __restrict__ fz, int H, int nc_buf, int nc_tile,
                                                       int b, int m, int fdim, int tx, int ty, int tz, float a_mid, float G, float dt) {
  const int ry = blockIdx.x % nc_tile, rz = blockIdx.x / nc_tile;
  const int cy = nc_buf + ty * nc_tile + ry, cz = nc_buf + tz * nc_tile + rz, cx0 = nc_buf + tx * nc_tile;
  const long long k0 = ((long long)(cz * H + cy) * H + cx0) * 64;
  const int s0 = fstart[k0], s1 = fstart[k0 + (long long)nc_tile * 64];
  const float offx = (float)b - (float)(tx * m), offy = (float)b - (float)(ty * m), offz = (float)b - (float)(tz * m);
  // two particles per iteration, every load of both issued before the first dependent use (the kernel is a chain of three dependent
  // DRAM/L2 latencies: fstart -> record -> force; ncu: long-scoreboard 88 %)
  auto cell = [&](float2 a, float z) {
    // 0-based index into the cropped cube: (i1 - (nf_buf-1)) with i1 = floor(x_t)+1
    const int ix = (int)floorf(__fadd_rn(a.x, offx)) + 2 - b;
    const int iy = (int)floorf(__fadd_rn(a.y, offy)) + 2 - b;
    const int iz = (int)floorf(__fadd_rn(z, offz)) + 2 - b;
    return ((long long)iz * fdim + iy) * fdim + ix;
  };
  for (int i = s0 + threadIdx.x; i < s1; i += 2 * TPB) {
    const bool two = i + TPB < s1;
    float2* p = reinterpret_cast<float2*>(xv) + 3LL * i;
    float2* p2 = two ? p + 3LL * TPB : p;
    const float2 a = p[0], a2 = p2[0];
    float2 bb = p[1], c = p[2], bb2 = p2[1], c2 = p2[2];
    const long long q = cell(a, bb.x), q2 = cell(a2, bb2.x);
    const float f0 = fx[q], f1 = fy[q], f2 = fz[q], g0 = fx[q2], g1 = fy[q2], g2 = fz[q2];
    bb.y += ((f0 * a_mid) * G) * dt;
    c.x += ((f1 * a_mid) * G) * dt;
    c.y += ((f2 * a_mid) * G) * dt;
    p[1] = bb; p[2] = c;
    if (two) {
      bb2.y += ((g0 * a_mid) * G) * dt;
      c2.x += ((g1 * a_mid) * G) * dt;
      c2.y += ((g2 * a_mid) * G) * dt;
      p2[1] = bb2; p2[2] = c2;
    }
  }
}

}  // namespace fine
