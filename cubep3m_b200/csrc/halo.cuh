// Density-peak pass of the halo finder (halofind.f90:564-672, find_halos): on a halofind step the driver has just run link_list and
// particle_pass (cubepm.f90:193-198); per tile the reference deposits the fine density (fine_ngp_mass with -DNGPH, else fine_cic_mass, over the
// coarse cells cic_l..cic_h of :597-598), then scans the physical cells for local maxima of the 3^3 neighbourhood above den_peak_cutoff
// (:620-627), refines each peak's position by parabolic interpolation per axis (:634-655, para_inter :770-778) and accumulates the clumping sums
// cftmass / cftmass2 (:621-622). Here: the deposit kernels of fine.cuh write the tile's density, peak_kernel does the scan (one thread per
// physical cell; the candidates are appended through a warp-aggregated counter, the host sorts them by density as :676-679 does).
#pragma once
#include "common.cuh"

namespace halo {

constexpr int TPB = 256;

// para_inter (halofind.f90:770-778) with the reference's operation order, no contraction
__device__ __forceinline__ float para_inter(float x1, float x2, float x3, float f1, float f2, float f3) {
  const float a = __fsub_rn(x2, x1), c = __fsub_rn(x2, x3), d23 = __fsub_rn(f2, f3), d21 = __fsub_rn(f2, f1);
  const float num = __fsub_rn(__fmul_rn(__fmul_rn(a, a), d23), __fmul_rn(__fmul_rn(c, c), d21));
  const float den = __fsub_rn(__fmul_rn(a, d23), __fmul_rn(c, d21));
  return __fsub_rn(x2, __fdiv_rn(__fmul_rn(0.5f, num), den));
}

struct Peak { int i, j, k, tile; float den, px, py, pz; };

// rho: the tile's density, [n][n][n + 2] floats (x fastest). Physical cells: 0-based [b, n - b) on every axis.
__global__ void __launch_bounds__(TPB) peak_kernel(const float* __restrict__ rho, int n, int b, int m, float cutoff, int para, int tile, float offx, float offy,
                                                   float offz, Peak* __restrict__ peaks, int cap, int* __restrict__ n_peaks, double* __restrict__ cft) {
  const long long total = (long long)m * m * m;
  const int n2 = n + 2;
  double s1 = 0.0, s2 = 0.0;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < (total + 31) / 32 * 32; t += (long long)gridDim.x * TPB) {
    bool is_peak = false;
    int i = 0, j = 0, k = 0;
    float c = 0.f;
    const float* p = rho;
    if (t < total) {
      i = (int)(t % m) + b; j = (int)((t / m) % m) + b; k = (int)(t / ((long long)m * m)) + b;
      p = rho + ((long long)k * n + j) * n2 + i;
      c = *p;
      s1 += (double)c; s2 += (double)__fmul_rn(c, c);                   // :621-622 (rho**2 in real(4), summed in real(8))
      if (c > cutoff) {
        float mx = c;
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) mx = fmaxf(mx, p[((long long)dz * n + dy) * n2 + dx]);
        is_peak = (mx == c);                                             // :623-624
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, is_peak);
    if (bal) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(n_peaks, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0) + __popc(bal & ((1u << lane) - 1));
      if (is_peak && base < cap) {
        Peak q;
        q.i = i + 1; q.j = j + 1; q.k = k + 1; q.tile = tile; q.den = c;  // 1-based tile-local cell as ipeak (:631)
        float px = (float)(i + 1) - 0.5f, py = (float)(j + 1) - 0.5f, pz = (float)(k + 1) - 0.5f;
        if (para) {                                                       // :634-655
          px = para_inter((float)i - 0.5f, (float)(i + 1) - 0.5f, (float)(i + 2) - 0.5f, p[-1], c, p[1]);
          py = para_inter((float)j - 0.5f, (float)(j + 1) - 0.5f, (float)(j + 2) - 0.5f, p[-n2], c, p[n2]);
          pz = para_inter((float)k - 0.5f, (float)(k + 1) - 0.5f, (float)(k + 2) - 0.5f, p[-(long long)n * n2], c, p[(long long)n * n2]);
        }
        q.px = __fadd_rn(px, offx); q.py = __fadd_rn(py, offy); q.pz = __fadd_rn(pz, offz);   // + offset as halo_pos (:723)
        peaks[base] = q;
      }
    }
  }
  s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
  if ((threadIdx.x & 31) == 0 && (s1 != 0.0 || s2 != 0.0)) { atomicAdd(&cft[0], s1); atomicAdd(&cft[1], s2); }
}

}  // namespace halo
