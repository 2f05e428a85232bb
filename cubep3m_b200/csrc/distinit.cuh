// dist_init on the device (utils/dist_init/dist_init_dm.f90): Zel'dovich initial conditions generated in place
// in the library's particle array, single rank, nc = nf_physical_dim <= the largest transform length of the library.
//   noise_kernel     : Gaussian white noise, Box-Muller on consecutive pairs along x (x2 cos x1, x2 sin x1 with x1 = 2 pi u_i, x2 = sqrt(-2 ln u_{i+1}),
//                      :617-629) from Philox4x32-10 keyed by (seed, cell pair) — the Fortran random_number stream is compiler-specific and not
//                      reproducible anyway; a host noise field can be supplied instead (parity tests)
//   (forward FFT with the library's own kernels, :653)
//   phi_k_kernel     : delta(k) = sqrt(Delta^2(2 pi kr / box) / (4 pi kr^3) nc^3) noise(k), zero mode 0 (:685-712), Delta^2 interpolated log-log
//                      in the caller's table by bisection (`power`, :1270-1299); times the potential kernel K(k) = -4 pi / sum_d (2 sin(pi k_d/nc))^2
//                      (:814-832); the short-range `correct_kernel` patch (:850-903) is not applied (as in the host twin, cubep3m_b200/ic.py)
//   (inverse FFT, :958-969, 1/nc^3 folded into the store)
//   particles_kernel : lattice point i1 = 2(i-1)+1, dis_d = (phi(i1 - e_d) - phi(i1 + e_d)) / 2 / (4 pi), x = dis + (i1 - 0.5), v = dis vfactor(a)
//                      (:1011-1036), written in dist_init's file order (i fastest); `reps` > 1 replicates the periodic box reps^3 times
#pragma once
#include "common.cuh"

namespace distinit {

constexpr int TPB = 256;

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// padded real mesh (nc+2, nc, nc); one thread per PAIR of consecutive x cells
__global__ void __launch_bounds__(TPB) noise_kernel(float* __restrict__ mesh, int nc, unsigned long long seed) {
  const long long npair = (long long)(nc / 2) * nc * nc;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < npair; t += (long long)gridDim.x * TPB) {
    const int xp = (int)(t % (nc / 2));
    const long long r = t / (nc / 2);
    unsigned rnd[4];
    philox4x32_10((unsigned)t, (unsigned)(t >> 32), 0x5eedu, 0u, (unsigned)seed, (unsigned)(seed >> 32), rnd);
    const float u1 = ((float)(rnd[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(rnd[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
    const float x1 = 6.2831853f * u1, x2 = sqrtf(-2.0f * logf(u2));
    float s, c;
    sincosf(x1, &s, &c);
    float* row = mesh + r * (nc + 2);
    row[2 * xp] = x2 * c; row[2 * xp + 1] = x2 * s;
  }
}

// log-log linear interpolation in (kt, d2t)[nt] (ascending k), clamped at the ends  (dist_init_dm.f90:1270-1299)
__device__ __forceinline__ float power_interp(const float* __restrict__ kt, const float* __restrict__ d2t, int nt, float k) {
  if (k <= kt[0]) return d2t[0];
  if (k >= kt[nt - 1]) return d2t[nt - 1];
  int lo = 0, hi = nt - 1;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (kt[mid] <= k) lo = mid; else hi = mid; }
  const float w = (logf(k) - logf(kt[lo])) / (logf(kt[hi]) - logf(kt[lo]));
  return expf(logf(d2t[lo]) * (1.0f - w) + logf(d2t[hi]) * w);
}

// spectrum (hc, nc, nc) complex in place: noise(k) -> phi(k)
__global__ void __launch_bounds__(TPB) phi_k_kernel(float2* __restrict__ spec, int nc, float box, const float* __restrict__ kt, const float* __restrict__ d2t, int nt) {
  const int hc = nc / 2 + 1;
  const long long total = (long long)hc * nc * nc;
  const float pi = 3.14159265358979f;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int kx = (int)(t % hc);
    const long long r = t / hc;
    const int jy = (int)(r % nc), jz = (int)(r / nc);
    const int ky = jy < nc / 2 ? jy : jy - nc, kz = jz < nc / 2 ? jz : jz - nc;
    float2 v = spec[t];
    if (kx == 0 && ky == 0 && kz == 0) v = make_float2(0.f, 0.f);
    else {
      const float kr = sqrtf((float)(kx * kx + ky * ky + kz * kz));
      const float d2 = power_interp(kt, d2t, nt, 2.0f * pi * kr / box);
      const float amp = sqrtf(d2 / (4.0f * pi * kr * kr * kr) * ((float)nc * (float)nc * (float)nc));
      const float sx = 2.0f * sinf(pi * (float)kx / (float)nc), sy = 2.0f * sinf(pi * (float)ky / (float)nc), sz = 2.0f * sinf(pi * (float)kz / (float)nc);
      const float kern = -4.0f * pi / (sx * sx + sy * sy + sz * sz);
      const float f = amp * kern;
      v.x *= f; v.y *= f;
    }
    spec[t] = v;
  }
}

// phi: real mesh (pitch nc), already scaled by 1/nc^3
__global__ void __launch_bounds__(TPB) particles_kernel(const float* __restrict__ phi, int nc, int reps, float vf, float* __restrict__ xv) {
  const int npd = nc / 2;
  const long long nbox = (long long)npd * npd * npd, total = nbox * reps * reps * reps;
  const float fourpi = 12.5663706f;
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const long long rep = t / nbox, q = t - rep * nbox;
    const int i = (int)(q % npd), j = (int)((q / npd) % npd), k = (int)(q / ((long long)npd * npd));
    const int c[3] = {2 * i, 2 * j, 2 * k};                         // 0-based lattice cell (Fortran i1 = 2(i-1)+1)
    const int rx = (int)(rep % reps), ry = (int)((rep / reps) % reps), rz = (int)(rep / (reps * reps));
    const int ro[3] = {rx, ry, rz};
    float out[6];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      int lo[3] = {c[0], c[1], c[2]}, hi[3] = {c[0], c[1], c[2]};
      lo[d] = (c[d] - 1 + nc) % nc; hi[d] = (c[d] + 1) % nc;
      const float pl = phi[((long long)lo[2] * nc + lo[1]) * nc + lo[0]], ph = phi[((long long)hi[2] * nc + hi[1]) * nc + hi[0]];
      const float dis = (pl - ph) / 2.0f / fourpi;
      out[d] = dis + ((float)(c[d] + 1) - 0.5f) + (float)(ro[d] * nc);
      out[3 + d] = dis * vf;
    }
    float2* p = reinterpret_cast<float2*>(xv) + 3 * t;
    p[0] = make_float2(out[0], out[1]); p[1] = make_float2(out[2], out[3]); p[2] = make_float2(out[4], out[5]);
  }
}

}  // namespace distinit
