// cic_power on the device (utils/cic_power/cic_power.f90): the 0.1 % P(k) acceptance metric.
//   cic_density_kernel : CIC deposit of the resident particles on the global nc^3 mesh, cell-centre convention x - 0.5, particle mass
//                        (nc/np)^3, after undoing the shake offset as checkpoint.f90:92 does                     cic_power.f90:1496-1539
//   (forward FFT with the library's own kernels; the "- 1" of delta = rho - 1 only changes the k = 0 mode, which is skipped)  :918
//   shell_bin_kernel   : per mode skip the redundant half of the kx = 0 plane (:1583-1584), pow = |delta_k / nc^3|^2 / (sinc sinc sinc)^4
//                        (:1590-1615), shells k1 = ceil(kr), k2 = k1 + 1 with weights w1 = k1 - kr, w2 = 1 - w1 (NGP build: 1, 0)   :1586-1589
// Single rank (the mesh lives on one GPU); the host part (lib.cu) turns the four shell sums into k, Delta^2, sigma (:1649-1660).
#pragma once
#include "common.cuh"

namespace power {

constexpr int TPB = 256;

__device__ __forceinline__ float unshake(float x, float s, float ncf) {   // np.mod(x - s, nc) in float32
  float r = fmodf(__fsub_rn(x, s), ncf);
  if (r < 0.f) r = __fadd_rn(r, ncf);
  return r;
}

__global__ void __launch_bounds__(TPB) cic_density_kernel(const float* __restrict__ xv, int np, int nc, float sx, float sy, float sz, float mp,
                                                          float* __restrict__ rho) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= np) return;
  const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
  const float2 a = p[0];
  const float z = p[1].x;
  const float ncf = (float)nc;
  const float q[3] = {unshake(a.x, sx, ncf) - 0.5f, unshake(a.y, sy, ncf) - 0.5f, unshake(z, sz, ncf) - 0.5f};
  int i1[3], i2[3];
  float d1[3], d2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float f = floorf(q[c]);
    d2[c] = q[c] - f; d1[c] = 1.0f - d2[c];
    int j = (int)f % nc; if (j < 0) j += nc;
    i1[c] = j; i2[c] = (j + 1 == nc) ? 0 : j + 1;
  }
  const long long P = nc + 2;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int jx = (c & 1) ? i2[0] : i1[0], jy = (c & 2) ? i2[1] : i1[1], jz = (c & 4) ? i2[2] : i1[2];
    const float w = mp * ((c & 1) ? d2[0] : d1[0]) * ((c & 2) ? d2[1] : d1[1]) * ((c & 4) ? d2[2] : d1[2]);
    atomicAdd(&rho[((long long)jz * nc + jy) * P + jx], w);
  }
}

// The same deposit for a mesh that is distributed over the ranks as z-slabs (rank q owns planes [q*zs, (q+1)*zs), pitch nc+2): the particle's global
// cell decides the owner and the eight contributions are added with atomics straight into the owners' slabs over NVLink (cubep3m_b200_cic_power, several ranks;
// the reference exchanges a one-cell buffer layer instead, cic_power.f90:1081-1495). (ox,oy,oz) = this rank's offset in the global box, in cells.
__global__ void __launch_bounds__(TPB) cic_density_dist_kernel(const float* __restrict__ xv, int np, int nc, float ox, float oy, float oz, float sx, float sy, float sz,
                                                               float mp, PeerTable P, long long off_slab, int zs) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= np) return;
  const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
  const float2 a = p[0];
  const float z = p[1].x;
  const float ncf = (float)nc;
  const float q[3] = {unshake(__fadd_rn(a.x, ox), sx, ncf) - 0.5f, unshake(__fadd_rn(a.y, oy), sy, ncf) - 0.5f, unshake(__fadd_rn(z, oz), sz, ncf) - 0.5f};
  int i1[3], i2[3];
  float d1[3], d2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float f = floorf(q[c]);
    d2[c] = q[c] - f; d1[c] = 1.0f - d2[c];
    int j = (int)f % nc; if (j < 0) j += nc;
    i1[c] = j; i2[c] = (j + 1 == nc) ? 0 : j + 1;
  }
  const long long Pp = nc + 2;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int jx = (c & 1) ? i2[0] : i1[0], jy = (c & 2) ? i2[1] : i1[1], jz = (c & 4) ? i2[2] : i1[2];
    const float w = mp * ((c & 1) ? d2[0] : d1[0]) * ((c & 2) ? d2[1] : d1[1]) * ((c & 4) ? d2[2] : d1[2]);
    float* slab = P.base[jz / zs] + off_slab;
    atomicAdd(&slab[((long long)(jz % zs) * nc + jy) * Pp + jx], w);
  }
}

// sums: [4][nb+2] doubles = P, P2, W, K per shell. sinc4[i] = (sin(pi k/nc)/(pi k/nc))^4 for the signed frequency index i (natural order).
// dk holds [jz][yl][kx] with yl = 0..ny_loc-1 the rows y0 + yl of the global mesh (the whole mesh: ny_loc = nc, y0 = 0). An axis transformed by the
// four-step passes of bigfft.cuh is stored digit-transposed: memory index m holds frequency m / n1 + n2 (m % n1) (n1 = 0: natural order).
__global__ void __launch_bounds__(TPB) shell_bin_kernel(const float2* __restrict__ dk, int nc, int ny_loc, int y0, int n1, int n2, const double* __restrict__ sinc4,
                                                        int ngp_binning, int nb, double* __restrict__ sums) {
  extern __shared__ double sh[];                 // [4][nb+2]
  const int hc = nc / 2 + 1, stride = nb + 2;
  for (int t = threadIdx.x; t < 4 * stride; t += TPB) sh[t] = 0.0;
  __syncthreads();
  const long long total = (long long)hc * ny_loc * nc;
  const double inv = 1.0 / ((double)nc * nc * nc);
  for (long long t = (long long)blockIdx.x * TPB + threadIdx.x; t < total; t += (long long)gridDim.x * TPB) {
    const int kx = (int)(t % hc);
    const long long r = t / hc;
    const int my = (int)(r % ny_loc) + y0, mz = (int)(r / ny_loc);
    const int jy = n1 ? my / n1 + n2 * (my % n1) : my, jz = n1 ? mz / n1 + n2 * (mz % n1) : mz;
    const int ky = jy < nc / 2 ? jy : jy - nc, kz = jz < nc / 2 ? jz : jz - nc;
    if (kx == 0 && !(ky > 0 || (ky == 0 && kz > 0))) continue;     // k = 0 and the redundant half of the kx = 0 plane
    const double kr = sqrt((double)kx * kx + (double)ky * ky + (double)kz * kz);
    const float2 v = dk[t];
    const double re = (double)v.x * inv, im = (double)v.y * inv;
    const double pw = (re * re + im * im) / (sinc4[kx] * sinc4[jy] * sinc4[jz]);
    const int k1 = (int)ceil(kr), k2 = k1 + 1;
    const double w1 = ngp_binning ? 1.0 : (double)k1 - kr, w2 = 1.0 - w1;
    if (k1 <= nb) {
      atomicAdd(&sh[k1], w1 * pw); atomicAdd(&sh[stride + k1], w1 * pw * pw); atomicAdd(&sh[2 * stride + k1], w1); atomicAdd(&sh[3 * stride + k1], w1 * kr);
    }
    if (k2 <= nb && w2 != 0.0) {
      atomicAdd(&sh[k2], w2 * pw); atomicAdd(&sh[stride + k2], w2 * pw * pw); atomicAdd(&sh[2 * stride + k2], w2); atomicAdd(&sh[3 * stride + k2], w2 * kr);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 4 * stride; t += TPB) if (sh[t] != 0.0) atomicAdd(&sums[t], sh[t]);
}

}  // namespace power
