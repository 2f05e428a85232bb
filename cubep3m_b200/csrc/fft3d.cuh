// 3-D real<->complex FFT passes on a padded (N+2, N, N) array (Fortran order = C order [z][y][x]).
// Same data layout and normalisation as the reference's FFTW calls (fft_fine.f90:47-51): forward is
// unnormalised exp(-i..), backward unnormalised; the 1/N^3 is folded into the last backward pass.
//
//   pass X  (contiguous axis): two real rows are packed into one complex sequence (a + i b), one
//           length-N complex FFT gives both half-spectra; 16 columns = 32 rows per CTA.
//   pass Y/Z (strided axes): a CTA owns 16 consecutive kx (128 contiguous bytes per element) and the
//           full extent of the transformed axis.
//   The Green's-function multiply  rho_hat * (i * kern_f(d))  (particle_mesh_threaded.f90:183-192) is fused
//   into the load of the first backward pass; the crop to force_f (:202-203) and the 1/n^3 (fft_fine.f90:51)
//   are fused into the store of the last one.
#pragma once
#include "common.cuh"
#include "fft_smem.cuh"

namespace fftk {

struct Smem {
  float *re0, *im0, *re1, *im1;
  float2* tw;
};
template <int N> __device__ __forceinline__ Smem carve(unsigned char* raw, const float2* __restrict__ tw_g) {
  Smem s;
  s.re0 = reinterpret_cast<float*>(raw);
  s.im0 = s.re0 + N * LXP;
  s.re1 = s.im0 + N * LXP;
  s.im1 = s.re1 + N * LXP;
  s.tw = reinterpret_cast<float2*>(s.im1 + N * LXP);
  for (int t = threadIdx.x; t < N; t += NT) s.tw[t] = tw_g[t];
  return s;
}

// ---- pass X forward: real rows -> half spectra, in place. rows are consecutive with pitch N+2 floats.
template <int N>
__global__ void __launch_bounds__(NT) fft_x_r2c(float* __restrict__ data, int nrows, const float2* __restrict__ tw_g) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem s = carve<N>(raw, tw_g);
  constexpr int P = N + 2;
  const long long r0 = (long long)blockIdx.x * (2 * LX);
  for (int q = threadIdx.x; q < 2 * LX * N; q += NT) {
    const int row = q / N, x = q - row * N;
    const long long gr = r0 + row;
    const float v = (gr < nrows) ? data[gr * P + x] : 0.f;
    ((row & 1) ? s.im0 : s.re0)[x * LXP + (row >> 1)] = v;
  }
  __syncthreads();
  fft_columns<N, false>(s.re0, s.im0, s.re1, s.im1, s.tw);
  const float* zr = result_buffer<N>() ? s.re1 : s.re0;
  const float* zi = result_buffer<N>() ? s.im1 : s.im0;
  for (int q = threadIdx.x; q < 2 * LX * P; q += NT) {
    const int row = q / P, f = q - row * P;
    const long long gr = r0 + row;
    if (gr >= nrows) continue;
    const int k = f >> 1, col = row >> 1;
    const int km = (k == 0) ? 0 : N - k;
    const float ar = zr[k * LXP + col], ai = zi[k * LXP + col], br = zr[km * LXP + col], bi = zi[km * LXP + col];
    float v;
    if ((row & 1) == 0) v = (f & 1) ? 0.5f * (ai - bi) : 0.5f * (ar + br);     // A = (Z[k] + conj Z[N-k]) / 2
    else                v = (f & 1) ? -0.5f * (ar - br) : 0.5f * (ai + bi);    // B = (Z[k] - conj Z[N-k]) / (2i)
    data[gr * P + f] = v;
  }
}

// ---- pass X forward fused with the NGP mass assignment (particle_mesh_threaded.f90:100-151): the tile's density row is not read from
// memory but produced on the fly from the fine-cell occupancy table, rho = mass_p * (fstart[k+1] - fstart[k]) for tile-local cells
// [4, n-5] and 0 elsewhere; `deltas` moves the mass of the few particles whose reference cell floor(fl(x+offset)) differs (fine.cuh).
// One warp handles one row at a time (lanes along x): no per-element divisions, coalesced gathers within a coarse cell.
template <int N>
__global__ void __launch_bounds__(NT) fft_x_r2c_ngp(float2* __restrict__ data, int cp, const int* __restrict__ fstart, int H, int b, int ox, int oy, int oz,
                                                    float mass_p, const int2* __restrict__ deltas, const int* __restrict__ ndelta_ptr, int delta_cap,
                                                    double* __restrict__ sum_phys, const float2* __restrict__ tw_g) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem s = carve<N>(raw, tw_g);
  constexpr int P = N + 2;
  const int r0 = blockIdx.x * (2 * LX);          // first row (= z*N + y) of this CTA; N*N is a multiple of 32
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double msum = 0.0;
  for (int row = warp; row < 2 * LX; row += NT / 32) {
    float rsum = 0.f;                               // per-row partial in fp32 (a handful of small multiples of mass_p), then fp64
    const int gr = r0 + row;
    const int y = gr % N, z = gr / N;
    float* dst = ((row & 1) ? s.im0 : s.re0) + (row >> 1);
    const bool yz_in = (y >= 4 && y <= N - 5 && z >= 4 && z <= N - 5);
    const bool yz_phys = (y >= b && y < N - b && z >= b && z < N - b);
    const int gy = y + oy, gz = z + oz;
    const long long rowkey = ((long long)((gz >> 2) * H + (gy >> 2)) * H) * 64 + (((gz & 3) << 4) | ((gy & 3) << 2));
    for (int x = lane; x < N; x += 32) {
      float v = 0.f;
      if (yz_in && x >= 4 && x <= N - 5) {
        const int gx = x + ox;
        const long long k = rowkey + (long long)(gx >> 2) * 64 + (gx & 3);
        v = mass_p * (float)(fstart[k + 1] - fstart[k]);
        if (yz_phys && x >= b && x < N - b) rsum += v;
      }
      dst[x * LXP] = v;
    }
    msum += (double)rsum;
  }
  msum = warp_sum_d(msum);
  if (lane == 0 && msum != 0.0) atomicAdd(sum_phys, msum);
  __syncthreads();
  const int nd = min(*ndelta_ptr, delta_cap);
  if (nd > 0) {
    for (int i = threadIdx.x; i < nd; i += NT) {
      const int2 d = deltas[i];
      const int c[2] = {d.x, d.y};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int x = c[u] % N, gr = c[u] / N;     // gr = z*N + y
        const int row = gr - r0;
        if (row >= 0 && row < 2 * LX) atomicAdd(&((row & 1) ? s.im0 : s.re0)[x * LXP + (row >> 1)], u == 0 ? -mass_p : mass_p);
      }
    }
    __syncthreads();
  }
  fft_columns<N, false>(s.re0, s.im0, s.re1, s.im1, s.tw);
  const float* zr = result_buffer<N>() ? s.re1 : s.re0;
  const float* zi = result_buffer<N>() ? s.im1 : s.im0;
  // untangle the two real rows of each column: A = (Z[k] + conj Z[N-k]) / 2, B = (Z[k] - conj Z[N-k]) / (2i); one warp per row
  for (int row = warp; row < 2 * LX; row += NT / 32) {
    float2* orow = data + (long long)(r0 + row) * cp;
    const int col = row >> 1;
    for (int k = lane; k < N / 2 + 1; k += 32) {
      const int km = (k == 0) ? 0 : N - k;
      const float ar = zr[k * LXP + col], ai = zi[k * LXP + col], br = zr[km * LXP + col], bi = zi[km * LXP + col];
      orow[k] = (row & 1) ? make_float2(0.5f * (ai + bi), -0.5f * (ar - br)) : make_float2(0.5f * (ar + br), 0.5f * (ai - bi));
    }
  }
}

// ---- pass Y / Z: strided complex columns. Persistent CTAs: each loops over work items (kx block, outer index, batch) and
// prefetches the NEXT item's column block into registers while the current one is transformed, so the global/L2 latency
// (the dominant stall in the first version: ~50 % long_scoreboard) overlaps the butterflies. Shared-memory layout is AoS
// float2 [N][16] (fft_smem.cuh: stage_aos) — one 64-bit access per point and constant offsets from one base per thread.
// element e of column c of item (bx, by, bz): in[bz*bstride + (by + outer0)*ostride + bx*LX + e*estride + c]
// MUL: multiply the loaded value by i*kern[e*kes + (by+outer0)*kos + kx]   (Z pass: e=z, outer=y; kern = one component)
// stores only elements e in [elo, ehi]
template <int N> struct StridedCfg {
  static constexpr int EPT = (N + (NT / LX) - 1) / (NT / LX);   // elements per thread (11 for N=176)
  static constexpr bool PREFETCH = EPT <= 19;                  // keep the prefetch registers bounded (N <= 304)
};
#ifndef FFTK_MINB
#define FFTK_MINB 2
#endif

template <int N, bool INV, bool MUL>
__global__ void __launch_bounds__(NT, StridedCfg<N>::PREFETCH ? FFTK_MINB : 1) fft_strided(const float2* __restrict__ in, float2* __restrict__ out, int hc,
                                                  long long estride, long long ostride, int outer0, int nouter, int nbatch,
                                                  const float* __restrict__ kern, long long kes, long long kos,
                                                  int elo, int ehi, const float2* __restrict__ tw_g, long long bstride) {
  extern __shared__ __align__(16) unsigned char raw[];
  float2* b0 = reinterpret_cast<float2*>(raw);
  float2* b1 = b0 + N * LX;
  float2* tw = b1 + N * LX;
  for (int t = threadIdx.x; t < N; t += NT) tw[t] = tw_g[t];
  constexpr int EPT = StridedCfg<N>::EPT;
  constexpr bool PF = StridedCfg<N>::PREFETCH;
  constexpr int ES = NT / LX;                      // element step between a thread's consecutive elements
  const int nbx = (hc + LX - 1) / LX;
  const long long total = (long long)nbx * nouter * nbatch;
  const int col = threadIdx.x % LX, e0 = threadIdx.x / LX;
  const int sidx = e0 * LX + col;
  const long long gstep = (long long)ES * estride, kstep = (long long)ES * kes;
  float2 pf[PF ? EPT : 1];
  auto decode = [&](long long item, long long& base, long long& kbase, bool& colok) {
    const int bx = (int)(item % nbx);
    const long long t = item / nbx;
    const int outer = (int)(t % nouter) + outer0, bz = (int)(t / nouter);
    const int kx0 = bx * LX;
    base = (long long)bz * bstride + (long long)outer * ostride + kx0 + (long long)e0 * estride + col;
    kbase = (long long)outer * kos + kx0 + (long long)e0 * kes + col;
    colok = (kx0 + col) < hc;
  };
  auto fetch = [&](long long base, long long kbase, bool colok, int it) -> float2 {
    float2 v = make_float2(0.f, 0.f);
    if (colok && e0 + it * ES < N) {
      v = in[base + it * gstep];
      if (MUL) {
        const float kv = kern[kbase + it * kstep];
        v = make_float2(-v.y * kv, v.x * kv);
      }
    }
    return v;
  };
  long long item = blockIdx.x;
  long long base = 0, kbase = 0; bool colok = false;
  if (item < total) {
    decode(item, base, kbase, colok);
    if (PF) {
#pragma unroll
      for (int it = 0; it < EPT; ++it) pf[it] = fetch(base, kbase, colok, it);
    }
  }
  const float2* res = result_buffer<N>() ? b1 : b0;
  while (item < total) {
#pragma unroll
    for (int it = 0; it < EPT; ++it) {
      const float2 v = PF ? pf[it] : fetch(base, kbase, colok, it);
      if (e0 + it * ES < N) b0[sidx + it * ES * LX] = v;
    }
    __syncthreads();
    const long long cur_base = base; const bool cur_ok = colok;
    const long long next = item + gridDim.x;
    if (next < total) {
      decode(next, base, kbase, colok);
      if (PF) {
#pragma unroll
        for (int it = 0; it < EPT; ++it) pf[it] = fetch(base, kbase, colok, it);
      }
    }
    fft_columns_aos<N, INV>(b0, b1, tw);
    if (cur_ok) {
      float2* op = out + cur_base;
#pragma unroll
      for (int it = 0; it < EPT; ++it) {
        const int e = e0 + it * ES;
        if (e >= elo && e <= ehi) op[it * gstep] = res[sidx + it * ES * LX];
      }
    }
    __syncthreads();
    item = next;
  }
}

// ---- fused pass Z: forward FFT along z, then for each of the three force components: multiply by i*kern_f(comp)
// (particle_mesh_threaded.f90:183-192) and inverse FFT along z, storing only the cropped z range. One load of the spectrum
// column block feeds four transforms; the forward-z result never goes back to memory. Persistent CTAs with register prefetch of
// the next block's spectrum and of the next component's Green's function values. AoS float2 shared-memory layout.
template <int N>
__global__ void __launch_bounds__(NT, StridedCfg<N>::PREFETCH ? FFTK_MINB : 1) fft_z_sandwich(const float2* __restrict__ spec, float2* __restrict__ g, long long gstride, int hc, int cp, int ny,
                                                     const float* __restrict__ kern, long long kstride, int kp, int elo, int ehi,
                                                     const float2* __restrict__ tw_g) {
  extern __shared__ __align__(16) unsigned char raw[];
  float2* b0 = reinterpret_cast<float2*>(raw);
  float2* b1 = b0 + N * LX;
  float2* b2 = b1 + N * LX;
  float2* tw = b2 + N * LX;
  for (int t = threadIdx.x; t < N; t += NT) tw[t] = tw_g[t];
  constexpr int EPT = StridedCfg<N>::EPT;
  constexpr bool PF = StridedCfg<N>::PREFETCH;
  constexpr int ES = NT / LX;
  const int nbx = (hc + LX - 1) / LX;
  const long long total = (long long)nbx * ny;
  const long long estride = (long long)ny * cp;
  const long long gstep = (long long)ES * estride;
  const int col = threadIdx.x % LX, e0 = threadIdx.x / LX;
  const int sidx = e0 * LX + col;
  // S = forward result, X = the other of {b0,b1}; the inverse transforms ping-pong between X and b2
  const float2* S = result_buffer<N>() ? b1 : b0;
  float2* X = result_buffer<N>() ? b0 : b1;
  const float2* Rr = result_buffer<N>() ? b2 : X;
  float2 pf[PF ? EPT : 1];
  float kf[PF ? EPT : 1];
  auto decode = [&](long long item, long long& base, bool& colok) {
    const int bx = (int)(item % nbx), y = (int)(item / nbx);
    base = (long long)y * cp + bx * LX + (long long)e0 * estride + col;
    colok = (bx * LX + col) < hc;
  };
  auto fetch = [&](long long base, bool colok, int it) -> float2 {
    return (colok && e0 + it * ES < N) ? spec[base + it * gstep] : make_float2(0.f, 0.f);
  };
  const long long kestride = (long long)ny * kp, kgstep = (long long)ES * kestride;   // the Green's function table has row pitch kp
  auto fetchk = [&](long long base, bool colok, int comp, int it) -> float {
    const long long y = (base % estride) / cp, kx = (base % estride) % cp;   // base = y*cp + kx + e0*estride
    return (colok && e0 + it * ES < N) ? kern[(long long)comp * kstride + y * kp + kx + (long long)e0 * kestride + it * kgstep] : 0.f;
  };
  long long item = blockIdx.x, base = 0; bool colok = false;
  if (item < total) {
    decode(item, base, colok);
    if (PF) {
#pragma unroll
      for (int it = 0; it < EPT; ++it) pf[it] = fetch(base, colok, it);
    }
  }
  while (item < total) {
#pragma unroll
    for (int it = 0; it < EPT; ++it) {
      const float2 v = PF ? pf[it] : fetch(base, colok, it);
      if (e0 + it * ES < N) b0[sidx + it * ES * LX] = v;
    }
    __syncthreads();
    const long long cur_base = base; const bool cur_ok = colok;
    if (PF) {
#pragma unroll
      for (int it = 0; it < EPT; ++it) kf[it] = fetchk(cur_base, cur_ok, 0, it);
    }
    fft_columns_aos<N, false>(b0, b1, tw);
    const long long next = item + gridDim.x;
#pragma unroll 1
    for (int comp = 0; comp < 3; ++comp) {
#pragma unroll
      for (int it = 0; it < EPT; ++it) {
        const float kv = PF ? kf[it] : fetchk(cur_base, cur_ok, comp, it);
        if (e0 + it * ES < N) { const float2 sv = S[sidx + it * ES * LX]; X[sidx + it * ES * LX] = make_float2(-sv.y * kv, sv.x * kv); }
      }
      __syncthreads();
      if (PF) {
        if (comp < 2) {
#pragma unroll
          for (int it = 0; it < EPT; ++it) kf[it] = fetchk(cur_base, cur_ok, comp + 1, it);
        } else if (next < total) {
          decode(next, base, colok);
#pragma unroll
          for (int it = 0; it < EPT; ++it) pf[it] = fetch(base, colok, it);
        }
      }
      fft_columns_aos<N, true>(X, b2, tw);
      if (cur_ok) {
        float2* go = g + (long long)comp * gstride + cur_base;
#pragma unroll
        for (int it = 0; it < EPT; ++it) {
          const int e = e0 + it * ES;
          if (e >= elo && e <= ehi) go[it * gstep] = Rr[sidx + it * ES * LX];
        }
      }
      __syncthreads();
    }
    if (!PF && next < total) decode(next, base, colok);
    item = next;
  }
}
constexpr size_t smem_bytes_sandwich(int n) { return smem_bytes_aos(n, 3); }

// ---- pass X backward: half spectra -> real rows with crop + scale.
// Row index space: ridx in [0, cnt_z*cnt_y): zc = ridx / cnt_y, yc = ridx % cnt_y, source row (z = zc+lo_z, y = yc+lo_y) of an
// array with ny_src rows per plane. Output: out[(zc*opitch_y + yc)*opitch_x + xc], xc in [0,cnt_x) <- x = xc + lo_x.
template <int N>
__global__ void __launch_bounds__(NT) fft_x_c2r(const float2* __restrict__ in, float* __restrict__ out, int lo_x, int cnt_x, int lo_y, int cnt_y,
                                                int lo_z, int cnt_z, int ny_src, long long opitch_x, long long opitch_y, float scale,
                                                const float2* __restrict__ tw_g, long long in_bstride, long long out_bstride) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem s = carve<N>(raw, tw_g);
  in += (long long)blockIdx.y * in_bstride;     // batch (force component) offset
  out += (long long)blockIdx.y * out_bstride;
  constexpr int HC = N / 2 + 1;
  const long long nrows = (long long)cnt_z * cnt_y;
  const long long r0 = (long long)blockIdx.x * (2 * LX);
  // per-row source / destination offsets once per CTA (the runtime divisions are too expensive per element)
  __shared__ long long srow[2 * LX], drow[2 * LX];
  if (threadIdx.x < 2 * LX) {
    const long long ridx = r0 + threadIdx.x;
    long long so = -1, dof = -1;
    if (ridx < nrows) {
      const int zc = (int)(ridx / cnt_y), yc = (int)(ridx - (long long)zc * cnt_y);
      so = ((long long)(zc + lo_z) * ny_src + (yc + lo_y)) * HC;
      dof = ((long long)zc * opitch_y + yc) * opitch_x;
    }
    srow[threadIdx.x] = so; drow[threadIdx.x] = dof;
  }
  __syncthreads();
  // stage A (even rows) into buffer 0, B (odd rows) into buffer 1; one warp per row, lanes along k
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = warp; row < 2 * LX; row += NT / 32) {
    const long long so = srow[row];
    float* dr = ((row & 1) ? s.re1 : s.re0) + (row >> 1);
    float* di = ((row & 1) ? s.im1 : s.im0) + (row >> 1);
    for (int k = lane; k < HC; k += 32) {
      const float2 v = (so >= 0) ? in[so + k] : make_float2(0.f, 0.f);
      dr[k * LXP] = v.x; di[k * LXP] = v.y;
    }
  }
  __syncthreads();
  // Z[k] = A[k] + i B[k];  Z[N-k] = conj(A[k]) + i conj(B[k]);  imaginary parts of the k=0 and k=N/2 bins are dropped (c2r)
  for (int q = threadIdx.x; q < LX * HC; q += NT) {
    const int k = q / LX, col = q - k * LX;
    float ar = s.re0[k * LXP + col], ai = s.im0[k * LXP + col], br = s.re1[k * LXP + col], bi = s.im1[k * LXP + col];
    if (k == 0 || 2 * k == N) { ai = 0.f; bi = 0.f; }
    s.re0[k * LXP + col] = ar - bi;
    s.im0[k * LXP + col] = ai + br;
    if (k != 0 && 2 * k != N) {
      s.re0[(N - k) * LXP + col] = ar + bi;
      s.im0[(N - k) * LXP + col] = br - ai;
    }
  }
  __syncthreads();
  fft_columns<N, true>(s.re0, s.im0, s.re1, s.im1, s.tw);
  const float* zr = result_buffer<N>() ? s.re1 : s.re0;
  const float* zi = result_buffer<N>() ? s.im1 : s.im0;
  // one warp per output row at a time: lanes run along x (coalesced stores, stride-17 conflict-free smem reads)
  for (int row = warp; row < 2 * LX; row += NT / 32) {
    const long long dof = drow[row];
    if (dof < 0) continue;
    const float* src = (row & 1) ? zi : zr;
    const int col = row >> 1;
    for (int xc = lane; xc < cnt_x; xc += 32) out[dof + xc] = src[(xc + lo_x) * LXP + col] * scale;
  }
}

// ---- pass X backward for the three force components of a fine tile, fused with the crop, the 1/n^3 scale and the max |F|^2
// (particle_mesh_threaded.f90:202-223). One CTA owns 32 cropped rows and loops over the components; the next component's rows
// are prefetched into registers during the current FFT; |F|^2 is accumulated per output element in registers.
template <int N>
__global__ void __launch_bounds__(NT, 2) fft_x_c2r3(const float2* __restrict__ in, int cp, float* __restrict__ out, int lo, int cnt, long long in_bstride,
                                                    long long out_bstride, float scale, unsigned int* __restrict__ fmax_bits,
                                                    const float2* __restrict__ tw_g) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem s = carve<N>(raw, tw_g);
  constexpr int HC = N / 2 + 1;
  constexpr int KIT = (HC + 31) / 32;          // k iterations per lane
  constexpr int RPW = 2 * LX / (NT / 32);      // rows per warp (4)
  constexpr int XIT_MAX = (N + 31) / 32;
  const long long nrows = (long long)cnt * cnt;
  const long long r0 = (long long)blockIdx.x * (2 * LX);
  __shared__ long long srow[2 * LX], drow[2 * LX];
  if (threadIdx.x < 2 * LX) {
    const long long ridx = r0 + threadIdx.x;
    long long so = -1, dof = -1;
    if (ridx < nrows) {
      const int zc = (int)(ridx / cnt), yc = (int)(ridx - (long long)zc * cnt);
      so = ((long long)(zc + lo) * N + (yc + lo)) * cp;
      dof = ((long long)zc * cnt + yc) * cnt;
    }
    srow[threadIdx.x] = so; drow[threadIdx.x] = dof;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 pf[RPW * KIT];
  float fsq[RPW * XIT_MAX];
#pragma unroll
  for (int i = 0; i < RPW * XIT_MAX; ++i) fsq[i] = 0.f;
  auto prefetch = [&](int comp) {
    const float2* src = in + (long long)comp * in_bstride;
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) {
      const long long so = srow[warp + ri * (NT / 32)];
#pragma unroll
      for (int ki = 0; ki < KIT; ++ki) {
        const int k = lane + 32 * ki;
        pf[ri * KIT + ki] = (so >= 0 && k < HC) ? src[so + k] : make_float2(0.f, 0.f);
      }
    }
  };
  prefetch(0);
  const float* zr = result_buffer<N>() ? s.re1 : s.re0;
  const float* zi = result_buffer<N>() ? s.im1 : s.im0;
#pragma unroll 1
  for (int comp = 0; comp < 3; ++comp) {
    // stage A (even rows) into buffer 0, B (odd rows) into buffer 1
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) {
      const int row = warp + ri * (NT / 32);
      float* dr = ((row & 1) ? s.re1 : s.re0) + (row >> 1);
      float* di = ((row & 1) ? s.im1 : s.im0) + (row >> 1);
#pragma unroll
      for (int ki = 0; ki < KIT; ++ki) {
        const int k = lane + 32 * ki;
        if (k < HC) { dr[k * LXP] = pf[ri * KIT + ki].x; di[k * LXP] = pf[ri * KIT + ki].y; }
      }
    }
    __syncthreads();
    if (comp < 2) prefetch(comp + 1);
    // Z[k] = A[k] + i B[k];  Z[N-k] = conj(A[k]) + i conj(B[k]);  imaginary parts of the k=0 and k=N/2 bins are dropped (c2r)
    for (int q = threadIdx.x; q < LX * HC; q += NT) {
      const int k = q / LX, col = q - k * LX;
      float ar = s.re0[k * LXP + col], ai = s.im0[k * LXP + col], br = s.re1[k * LXP + col], bi = s.im1[k * LXP + col];
      if (k == 0 || 2 * k == N) { ai = 0.f; bi = 0.f; }
      s.re0[k * LXP + col] = ar - bi;
      s.im0[k * LXP + col] = ai + br;
      if (k != 0 && 2 * k != N) {
        s.re0[(N - k) * LXP + col] = ar + bi;
        s.im0[(N - k) * LXP + col] = br - ai;
      }
    }
    __syncthreads();
    fft_columns<N, true>(s.re0, s.im0, s.re1, s.im1, s.tw);
    float* o = out + (long long)comp * out_bstride;
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) {
      const int row = warp + ri * (NT / 32);
      const long long dof = drow[row];
      const float* src = (row & 1) ? zi : zr;
      const int col = row >> 1;
#pragma unroll
      for (int xi = 0; xi < XIT_MAX; ++xi) {
        const int xc = lane + 32 * xi;
        if (dof >= 0 && xc < cnt) {
          const float v = src[(xc + lo) * LXP + col] * scale;
          o[dof + xc] = v;
          fsq[ri * XIT_MAX + xi] = fmaf(v, v, fsq[ri * XIT_MAX + xi]);
        }
      }
    }
    __syncthreads();
  }
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < RPW * XIT_MAX; ++i) mx = fmaxf(mx, fsq[i]);
  mx = warp_max(mx);
  if (lane == 0 && mx > 0.f) atomic_max_float_nonneg(fmax_bits, mx);
}

}  // namespace fftk
#include "fft3d2.cuh"
namespace fftk {

// ------------------------------------------------------------------------------------------------
// host-side dispatch on N (each axis of a mesh may have its own length: the fine tile is cubic, the global coarse mesh of a
// (Dx,Dy,Dz) rank grid need not be)
// ------------------------------------------------------------------------------------------------
#define FFTK_FOR_ALL_N(X) X(16) X(32) X(48) X(64) X(80) X(112) X(128) X(176) X(256) X(304) X(512) X(560)

inline bool supported(int n) {
  switch (n) {
#define X(N) case N:
    FFTK_FOR_ALL_N(X)
#undef X
    return true;
  }
  return false;
}

// cudaFuncSetAttribute is per DEVICE: a process that opens contexts on several GPUs (cfg.local_gpu) must opt in on each of them
// (ADVICE r1). The flags are only written from the thread that owns the handle (the library is not re-entrant per handle);
// the cached occupancy numbers further down are the same on every B200 and therefore stay process-wide.
template <int N> int set_smem_attr() {
  static bool done_dev[64] = {false};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return 0;
  const int bytes = (int)smem_bytes(N);
  CK(cudaFuncSetAttribute(fft_x_r2c<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(cudaFuncSetAttribute(fft_x_c2r<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(cudaFuncSetAttribute(fft_x_c2r3<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));      // also the fine-CIC path of the two-stage sizes (odd row pitch)
  if constexpr (!Plan2<N>::ok) {      // the first-generation strided / fused kernels only exist for the sizes the two-stage plan cannot do (16, 512, 560)
    CK(cudaFuncSetAttribute(fft_x_r2c_ngp<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute((fft_strided<N, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute((fft_strided<N, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute((fft_strided<N, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(fft_z_sandwich<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_sandwich(N)));
  }
  if constexpr (Plan2<N>::ok) {
    const int b2 = (int)smem_bytes2(N, 2, LX);
    CK(cudaFuncSetAttribute((fft_strided2<N, false, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, b2));
    CK(cudaFuncSetAttribute((fft_strided2<N, true, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, b2));
    CK(cudaFuncSetAttribute((fft_strided2<N, true, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, b2));
    CK(cudaFuncSetAttribute((fft_strided2<N, false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, b2));
    CK(cudaFuncSetAttribute((fft_strided2<N, true, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, b2));
    CK(cudaFuncSetAttribute((fft_z_sandwich2<N, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_sandwich2(N)));
    CK(cudaFuncSetAttribute((fft_z_sandwich2<N, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_sandwich2(N)));
    CK(cudaFuncSetAttribute(fft_z_sandwich3<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sand3<N>::smem));
    CK(cudaFuncSetAttribute(fft_x_r2c_ngp2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes2(N, 1, XP)));
    CK(cudaFuncSetAttribute(fft_x_c2r3_v2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_c2r3_v2(N)));
    CK(cudaFuncSetAttribute(fft_x_c2r3_v3<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2R3<N>::smem));
    CK(cudaFuncSetAttribute(fft_x_c2r3_v4<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2R4<N>::smem));
  }
  done = true;
  return 0;
}

struct Mesh3 {            // a padded real mesh (nx+2, ny, nz) / complex (nx/2+1, ny, nz)
  int nx, ny, nz;
  const float2 *twx, *twy, *twz;
  int hc() const { return nx / 2 + 1; }
};

template <int N> int launch_x_r2c_t(cubep3m_b200_ctx* ctx, int kc, float* data, int nrows, const float2* tw) {
  if (int st = set_smem_attr<N>()) return st;
  LAUNCH(ctx, kc, fft_x_r2c<N>, dim3((nrows + 2 * LX - 1) / (2 * LX)), dim3(NT), (int)smem_bytes(N), data, nrows, tw);
  return 0;
}
struct NgpSource {            // what fft_x_r2c_ngp needs to produce the tile's density rows itself
  const int* fstart; int H, b, ox, oy, oz; float mass_p; const int2* deltas; const int* ndelta; int delta_cap; double* sum_phys;
};
template <int N> int launch_x_r2c_ngp_t(cubep3m_b200_ctx* ctx, int kc, float2* data, int cp, const NgpSource& g, const float2* tw) {
  if (int st = set_smem_attr<N>()) return st;
  static_assert((N * N) % (2 * LX) == 0, "rows per tile must be a multiple of the rows per CTA");
  if constexpr (Plan2<N>::ok && N % 4 == 0) {
    LAUNCH(ctx, kc, fft_x_r2c_ngp2<N>, dim3(N * N / (2 * LX)), dim3(Plan2<N>::NT), (int)smem_bytes2(N, 1, XP), data, cp, g.fstart, g.H, g.b, g.ox, g.oy, g.oz, g.mass_p,
           g.deltas, g.ndelta, g.delta_cap, g.sum_phys, tw);
    return 0;
  } else {
    LAUNCH(ctx, kc, fft_x_r2c_ngp<N>, dim3(N * N / (2 * LX)), dim3(NT), (int)smem_bytes(N), data, cp, g.fstart, g.H, g.b, g.ox, g.oy, g.oz, g.mass_p, g.deltas,
           g.ndelta, g.delta_cap, g.sum_phys, tw);
    return 0;
  }
}
template <int N> int launch_strided_t(cubep3m_b200_ctx* ctx, int kc, bool inv, const float2* in, float2* out, int hc, long long estride,
                                      long long ostride, int outer0, int nouter, const float* kern, long long kes, long long kos, int elo,
                                      int ehi, const float2* tw, int nbatch, long long bstride) {
  if (int st = set_smem_attr<N>()) return st;
  if constexpr (Plan2<N>::ok) {
    const int sm2 = (int)smem_bytes2(N, 2, LX);
    static int occ2[3] = {0, 0, 0};
    if (!occ2[0]) {
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2[0], (fft_strided2<N, false, false, false>), Plan2<N>::NT, sm2));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2[1], (fft_strided2<N, true, true, false>), Plan2<N>::NT, sm2));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2[2], (fft_strided2<N, true, false, false>), Plan2<N>::NT, sm2));
    }
    // 16-byte staging needs 128-byte aligned column blocks: even strides (the padded spectra of the fused fine-tile path)
    const bool a16 = !kern && estride % 2 == 0 && ostride % 2 == 0 && bstride % 2 == 0 && ((uintptr_t)in & 15) == 0 &&
                     std::min(estride, ostride) >= (long long)((hc + LX - 1) / LX) * LX;   // the row pitch covers whole 16-column blocks
    const long long total2 = (long long)((hc + LX - 1) / LX) * nouter * nbatch;
    auto grid2 = [&](int o) { return dim3((unsigned)std::min<long long>(total2, (long long)NUM_SMS * std::max(o, 1))); };
    const dim3 blk(Plan2<N>::NT);
    // 32-bit element offsets (loads) and 32-bit BYTE offsets (stage-B stores) inside the kernel: the largest offset touched must stay below 2^29 elements
    const long long span = (long long)(nbatch - 1) * bstride + (long long)(outer0 + nouter) * ostride + (long long)N * estride + hc;
    const long long kspan = kern ? (long long)(outer0 + nouter) * kos + (long long)N * kes + hc : 0;
    if constexpr (N == 32 || N == 64) {
      // arrays beyond 4 GB (bigfft.cuh: the four-step passes of cic_power's 1024^3 / 2048^3 meshes): forward, no multiply, 64-bit offsets
      if (span >= (1LL << 29) && !inv && !kern && total2 < (1LL << 31)) {
        static bool attr64 = false;
        if (!attr64) { CK(cudaFuncSetAttribute((fft_strided2<N, false, false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, sm2)); attr64 = true; }
        LAUNCH(ctx, kc, (fft_strided2<N, false, false, false, true>), grid2(occ2[0]), blk, sm2, in, out, hc, estride, ostride, outer0, nouter, nbatch, nullptr, 0, 0, elo, ehi, tw, bstride);
        return 0;
      }
    }
    if (span >= (1LL << 29) || kspan >= (1LL << 31) || total2 >= (1LL << 31)) return CUBEP3M_B200_EINVAL;
    const int es = (int)estride, os = (int)ostride, bs = (int)bstride, ke = (int)kes, ko = (int)kos;
    if (!inv && a16) LAUNCH(ctx, kc, (fft_strided2<N, false, false, true>), grid2(occ2[0]), blk, sm2, in, out, hc, es, os, outer0, nouter, nbatch, nullptr, 0, 0, elo, ehi, tw, bs);
    else if (!inv) LAUNCH(ctx, kc, (fft_strided2<N, false, false, false>), grid2(occ2[0]), blk, sm2, in, out, hc, es, os, outer0, nouter, nbatch, nullptr, 0, 0, elo, ehi, tw, bs);
    else if (kern) LAUNCH(ctx, kc, (fft_strided2<N, true, true, false>), grid2(occ2[1]), blk, sm2, in, out, hc, es, os, outer0, nouter, nbatch, kern, ke, ko, elo, ehi, tw, bs);
    else if (a16) LAUNCH(ctx, kc, (fft_strided2<N, true, false, true>), grid2(occ2[2]), blk, sm2, in, out, hc, es, os, outer0, nouter, nbatch, nullptr, 0, 0, elo, ehi, tw, bs);
    else LAUNCH(ctx, kc, (fft_strided2<N, true, false, false>), grid2(occ2[2]), blk, sm2, in, out, hc, es, os, outer0, nouter, nbatch, nullptr, 0, 0, elo, ehi, tw, bs);
    return 0;
  } else {
  const int sm = (int)smem_bytes_aos(N, 2);
  static int occ[3] = {0, 0, 0};    // resident CTAs per SM of the three instantiations
  if (!occ[0]) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], (fft_strided<N, false, false>), NT, sm));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], (fft_strided<N, true, true>), NT, sm));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], (fft_strided<N, true, false>), NT, sm));
  }
  const long long total = (long long)((hc + LX - 1) / LX) * nouter * nbatch;
  auto grid = [&](int o) { return dim3((unsigned)std::min<long long>(total, (long long)NUM_SMS * std::max(o, 1))); };
  if (!inv) LAUNCH(ctx, kc, (fft_strided<N, false, false>), grid(occ[0]), dim3(NT), sm, in, out, hc, estride, ostride, outer0, nouter, nbatch, nullptr, 0LL, 0LL, elo, ehi, tw, bstride);
  else if (kern) LAUNCH(ctx, kc, (fft_strided<N, true, true>), grid(occ[1]), dim3(NT), sm, in, out, hc, estride, ostride, outer0, nouter, nbatch, kern, kes, kos, elo, ehi, tw, bstride);
  else LAUNCH(ctx, kc, (fft_strided<N, true, false>), grid(occ[2]), dim3(NT), sm, in, out, hc, estride, ostride, outer0, nouter, nbatch, nullptr, 0LL, 0LL, elo, ehi, tw, bstride);
  return 0;
  }
}
template <int N> int launch_sandwich_t(cubep3m_b200_ctx* ctx, int kc, const float2* spec, float2* g, long long gstride, int hc, int cp, int ny, const float* kern,
                                       long long kstride, int kp, int elo, int ehi, const float2* tw) {
  if (int st = set_smem_attr<N>()) return st;
  if constexpr (Plan2<N>::ok) {
    const int sm2 = (int)smem_bytes_sandwich2(N);
    static int occ2 = 0;
    if (!occ2) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, (fft_z_sandwich2<N, false>), Plan2<N>::NT, sm2));
    const long long total2 = (long long)((hc + LX - 1) / LX) * ny;
    if (kp % 16 != 0 || 3 * gstride >= (1LL << 29)) return CUBEP3M_B200_EINVAL;   // stage-B stores use 32-bit byte offsets from g
    const dim3 grd((unsigned)std::min<long long>(total2, (long long)NUM_SMS * std::max(occ2, 1)));
    const bool a16 = cp % 2 == 0 && cp >= (hc + LX - 1) / LX * LX && ((uintptr_t)spec & 15) == 0;
    static const bool use_v3 = [] { const char* e = getenv("CUBEP3M_B200_SANDWICH"); return e && !strcmp(e, "v3"); }();   // A/B knob: 8-column items (measured: 144 vs 141 us per tile at n = 304, so off)
    if (a16 && use_v3) {
      static int occ3 = 0;
      if (!occ3) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, fft_z_sandwich3<N>, Sand3<N>::NT, (int)Sand3<N>::smem));
      const long long total3 = (long long)((hc + Sand3<N>::CW - 1) / Sand3<N>::CW) * ny;
      LAUNCH(ctx, kc, fft_z_sandwich3<N>, dim3((unsigned)std::min<long long>(total3, (long long)NUM_SMS * std::max(occ3, 1))), dim3(Sand3<N>::NT), (int)Sand3<N>::smem, spec, g,
             (int)gstride, hc, cp, ny, kern, kstride, kp, elo, ehi, tw);
      return 0;
    }
    if (a16) LAUNCH(ctx, kc, (fft_z_sandwich2<N, true>), grd, dim3(Plan2<N>::NT), sm2, spec, g, (int)gstride, hc, cp, ny, kern, kstride, kp, elo, ehi, tw);
    else LAUNCH(ctx, kc, (fft_z_sandwich2<N, false>), grd, dim3(Plan2<N>::NT), sm2, spec, g, (int)gstride, hc, cp, ny, kern, kstride, kp, elo, ehi, tw);
    return 0;
  } else {
  static int occ = 0;
  if (!occ) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_z_sandwich<N>, NT, (int)smem_bytes_sandwich(N)));
  const long long total = (long long)((hc + LX - 1) / LX) * ny;
  LAUNCH(ctx, kc, fft_z_sandwich<N>, dim3((unsigned)std::min<long long>(total, (long long)NUM_SMS * std::max(occ, 1))), dim3(NT), (int)smem_bytes_sandwich(N), spec,
         g, gstride, hc, cp, ny, kern, kstride, kp, elo, ehi, tw);
  return 0;
  }
}
template <int N> int launch_x_c2r_t(cubep3m_b200_ctx* ctx, int kc, const float2* in, float* out, int lo_x, int cnt_x, int lo_y, int cnt_y, int lo_z,
                                    int cnt_z, int ny_src, long long opx, long long opy, float scale, const float2* tw, int nbatch, long long ibs,
                                    long long obs) {
  if (int st = set_smem_attr<N>()) return st;
  const long long nrows = (long long)cnt_z * cnt_y;
  LAUNCH(ctx, kc, fft_x_c2r<N>, dim3((unsigned)((nrows + 2 * LX - 1) / (2 * LX)), nbatch), dim3(NT), (int)smem_bytes(N), in, out, lo_x, cnt_x, lo_y, cnt_y,
         lo_z, cnt_z, ny_src, opx, opy, scale, tw, ibs, obs);
  return 0;
}
#ifndef FFTK_C2R3_V4
#define FFTK_C2R3_V4 1   // row-major lanes, tangle fused into stage A, results stored from registers (fft3d2.cuh)
#endif
#ifndef FFTK_C2R3_V3
#define FFTK_C2R3_V3 1
#endif
#ifndef FFTK_C2R3_V2
#define FFTK_C2R3_V2 0   // measured on B200 (n = 304): first-generation fft_x_c2r3 175 us, fft_x_c2r3_v2 222 us per tile
#endif
template <int N> int launch_x_c2r3_t(cubep3m_b200_ctx* ctx, int kc, const float2* in, int cp, float* out, int lo, int cnt, long long ibs, long long obs, float scale,
                                     unsigned int* fmax_bits, const float2* tw) {
  if (int st = set_smem_attr<N>()) return st;
  const long long nrows = (long long)cnt * cnt;
  if constexpr (Plan2<N>::ok) {
    // bulk-copy (TMA engine) staging needs 16-byte aligned rows of at least RP elements
    static const bool use_v4 = [] { const char* e = getenv("CUBEP3M_B200_C2R"); return !(e && !strcmp(e, "v3")); }();   // A/B knob
    if (FFTK_C2R3_V4 && use_v4 && cp % 2 == 0 && cp >= C2R4<N>::X::RP && ((uintptr_t)in & 15) == 0 && ibs % 2 == 0 && 3 * ibs < (1LL << 31) && 3 * obs < (1LL << 31)) {
      static int occ4 = 0;
      if (!occ4) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, fft_x_c2r3_v4<N>, C2R4<N>::NT, (int)C2R4<N>::smem));
      const long long nblk = (nrows + 2 * C2R4<N>::CW - 1) / (2 * C2R4<N>::CW);
      LAUNCH(ctx, kc, fft_x_c2r3_v4<N>, dim3((unsigned)std::min<long long>(nblk, (long long)NUM_SMS * std::max(occ4, 1))), dim3(C2R4<N>::NT), (int)C2R4<N>::smem, in, cp, out,
             lo, cnt, (int)ibs, (int)obs, scale, fmax_bits, tw);
      return 0;
    }
    if (FFTK_C2R3_V3 && cp % 2 == 0 && cp >= C2R3<N>::RP && ((uintptr_t)in & 15) == 0 && ibs % 2 == 0 && 3 * ibs < (1LL << 31) && 3 * obs < (1LL << 31)) {
      static int occ3 = 0;
      if (!occ3) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, fft_x_c2r3_v3<N>, Plan2<N>::NT, (int)C2R3<N>::smem));
      const long long nblk = (nrows + 2 * LX - 1) / (2 * LX);
      LAUNCH(ctx, kc, fft_x_c2r3_v3<N>, dim3((unsigned)std::min<long long>(nblk, (long long)NUM_SMS * std::max(occ3, 1))), dim3(Plan2<N>::NT), (int)C2R3<N>::smem, in, cp, out,
             lo, cnt, (int)ibs, (int)obs, scale, fmax_bits, tw);
      return 0;
    }
  }
  if constexpr (Plan2<N>::ok && FFTK_C2R3_V2) {
    if (3 * ibs >= (1LL << 31) || 3 * obs >= (1LL << 31)) return CUBEP3M_B200_EINVAL;
    LAUNCH(ctx, kc, fft_x_c2r3_v2<N>, dim3((unsigned)((nrows + 2 * LX - 1) / (2 * LX))), dim3(Plan2<N>::NT), (int)smem_bytes_c2r3_v2(N), in, cp, out, lo, cnt, (int)ibs, (int)obs,
           scale, fmax_bits, tw);
    return 0;
  }
  // first-generation kernel: the three-factor sizes, and the fine-CIC path of every size (its in-place r2c layout has an odd row pitch n/2+1)
  LAUNCH(ctx, kc, fft_x_c2r3<N>, dim3((unsigned)((nrows + 2 * LX - 1) / (2 * LX))), dim3(NT), (int)smem_bytes(N), in, cp, out, lo, cnt, ibs, obs, scale, fmax_bits, tw);
  return 0;
}
#define FFTK_SWITCH(N_, CALL)                                 \
  switch (N_) {                                               \
    FFTK_FOR_ALL_N(CALL)                                      \
    default: return CUBEP3M_B200_EINVAL;                      \
  }

inline int launch_x_r2c(cubep3m_b200_ctx* ctx, int kc, int n, float* data, int nrows, const float2* tw) {
#define X(N) case N: return launch_x_r2c_t<N>(ctx, kc, data, nrows, tw);
  FFTK_SWITCH(n, X)
#undef X
}
inline int launch_x_r2c_ngp(cubep3m_b200_ctx* ctx, int kc, int n, float2* data, int cp, const NgpSource& g, const float2* tw) {
#define X(N) case N: return launch_x_r2c_ngp_t<N>(ctx, kc, data, cp, g, tw);
  FFTK_SWITCH(n, X)
#undef X
}
inline int launch_strided(cubep3m_b200_ctx* ctx, int kc, int n, bool inv, const float2* in, float2* out, int hc, long long estride, long long ostride,
                          int outer0, int nouter, const float* kern, long long kes, long long kos, int elo, int ehi, const float2* tw,
                          int nbatch = 1, long long bstride = 0) {
#define X(N) case N: return launch_strided_t<N>(ctx, kc, inv, in, out, hc, estride, ostride, outer0, nouter, kern, kes, kos, elo, ehi, tw, nbatch, bstride);
  FFTK_SWITCH(n, X)
#undef X
}
inline int launch_sandwich(cubep3m_b200_ctx* ctx, int kc, int n, const float2* spec, float2* g, long long gstride, int hc, int cp, int ny, const float* kern,
                           long long kstride, int kp, int elo, int ehi, const float2* tw) {
#define X(N) case N: return launch_sandwich_t<N>(ctx, kc, spec, g, gstride, hc, cp, ny, kern, kstride, kp, elo, ehi, tw);
  FFTK_SWITCH(n, X)
#undef X
}
inline int launch_x_c2r(cubep3m_b200_ctx* ctx, int kc, int n, const float2* in, float* out, int lo_x, int cnt_x, int lo_y, int cnt_y, int lo_z,
                        int cnt_z, int ny_src, long long opx, long long opy, float scale, const float2* tw, int nbatch = 1, long long ibs = 0,
                        long long obs = 0) {
#define X(N) case N: return launch_x_c2r_t<N>(ctx, kc, in, out, lo_x, cnt_x, lo_y, cnt_y, lo_z, cnt_z, ny_src, opx, opy, scale, tw, nbatch, ibs, obs);
  FFTK_SWITCH(n, X)
#undef X
}

inline int launch_x_c2r3(cubep3m_b200_ctx* ctx, int kc, int n, const float2* in, int cp, float* out, int lo, int cnt, long long ibs, long long obs, float scale,
                         unsigned int* fmax_bits, const float2* tw) {
#define X(N) case N: return launch_x_c2r3_t<N>(ctx, kc, in, cp, out, lo, cnt, ibs, obs, scale, fmax_bits, tw);
  FFTK_SWITCH(n, X)
#undef X
}

// Fine-tile solve: forward x, y; fused z (forward, 3 x kernel multiply + inverse); then inverse y and inverse x (crop + scale) for the
// three components in one launch each.
// data: the tile's density (n+2,n,n) reals transformed in place (complex pitch cp = hc), or, when ngp != nullptr, only the destination of
// the spectrum, which the first pass generates from the fine-cell table; then cp may be any pitch >= hc (the library pads it to a multiple
// of 16 so that every 16-column block starts on a 128-byte boundary). g3: scratch of 3 complex tiles of the same pitch;
// force3: 3 x cnt^3 outputs (component-major).
inline int fine_solve(cubep3m_b200_ctx* ctx, const Mesh3& m, float* data, float* g3, int cp, const float* kern3, long long kstride, int kp, float* force3, int lo, int cnt,
                      float scale, unsigned int* fmax_bits, const NgpSource* ngp = nullptr) {
  const int n = m.nx, hc = m.hc();
  if (!ngp && cp != hc) return CUBEP3M_B200_EINVAL;
  const long long cplx = (long long)cp * n * n;     // complex elements per tile
  float2* c = reinterpret_cast<float2*>(data);
  float2* g = reinterpret_cast<float2*>(g3);
  if (ngp) { if (int st = launch_x_r2c_ngp(ctx, KC_FFT_X_R2C, n, c, cp, *ngp, m.twx)) return st; }
  else if (int st = launch_x_r2c(ctx, KC_FFT_X_R2C, n, data, n * n, m.twx)) return st;
  if (int st = launch_strided(ctx, KC_FFT_FWD_STRIDED, n, false, c, c, hc, (long long)cp, (long long)n * cp, 0, n, nullptr, 0, 0, 0, n - 1, m.twy)) return st;
  if (int st = launch_sandwich(ctx, KC_FFT_INV_Z_MUL, n, c, g, cplx, hc, cp, n, kern3, kstride, kp, lo, lo + cnt - 1, m.twz)) return st;
  if (int st = launch_strided(ctx, KC_FFT_INV_Y, n, true, g, g, hc, (long long)cp, (long long)n * cp, lo, cnt, nullptr, 0, 0, lo, lo + cnt - 1, m.twy, 3, cplx)) return st;
  if (int st = launch_x_c2r3(ctx, KC_FFT_X_C2R, n, g, cp, force3, lo, cnt, cplx, (long long)cnt * cnt * cnt, scale, fmax_bits, m.twx)) return st;
  CK(cudaGetLastError());
  return 0;
}

// forward 3-D r2c in place on data (nx+2, ny, nz)
inline int forward3d(cubep3m_b200_ctx* ctx, const Mesh3& g, float* data) {
  const int hc = g.hc(), cb = ctx->fft_class_base;
  if (int st = launch_x_r2c(ctx, cb ? cb : KC_FFT_X_R2C, g.nx, data, g.ny * g.nz, g.twx)) return st;
  float2* c = reinterpret_cast<float2*>(data);
  // Y: outer = z, element stride hc
  if (int st = launch_strided(ctx, cb ? cb : KC_FFT_FWD_STRIDED, g.ny, false, c, c, hc, (long long)hc, (long long)g.ny * hc, 0, g.nz, nullptr, 0, 0, 0,
                              g.ny - 1, g.twy)) return st;
  // Z: outer = y, element stride ny*hc
  if (int st = launch_strided(ctx, cb ? cb : KC_FFT_FWD_STRIDED, g.nz, false, c, c, hc, (long long)g.ny * hc, (long long)hc, 0, g.ny, nullptr, 0, 0, 0,
                              g.nz - 1, g.twz)) return st;
  CK(cudaGetLastError());
  return 0;
}

// backward 3-D c2r: src (hc,ny,nz) complex is read; work (same size) is scratch (may equal src when kern == nullptr);
// out receives the window [lo, lo+cnt) per axis scaled by `scale`, with pitches opx (x row length) and opy (rows per plane).
// If kern != nullptr the spectrum is first multiplied by i*kern (one component, layout [z][y][kx]).
inline int backward3d(cubep3m_b200_ctx* ctx, const Mesh3& g, const float* src, float* work, const float* kern, float* out, const int lo[3],
                      const int cnt[3], long long opx, long long opy, float scale) {
  const int hc = g.hc(), cb = ctx->fft_class_base;
  const float2* s = reinterpret_cast<const float2*>(src);
  float2* w = reinterpret_cast<float2*>(work);
  // Z backward (+ multiply): outer = y (all), keep only z in the window
  if (int st = launch_strided(ctx, cb ? cb : KC_FFT_INV_Z_MUL, g.nz, true, s, w, hc, (long long)g.ny * hc, (long long)hc, 0, g.ny, kern,
                              (long long)g.ny * hc, (long long)hc, lo[2], lo[2] + cnt[2] - 1, g.twz)) return st;
  // Y backward: outer = z in the window only, keep only y in the window
  if (int st = launch_strided(ctx, cb ? cb : KC_FFT_INV_Y, g.ny, true, w, w, hc, (long long)hc, (long long)g.ny * hc, lo[2], cnt[2], nullptr, 0, 0, lo[1],
                              lo[1] + cnt[1] - 1, g.twy)) return st;
  // X backward: window rows only
  if (int st = launch_x_c2r(ctx, cb ? cb : KC_FFT_X_C2R, g.nx, w, out, lo[0], cnt[0], lo[1], cnt[1], lo[2], cnt[2], g.ny, opx, opy, scale, g.twx)) return st;
  CK(cudaGetLastError());
  return 0;
}

// twiddle table exp(-2 pi i t / n) on the device; constant radix tables
inline int make_twiddles(int n, float2** out) {
  std::vector<float2> h(n);
  for (int t = 0; t < n; ++t) {
    const double a = -2.0 * M_PI * (double)t / (double)n;
    h[t] = make_float2((float)cos(a), (float)sin(a));
  }
  CK(cudaMalloc(out, sizeof(float2) * n));
  CK(cudaMemcpy(*out, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
  return 0;
}
inline int init_constants() {
  for (int R = 1; R < MAXR; ++R)
    for (int t = 0; t < R; ++t) {
      const double a = -2.0 * M_PI * (double)t / (double)R;
      h_w[R][t] = make_float2((float)cos(a), (float)sin(a));
    }
  CK(cudaMemcpyToSymbol(c_w, h_w, sizeof(h_w)));
  return 0;
}

}  // namespace fftk
