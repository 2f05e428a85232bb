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

// ---- pass Y / Z: strided complex columns.
// element e of column c of block (bx, by): in[base + e*estride + c], base = (by + outer0)*ostride + bx*LX
// MUL: multiply the loaded value by i*kern[e*kes + (by+outer0)*kos + kx]   (Z pass: e=z, outer=y; kern = one component)
// stores only elements e in [elo, ehi]
template <int N, bool INV, bool MUL>
__global__ void __launch_bounds__(NT) fft_strided(const float2* __restrict__ in, float2* __restrict__ out, int hc,
                                                  long long estride, long long ostride, int outer0,
                                                  const float* __restrict__ kern, long long kes, long long kos,
                                                  int elo, int ehi, const float2* __restrict__ tw_g) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem s = carve<N>(raw, tw_g);
  const int outer = blockIdx.y + outer0;
  const int kx0 = blockIdx.x * LX;
  const long long base = (long long)outer * ostride + kx0;
  const int col = threadIdx.x % LX;
  const bool colok = (kx0 + col) < hc;
  for (int e = threadIdx.x / LX; e < N; e += NT / LX) {
    float2 v = make_float2(0.f, 0.f);
    if (colok) {
      v = in[base + (long long)e * estride + col];
      if (MUL) {
        const float kv = kern[(long long)e * kes + (long long)outer * kos + kx0 + col];
        v = make_float2(-v.y * kv, v.x * kv);
      }
    }
    s.re0[e * LXP + col] = v.x;
    s.im0[e * LXP + col] = v.y;
  }
  __syncthreads();
  fft_columns<N, INV>(s.re0, s.im0, s.re1, s.im1, s.tw);
  const float* zr = result_buffer<N>() ? s.re1 : s.re0;
  const float* zi = result_buffer<N>() ? s.im1 : s.im0;
  if (colok)
    for (int e = elo + threadIdx.x / LX; e <= ehi; e += NT / LX)
      out[base + (long long)e * estride + col] = make_float2(zr[e * LXP + col], zi[e * LXP + col]);
}

// ---- pass X backward: half spectra -> real rows with crop + scale.
// Row index space: ridx in [0, cnt*cnt): zc = ridx / cnt, yc = ridx % cnt, source row (z = zc+lo, y = yc+lo).
// Output: out[(zc*opitch_y + yc)*opitch_x + xc], xc in [0,cnt) <- x = xc + lo.
template <int N>
__global__ void __launch_bounds__(NT) fft_x_c2r(const float2* __restrict__ in, float* __restrict__ out, int lo, int cnt,
                                                long long opitch_x, long long opitch_y, float scale,
                                                const float2* __restrict__ tw_g) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem s = carve<N>(raw, tw_g);
  constexpr int HC = N / 2 + 1;
  const long long nrows = (long long)cnt * cnt;
  const long long r0 = (long long)blockIdx.x * (2 * LX);
  // stage A (even rows) into buffer 0, B (odd rows) into buffer 1
  for (int q = threadIdx.x; q < 2 * LX * HC; q += NT) {
    const int row = q / HC, k = q - row * HC;
    const long long ridx = r0 + row;
    float2 v = make_float2(0.f, 0.f);
    if (ridx < nrows) {
      const int zc = (int)(ridx / cnt), yc = (int)(ridx - (long long)zc * cnt);
      v = in[((long long)(zc + lo) * N + (yc + lo)) * HC + k];
    }
    const int col = row >> 1;
    if (row & 1) { s.re1[k * LXP + col] = v.x; s.im1[k * LXP + col] = v.y; }
    else         { s.re0[k * LXP + col] = v.x; s.im0[k * LXP + col] = v.y; }
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
  for (int q = threadIdx.x; q < 2 * LX * cnt; q += NT) {
    const int row = q / cnt, xc = q - row * cnt;
    const long long ridx = r0 + row;
    if (ridx >= nrows) continue;
    const int zc = (int)(ridx / cnt), yc = (int)(ridx - (long long)zc * cnt);
    const int x = xc + lo, col = row >> 1;
    const float v = (row & 1) ? zi[x * LXP + col] : zr[x * LXP + col];
    out[((long long)zc * opitch_y + yc) * opitch_x + xc] = v * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch on N
// ------------------------------------------------------------------------------------------------
template <int N> int set_smem_attr() {
  static bool done = false;
  if (done) return 0;
  const int bytes = (int)smem_bytes(N);
  CK(cudaFuncSetAttribute(fft_x_r2c<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(cudaFuncSetAttribute(fft_x_c2r<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(cudaFuncSetAttribute((fft_strided<N, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(cudaFuncSetAttribute((fft_strided<N, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(cudaFuncSetAttribute((fft_strided<N, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done = true;
  return 0;
}

// forward 3-D r2c in place on data (N+2, N, N)
template <int N> int forward3d_t(cubep3m_b200_ctx* ctx, float* data, const float2* tw) {
  if (int st = set_smem_attr<N>()) return st;
  const int hc = N / 2 + 1, sm = (int)smem_bytes(N);
  const int nrows = N * N;
  const int cb = ctx->fft_class_base;
  LAUNCH(ctx, cb ? cb : KC_FFT_X_R2C, fft_x_r2c<N>, dim3((nrows + 2 * LX - 1) / (2 * LX)), dim3(NT), sm, data, nrows, tw);
  float2* c = reinterpret_cast<float2*>(data);
  const int chunks = (hc + LX - 1) / LX;
  // Y: outer = z, element stride hc
  LAUNCH(ctx, cb ? cb : KC_FFT_FWD_STRIDED, (fft_strided<N, false, false>), dim3(chunks, N), dim3(NT), sm, c, c, hc, (long long)hc, (long long)N * hc, 0,
         nullptr, 0LL, 0LL, 0, N - 1, tw);
  // Z: outer = y, element stride N*hc
  LAUNCH(ctx, cb ? cb : KC_FFT_FWD_STRIDED, (fft_strided<N, false, false>), dim3(chunks, N), dim3(NT), sm, c, c, hc, (long long)N * hc, (long long)hc, 0,
         nullptr, 0LL, 0LL, 0, N - 1, tw);
  CK(cudaGetLastError());
  return 0;
}

// backward 3-D c2r: src (hc,N,N) complex is read; work (same size) is scratch (may equal src when kern == nullptr);
// out receives the cropped cube [lo, lo+cnt)^3 scaled by `scale`.
// If kern != nullptr the spectrum is first multiplied by i*kern (one component, layout [z][y][kx]).
template <int N> int backward3d_t(cubep3m_b200_ctx* ctx, const float* src, float* work, const float* kern, float* out,
                                  int lo, int cnt, long long opitch_x, long long opitch_y, float scale, const float2* tw) {
  if (int st = set_smem_attr<N>()) return st;
  const int hc = N / 2 + 1, sm = (int)smem_bytes(N);
  const float2* s = reinterpret_cast<const float2*>(src);
  float2* w = reinterpret_cast<float2*>(work);
  const int chunks = (hc + LX - 1) / LX;
  // Z backward (+ multiply): outer = y (all), keep only z in the crop
  const int cb = ctx->fft_class_base;
  if (kern) {
    LAUNCH(ctx, cb ? cb : KC_FFT_INV_Z_MUL, (fft_strided<N, true, true>), dim3(chunks, N), dim3(NT), sm, s, w, hc, (long long)N * hc, (long long)hc, 0,
           kern, (long long)N * hc, (long long)hc, lo, lo + cnt - 1, tw);
  } else {
    LAUNCH(ctx, cb ? cb : KC_FFT_INV_Z_MUL, (fft_strided<N, true, false>), dim3(chunks, N), dim3(NT), sm, s, w, hc, (long long)N * hc, (long long)hc, 0,
           nullptr, 0LL, 0LL, lo, lo + cnt - 1, tw);
  }
  // Y backward: outer = z in the crop only, keep only y in the crop
  LAUNCH(ctx, cb ? cb : KC_FFT_INV_Y, (fft_strided<N, true, false>), dim3(chunks, cnt), dim3(NT), sm, w, w, hc, (long long)hc, (long long)N * hc, lo,
         nullptr, 0LL, 0LL, lo, lo + cnt - 1, tw);
  // X backward: cropped rows only
  const long long nrows = (long long)cnt * cnt;
  LAUNCH(ctx, cb ? cb : KC_FFT_X_C2R, fft_x_c2r<N>, dim3((unsigned)((nrows + 2 * LX - 1) / (2 * LX))), dim3(NT), sm, w, out, lo, cnt, opitch_x, opitch_y, scale, tw);
  CK(cudaGetLastError());
  return 0;
}

#define FFTK_DISPATCH(N_, CALL)                \
  switch (N_) {                                \
    case 16: return CALL(16);                  \
    case 32: return CALL(32);                  \
    case 48: return CALL(48);                  \
    case 64: return CALL(64);                  \
    case 80: return CALL(80);                  \
    case 112: return CALL(112);                \
    case 128: return CALL(128);                \
    case 176: return CALL(176);                \
    case 256: return CALL(256);                \
    case 304: return CALL(304);                \
    case 512: return CALL(512);                \
    case 560: return CALL(560);                \
    default: return CUBEP3M_B200_EINVAL;       \
  }

inline bool supported(int n) {
  switch (n) { case 16: case 32: case 48: case 64: case 80: case 112: case 128: case 176: case 256: case 304: case 512: case 560: return true; }
  return false;
}

inline int forward3d(cubep3m_b200_ctx* ctx, int n, float* data, const float2* tw) {
#define CALL_(N) forward3d_t<N>(ctx, data, tw)
  FFTK_DISPATCH(n, CALL_)
#undef CALL_
}
inline int backward3d(cubep3m_b200_ctx* ctx, int n, const float* src, float* work, const float* kern, float* out, int lo,
                      int cnt, long long opx, long long opy, float scale, const float2* tw) {
#define CALL_(N) backward3d_t<N>(ctx, src, work, kern, out, lo, cnt, opx, opy, scale, tw)
  FFTK_DISPATCH(n, CALL_)
#undef CALL_
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
