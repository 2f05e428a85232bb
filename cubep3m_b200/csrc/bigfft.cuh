// Forward transforms longer than one CTA's shared memory holds (N = 1024, 2048: the nf_physical_dim^3 meshes of cic_power at BASELINE configs[2-3]):
// four-step decomposition on top of the library's own passes.
//   strided axis, N = N1 * N2 (N1 = 32):  x[n1 + N1 n2]
//       1. N2-point transforms over n2 (element stride N1 * es) for every n1           -> Y[n1][k2]     (fft_strided2<N2>, n1 as the batch level)
//       2. Y[n1][k2] *= w_N^(n1 k2)                                                                      (big_twiddle_kernel)
//       3. N1-point transforms over n1 (element stride es) for every k2                -> X[k2 + N2 k1] at memory index k1 + N1 k2   (fft_strided2<N1>)
//      The result stays in that digit-transposed order: memory index m = k1 + N1 k2 holds frequency k = k2 + N2 k1 (big_freq). cic_power only bins
//      |delta_k|^2 by |k|, so nothing is ever un-permuted.
//   contiguous axis (real rows of N values = M = N/2 complex z[n] = x[2n] + i x[2n+1], M = 16 * N2):
//       1. N2-point transforms over n2 (stride 16) for the 16 contiguous n1 columns of every row                 (fft_strided2<N2>, rows as the outer level)
//       2. twiddle w_M^(n1 k2) and a radix-16 transform over the 16 contiguous n1, one thread per (row, k2)      (big_row16_kernel)
//       3. untangle Z -> the N/2+1 half-spectrum values of the real row, natural order, into a second array      (big_untangle_kernel)
// Arrays may exceed 4 GB (a 1024^3 half spectrum is 4.3 GB): the strided passes then take their 64-bit-offset instantiation (fft3d2.cuh: W64).
#pragma once
#include "fft3d.cuh"

namespace fftk {

constexpr int BIG_N1 = 32;
inline bool big_supported(int n) {       // both axes kinds
  if (n % (2 * 16) != 0 || n % BIG_N1 != 0) return false;
  const int n2x = n / 2 / 16, n2s = n / BIG_N1;
  auto okf = [](int f) { return f == 16 || f == 32 || f == 64; };        // 512 (tests: the same mesh both ways), 1024, 2048
  return okf(n2x) && okf(n2s);
}
// frequency index held at memory index m of a four-step axis of length n (identity for a directly transformed axis: n1 == 0)
__host__ __device__ __forceinline__ int big_freq(int m, int n1, int n2) { return n1 ? (m / n1) + n2 * (m % n1) : m; }

// element (b, o, e, c): data[o*ostride + e*estride + c], e = n1 + N1*k2, c < hc
__global__ void __launch_bounds__(256) big_twiddle_kernel(float2* __restrict__ data, int hc, long long estride, long long ostride, long long nouter, int N1, int N2,
                                                          const float2* __restrict__ twN) {
  const int N = N1 * N2;
  const long long total = nouter * N * hc;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const int c = (int)(t % hc);
    const long long r = t / hc;
    const int e = (int)(r % N);
    const long long o = r / N;
    const int n1 = e % N1, k2 = e / N1;
    if (n1 == 0 || k2 == 0) continue;
    const float2 w = twN[n1 * k2];
    float2* p = data + o * ostride + (long long)e * estride + c;
    const float2 v = *p;
    *p = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
  }
}

// rows of M = 16*N2 complex (pitch `pitch` complex); one thread per (row, k2): 16 contiguous values, twiddle w_M^(n1 k2), radix 16 over n1
__global__ void __launch_bounds__(256) big_row16_kernel(float2* __restrict__ data, long long nrows, long long pitch, int N2, const float2* __restrict__ twM) {
  const long long total = nrows * N2;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const int k2 = (int)(t % N2);
    const long long row = t / N2;
    float2* p = data + row * pitch + 16 * k2;
    float2 v[16];                                     // (rows have an odd pitch of M + 1 values: 8-byte accesses)
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = p[q];
#pragma unroll
    for (int n1 = 1; n1 < 16; ++n1) { const float2 w = twM[n1 * k2]; const float2 a = v[n1]; v[n1] = make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
    Radix<16, false>::run(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = v[q];
  }
}

// Z (digit-transposed: Z[k2 + N2 k1] at 16 k2 + k1) of the packed row -> X[k], k = 0..M, of the real row:
//   X[k] = (Z[k] + conj Z[M-k]) / 2 - i/2 w_N^k (Z[k] - conj Z[M-k]),  Z[M] = Z[0]
__global__ void __launch_bounds__(256) big_untangle_kernel(const float2* __restrict__ z, float2* __restrict__ out, long long nrows, long long pitch, int N2,
                                                           const float2* __restrict__ twN) {
  const int M = 16 * N2, hc = M + 1;
  const long long total = nrows * hc;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const int k = (int)(t % hc);
    const long long row = t / hc;
    const float2* zr = z + row * pitch;
    const int ka = k % M, kb = (M - k) % M;
    const float2 a = zr[16 * (ka % N2) + ka / N2], b0 = zr[16 * (kb % N2) + kb / N2];
    const float2 b = make_float2(b0.x, -b0.y);                                   // conj Z[M-k]
    const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));        // even part
    const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));
    // -i w d, w = exp(-2 pi i k / N) = twN[k] for k < N (k = M gives w = -1)
    const float2 w = (k == M) ? make_float2(-1.f, 0.f) : twN[k];
    const float2 wd = make_float2(w.x * d.x - w.y * d.y, w.x * d.y + w.y * d.x);
    out[row * pitch + k] = make_float2(e.x + wd.y, e.y - wd.x);
  }
}

struct BigTwiddles { float2 *t16 = nullptr, *t32 = nullptr, *t64 = nullptr, *tN = nullptr, *tM = nullptr; int N = 0; };
inline void big_free(BigTwiddles& T) { for (float2* p : {T.t16, T.t32, T.t64, T.tN, T.tM}) if (p) cudaFree(p); T = BigTwiddles(); }
inline int big_init(int N, BigTwiddles& T) {
  T.N = N;
  if (int st = make_twiddles(16, &T.t16)) return st;
  if (int st = make_twiddles(32, &T.t32)) return st;
  if (int st = make_twiddles(64, &T.t64)) return st;
  if (int st = make_twiddles(N, &T.tN)) return st;
  if (int st = make_twiddles(N / 2, &T.tM)) return st;
  return 0;
}
inline const float2* big_tw(const BigTwiddles& T, int n) { return n == 16 ? T.t16 : (n == 32 ? T.t32 : T.t64); }

// forward along a strided axis of length T.N, in place: lines data[o*ostride + e*estride + c], c < hc, o < nouter
inline int big_forward_strided(cubep3m_b200_ctx* ctx, int kc, const BigTwiddles& T, float2* data, int hc, long long estride, long long ostride, int nouter) {
  const int N1 = BIG_N1, N2 = T.N / BIG_N1;
  if (int st = launch_strided(ctx, kc, N2, false, data, data, hc, (long long)N1 * estride, ostride, 0, nouter, nullptr, 0, 0, 0, N2 - 1, big_tw(T, N2), N1, estride)) return st;
  LAUNCH(ctx, kc, big_twiddle_kernel, NUM_SMS * 16, 256, 0, data, hc, estride, ostride, (long long)nouter, N1, N2, T.tN);
  if (int st = launch_strided(ctx, kc, N1, false, data, data, hc, estride, ostride, 0, nouter, nullptr, 0, 0, 0, N1 - 1, big_tw(T, N1), N2, (long long)N1 * estride)) return st;
  return 0;
}
// r2c along the contiguous axis: `real` holds nrows rows of T.N reals with pitch T.N + 2 floats (destroyed), `out` receives T.N/2 + 1 complex per row, same pitch
inline int big_forward_x(cubep3m_b200_ctx* ctx, int kc, const BigTwiddles& T, float* real, float2* out, long long nrows) {
  const int M = T.N / 2, N2 = M / 16;
  const long long pitch = M + 1;
  float2* z = reinterpret_cast<float2*>(real);
  // rows in chunks: the row index is the strided kernels' 32-bit `outer` level
  const long long chunk = 1 << 20;
  for (long long r0 = 0; r0 < nrows; r0 += chunk) {
    const int nr = (int)std::min(chunk, nrows - r0);
    if (int st = launch_strided(ctx, kc, N2, false, z + r0 * pitch, z + r0 * pitch, 16, 16, pitch, 0, nr, nullptr, 0, 0, 0, N2 - 1, big_tw(T, N2))) return st;
  }
  LAUNCH(ctx, kc, big_row16_kernel, NUM_SMS * 16, 256, 0, z, nrows, pitch, N2, T.tM);
  LAUNCH(ctx, kc, big_untangle_kernel, NUM_SMS * 16, 256, 0, z, out, nrows, pitch, N2, T.tN);
  return 0;
}

}  // namespace fftk
