// TEST INFRASTRUCTURE — CPU oracle FFT. Not part of the product; only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// The reference calls FFTW (un-vendored; FFTW 2.1.5 `rfftw3d_f77_*` in fftw2.f90:10-25 or FFTW 3.x
// `sfftw_plan_dft_r2c_3d` / `sfftw_plan_dft_c2r_3d` in fft_fine.f90:28-51, fft_coarse.f90:134-209).
// What those call sites define is the *unnormalised* DFT with FFTW_FORWARD = exp(-i...), in place on a
// Fortran array (n+2, n, n): the first (contiguous) axis is the halved one and holds n/2+1 interleaved
// (re,im) pairs.  This header restates that published definition in float32 (twiddles computed in double)
// — any correct fp32 FFT agrees with FFTW to ~1e-6 relative, which is what the parity tolerance (1e-4)
// assumes.  Parity of this restatement is pinned against numpy.fft (pocketfft) in tests/test_oracle.py.
//
// Round 2: the transform is a batched Stockham autosort FFT (radix 4 / 2 / odd prime with the symmetric
// half-sum butterfly) over BATCH sequences at a time in split re/im [n][BATCH] buffers, so that every inner
// loop is a unit-stride loop the compiler vectorises, and the batches of one 3-D pass are spread over OpenMP
// threads (nested under the tile loop when there are fewer tiles than threads).  Round 1's recursive scalar
// Cooley-Tukey made the CPU baseline ~10x slower than an FFTW build of the reference would be.
#pragma once
#include <complex>
#include <vector>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <omp.h>

namespace oracle {

typedef std::complex<float> cf;
constexpr int BATCH = 16;

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORACLE_SIMD __attribute__((target_clones("avx2", "default")))
#else
#define ORACLE_SIMD
#endif

struct Fft1d {
  int n = 0;
  std::vector<int> factors;
  std::vector<float> twr, twi;   // exp(-2 pi i j / n)
  void init(int n_) {
    n = n_;
    factors.clear();
    int r = n;
    while (r % 4 == 0) { factors.push_back(4); r /= 4; }
    while (r % 2 == 0) { factors.push_back(2); r /= 2; }
    for (int p = 3; p * p <= r; p += 2)
      while (r % p == 0) { factors.push_back(p); r /= p; }
    if (r > 1) factors.push_back(r);
    twr.resize(n); twi.resize(n);
    for (int j = 0; j < n; ++j) {
      const double a = -2.0 * M_PI * (double)j / (double)n;
      twr[j] = (float)cos(a); twi[j] = (float)sin(a);
    }
  }
};

// one Stockham stage (decimation in frequency): sequence length len = p*m at stride s (s*len == n), L = s*BATCH contiguous floats per (q, r):
//   y[s*(p*q + u) + k] = ( sum_r x[s*(q + m*r) + k] W_p^{ru} ) * w_len^{qu}
// sgn = -1 forward, +1 backward (conjugate twiddles).
ORACLE_SIMD static void stockham_stage(int n, int p, int m, int s, float sgn, const float* __restrict__ twr, const float* __restrict__ twi,
                                       const float* __restrict__ xr, const float* __restrict__ xi, float* __restrict__ yr, float* __restrict__ yi) {
  const int L = s * BATCH;
  const int tstep = n / (p * m);          // w_len^j = tw[j * n/len]
  if (p == 2) {
    for (int q = 0; q < m; ++q) {
      const float wr = twr[(size_t)q * tstep], wi = -sgn * twi[(size_t)q * tstep];
      const float* ar = xr + (size_t)q * L; const float* ai = xi + (size_t)q * L;
      const float* br = xr + (size_t)(q + m) * L; const float* bi = xi + (size_t)(q + m) * L;
      float* o0r = yr + (size_t)(2 * q) * L; float* o0i = yi + (size_t)(2 * q) * L;
      float* o1r = yr + (size_t)(2 * q + 1) * L; float* o1i = yi + (size_t)(2 * q + 1) * L;
      for (int k = 0; k < L; ++k) {
        const float sr = ar[k] + br[k], si = ai[k] + bi[k], dr = ar[k] - br[k], di = ai[k] - bi[k];
        o0r[k] = sr; o0i[k] = si;
        o1r[k] = dr * wr - di * wi; o1i[k] = dr * wi + di * wr;
      }
    }
  } else if (p == 4) {
    for (int q = 0; q < m; ++q) {
      const float w1r = twr[(size_t)q * tstep], w1i = -sgn * twi[(size_t)q * tstep];
      const float w2r = twr[(size_t)2 * q * tstep], w2i = -sgn * twi[(size_t)2 * q * tstep];
      const float w3r = twr[(size_t)3 * q * tstep], w3i = -sgn * twi[(size_t)3 * q * tstep];
      const float* a0r = xr + (size_t)q * L; const float* a0i = xi + (size_t)q * L;
      const float* a1r = xr + (size_t)(q + m) * L; const float* a1i = xi + (size_t)(q + m) * L;
      const float* a2r = xr + (size_t)(q + 2 * m) * L; const float* a2i = xi + (size_t)(q + 2 * m) * L;
      const float* a3r = xr + (size_t)(q + 3 * m) * L; const float* a3i = xi + (size_t)(q + 3 * m) * L;
      float* o0r = yr + (size_t)(4 * q) * L; float* o0i = yi + (size_t)(4 * q) * L;
      float* o1r = o0r + L; float* o1i = o0i + L; float* o2r = o1r + L; float* o2i = o1i + L; float* o3r = o2r + L; float* o3i = o2i + L;
      for (int k = 0; k < L; ++k) {
        const float s0r = a0r[k] + a2r[k], s0i = a0i[k] + a2i[k], s1r = a0r[k] - a2r[k], s1i = a0i[k] - a2i[k];
        const float s2r = a1r[k] + a3r[k], s2i = a1i[k] + a3i[k], s3r = a1r[k] - a3r[k], s3i = a1i[k] - a3i[k];
        // forward: W_4 = -i: (s3) * (-i) = (s3i, -s3r); backward: (+i) = (-s3i, s3r)
        const float tr = -sgn * s3i, ti = sgn * s3r;
        const float b1r = s1r + tr, b1i = s1i + ti, b3r = s1r - tr, b3i = s1i - ti, b2r = s0r - s2r, b2i = s0i - s2i;
        o0r[k] = s0r + s2r; o0i[k] = s0i + s2i;
        o1r[k] = b1r * w1r - b1i * w1i; o1i[k] = b1r * w1i + b1i * w1r;
        o2r[k] = b2r * w2r - b2i * w2i; o2i[k] = b2r * w2i + b2i * w2r;
        o3r[k] = b3r * w3r - b3i * w3i; o3i[k] = b3r * w3i + b3i * w3r;
      }
    }
  } else {
    // odd prime p: with S_r = a_r + a_{p-r}, D_r = a_r - a_{p-r} (r = 1..h, h = (p-1)/2):
    //   b_u = a_0 + sum_r S_r cos(2 pi r u / p) + sgn * i * sum_r D_r sin(2 pi r u / p),  b_{p-u} = the same with the opposite sign of the second sum
    const int h = (p - 1) / 2;
    const int pstep = n / p;              // W_p^j = tw[j * n/p]
    static thread_local std::vector<float> scratch;
    if (scratch.size() < (size_t)4 * h * L) scratch.resize((size_t)4 * h * L);
    float *Sr = scratch.data(), *Si = Sr + (size_t)h * L, *Dr = Si + (size_t)h * L, *Di = Dr + (size_t)h * L;
    for (int q = 0; q < m; ++q) {
      const float* a0r = xr + (size_t)q * L; const float* a0i = xi + (size_t)q * L;
      for (int r = 1; r <= h; ++r) {
        const float* ar = xr + (size_t)(q + m * r) * L; const float* ai = xi + (size_t)(q + m * r) * L;
        const float* br = xr + (size_t)(q + m * (p - r)) * L; const float* bi = xi + (size_t)(q + m * (p - r)) * L;
        float* sr = &Sr[(size_t)(r - 1) * L]; float* si = &Si[(size_t)(r - 1) * L]; float* dr = &Dr[(size_t)(r - 1) * L]; float* di = &Di[(size_t)(r - 1) * L];
        for (int k = 0; k < L; ++k) { sr[k] = ar[k] + br[k]; si[k] = ai[k] + bi[k]; dr[k] = ar[k] - br[k]; di[k] = ai[k] - bi[k]; }
      }
      {
        float* o0r = yr + (size_t)(p * q) * L; float* o0i = yi + (size_t)(p * q) * L;
        for (int k = 0; k < L; ++k) { o0r[k] = a0r[k]; o0i[k] = a0i[k]; }
        for (int r = 0; r < h; ++r) {
          const float* sr = &Sr[(size_t)r * L]; const float* si = &Si[(size_t)r * L];
          for (int k = 0; k < L; ++k) { o0r[k] += sr[k]; o0i[k] += si[k]; }
        }
      }
      for (int u = 1; u <= h; ++u) {
        float* our = yr + (size_t)(p * q + u) * L; float* oui = yi + (size_t)(p * q + u) * L;
        float* ovr = yr + (size_t)(p * q + p - u) * L; float* ovi = yi + (size_t)(p * q + p - u) * L;
        // X = a_0 + sum S_r c ; Y = sum D_r s  (s = sin(2 pi r u / p) > or < 0)
        for (int k = 0; k < L; ++k) { our[k] = a0r[k]; oui[k] = a0i[k]; ovr[k] = 0.f; ovi[k] = 0.f; }
        for (int r = 1; r <= h; ++r) {
          const int j = (r * u) % p;
          const float c = twr[(size_t)j * pstep], sn = -twi[(size_t)j * pstep];     // cos, sin of +2 pi j / p
          const float* sr = &Sr[(size_t)(r - 1) * L]; const float* si = &Si[(size_t)(r - 1) * L];
          const float* dr = &Dr[(size_t)(r - 1) * L]; const float* di = &Di[(size_t)(r - 1) * L];
          for (int k = 0; k < L; ++k) { our[k] += sr[k] * c; oui[k] += si[k] * c; ovr[k] += dr[k] * sn; ovi[k] += di[k] * sn; }
        }
        const float w1r = twr[(size_t)q * u * tstep % n], w1i = -sgn * twi[(size_t)q * u * tstep % n];
        const float w2r = twr[(size_t)q * (p - u) * tstep % n], w2i = -sgn * twi[(size_t)q * (p - u) * tstep % n];
        for (int k = 0; k < L; ++k) {
          // sgn * i * Y = sgn * (-Yi, Yr)
          const float xr_ = our[k], xi_ = oui[k], tr = -sgn * ovi[k], ti = sgn * ovr[k];
          const float bur = xr_ + tr, bui = xi_ + ti, bvr = xr_ - tr, bvi = xi_ - ti;
          our[k] = bur * w1r - bui * w1i; oui[k] = bur * w1i + bui * w1r;
          ovr[k] = bvr * w2r - bvi * w2i; ovi[k] = bvr * w2i + bvi * w2r;
        }
      }
    }
  }
}

// BATCH sequences in split re/im [n][BATCH] buffers; the result ends in (ar, ai) (a copy is made when the stage count is odd)
inline void fft_batch(const Fft1d& P, bool forward, float* ar, float* ai, float* br, float* bi) {
  const float sgn = forward ? -1.f : 1.f;
  int len = P.n, s = 1;
  float *xr = ar, *xi = ai, *yr = br, *yi = bi;
  for (int p : P.factors) {
    const int m = len / p;
    stockham_stage(P.n, p, m, s, sgn, P.twr.data(), P.twi.data(), xr, xi, yr, yi);
    std::swap(xr, yr); std::swap(xi, yi);
    len = m; s *= p;
  }
  if (xr != ar) { std::memcpy(ar, xr, sizeof(float) * (size_t)P.n * BATCH); std::memcpy(ai, xi, sizeof(float) * (size_t)P.n * BATCH); }
}

// threads to use for the batches of one pass: everything when called from serial code, the idle share when nested under the tile loop
// (g_threads = the thread budget of the step, set by the caller of the tile loop before it opens its parallel region)
static int g_threads = 0;
inline int inner_threads() {
  if (!omp_in_parallel()) return omp_get_max_threads();
  if (omp_get_max_active_levels() < 2 || g_threads <= 0) return 1;
  return std::max(1, g_threads / std::max(omp_get_num_threads(), 1));
}

// In-place 3-D real<->complex transform on a Fortran-ordered padded array a(nx+2, ny, nz).
// (The reference only ever transforms cubes; the three lengths differ only for the non-cubic rank grids used at 2/4 GPUs.)
struct Fft3dR2C {
  int nx = 0, ny = 0, nz = 0;
  Fft1d px, py, pz;
  void init(int n_) { init(n_, n_, n_); }
  void init(int nx_, int ny_, int nz_) { nx = nx_; ny = ny_; nz = nz_; px.init(nx); py.init(ny); pz.init(nz); }

  // complex columns along an axis of c(hc, ny, nz): `count` lines of length P.n, element stride `es`; line l starts at base(l) and the
  // BATCH lines of one batch are consecutive kx (unit stride)
  template <typename BaseFn> void strided_pass(const Fft1d& P, bool forward, cf* c, long long es, long long nlines_x, long long nouter, BaseFn base) const {
    const long long nbx = (nlines_x + BATCH - 1) / BATCH, total = nbx * nouter;
    const int n = P.n, nth = inner_threads();
#pragma omp parallel num_threads(nth)
    {
      std::vector<float> buf((size_t)4 * n * BATCH);
      float *ar = buf.data(), *ai = ar + (size_t)n * BATCH, *br = ai + (size_t)n * BATCH, *bi = br + (size_t)n * BATCH;
#pragma omp for schedule(static)
      for (long long it = 0; it < total; ++it) {
        const long long o = it / nbx, x0 = (it - o * nbx) * BATCH;
        const int nb = (int)std::min<long long>(BATCH, nlines_x - x0);
        cf* b0 = c + base(o) + x0;
        for (int e = 0; e < n; ++e) {
          const cf* src = b0 + (long long)e * es;
          float* dr = ar + (size_t)e * BATCH; float* di = ai + (size_t)e * BATCH;
          for (int b = 0; b < nb; ++b) { dr[b] = src[b].real(); di[b] = src[b].imag(); }
          for (int b = nb; b < BATCH; ++b) { dr[b] = 0.f; di[b] = 0.f; }
        }
        fft_batch(P, forward, ar, ai, br, bi);
        for (int e = 0; e < n; ++e) {
          cf* dst = b0 + (long long)e * es;
          const float* sr = ar + (size_t)e * BATCH; const float* si = ai + (size_t)e * BATCH;
          for (int b = 0; b < nb; ++b) dst[b] = cf(sr[b], si[b]);
        }
      }
    }
  }

  // forward: real (first nx of each padded row) -> nx/2+1 complex per row; unnormalised. Two real rows ride one complex transform.
  void forward(float* a) const {
    const int n2 = nx + 2, hc = nx / 2 + 1, n = nx;
    const long long nrows = (long long)ny * nz, nbatch = (nrows + 2 * BATCH - 1) / (2 * BATCH);
    const int nth = inner_threads();
#pragma omp parallel num_threads(nth)
    {
      std::vector<float> buf((size_t)4 * n * BATCH);
      float *ar = buf.data(), *ai = ar + (size_t)n * BATCH, *br = ai + (size_t)n * BATCH, *bi = br + (size_t)n * BATCH;
#pragma omp for schedule(static)
      for (long long it = 0; it < nbatch; ++it) {
        const long long r0 = it * 2 * BATCH;
        for (int b = 0; b < BATCH; ++b) {
          const long long ra = r0 + 2 * b, rb = ra + 1;
          const float* pa = ra < nrows ? a + (size_t)n2 * ra : nullptr;
          const float* pb = rb < nrows ? a + (size_t)n2 * rb : nullptr;
          for (int i = 0; i < n; ++i) { ar[(size_t)i * BATCH + b] = pa ? pa[i] : 0.f; ai[(size_t)i * BATCH + b] = pb ? pb[i] : 0.f; }
        }
        fft_batch(px, true, ar, ai, br, bi);
        for (int b = 0; b < BATCH; ++b) {
          const long long ra = r0 + 2 * b, rb = ra + 1;
          float* pa = ra < nrows ? a + (size_t)n2 * ra : nullptr;
          float* pb = rb < nrows ? a + (size_t)n2 * rb : nullptr;
          for (int k = 0; k < hc; ++k) {
            const int km = (k == 0) ? 0 : n - k;
            const float zr = ar[(size_t)k * BATCH + b], zi = ai[(size_t)k * BATCH + b], wr = ar[(size_t)km * BATCH + b], wi = ai[(size_t)km * BATCH + b];
            // A = (Z[k] + conj Z[n-k]) / 2,  B = (Z[k] - conj Z[n-k]) / (2i)
            if (pa) { pa[2 * k] = 0.5f * (zr + wr); pa[2 * k + 1] = 0.5f * (zi - wi); }
            if (pb) { pb[2 * k] = 0.5f * (zi + wi); pb[2 * k + 1] = -0.5f * (zr - wr); }
          }
        }
      }
    }
    cf* c = reinterpret_cast<cf*>(a);  // c(hc, ny, nz)
    strided_pass(py, true, c, hc, hc, nz, [&](long long k) { return (long long)hc * ny * k; });
    strided_pass(pz, true, c, (long long)hc * ny, hc, ny, [&](long long j) { return (long long)hc * j; });
  }

  // backward: complex -> real, unnormalised (the caller divides by n^3: fftw2.f90:22, fft_fine.f90:51)
  void backward(float* a) const {
    const int n2 = nx + 2, hc = nx / 2 + 1, n = nx;
    cf* c = reinterpret_cast<cf*>(a);
    strided_pass(pz, false, c, (long long)hc * ny, hc, ny, [&](long long j) { return (long long)hc * j; });
    strided_pass(py, false, c, hc, hc, nz, [&](long long k) { return (long long)hc * ny * k; });
    const long long nrows = (long long)ny * nz, nbatch = (nrows + 2 * BATCH - 1) / (2 * BATCH);
    const int nth = inner_threads();
#pragma omp parallel num_threads(nth)
    {
      std::vector<float> buf((size_t)4 * n * BATCH);
      float *ar = buf.data(), *ai = ar + (size_t)n * BATCH, *br = ai + (size_t)n * BATCH, *bi = br + (size_t)n * BATCH;
#pragma omp for schedule(static)
      for (long long it = 0; it < nbatch; ++it) {
        const long long r0 = it * 2 * BATCH;
        // Z[k] = A[k] + i B[k], Z[n-k] = conj A[k] + i conj B[k]; c2r ignores the imaginary parts of the self-conjugate bins, as FFTW does
        for (int b = 0; b < BATCH; ++b) {
          const long long ra = r0 + 2 * b, rb = ra + 1;
          const float* pa = ra < nrows ? a + (size_t)n2 * ra : nullptr;
          const float* pb = rb < nrows ? a + (size_t)n2 * rb : nullptr;
          for (int k = 0; k < hc; ++k) {
            float Ar = pa ? pa[2 * k] : 0.f, Ai = pa ? pa[2 * k + 1] : 0.f, Br = pb ? pb[2 * k] : 0.f, Bi = pb ? pb[2 * k + 1] : 0.f;
            if (k == 0 || 2 * k == n) { Ai = 0.f; Bi = 0.f; }
            ar[(size_t)k * BATCH + b] = Ar - Bi; ai[(size_t)k * BATCH + b] = Ai + Br;
            if (k != 0 && 2 * k != n) { ar[(size_t)(n - k) * BATCH + b] = Ar + Bi; ai[(size_t)(n - k) * BATCH + b] = Br - Ai; }
          }
        }
        fft_batch(px, false, ar, ai, br, bi);
        for (int b = 0; b < BATCH; ++b) {
          const long long ra = r0 + 2 * b, rb = ra + 1;
          float* pa = ra < nrows ? a + (size_t)n2 * ra : nullptr;
          float* pb = rb < nrows ? a + (size_t)n2 * rb : nullptr;
          if (pa) { for (int i = 0; i < n; ++i) pa[i] = ar[(size_t)i * BATCH + b]; pa[n] = 0.f; pa[n + 1] = 0.f; }
          if (pb) { for (int i = 0; i < n; ++i) pb[i] = ai[(size_t)i * BATCH + b]; pb[n] = 0.f; pb[n + 1] = 0.f; }
        }
      }
    }
  }
};

}  // namespace oracle
