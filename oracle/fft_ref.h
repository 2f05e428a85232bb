// TEST INFRASTRUCTURE — CPU oracle FFT. Not part of the product; only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// The reference calls FFTW (un-vendored; FFTW 2.1.5 `rfftw3d_f77_*` in fftw2.f90:10-25 or FFTW 3.x
// `sfftw_plan_dft_r2c_3d` / `sfftw_plan_dft_c2r_3d` in fft_fine.f90:28-51, fft_coarse.f90:134-209).
// What those call sites define is the *unnormalised* DFT with FFTW_FORWARD = exp(-i...), in place on a
// Fortran array (n+2, n, n): the first (contiguous) axis is the halved one and holds n/2+1 interleaved
// (re,im) pairs.  This header restates that published definition with a plain mixed-radix
// Cooley-Tukey in float32 (twiddles computed in double) — any correct fp32 FFT agrees with FFTW to
// ~1e-6 relative, which is what the parity tolerance (1e-4) assumes.  Parity of this restatement is
// pinned against numpy.fft (pocketfft) in tests/test_oracle_fft.py.
#pragma once
#include <complex>
#include <vector>
#include <cmath>
#include <cstring>
#include <algorithm>

namespace oracle {

typedef std::complex<float> cf;

struct Fft1d {
  int n = 0;
  std::vector<int> factors;
  std::vector<cf> tw_f, tw_b;  // exp(-/+ 2 pi i j / n)
  void init(int n_) {
    n = n_;
    factors.clear();
    int r = n;
    while (r % 4 == 0) { factors.push_back(4); r /= 4; }
    while (r % 2 == 0) { factors.push_back(2); r /= 2; }
    for (int p = 3; p * p <= r; p += 2)
      while (r % p == 0) { factors.push_back(p); r /= p; }
    if (r > 1) factors.push_back(r);
    tw_f.resize(n); tw_b.resize(n);
    for (int j = 0; j < n; ++j) {
      double a = -2.0 * M_PI * (double)j / (double)n;
      tw_f[j] = cf((float)cos(a), (float)sin(a));
      tw_b[j] = cf((float)cos(a), (float)-sin(a));
    }
  }
  // out[k] = sum_j in[j*istride] * w^(jk), recursive decimation in time; fi = index into factors
  void rec(int m, int fi, const cf* in, int istride, cf* out, const cf* tw) const {
    if (m == 1) { out[0] = in[0]; return; }
    const int p = factors[fi];
    const int q = m / p;
    for (int r = 0; r < p; ++r) rec(q, fi + 1, in + (size_t)r * istride, istride * p, out + (size_t)r * q, tw);
    const int tstep = n / m;  // w_m^j = tw[j * n/m]
    cf t[64];
    if (p == 2) {
      for (int k = 0; k < q; ++k) {
        cf a = out[k], b = out[q + k] * tw[(size_t)k * tstep];
        out[k] = a + b; out[q + k] = a - b;
      }
    } else if (p == 4) {
      const bool fwd = (tw == tw_f.data());
      for (int k = 0; k < q; ++k) {
        cf a = out[k];
        cf b = out[q + k] * tw[(size_t)k * tstep];
        cf c = out[2 * q + k] * tw[(size_t)2 * k * tstep];
        cf d = out[3 * q + k] * tw[(size_t)3 * k * tstep];
        cf s0 = a + c, s1 = a - c, s2 = b + d, s3 = b - d;
        // multiply s3 by -i (forward) or +i (backward)
        cf s3r = fwd ? cf(s3.imag(), -s3.real()) : cf(-s3.imag(), s3.real());
        out[k] = s0 + s2; out[q + k] = s1 + s3r; out[2 * q + k] = s0 - s2; out[3 * q + k] = s1 - s3r;
      }
    } else {
      const int pstep = n / p;  // w_p^j = tw[j * n/p]
      for (int k = 0; k < q; ++k) {
        for (int r = 0; r < p; ++r) t[r] = out[(size_t)r * q + k] * tw[(size_t)r * k * tstep];
        for (int u = 0; u < p; ++u) {
          cf acc = t[0];
          for (int r = 1; r < p; ++r) acc += t[r] * tw[(size_t)((r * u) % p) * pstep];
          out[(size_t)u * q + k] = acc;
        }
      }
    }
  }
  void exec(const cf* in, int istride, cf* out, bool forward) const {
    rec(n, 0, in, istride, out, forward ? tw_f.data() : tw_b.data());
  }
};

// In-place 3-D real<->complex transform on a Fortran-ordered padded array a(nx+2, ny, nz).
// (The reference only ever transforms cubes; the three lengths differ only for the non-cubic rank grids used at 2/4 GPUs.)
struct Fft3dR2C {
  int nx = 0, ny = 0, nz = 0;
  Fft1d px, py, pz;
  void init(int n_) { init(n_, n_, n_); }
  void init(int nx_, int ny_, int nz_) { nx = nx_; ny = ny_; nz = nz_; px.init(nx); py.init(ny); pz.init(nz); }
  // forward: real (first nx of each padded row) -> nx/2+1 complex per row; unnormalised.
  void forward(float* a) const {
    const int n2 = nx + 2, hc = nx / 2 + 1;
    std::vector<cf> in(std::max(nx, std::max(ny, nz))), out(in.size());
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j) {
        float* row = a + (size_t)n2 * (j + (size_t)ny * k);
        for (int i = 0; i < nx; ++i) in[i] = cf(row[i], 0.f);
        px.exec(in.data(), 1, out.data(), true);
        for (int i = 0; i < hc; ++i) { row[2 * i] = out[i].real(); row[2 * i + 1] = out[i].imag(); }
      }
    cf* c = reinterpret_cast<cf*>(a);  // c(hc, ny, nz)
    for (int k = 0; k < nz; ++k)
      for (int i = 0; i < hc; ++i) {
        cf* base = c + i + (size_t)hc * ny * k;
        py.exec(base, hc, out.data(), true);
        for (int j = 0; j < ny; ++j) base[(size_t)j * hc] = out[j];
      }
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < hc; ++i) {
        cf* base = c + i + (size_t)hc * j;
        pz.exec(base, hc * ny, out.data(), true);
        for (int k = 0; k < nz; ++k) base[(size_t)k * hc * ny] = out[k];
      }
  }
  // backward: complex -> real, unnormalised (the caller divides by n^3: fftw2.f90:22, fft_fine.f90:51)
  void backward(float* a) const {
    const int n2 = nx + 2, hc = nx / 2 + 1;
    std::vector<cf> in(std::max(nx, std::max(ny, nz))), out(in.size());
    cf* c = reinterpret_cast<cf*>(a);
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < hc; ++i) {
        cf* base = c + i + (size_t)hc * j;
        pz.exec(base, hc * ny, out.data(), false);
        for (int k = 0; k < nz; ++k) base[(size_t)k * hc * ny] = out[k];
      }
    for (int k = 0; k < nz; ++k)
      for (int i = 0; i < hc; ++i) {
        cf* base = c + i + (size_t)hc * ny * k;
        py.exec(base, hc, out.data(), false);
        for (int j = 0; j < ny; ++j) base[(size_t)j * hc] = out[j];
      }
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j) {
        float* row = a + (size_t)n2 * (j + (size_t)ny * k);
        for (int i = 0; i < hc; ++i) in[i] = cf(row[2 * i], row[2 * i + 1]);
        for (int i = hc; i < nx; ++i) in[i] = std::conj(in[nx - i]);
        // c2r ignores the imaginary parts of the self-conjugate bins, as FFTW does
        in[0] = cf(in[0].real(), 0.f);
        if (nx % 2 == 0) in[nx / 2] = cf(in[nx / 2].real(), 0.f);
        px.exec(in.data(), 1, out.data(), false);
        for (int i = 0; i < nx; ++i) row[i] = out[i].real();
        row[nx] = 0.f; row[nx + 1] = 0.f;
      }
  }
};

}  // namespace oracle
