// Shared-memory Stockham FFT building blocks (hand-written; replaces the FFTW calls of
// fftw2.f90:19-22 / fft_fine.f90:47-51 / fft_coarse.f90:181-186).
//
// One CTA transforms LX independent complex sequences of length N ("columns") that live in shared
// memory as two float planes re[N][LXP], im[N][LXP] (LXP = LX+1: column index is the fast one, so the
// butterflies — lanes along columns — are conflict-free, and the transposing loads/stores of the
// contiguous-axis pass — lanes along the sequence — hit stride-17 banks, also conflict-free).
// N is factored at compile time into up to three radices; each radix is a register butterfly.
#pragma once
#include <cuda_runtime.h>

namespace fftk {

constexpr int LX = 16;        // columns per CTA
constexpr int LXP = LX + 1;   // padded pitch (floats)
constexpr int NT = 256;       // threads per CTA

// exp(-2 pi i t / R) for every supported radix R (row R of the table), filled by fft_init_constants().
constexpr int MAXR = 20;
// The library is built as ONE translation unit (lib.cu), so these are plain definitions.
__constant__ float2 c_w[MAXR][MAXR];
static float2 h_w[MAXR][MAXR];   // host mirror (also used by the CPU emulation test of the butterflies)
#ifdef __CUDA_ARCH__
#define FFTK_W(R, t) c_w[R][t]
#define FFTK_HD __device__ __forceinline__
#else
#define FFTK_W(R, t) h_w[R][t]
#define FFTK_HD __host__ __device__ inline
#endif

template <int N> struct Factors;  // r0*r1*r2 == N, r2 may be 1
#define FFTK_FACTORS(N_, A, B, C)                                     \
  template <> struct Factors<N_> { static constexpr int r0 = A, r1 = B, r2 = C; };
FFTK_FACTORS(16, 16, 1, 1)
FFTK_FACTORS(32, 16, 2, 1)
FFTK_FACTORS(48, 16, 3, 1)
FFTK_FACTORS(64, 16, 4, 1)
FFTK_FACTORS(80, 16, 5, 1)
FFTK_FACTORS(112, 16, 7, 1)
FFTK_FACTORS(128, 16, 8, 1)
FFTK_FACTORS(176, 16, 11, 1)
FFTK_FACTORS(256, 16, 16, 1)
FFTK_FACTORS(304, 16, 19, 1)
FFTK_FACTORS(512, 16, 16, 2)
FFTK_FACTORS(560, 16, 7, 5)

FFTK_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FFTK_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FFTK_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> FFTK_HD float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
template <bool INV> FFTK_HD float2 wconst(int R, int t) {
  float2 w = FFTK_W(R, t);
  return INV ? make_float2(w.x, -w.y) : w;
}

template <int R, bool INV> struct Radix;

template <bool INV> struct Radix<1, INV> { static FFTK_HD void run(float2 (&)[1]) {} };
template <bool INV> struct Radix<2, INV> {
  static FFTK_HD void run(float2 (&v)[2]) { float2 a = v[0], b = v[1]; v[0] = cadd(a, b); v[1] = csub(a, b); }
};
template <bool INV> struct Radix<4, INV> {
  static FFTK_HD void run(float2 (&v)[4]) {
    float2 s0 = cadd(v[0], v[2]), s1 = csub(v[0], v[2]), s2 = cadd(v[1], v[3]), s3 = mul_mi<INV>(csub(v[1], v[3]));
    v[0] = cadd(s0, s2); v[1] = cadd(s1, s3); v[2] = csub(s0, s2); v[3] = csub(s1, s3);
  }
};
// composite R = R1*R2: n = R2*n1 + n2, k = k1 + R1*k2
template <int R1, int R2, bool INV> FFTK_HD void dft_composite(float2 (&v)[R1 * R2]) {
  constexpr int R = R1 * R2;
  float2 y[R2][R1];
#pragma unroll
  for (int n2 = 0; n2 < R2; ++n2) {
    float2 t[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) t[n1] = v[R2 * n1 + n2];
    Radix<R1, INV>::run(t);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) y[n2][k1] = (n2 * k1 == 0) ? t[k1] : cmul(t[k1], wconst<INV>(R, (n2 * k1) % R));
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) {
    float2 t[R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) t[n2] = y[n2][k1];
    Radix<R2, INV>::run(t);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
  }
}
template <bool INV> struct Radix<8, INV> { static FFTK_HD void run(float2 (&v)[8]) { dft_composite<2, 4, INV>(v); } };
template <bool INV> struct Radix<16, INV> { static FFTK_HD void run(float2 (&v)[16]) { dft_composite<4, 4, INV>(v); } };

// odd prime radix, O(R^2/2) with the (v_r +- v_{R-r}) symmetry
template <int R, bool INV> FFTK_HD void dft_odd(float2 (&v)[R]) {
  constexpr int H = (R - 1) / 2;
  float2 a[H], b[H];
#pragma unroll
  for (int r = 0; r < H; ++r) { a[r] = cadd(v[r + 1], v[R - 1 - r]); b[r] = csub(v[r + 1], v[R - 1 - r]); }
  float2 v0 = v[0];
  float2 s = v0;
#pragma unroll
  for (int r = 0; r < H; ++r) s = cadd(s, a[r]);
  v[0] = s;
#pragma unroll
  for (int u = 1; u <= H; ++u) {
    float cr = v0.x, ci = v0.y;   // sum a_r cos
    float sr = 0.f, si = 0.f;     // sum b_r sin
#pragma unroll
    for (int r = 1; r <= H; ++r) {
      const float2 w = FFTK_W(R, (r * u) % R);   // (cos, -sin) of 2 pi r u / R
      cr = fmaf(a[r - 1].x, w.x, cr); ci = fmaf(a[r - 1].y, w.x, ci);
      sr = fmaf(b[r - 1].x, w.y, sr); si = fmaf(b[r - 1].y, w.y, si);
    }
    // forward: X[u] = sum v_r (cos - i sin) => A + (-i)(sum b sin_pos) with w.y = -sin: i*w.y*b
    // i*(sr + i si) = (-si, sr)
    if (!INV) { v[u] = make_float2(cr - si, ci + sr); v[R - u] = make_float2(cr + si, ci - sr); }
    else      { v[u] = make_float2(cr + si, ci - sr); v[R - u] = make_float2(cr - si, ci + sr); }
  }
}
template <bool INV> struct Radix<3, INV> { static FFTK_HD void run(float2 (&v)[3]) { dft_odd<3, INV>(v); } };
template <bool INV> struct Radix<5, INV> { static FFTK_HD void run(float2 (&v)[5]) { dft_odd<5, INV>(v); } };
template <bool INV> struct Radix<7, INV> { static FFTK_HD void run(float2 (&v)[7]) { dft_odd<7, INV>(v); } };
template <bool INV> struct Radix<11, INV> { static FFTK_HD void run(float2 (&v)[11]) { dft_odd<11, INV>(v); } };
template <bool INV> struct Radix<19, INV> { static FFTK_HD void run(float2 (&v)[19]) { dft_odd<19, INV>(v); } };

// One Stockham stage: radix R, NS = product of the radices already applied.
// tw: exp(-2 pi i t / N), t in [0,N), in shared memory.
template <int N, int R, int NS, bool INV>
FFTK_HD void stage(const float* __restrict__ ire, const float* __restrict__ iim,
                                      float* __restrict__ ore, float* __restrict__ oim, const float2* __restrict__ tw, int tid) {
  constexpr int NB = N / R;
  for (int w = tid; w < NB * LX; w += NT) {
    const int col = w % LX, j = w / LX;
    const int k = j % NS;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { const int idx = (j + r * NB) * LXP + col; v[r] = make_float2(ire[idx], iim[idx]); }
    if (NS > 1) {
      constexpr int TS = N / (NS * R);
#pragma unroll
      for (int r = 1; r < R; ++r) {
        float2 t = tw[r * k * TS];
        if (INV) t.y = -t.y;
        v[r] = cmul(v[r], t);
      }
    }
    Radix<R, INV>::run(v);
    const int j0 = (j / NS) * NS * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) { const int idx = (j0 + r * NS) * LXP + col; ore[idx] = v[r].x; oim[idx] = v[r].y; }
  }
}

// Full transform of the LX columns held in buffer 0 (re0/im0). Buffers ping-pong; returns (at compile time)
// the index of the buffer holding the result via result_buffer<N>().
template <int N> __host__ __device__ constexpr int num_stages() { return 1 + (Factors<N>::r1 > 1) + (Factors<N>::r2 > 1); }
template <int N> __host__ __device__ constexpr int result_buffer() { return num_stages<N>() & 1; }

template <int N, bool INV>
FFTK_HD void fft_columns(float* re0, float* im0, float* re1, float* im1, const float2* tw) {
  using F = Factors<N>;
#ifdef __CUDA_ARCH__
  const int tid = threadIdx.x;
  stage<N, F::r0, 1, INV>(re0, im0, re1, im1, tw, tid);
  __syncthreads();
  if (F::r1 > 1) {
    stage<N, F::r1, F::r0, INV>(re1, im1, re0, im0, tw, tid);
    __syncthreads();
  }
  if (F::r2 > 1) {
    stage<N, F::r2, F::r0 * F::r1, INV>(re0, im0, re1, im1, tw, tid);
    __syncthreads();
  }
#else
  for (int tid = 0; tid < NT; ++tid) stage<N, F::r0, 1, INV>(re0, im0, re1, im1, tw, tid);
  if (F::r1 > 1) for (int tid = 0; tid < NT; ++tid) stage<N, F::r1, F::r0, INV>(re1, im1, re0, im0, tw, tid);
  if (F::r2 > 1) for (int tid = 0; tid < NT; ++tid) stage<N, F::r2, F::r0 * F::r1, INV>(re0, im0, re1, im1, tw, tid);
#endif
}


// ---- AoS variant for the strided passes: columns live as float2 buf[N][LX] (no padding). Lanes run along the 16 columns, so a
// warp touches two 128-byte rows per 64-bit access: conflict-free, and one LDS.64/STS.64 per point instead of two 32-bit ones.
template <int N, int R, int NS, bool INV>
FFTK_HD void stage_aos(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, int tid) {
  constexpr int NB = N / R;
  for (int w = tid; w < NB * LX; w += NT) {
    const int col = w % LX, j = w / LX;
    const int k = j % NS;
    const float2* p = in + (j * LX + col);
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = p[r * NB * LX];
    if (NS > 1) {
      constexpr int TS = N / (NS * R);
      const float2* t0 = tw + k * TS;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        float2 t = t0[(r - 1) * k * TS];      // tw[r*k*TS]
        if (INV) t.y = -t.y;
        v[r] = cmul(v[r], t);
      }
    }
    Radix<R, INV>::run(v);
    float2* q = out + (((j / NS) * NS * R + k) * LX + col);
#pragma unroll
    for (int r = 0; r < R; ++r) q[r * NS * LX] = v[r];
  }
}

template <int N, bool INV>
FFTK_HD void fft_columns_aos(float2* b0, float2* b1, const float2* tw) {
  using F = Factors<N>;
#ifdef __CUDA_ARCH__
  const int tid = threadIdx.x;
  stage_aos<N, F::r0, 1, INV>(b0, b1, tw, tid);
  __syncthreads();
  if (F::r1 > 1) {
    stage_aos<N, F::r1, F::r0, INV>(b1, b0, tw, tid);
    __syncthreads();
  }
  if (F::r2 > 1) {
    stage_aos<N, F::r2, F::r0 * F::r1, INV>(b0, b1, tw, tid);
    __syncthreads();
  }
#else
  for (int tid = 0; tid < NT; ++tid) stage_aos<N, F::r0, 1, INV>(b0, b1, tw, tid);
  if (F::r1 > 1) for (int tid = 0; tid < NT; ++tid) stage_aos<N, F::r1, F::r0, INV>(b1, b0, tw, tid);
  if (F::r2 > 1) for (int tid = 0; tid < NT; ++tid) stage_aos<N, F::r2, F::r0 * F::r1, INV>(b0, b1, tw, tid);
#endif
}
constexpr size_t smem_bytes_aos(int n, int nbuf) { return (size_t)nbuf * n * LX * sizeof(float2) + (size_t)n * sizeof(float2); }

constexpr size_t smem_bytes(int n) { return (size_t)4 * n * LXP * sizeof(float) + (size_t)n * sizeof(float2); }

}  // namespace fftk
