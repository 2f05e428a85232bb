// Second-generation shared-memory FFT core for two-factor lengths N = R0 * R1 (all fine-tile sizes: 176 = 16*11, 304 = 16*19,
// 112 = 16*7, 80 = 16*5, and the power-of-two coarse sizes up to 256).
//
//  * packed arithmetic: every complex add/sub/rotate is ONE FADD2, every complex multiply two (FMUL2 + FFMA2) — sm_100's
//    f32x2 instructions take operand swizzles (.LO_HI), per-half negation (.NP) and scalar broadcast (.F32) for free. Measured on
//    B200 (profiles/r1_fp32_peak.json) FFMA2 has the FLOP rate of FFMA, so this halves ISSUE SLOTS, which is what bounds these
//    kernels (ncu: issue-active ~55 %, fma pipe ~30 %).
//  * in place on ONE buffer: stage A (radix R0, Stockham NS=1) reads x[j + r*R1] and writes y[j*R0 + r] after a barrier; stage B
//    (radix R1, NS=R0) reads y[j + r*R0] and produces X[j + r*R0] — the same slots — so its results go straight to their consumer
//    (global memory, or back to the same slots) with no further barrier.
//  * one butterfly per thread per stage: a CTA of 16 * max(R0,R1) threads owns 16 columns (lanes along the columns).
// Buffers are AoS float2 [N][PITCH]: PITCH = 16 for the strided passes, 17 for the contiguous-axis passes whose transposing accesses
// run lanes along the sequence (stride 17 float2 = 34 words: conflict-free per half-warp for 64-bit accesses).
#pragma once
#include <type_traits>
#include "fft_smem.cuh"

namespace fftk {

#ifdef __CUDA_ARCH__
FFTK_HD float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
FFTK_HD float2 psub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
FFTK_HD float2 add_irot(float2 a, float2 d) { return __fadd2_rn(a, make_float2(-d.y, d.x)); }   // a + i d
FFTK_HD float2 sub_irot(float2 a, float2 d) { return __fadd2_rn(a, make_float2(d.y, -d.x)); }   // a - i d
FFTK_HD float2 pfma_s(float s, float2 a, float2 c) { return __ffma2_rn(make_float2(s, s), a, c); }   // s*a + c
FFTK_HD float2 pmul_s(float s, float2 a) { return __fmul2_rn(make_float2(s, s), a); }
FFTK_HD float2 pfma_v(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// v * t  and  v * conj(t)
FFTK_HD float2 cmulp(float2 v, float2 t) {
  const float2 p = __fmul2_rn(make_float2(t.y, t.y), make_float2(v.y, v.x));
  return __ffma2_rn(make_float2(t.x, t.x), v, make_float2(-p.x, p.y));
}
FFTK_HD float2 cmulp_conj(float2 v, float2 t) {
  const float2 p = __fmul2_rn(make_float2(t.y, t.y), make_float2(v.y, v.x));
  return __ffma2_rn(make_float2(t.x, t.x), v, make_float2(p.x, -p.y));
}
#else
FFTK_HD float2 padd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FFTK_HD float2 psub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FFTK_HD float2 add_irot(float2 a, float2 d) { return make_float2(a.x - d.y, a.y + d.x); }
FFTK_HD float2 sub_irot(float2 a, float2 d) { return make_float2(a.x + d.y, a.y - d.x); }
FFTK_HD float2 pfma_s(float s, float2 a, float2 c) { return make_float2(s * a.x + c.x, s * a.y + c.y); }
FFTK_HD float2 pmul_s(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
FFTK_HD float2 pfma_v(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
FFTK_HD float2 cmulp(float2 v, float2 t) { return make_float2(v.x * t.x - v.y * t.y, v.x * t.y + v.y * t.x); }
FFTK_HD float2 cmulp_conj(float2 v, float2 t) { return make_float2(v.x * t.x + v.y * t.y, v.y * t.x - v.x * t.y); }
#endif
template <bool INV> FFTK_HD float2 twmul(float2 v, float2 t) { return INV ? cmulp_conj(v, t) : cmulp(v, t); }   // t = exp(-i..)
// a -/+ i d: multiplication of d by the primitive 4th root of the transform direction
template <bool INV> FFTK_HD float2 add_w4(float2 a, float2 d) { return INV ? add_irot(a, d) : sub_irot(a, d); }
template <bool INV> FFTK_HD float2 sub_w4(float2 a, float2 d) { return INV ? sub_irot(a, d) : add_irot(a, d); }

// 16th roots of unity exp(-2 pi i k / 16), k = 0..15, as literals (compile-time operands of the composite butterflies)
FFTK_HD float2 w16(int k) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  switch (k & 15) {
    case 0: return make_float2(1.f, 0.f);
    case 1: return make_float2(c1, -s1);
    case 2: return make_float2(h, -h);
    case 3: return make_float2(s1, -c1);
    case 4: return make_float2(0.f, -1.f);
    case 5: return make_float2(-s1, -c1);
    case 6: return make_float2(-h, -h);
    case 7: return make_float2(-c1, -s1);
    case 8: return make_float2(-1.f, 0.f);
    case 9: return make_float2(-c1, s1);
    case 10: return make_float2(-h, h);
    case 11: return make_float2(-s1, c1);
    case 12: return make_float2(0.f, 1.f);
    case 13: return make_float2(s1, c1);
    case 14: return make_float2(h, h);
    default: return make_float2(c1, s1);
  }
}

// exp(-2 pi i t / R) for the odd prime radices as literals: after unrolling they become immediate operands of FFMA2 (no constant loads)
template <int R> FFTK_HD float2 wodd(int t) {
  if constexpr (R == 3) {
    switch (t) {
      case 0: return make_float2(1.000000000e+00f, -0.000000000e+00f);
      case 1: return make_float2(-5.000000000e-01f, -8.660254038e-01f);
      case 2: return make_float2(-5.000000000e-01f, 8.660254038e-01f);
    }
  }
  if constexpr (R == 5) {
    switch (t) {
      case 0: return make_float2(1.000000000e+00f, -0.000000000e+00f);
      case 1: return make_float2(3.090169944e-01f, -9.510565163e-01f);
      case 2: return make_float2(-8.090169944e-01f, -5.877852523e-01f);
      case 3: return make_float2(-8.090169944e-01f, 5.877852523e-01f);
      case 4: return make_float2(3.090169944e-01f, 9.510565163e-01f);
    }
  }
  if constexpr (R == 7) {
    switch (t) {
      case 0: return make_float2(1.000000000e+00f, -0.000000000e+00f);
      case 1: return make_float2(6.234898019e-01f, -7.818314825e-01f);
      case 2: return make_float2(-2.225209340e-01f, -9.749279122e-01f);
      case 3: return make_float2(-9.009688679e-01f, -4.338837391e-01f);
      case 4: return make_float2(-9.009688679e-01f, 4.338837391e-01f);
      case 5: return make_float2(-2.225209340e-01f, 9.749279122e-01f);
      case 6: return make_float2(6.234898019e-01f, 7.818314825e-01f);
    }
  }
  if constexpr (R == 11) {
    switch (t) {
      case 0: return make_float2(1.000000000e+00f, -0.000000000e+00f);
      case 1: return make_float2(8.412535328e-01f, -5.406408175e-01f);
      case 2: return make_float2(4.154150130e-01f, -9.096319954e-01f);
      case 3: return make_float2(-1.423148383e-01f, -9.898214419e-01f);
      case 4: return make_float2(-6.548607339e-01f, -7.557495744e-01f);
      case 5: return make_float2(-9.594929736e-01f, -2.817325568e-01f);
      case 6: return make_float2(-9.594929736e-01f, 2.817325568e-01f);
      case 7: return make_float2(-6.548607339e-01f, 7.557495744e-01f);
      case 8: return make_float2(-1.423148383e-01f, 9.898214419e-01f);
      case 9: return make_float2(4.154150130e-01f, 9.096319954e-01f);
      case 10: return make_float2(8.412535328e-01f, 5.406408175e-01f);
    }
  }
  if constexpr (R == 19) {
    switch (t) {
      case 0: return make_float2(1.000000000e+00f, -0.000000000e+00f);
      case 1: return make_float2(9.458172417e-01f, -3.246994692e-01f);
      case 2: return make_float2(7.891405094e-01f, -6.142127127e-01f);
      case 3: return make_float2(5.469481581e-01f, -8.371664783e-01f);
      case 4: return make_float2(2.454854871e-01f, -9.694002659e-01f);
      case 5: return make_float2(-8.257934547e-02f, -9.965844930e-01f);
      case 6: return make_float2(-4.016954247e-01f, -9.157733267e-01f);
      case 7: return make_float2(-6.772815716e-01f, -7.357239107e-01f);
      case 8: return make_float2(-8.794737512e-01f, -4.759473930e-01f);
      case 9: return make_float2(-9.863613034e-01f, -1.645945903e-01f);
      case 10: return make_float2(-9.863613034e-01f, 1.645945903e-01f);
      case 11: return make_float2(-8.794737512e-01f, 4.759473930e-01f);
      case 12: return make_float2(-6.772815716e-01f, 7.357239107e-01f);
      case 13: return make_float2(-4.016954247e-01f, 9.157733267e-01f);
      case 14: return make_float2(-8.257934547e-02f, 9.965844930e-01f);
      case 15: return make_float2(2.454854871e-01f, 9.694002659e-01f);
      case 16: return make_float2(5.469481581e-01f, 8.371664783e-01f);
      case 17: return make_float2(7.891405094e-01f, 6.142127127e-01f);
      case 18: return make_float2(9.458172417e-01f, 3.246994692e-01f);
    }
  }
  return make_float2(1.f, 0.f);
}

template <int R, bool INV> struct PRadix;
template <bool INV> struct PRadix<1, INV> { static FFTK_HD void run(float2 (&)[1]) {} };
template <bool INV> struct PRadix<2, INV> {
  static FFTK_HD void run(float2 (&v)[2]) { const float2 a = v[0], b = v[1]; v[0] = padd(a, b); v[1] = psub(a, b); }
};
template <bool INV> struct PRadix<4, INV> {
  static FFTK_HD void run(float2 (&v)[4]) {
    const float2 s0 = padd(v[0], v[2]), s1 = psub(v[0], v[2]), s2 = padd(v[1], v[3]), d = psub(v[1], v[3]);
    v[0] = padd(s0, s2); v[2] = psub(s0, s2); v[1] = add_w4<INV>(s1, d); v[3] = sub_w4<INV>(s1, d);
  }
};
// composite R = R1*R2 with 16th-root twiddles (R divides 16): n = R2*n1 + n2, k = k1 + R1*k2
template <int R1, int R2, bool INV> FFTK_HD void pdft_composite(float2 (&v)[R1 * R2]) {
  constexpr int R = R1 * R2;
  float2 y[R2][R1];
#pragma unroll
  for (int n2 = 0; n2 < R2; ++n2) {
    float2 t[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) t[n1] = v[R2 * n1 + n2];
    PRadix<R1, INV>::run(t);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      const int e = ((n2 * k1) % R) * (16 / R);
      if (e == 0) y[n2][k1] = t[k1];
      else if (e == 4) y[n2][k1] = INV ? make_float2(-t[k1].y, t[k1].x) : make_float2(t[k1].y, -t[k1].x);
      else if (e == 8) y[n2][k1] = make_float2(-t[k1].x, -t[k1].y);
      else y[n2][k1] = twmul<INV>(t[k1], w16(e));
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) {
    float2 t[R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) t[n2] = y[n2][k1];
    PRadix<R2, INV>::run(t);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
  }
}
template <bool INV> struct PRadix<8, INV> { static FFTK_HD void run(float2 (&v)[8]) { pdft_composite<2, 4, INV>(v); } };
template <bool INV> struct PRadix<16, INV> { static FFTK_HD void run(float2 (&v)[16]) { pdft_composite<4, 4, INV>(v); } };

// odd prime radix with the (v_r +- v_{R-r}) symmetry, packed: C = v0 + sum a_r cos, S = sum b_r (-sin); X[u] = C + i S, X[R-u] = C - i S
// (forward; swapped for the inverse). Results are handed to `emit(index, value)` as they are produced, so that only the a/b terms stay
// live (radix 19: 36 + 4 registers instead of 76).
template <int R, bool INV, typename Emit> FFTK_HD void pdft_odd_emit(float2 (&v)[R], Emit emit) {
  constexpr int H = (R - 1) / 2;
  const float2 v0 = v[0];
  float2 s = v0;
#pragma unroll
  for (int r = 0; r < H; ++r) {
    const float2 p = v[r + 1], q = v[R - 1 - r];
    v[r + 1] = padd(p, q);        // a_r
    v[R - 1 - r] = psub(p, q);    // b_r
    s = padd(s, v[r + 1]);
  }
  emit(0, s);
#pragma unroll
  for (int u = 1; u <= H; ++u) {
    float2 C = v0, S = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 1; r <= H; ++r) {
      const float2 w = wodd<R>((r * u) % R);   // (cos, -sin) of 2 pi r u / R
      C = pfma_s(w.x, v[r], C);
      S = (r == 1) ? pmul_s(w.y, v[R - r]) : pfma_s(w.y, v[R - r], S);
    }
    emit(u, INV ? sub_irot(C, S) : add_irot(C, S));
    emit(R - u, INV ? add_irot(C, S) : sub_irot(C, S));
  }
}
template <int R> struct IsOddPrime { static constexpr bool value = (R == 3 || R == 5 || R == 7 || R == 11 || R == 19); };

// generic "run and emit": array radices compute in registers then emit every output
template <int R, bool INV, typename Emit> FFTK_HD void pradix_emit(float2 (&v)[R], Emit emit) {
  if constexpr (IsOddPrime<R>::value) {
    pdft_odd_emit<R, INV>(v, emit);
  } else {
    PRadix<R, INV>::run(v);
#pragma unroll
    for (int r = 0; r < R; ++r) emit(r, v[r]);
  }
}

template <int N> struct Plan2 {
  using F = Factors<N>;
  static constexpr int R0 = F::r0, R1 = F::r1;
  static constexpr bool ok = (F::r1 > 1 && F::r2 == 1);
  static constexpr int RMAX = R0 > R1 ? R0 : R1;
  static constexpr int NT = ((LX * RMAX + 31) / 32) * 32;    // threads per CTA (320 for N = 304, else 256)
  static constexpr int NA = LX * R1;                          // threads with a stage-A butterfly (N/R0 = R1 items per column)
  static constexpr int NB = LX * R0;                          // threads with a stage-B butterfly
};

// stage A, part 1: gather x[j + r*R1] of column `col` and transform (radix R0, no twiddles). j in [0, R1).
template <int N, bool INV, int PITCH> FFTK_HD void stageA_load(const float2* __restrict__ buf, int j, int col, float2 (&v)[Plan2<N>::R0]) {
  constexpr int R0 = Plan2<N>::R0, R1 = Plan2<N>::R1;
  const float2* p = buf + j * PITCH + col;
#pragma unroll
  for (int r = 0; r < R0; ++r) v[r] = p[r * R1 * PITCH];
}
// stage A, part 2 (after a barrier): y[j*R0 + r] = v[r]
template <int N, int PITCH> FFTK_HD void stageA_store(float2* __restrict__ buf, int j, int col, const float2 (&v)[Plan2<N>::R0]) {
  constexpr int R0 = Plan2<N>::R0;
  float2* q = buf + (j * R0) * PITCH + col;
#pragma unroll
  for (int r = 0; r < R0; ++r) q[r * PITCH] = v[r];
}
// stage B: read y[j + r*R0], twiddle by tw[r*j] (tw[t] = exp(-2 pi i t/N), conjugated for the inverse), radix R1; output index
// j + r*R0 is handed to emit(r, value). j in [0, R0).
template <int N, bool INV, int PITCH, typename Emit>
FFTK_HD void stageB(const float2* __restrict__ buf, const float2* __restrict__ tw, int j, int col, Emit emit) {
  constexpr int R0 = Plan2<N>::R0, R1 = Plan2<N>::R1;
  float2 v[R1];
  const float2* p = buf + j * PITCH + col;
#pragma unroll
  for (int r = 0; r < R1; ++r) v[r] = p[r * R0 * PITCH];
  const float2* t0 = tw + j;
#pragma unroll
  for (int r = 1; r < R1; ++r) v[r] = twmul<INV>(v[r], t0[(r - 1) * j]);   // tw[r*j]
  pradix_emit<R1, INV>(v, emit);
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Contiguous-axis (x) c2r with lanes along the SEQUENCE: the half-spectrum rows arrive row-contiguous (bulk copies), so running the lanes
// of a warp along k makes every shared-memory access of the transform contiguous and removes the separate "tangle" pass (and its two
// barriers) of the column-major kernels: stage A reads A[k], B[k] of the two real rows packed into one complex column straight from the
// staged rows, forms x[idx] = A + iB (idx <= N/2) or conj(A) + i conj(B) (idx > N/2) in registers, and stores its radix-16 results to a
// column-major work buffer Y; stage B reads Y and hands its results to the caller (global stores), so Y is never written twice.
//   thread (c, j) of stage A: j in [0, R1), elements idx = j + r*R1, r = 0..15; k = idx (r < 8, or r = 8 and j = 0) else N - idx
//   Y[c][pos(i)], pos(i) = i + (i >> 4): stage A stores y[16 j + r] at 17 j + r, stage B loads y[j + 16 r] at j + 17 r — both run the
//   lanes along j with a stride of 17 (A) or 1 (B) float2, conflict-free; the column pitch 17*R1 == R1 (mod 16) keeps consecutive
//   threads on consecutive 8-byte bank slots across a column change.
template <int N> struct XRow {
  using P = Plan2<N>;
  static_assert(P::R0 == 16, "XRow needs N = 16 * R1");
  static constexpr int R1 = P::R1;
  static constexpr int HC = N / 2 + 1;
  static constexpr int RP = (HC + 1) / 2 * 2;      // staged row length (float2): even, so that a row is a multiple of 16 bytes
  static constexpr int YP = 17 * R1;               // column pitch of Y (float2)
  template <int CW> static constexpr int NA_ = CW * R1;      // threads with a stage-A / stage-B butterfly for CW columns per item
  template <int CW> static constexpr int NB_ = CW * 16;
  static constexpr int NA = LX * R1, NB = LX * 16;
  static constexpr int NTW = (R1 - 1) * 16;        // transposed twiddles twT[(r-1)*16 + j] = exp(-2 pi i r j / N)
};

template <int N> FFTK_HD void c2r_stageA_load(const float2* __restrict__ rowA, const float2* __restrict__ rowB, int j, float2 (&v)[16]) {
  constexpr int R1 = XRow<N>::R1;
  const bool j0 = (j == 0);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    if (r < 8) {
      float2 a = rowA[j + r * R1], b = rowB[j + r * R1];
      if (r == 0 && j0) { a.y = 0.f; b.y = 0.f; }              // imaginary part of the k = 0 bin is dropped (c2r)
      v[r] = add_irot(a, b);                                     // A + iB
    } else if (r == 8) {
      const int k = j0 ? 8 * R1 : 8 * R1 - j;                    // j = 0: the k = N/2 bin itself (imaginary part dropped)
      const float2 a = rowA[k], b = rowB[k];
      v[r] = j0 ? make_float2(a.x, b.x) : make_float2(a.x + b.y, b.x - a.y);
    } else {
      const int k = (16 - r) * R1 - j;                           // N - idx
      const float2 a = rowA[k], b = rowB[k];
      v[r] = make_float2(a.x + b.y, b.x - a.y);                  // conj(A) + i conj(B)
    }
  }
}
template <int N> FFTK_HD void c2r_stageA_store(float2* __restrict__ ycol, int j, const float2 (&v)[16]) {
  float2* q = ycol + 17 * j;
#pragma unroll
  for (int r = 0; r < 16; ++r) q[r] = v[r];
}
template <int N> FFTK_HD void c2r_stageB_load(const float2* __restrict__ ycol, int j, float2 (&u)[XRow<N>::R1]) {
#pragma unroll
  for (int r = 0; r < XRow<N>::R1; ++r) u[r] = ycol[j + 17 * r];
}
// inverse twiddles + radix R1; output x = j + 16 r goes to emit(r, value)
template <int N, typename Emit> FFTK_HD void c2r_stageB_finish(float2 (&u)[XRow<N>::R1], const float2* __restrict__ twT, int j, Emit emit) {
  constexpr int R1 = XRow<N>::R1;
#pragma unroll
  for (int r = 1; r < R1; ++r) u[r] = twmul<true>(u[r], twT[(r - 1) * 16 + j]);
  pradix_emit<R1, true>(u, emit);
}

constexpr size_t smem_bytes2(int n, int nbuf, int pitch) { return (size_t)nbuf * n * pitch * sizeof(float2) + (size_t)n * sizeof(float2); }

}  // namespace fftk
