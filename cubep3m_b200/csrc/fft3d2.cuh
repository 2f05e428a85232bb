// Second-generation FFT pass kernels (see fft2.cuh for the core): used for every two-factor N (Plan2<N>::ok); the first-generation
// kernels of fft3d.cuh remain for N = 16 and the three-factor sizes (512, 560).
//
//   fft_strided2     Y / Z pass. The next work item's column block is staged with cp.async (8-byte, the spectrum pitch hc*8 B is not
//                    16-byte aligned so TMA/bulk copies cannot be used) into the other of two buffers while the current one is
//                    transformed IN PLACE; stage-B results go straight from registers to global memory (128 B per half-warp).
//   fft_z_sandwich2  forward z + 3 x (multiply by i*kern_f(comp), inverse z), two buffers: the forward result stays in the landing
//                    buffer, components 0 and 1 run A: L -> W, B: W -> global, component 2 runs in place on L, which frees W for
//                    the cp.async prefetch of the next item during the whole third inverse transform. The Green's-function block of
//                    the NEXT component is staged with 16-byte cp.async into a third, small buffer (kern_f is stored with a pitch of
//                    a multiple of 16 floats for this) while the current component is transformed.
//   Addressing is 32-bit and incremental (one decode per item, pointer stepping, crop tests as a per-thread bitmask): the first
//   version of these kernels spent 75 % of its issue slots on integer work (profiles/r1_ncu_v2_first_try.txt).
//   fft_x_r2c_ngp2   NGP density from the fine-cell table + r2c along x. One thread produces a whole coarse cell's (fz fixed) 4x4 block
//                    of densities from 17 consecutive table entries (four 16-byte loads + one), all loads in flight before first use.
//   fft_x_c2r3_v2    c2r along x for the 3 force components with crop + 1/n^3 + max |F|^2; half-spectrum rows are staged with cp.async
//                    (next component prefetched during the current FFT).
#pragma once
#include "fft2.cuh"
// (included from the middle of fft3d.cuh, after the first-generation kernels and before the host-side dispatch)

namespace fftk {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(unsigned sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// bit r set iff stage-B output e = j + r*R0 lies in [elo, ehi]
template <int N> __device__ __forceinline__ unsigned crop_mask(int j, int elo, int ehi) {
  unsigned m = 0;
#pragma unroll
  for (int r = 0; r < Plan2<N>::R1; ++r) { const int e = j + r * Plan2<N>::R0; if (e >= elo && e <= ehi) m |= 1u << r; }
  return m;
}

// stage one column block (N rows of 16 float2, row stride `estride` elements) into dst[N][16].
// A16 = false: thread (j, col) copies element col of rows j, j+ES, ... with 8-byte cp.async (any pitch); columns beyond hc are zero-filled.
// A16 = true : the block starts on a 128-byte boundary and the pitch is even (padded spectra): 8 threads copy one row with 16-byte cp.async,
//              half as many copies; columns beyond hc are read from the (zero) padding.
template <int N, int NTH, bool A16, typename ES = int> __device__ __forceinline__ void stage_block(const float2* __restrict__ blk, ES estride, bool ok, unsigned sbuf,
                                                                                float2* __restrict__ gbuf) {
  if constexpr (A16) {
    constexpr int RPP = NTH / 8;                   // rows per pass
    const int row = threadIdx.x >> 3, q = threadIdx.x & 7;
    const float2* src = blk + (long long)row * estride + q * 2;
    const unsigned sd = sbuf + (row * LX + q * 2) * 8;
#pragma unroll
    for (int it = 0; it < (N + RPP - 1) / RPP; ++it) {
      if (row + it * RPP < N) cp_async16(sd + it * RPP * LX * 8, src);
      src += (long long)RPP * estride;
    }
  } else {
    constexpr int ES = NTH / LX;
    const int col = threadIdx.x % LX, j = threadIdx.x / LX;
    const float2* src = blk + (long long)j * estride + col;
    const unsigned sd = sbuf + (j * LX + col) * 8;
    float2* dgen = gbuf + j * LX + col;
#pragma unroll
    for (int it = 0; it < (N + ES - 1) / ES; ++it) {
      if (j + it * ES < N) {
        if (ok) cp_async8(sd + it * ES * LX * 8, src);
        else dgen[it * ES * LX] = make_float2(0.f, 0.f);
        src += (long long)ES * estride;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- strided pass
// element e of column c of item (bx, outer, bz): in[bz*bstride + (outer + outer0)*ostride + bx*16 + e*estride + c]   (all offsets < 2^31)
// W64: 64-bit element offsets for arrays beyond 4 GB (the 1024^3 / 2048^3-slab spectra of cic_power, bigfft.cuh); the 32-bit form is the hot path.
template <int N, bool INV, bool MUL, bool A16, bool W64 = false>
__global__ void __launch_bounds__(Plan2<N>::NT, 2) fft_strided2(const float2* __restrict__ in, float2* __restrict__ out, int hc,
                                                                typename std::conditional<W64, long long, int>::type estride,
                                                                typename std::conditional<W64, long long, int>::type ostride,
                                                                int outer0, int nouter, int nbatch, const float* __restrict__ kern, int kes, int kos,
                                                                int elo, int ehi, const float2* __restrict__ tw_g,
                                                                typename std::conditional<W64, long long, int>::type bstride) {
  using off_t = typename std::conditional<W64, long long, int>::type;
  using P = Plan2<N>;
  constexpr int NT2 = P::NT, R0 = P::R0, R1 = P::R1;
  extern __shared__ __align__(16) unsigned char raw[];
  float2* buf0 = reinterpret_cast<float2*>(raw);
  float2* buf1 = buf0 + N * LX;
  float2* tw = buf1 + N * LX;
  for (int t = threadIdx.x; t < N; t += NT2) tw[t] = tw_g[t];
  const int col = threadIdx.x % LX, j = threadIdx.x / LX;
  const unsigned nbx = (hc + LX - 1) / LX;
  const int total = (int)nbx * nouter * nbatch;
  const unsigned mask = crop_mask<N>(j, elo, ehi);
  const unsigned s0 = smem_u32(buf0), s1 = smem_u32(buf1);
  const off_t joff = (off_t)j * estride + col;
  auto decode = [&](int item, off_t& off, int& koff, bool& ok) {      // off: offset of the block's first element (column 0, element 0)
    const unsigned bx = (unsigned)item % nbx, t = (unsigned)item / nbx;
    const unsigned o = t % (unsigned)nouter, bz = t / (unsigned)nouter;
    const int kx = (int)bx * LX;
    off = (off_t)bz * bstride + (off_t)((int)o + outer0) * ostride + kx;
    koff = ((int)o + outer0) * kos + kx + col;
    ok = kx + col < hc;
  };
  int item = blockIdx.x, koff = 0;
  off_t off = 0;
  bool ok = false;
  int p = 0;
  if (item < total) {
    decode(item, off, koff, ok);
    stage_block<N, NT2, A16>(in + off, estride, ok, s0, buf0);
    cp_async_commit();
  }
  while (item < total) {
    float2* buf = p ? buf1 : buf0;
    const int next = item + gridDim.x;
    off_t noff = 0;
    int nkoff = 0;
    bool nok = false;
    cp_async_wait_all();
    __syncthreads();                               // item landed; every thread is done with the other buffer
    if (next < total) {
      decode(next, noff, nkoff, nok);
      stage_block<N, NT2, A16>(in + noff, estride, nok, p ? s0 : s1, p ? buf0 : buf1);
      cp_async_commit();
    }
    float2 v[R0];
    if (threadIdx.x < P::NA) {
      stageA_load<N, INV, LX>(buf, j, col, v);
      if (MUL) {
        const float* kp = kern + koff + j * kes;
#pragma unroll
        for (int r = 0; r < R0; ++r) {
          const float kv = ok ? kp[(long long)r * R1 * kes] : 0.f;
          v[r] = make_float2(-v[r].y * kv, v[r].x * kv);
        }
      }
      PRadix<R0, INV>::run(v);
    }
    __syncthreads();
    if (threadIdx.x < P::NA) stageA_store<N, LX>(buf, j, col, v);
    __syncthreads();
    if (threadIdx.x < P::NB) {
      // 32-bit byte offsets from the array base: one IMAD per store instead of a 64-bit address chain rebuilt for every output
      const unsigned m = ok ? mask : 0u;
      if constexpr (W64) {
        float2* op = out + (off + joff);
        const long long step = (long long)R0 * estride;
        stageB<N, INV, LX>(buf, tw, j, col, [&](int r, float2 val) { if (m & (1u << r)) op[(long long)r * step] = val; });
      } else {
        const unsigned ob = (unsigned)(off + joff) * 8u, stepb = (unsigned)(R0 * estride) * 8u;
        char* obase = reinterpret_cast<char*>(out);
        stageB<N, INV, LX>(buf, tw, j, col, [&](int r, float2 val) {
          if (m & (1u << r)) *reinterpret_cast<float2*>(obase + (ob + (unsigned)r * stepb)) = val;
        });
      }
    }
    item = next; off = noff; koff = nkoff; ok = nok;
    p ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------- fused z pass
// kern: [comp][z][y][kp] floats, kp a multiple of 16 (64-byte aligned 16-column blocks), comp stride kstride.
// spec / g: complex rows of pitch cp (>= hc) float2.
template <int N, bool A16>
__global__ void __launch_bounds__(Plan2<N>::NT, 2) fft_z_sandwich2(const float2* __restrict__ spec, float2* __restrict__ g, int gstride, int hc, int cp, int ny,
                                                                   const float* __restrict__ kern, long long kstride, int kp, int elo, int ehi,
                                                                   const float2* __restrict__ tw_g) {
  using P = Plan2<N>;
  constexpr int NT2 = P::NT, R0 = P::R0, R1 = P::R1;
  extern __shared__ __align__(16) unsigned char raw[];
  float2* L = reinterpret_cast<float2*>(raw);
  float2* W = L + N * LX;
  float* K = reinterpret_cast<float*>(W + N * LX);
  float2* tw = reinterpret_cast<float2*>(K + N * LX);
  for (int t = threadIdx.x; t < N; t += NT2) tw[t] = tw_g[t];
  const int col = threadIdx.x % LX, j = threadIdx.x / LX;
  const unsigned nbx = (hc + LX - 1) / LX;
  const int total = (int)nbx * ny;
  const int estride = ny * cp;
  const long long kes = (long long)ny * kp;         // z stride of the Green's function table
  const unsigned mask = crop_mask<N>(j, elo, ehi);
  const int slot = j * LX + col, joff = j * estride + col;
  unsigned sL = smem_u32(L), sW = smem_u32(W);
  const unsigned sK = smem_u32(K);
  const bool actA = threadIdx.x < P::NA, actB = threadIdx.x < P::NB;
  auto decode = [&](int item, int& off, int& kb, bool& ok) {
    const unsigned bx = (unsigned)item % nbx, y = (unsigned)item / nbx;
    off = (int)y * cp + (int)bx * LX;
    kb = (int)y * kp + (int)bx * LX;
    ok = (int)bx * LX + col < hc;
  };
  auto issueK = [&](int comp, int kb) {             // N rows of 16 floats = 4 x 16 bytes each
    const float* src = kern + comp * kstride + kb;
#pragma unroll
    for (int it = 0; it < (4 * N + NT2 - 1) / NT2; ++it) {
      const int t = threadIdx.x + it * NT2;
      if (t < 4 * N) cp_async16(sK + t * 16, src + (long long)(t >> 2) * kes + (t & 3) * 4);
    }
    cp_async_commit();
  };
  auto stageA_inv = [&](float2 (&v)[R0]) {          // gather S, multiply by i*kern_f (:188-189), radix R0
    stageA_load<N, true, LX>(L, j, col, v);
    const float* kq = K + slot;
#pragma unroll
    for (int r = 0; r < R0; ++r) { const float kv = kq[r * R1 * LX]; v[r] = make_float2(-v[r].y * kv, v[r].x * kv); }
    PRadix<R0, true>::run(v);
  };
  int item = blockIdx.x, off = 0, kb = 0;
  bool ok = false;
  if (item < total) {
    decode(item, off, kb, ok);
    stage_block<N, NT2, A16>(spec + off, estride, ok, sL, L);
    issueK(0, kb);
  }
  // stage-B stores: 32-bit byte offsets from g (one IMAD per store instead of a 64-bit address chain rebuilt for every output)
  const unsigned stepb = (unsigned)(R0 * estride) * 8u;
  char* gbase = reinterpret_cast<char*>(g);
  while (item < total) {
    const int next = item + gridDim.x;
    int noff = 0, nkb = 0;
    bool nok = false;
    if (next < total) decode(next, noff, nkb, nok);
    const unsigned m = ok ? mask : 0u;
    cp_async_wait_all();
    __syncthreads();                                // spectrum block and kern_f(0) block landed
    float2 v[R0];
    // ---- forward transform in place on L
    if (actA) { stageA_load<N, false, LX>(L, j, col, v); PRadix<R0, false>::run(v); }
    __syncthreads();
    if (actA) stageA_store<N, LX>(L, j, col, v);
    __syncthreads();
    if (actB) {
      float2* sp = L + slot;
      stageB<N, false, LX>(L, tw, j, col, [&](int r, float2 val) { sp[r * R0 * LX] = val; });   // same slots this thread read
    }
    __syncthreads();
    // ---- components 0 and 1: A: L -> W, B: W -> global
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
      if (actA) { stageA_inv(v); stageA_store<N, LX>(W, j, col, v); }
      __syncthreads();                              // W complete; K consumed
      issueK(comp + 1, kb);
      if (actB) {
        const unsigned ob = (unsigned)(comp * gstride + off + joff) * 8u;
        stageB<N, true, LX>(W, tw, j, col, [&](int r, float2 val) { if (m & (1u << r)) *reinterpret_cast<float2*>(gbase + (ob + (unsigned)r * stepb)) = val; });
      }
      cp_async_wait_all();
      __syncthreads();                              // next kern_f block landed; W free
    }
    if (next < total) { stage_block<N, NT2, A16>(spec + noff, estride, nok, sW, W); cp_async_commit(); }
    // ---- component 2 in place on L
    if (actA) stageA_inv(v);
    __syncthreads();                                // all gathers from L and K precede the scatter / the next kern_f block
    if (next < total) issueK(0, nkb);
    if (actA) stageA_store<N, LX>(L, j, col, v);
    __syncthreads();
    if (actB) {
      const unsigned ob = (unsigned)(2 * gstride + off + joff) * 8u;
      stageB<N, true, LX>(L, tw, j, col, [&](int r, float2 val) { if (m & (1u << r)) *reinterpret_cast<float2*>(gbase + (ob + (unsigned)r * stepb)) = val; });
    }
    { float2* t = L; L = W; W = t; const unsigned u = sL; sL = sW; sW = u; }
    item = next; off = noff; kb = nkb; ok = nok;
  }
}
constexpr size_t smem_bytes_sandwich2(int n) { return (size_t)2 * n * LX * sizeof(float2) + (size_t)n * LX * sizeof(float) + (size_t)n * sizeof(float2); }

// ---------------------------------------------------------------------------------------------- fused z pass, 8-column items
// Same pipeline as fft_z_sandwich2 on items of 8 columns (64-byte rows): 160-thread CTAs, FOUR per SM instead of two 320-thread ones — the same
// warps and registers per SM, but four independent barrier domains whose phases interleave (what the row-major c2r pass gained from its 8-column
// items), and half the tail of the persistent grid in time. Two layouts per buffer: X = natural order [N][8] (what the copies deliver and what
// stage A gathers from, x[j + r R1]: consecutive j are consecutive rows, conflict-free), Y = stage A's output y[j R0 + r] with every group of R0
// rows skewed by one row ((R0 + 1) * 8 float2 per group): a half-warp's two j then fall into different halves of a 128-byte bank line, and stage B
// reads y[j + r R0] as consecutive rows again. Because X and Y differ, the forward transform runs A: L(X) -> W(Y), B: W(Y) -> L(X) (one barrier
// fewer than in place), components 0 and 1 run A: L(X) -> W(Y), B: W(Y) -> global, component 2 runs A: L(X) -> registers -> L(Y), B: L(Y) -> global
// while W receives the next item.
template <int N> struct Sand3 {
  using P = Plan2<N>;
  static constexpr int CW = 8;
  static constexpr int R0 = P::R0, R1 = P::R1;
  static constexpr int NT = ((CW * P::RMAX + 31) / 32) * 32;      // 160 for N = 304
  static constexpr int NA = CW * R1, NB = CW * R0;
  static constexpr int YP = (R0 + 1) * CW;                        // float2 per skewed group of R0 rows
  static constexpr int BUF = R1 * YP;                             // float2 per buffer (>= N * CW)
  static constexpr int CPS = 4;                                   // CTAs per SM
  static constexpr size_t smem = (size_t)2 * BUF * sizeof(float2) + (size_t)N * CW * sizeof(float) + (size_t)N * sizeof(float2);
};

template <int N>
__global__ void __launch_bounds__(Sand3<N>::NT, Sand3<N>::CPS) fft_z_sandwich3(const float2* __restrict__ spec, float2* __restrict__ g, int gstride, int hc, int cp, int ny,
                                                                                const float* __restrict__ kern, long long kstride, int kp, int elo, int ehi,
                                                                                const float2* __restrict__ tw_g) {
  using S = Sand3<N>;
  constexpr int NT3 = S::NT, R0 = S::R0, R1 = S::R1, CW = S::CW, YP = S::YP;
  extern __shared__ __align__(16) unsigned char raw[];
  float2* L = reinterpret_cast<float2*>(raw);
  float2* W = L + S::BUF;
  float* K = reinterpret_cast<float*>(W + S::BUF);
  float2* tw = reinterpret_cast<float2*>(K + N * CW);
  for (int t = threadIdx.x; t < N; t += NT3) tw[t] = tw_g[t];
  const int col = threadIdx.x % CW, j = threadIdx.x / CW;
  const unsigned nbx = (hc + CW - 1) / CW;
  const int total = (int)nbx * ny;
  const int estride = ny * cp;
  const long long kes = (long long)ny * kp;         // z stride of the Green's function table
  const unsigned mask = crop_mask<N>(j, elo, ehi);
  const int joff = j * estride + col;
  unsigned sL = smem_u32(L), sW = smem_u32(W);
  const unsigned sK = smem_u32(K);
  const bool actA = threadIdx.x < S::NA, actB = threadIdx.x < S::NB;
  auto decode = [&](int item, int& off, int& kb, bool& ok) {
    const unsigned bx = (unsigned)item % nbx, y = (unsigned)item / nbx;
    off = (int)y * cp + (int)bx * CW;
    kb = (int)y * kp + (int)bx * CW;
    ok = (int)bx * CW + col < hc;
  };
  auto issueS = [&](int off, unsigned sdst) {        // N rows of 8 float2 = 4 x 16 bytes each, natural (X) order
    constexpr int RPP = NT3 / 4;
    const int row = threadIdx.x >> 2, q = threadIdx.x & 3;
    const float2* src = spec + off + (long long)row * estride + q * 2;
    const unsigned sd = sdst + (row * CW + q * 2) * 8;
#pragma unroll
    for (int it = 0; it < (N + RPP - 1) / RPP; ++it) {
      if (row + it * RPP < N) cp_async16(sd + it * RPP * CW * 8, src);
      src += (long long)RPP * estride;
    }
  };
  auto issueK = [&](int comp, int kb) {             // N rows of 8 floats = 2 x 16 bytes each
    const float* src = kern + comp * kstride + kb;
#pragma unroll
    for (int it = 0; it < (2 * N + NT3 - 1) / NT3; ++it) {
      const int t = threadIdx.x + it * NT3;
      if (t < 2 * N) cp_async16(sK + t * 16, src + (long long)(t >> 1) * kes + (t & 1) * 4);
    }
    cp_async_commit();
  };
  auto loadA = [&](const float2* __restrict__ X, float2 (&v)[R0]) {       // x[j + r R1]
    const float2* p = X + j * CW + col;
#pragma unroll
    for (int r = 0; r < R0; ++r) v[r] = p[r * R1 * CW];
  };
  auto storeA = [&](float2* __restrict__ Y, const float2 (&v)[R0]) {      // y[j R0 + r]
    float2* q = Y + j * YP + col;
#pragma unroll
    for (int r = 0; r < R0; ++r) q[r * CW] = v[r];
  };
  // stage B on the Y layout: y[j + r R0] -> twiddle -> radix R1 -> emit(r, X[j + r R0])
  auto runB = [&](const float2* __restrict__ Y, auto inv_tag, auto emit) {
    constexpr bool INV = decltype(inv_tag)::value;
    float2 u[R1];
    const float2* p = Y + j * CW + col;
#pragma unroll
    for (int r = 0; r < R1; ++r) u[r] = p[r * YP];
    const float2* t0 = tw + j;
#pragma unroll
    for (int r = 1; r < R1; ++r) u[r] = twmul<INV>(u[r], t0[(r - 1) * j]);
    pradix_emit<R1, INV>(u, emit);
  };
  auto mulK = [&](float2 (&v)[R0]) {                 // i * kern_f (:188-189)
    const float* kq = K + j * CW + col;
#pragma unroll
    for (int r = 0; r < R0; ++r) { const float kv = kq[r * R1 * CW]; v[r] = make_float2(-v[r].y * kv, v[r].x * kv); }
  };
  int item = blockIdx.x, off = 0, kb = 0;
  bool ok = false;
  if (item < total) {
    decode(item, off, kb, ok);
    issueS(off, sL);
    issueK(0, kb);
  }
  const unsigned stepb = (unsigned)(R0 * estride) * 8u;
  char* gbase = reinterpret_cast<char*>(g);
  while (item < total) {
    const int next = item + gridDim.x;
    int noff = 0, nkb = 0;
    bool nok = false;
    if (next < total) decode(next, noff, nkb, nok);
    const unsigned m = ok ? mask : 0u;
    cp_async_wait_all();
    __syncthreads();                                // spectrum block and kern_f(0) block landed; W free
    float2 v[R0];
    // ---- forward: A: L(X) -> W(Y), B: W(Y) -> L(X)
    if (actA) { loadA(L, v); PRadix<R0, false>::run(v); storeA(W, v); }
    __syncthreads();
    if (actB) {
      float2* sp = L + j * CW + col;
      runB(W, std::false_type{}, [&](int r, float2 val) { sp[r * R0 * CW] = val; });
    }
    __syncthreads();
    // ---- components 0 and 1: A: L(X) -> W(Y), B: W(Y) -> global
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
      if (actA) { loadA(L, v); mulK(v); PRadix<R0, true>::run(v); storeA(W, v); }
      __syncthreads();                              // W complete; K consumed
      issueK(comp + 1, kb);
      if (actB) {
        const unsigned ob = (unsigned)(comp * gstride + off + joff) * 8u;
        runB(W, std::true_type{}, [&](int r, float2 val) { if (m & (1u << r)) *reinterpret_cast<float2*>(gbase + (ob + (unsigned)r * stepb)) = val; });
      }
      cp_async_wait_all();
      __syncthreads();                              // next kern_f block landed; W free
    }
    if (next < total) { issueS(noff, sW); cp_async_commit(); }
    // ---- component 2: A: L(X) -> registers -> L(Y), B: L(Y) -> global
    if (actA) { loadA(L, v); mulK(v); PRadix<R0, true>::run(v); }
    __syncthreads();                                // all gathers from L and K precede the scatter / the next kern_f block
    if (next < total) issueK(0, nkb);
    if (actA) storeA(L, v);
    __syncthreads();
    if (actB) {
      const unsigned ob = (unsigned)(2 * gstride + off + joff) * 8u;
      runB(L, std::true_type{}, [&](int r, float2 val) { if (m & (1u << r)) *reinterpret_cast<float2*>(gbase + (ob + (unsigned)r * stepb)) = val; });
    }
    { float2* t = L; L = W; W = t; const unsigned u = sL; sL = sW; sW = u; }
    item = next; off = noff; kb = nkb; ok = nok;
  }
}

// ---------------------------------------------------------------------------------------------- x pass forward + NGP density
constexpr int XP = LX + 1;   // pitch (float2) of the contiguous-axis passes

template <int N>
__global__ void __launch_bounds__(Plan2<N>::NT, 3) fft_x_r2c_ngp2(float2* __restrict__ data, int cp, const int* __restrict__ fstart, int H, int b, int ox, int oy, int oz,
                                                                  float mass_p, const int2* __restrict__ deltas, const int* __restrict__ ndelta_ptr,
                                                                  int delta_cap, double* __restrict__ sum_phys, const float2* __restrict__ tw_g) {
  using P = Plan2<N>;
  constexpr int NT2 = P::NT, NW = NT2 / 32, R0 = P::R0, NCX = N / 4, UNITS = 8 * NCX, UPT = (UNITS + NT2 - 1) / NT2;
  constexpr int HC = N / 2 + 1, KCH = (HC + 31) / 32;
  extern __shared__ __align__(16) unsigned char raw[];
  float2* Z = reinterpret_cast<float2*>(raw);
  float2* tw = Z + N * XP;
  for (int t = threadIdx.x; t < N; t += NT2) tw[t] = tw_g[t];
  const int r0 = blockIdx.x * (2 * LX);          // first row (= z*N + y) of this CTA; N*N is a multiple of 32
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- density: unit u = (cyl, ccx): rows r0+4*cyl .. +3 (one coarse-cell row group, fy = 0..3), x cells 4*ccx .. +3
  int4 q[UPT][4];
  int qe[UPT];
  bool live[UPT];
#pragma unroll
  for (int s = 0; s < UPT; ++s) {
    const int u = threadIdx.x + s * NT2;
    live[s] = false;
    if (u < UNITS) {
      const int cyl = u & 7, ccx = u >> 3;            // lanes run over the 8 row groups first: 64-bit stores of a half-warp are 2-way conflicted at worst
      const int gr = r0 + 4 * cyl, y0 = gr % N, z = gr / N;
      if (z >= 4 && z <= N - 5 && y0 >= 4 && y0 <= N - 8 && ccx >= 1 && ccx <= NCX - 2) {   // deposit range [4, N-5] on every axis (:120-121)
        const int gz = z + oz, gy = y0 + oy, gx = 4 * ccx + ox;
        const long long k = ((long long)((gz >> 2) * H + (gy >> 2)) * H + (gx >> 2)) * 64 + ((gz & 3) << 4);
        const int4* src = reinterpret_cast<const int4*>(fstart + k);
        q[s][0] = src[0]; q[s][1] = src[1]; q[s][2] = src[2]; q[s][3] = src[3];
        qe[s] = fstart[k + 16];
        live[s] = true;
      }
    }
  }
  double msum = 0.0;
#pragma unroll
  for (int s = 0; s < UPT; ++s) {
    const int u = threadIdx.x + s * NT2;
    if (u < UNITS) {
      const int cyl = u & 7, ccx = u >> 3;
      float val[16];
      if (live[s]) {
        const int e[17] = {q[s][0].x, q[s][0].y, q[s][0].z, q[s][0].w, q[s][1].x, q[s][1].y, q[s][1].z, q[s][1].w, q[s][2].x,
                           q[s][2].y, q[s][2].z, q[s][2].w, q[s][3].x, q[s][3].y, q[s][3].z, q[s][3].w, qe[s]};
#pragma unroll
        for (int i = 0; i < 16; ++i) val[i] = mass_p * (float)(e[i + 1] - e[i]);
        const int gr = r0 + 4 * cyl, y0 = gr % N, z = gr / N, x0 = 4 * ccx;
        if (z >= b && z < N - b && y0 >= b && y0 < N - b && x0 >= b && x0 < N - b)           // b % 4 == 0; the 16 values are small multiples of mass_p
          msum += (double)(mass_p * (float)(e[16] - e[0]));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) val[i] = 0.f;
      }
      // rows 4*cyl + 2h (real part) and 4*cyl + 2h + 1 (imaginary part) -> column 2*cyl + h
      float2* zq = Z + (4 * ccx) * XP + 2 * cyl;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int fx = 0; fx < 4; ++fx) zq[fx * XP + h] = make_float2(val[8 * h + fx], val[8 * h + 4 + fx]);
    }
  }
  msum = warp_sum_d(msum);
  if (lane == 0 && msum != 0.0) atomicAdd(sum_phys, msum);
  __syncthreads();
  const int nd = min(*ndelta_ptr, delta_cap);
  if (nd > 0) {
    for (int i = threadIdx.x; i < nd; i += NT2) {
      const int2 d = deltas[i];
      const int c[2] = {d.x, d.y};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int x = c[u] % N, gr = c[u] / N;     // gr = z*N + y
        const int row = gr - r0;
        if (row >= 0 && row < 2 * LX) atomicAdd(reinterpret_cast<float*>(Z + x * XP + (row >> 1)) + (row & 1), u == 0 ? -mass_p : mass_p);
      }
    }
    __syncthreads();
  }
  // ---- FFT of the 16 packed columns, in place
  const int col = threadIdx.x % LX, j = threadIdx.x / LX;
  {
    float2 v[R0];
    if (threadIdx.x < P::NA) { stageA_load<N, false, XP>(Z, j, col, v); PRadix<R0, false>::run(v); }
    __syncthreads();
    if (threadIdx.x < P::NA) stageA_store<N, XP>(Z, j, col, v);
    __syncthreads();
    if (threadIdx.x < P::NB) {
      float2* sp = Z + j * XP + col;
      stageB<N, false, XP>(Z, tw, j, col, [&](int r, float2 val) { sp[r * R0 * XP] = val; });
    }
    __syncthreads();
  }
  // ---- untangle the two real rows of each column: A = (Z[k] + conj Z[N-k]) / 2, B = (Z[k] - conj Z[N-k]) / (2i).
  // Work unit = (column, chunk of 32 k): 16*KCH units dealt round-robin to the warps; each unit stores 256 contiguous bytes to two rows.
#pragma unroll
  for (int s = 0; s < (LX * KCH + NW - 1) / NW; ++s) {
    const int unit = warp + s * NW;
    if (unit < LX * KCH) {
      const int c = unit / KCH, k = (unit - c * KCH) * 32 + lane;
      if (k < HC) {
        const int km = (k == 0) ? 0 : N - k;
        const float2 zk = Z[k * XP + c], zm = Z[km * XP + c];
        const float2 sm = padd(zk, zm), df = psub(zk, zm);
        float2* rowA = data + (long long)(r0 + 2 * c) * cp;     // complex rows of pitch cp
        rowA[k] = make_float2(0.5f * sm.x, 0.5f * df.y);
        rowA[cp + k] = make_float2(0.5f * sm.y, -0.5f * df.x);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- x pass backward, 3 components
// Row staging buffer: Rw[32][HC] float2, filled with cp.async from the cropped rows of component `comp`.
template <int N>
__global__ void __launch_bounds__(Plan2<N>::NT, 2) fft_x_c2r3_v2(const float2* __restrict__ in, int cp, float* __restrict__ out, int lo, int cnt, int in_bstride,
                                                                 int out_bstride, float scale, unsigned int* __restrict__ fmax_bits,
                                                                 const float2* __restrict__ tw_g) {
  using P = Plan2<N>;
  constexpr int NT2 = P::NT, NW = NT2 / 32, R0 = P::R0, HC = N / 2 + 1, KCH = (HC + 31) / 32;
  extern __shared__ __align__(16) unsigned char raw[];
  float2* Z = reinterpret_cast<float2*>(raw);
  float2* Rw = Z + N * XP;
  float2* tw = Rw + 2 * LX * HC;
  __shared__ int srow[2 * LX], drow[2 * LX];
  for (int t = threadIdx.x; t < N; t += NT2) tw[t] = tw_g[t];
  const int nrows = cnt * cnt;
  const int r0 = blockIdx.x * (2 * LX);
  if (threadIdx.x < 2 * LX) {
    const int ridx = r0 + threadIdx.x;
    int so = -1, dof = -1;
    if (ridx < nrows) {
      const int zc = ridx / cnt, yc = ridx - zc * cnt;
      so = ((zc + lo) * N + (yc + lo)) * cp;
      dof = (zc * cnt + yc) * cnt;
    }
    srow[threadIdx.x] = so; drow[threadIdx.x] = dof;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned sR = smem_u32(Rw);
  auto issue = [&](int comp) {                     // unit = (row, chunk of 32 k)
    const float2* src = in + (long long)comp * in_bstride + lane;
#pragma unroll
    for (int s = 0; s < (2 * LX * KCH + NW - 1) / NW; ++s) {
      const int unit = warp + s * NW;
      if (unit < 2 * LX * KCH) {
        const int row = unit / KCH, k = (unit - row * KCH) * 32 + lane;
        if (k < HC) {
          const int so = srow[row];
          if (so >= 0) cp_async8(sR + (row * HC + k) * 8, src + so + (k - lane));
          else Rw[row * HC + k] = make_float2(0.f, 0.f);
        }
      }
    }
    cp_async_commit();
  };
  issue(0);
  const int col = threadIdx.x % LX, j = threadIdx.x / LX;
  constexpr int XCH = (N + 31) / 32, SU = (LX * XCH + NW - 1) / NW;
  float2 fsq[SU];
#pragma unroll
  for (int s = 0; s < SU; ++s) fsq[s] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int comp = 0; comp < 3; ++comp) {
    cp_async_wait_all();
    __syncthreads();                               // rows of `comp` landed; Z is free (previous store phase done)
    // Z[k] = A[k] + i B[k];  Z[N-k] = conj(A[k]) + i conj(B[k]);  imaginary parts of the k=0 and k=N/2 bins are dropped (c2r)
#pragma unroll
    for (int s = 0; s < (LX * KCH + NW - 1) / NW; ++s) {
      const int unit = warp + s * NW;
      if (unit < LX * KCH) {
        const int c = unit / KCH, k = (unit - c * KCH) * 32 + lane;
        if (k < HC) {
          float2 a = Rw[(2 * c) * HC + k], bq = Rw[(2 * c + 1) * HC + k];
          const bool edge = (k == 0 || 2 * k == N);
          if (edge) { a.y = 0.f; bq.y = 0.f; }
          Z[k * XP + c] = add_irot(a, bq);                               // (ar - bi, ai + br)
          if (!edge) Z[(N - k) * XP + c] = make_float2(a.x + bq.y, bq.x - a.y);
        }
      }
    }
    __syncthreads();                               // Rw consumed
    if (comp < 2) issue(comp + 1);
    {
      float2 v[R0];
      if (threadIdx.x < P::NA) { stageA_load<N, true, XP>(Z, j, col, v); PRadix<R0, true>::run(v); }
      __syncthreads();
      if (threadIdx.x < P::NA) stageA_store<N, XP>(Z, j, col, v);
      __syncthreads();
      if (threadIdx.x < P::NB) {
        float2* sp = Z + j * XP + col;
        stageB<N, true, XP>(Z, tw, j, col, [&](int r, float2 val) { sp[r * R0 * XP] = val; });
      }
      __syncthreads();
    }
    // unit = (column, chunk of 32 x): a column holds two output rows (real part -> even row, imaginary part -> odd row); one 64-bit
    // conflict-free shared load and two coalesced 32-bit stores per element. |F|^2 of the thread's elements accumulates in fsq (the
    // unit -> thread map is the same for the three components).
    float* o = out + (long long)comp * out_bstride;
#pragma unroll
    for (int s = 0; s < SU; ++s) {
      const int unit = warp + s * NW;
      if (unit < LX * XCH) {
        const int c = unit / XCH, xc = (unit - c * XCH) * 32 + lane;
        const int dofA = drow[2 * c], dofB = drow[2 * c + 1];
        if (dofA >= 0 && xc < cnt) {
          const float2 zz = pmul_s(scale, Z[(xc + lo) * XP + c]);
          o[dofA + xc] = zz.x;
          if (dofB >= 0) o[dofB + xc] = zz.y;
          fsq[s] = pfma_v(zz, zz, fsq[s]);
        }
      }
    }
  }
  float mx = 0.f;
#pragma unroll
  for (int s = 0; s < SU; ++s) mx = fmaxf(mx, fmaxf(fsq[s].x, fsq[s].y));   // rows beyond the crop contribute 0 (their inputs were zero-filled)
  mx = warp_max(mx);
  if (lane == 0 && mx > 0.f) atomic_max_float_nonneg(fmax_bits, mx);
}
constexpr size_t smem_bytes_c2r3_v2(int n) { return (size_t)n * XP * sizeof(float2) + (size_t)2 * LX * (n / 2 + 1) * sizeof(float2) + (size_t)n * sizeof(float2); }

// ---------------------------------------------------------------------------------------------- x pass backward, 3 components, v3
// Persistent CTAs; the 32 half-spectrum rows of the next (row block, component) are fetched by the TMA engine as 32 bulk copies
// (cp.async.bulk, one per lane of warp 0, completion on an mbarrier) while the current one is transformed: no per-element staging
// instructions at all. Needs rows that start on 16-byte boundaries and a pitch >= RP (the padded spectra of the fused path).
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <int N> struct C2R3 {
  static constexpr int HC = N / 2 + 1;
  static constexpr int RP = (HC + 1) / 2 * 2;      // staged row length (float2): even, so that a row is a multiple of 16 bytes
  static constexpr size_t smem = (size_t)N * XP * sizeof(float2) + (size_t)2 * LX * RP * sizeof(float2) + (size_t)N * sizeof(float2) + 2 * 2 * LX * sizeof(int) + 16;
};

template <int N>
__global__ void __launch_bounds__(Plan2<N>::NT, 2) fft_x_c2r3_v3(const float2* __restrict__ in, int cp, float* __restrict__ out, int lo, int cnt, int in_bstride,
                                                                 int out_bstride, float scale, unsigned int* __restrict__ fmax_bits,
                                                                 const float2* __restrict__ tw_g) {
  using P = Plan2<N>;
  constexpr int NT2 = P::NT, NW = NT2 / 32, R0 = P::R0, HC = C2R3<N>::HC, RP = C2R3<N>::RP, KCH = (HC + 31) / 32;
  constexpr int XCH = (N + 31) / 32, SU = (LX * XCH + NW - 1) / NW;
  extern __shared__ __align__(16) unsigned char raw[];
  float2* Z = reinterpret_cast<float2*>(raw);
  float2* Rw = Z + N * XP;
  float2* tw = Rw + 2 * LX * RP;
  int* drow = reinterpret_cast<int*>(tw + N);        // [2][32]
  const unsigned bar = smem_u32(drow + 2 * 2 * LX);  // 8-byte aligned: every array before it is a multiple of 8 bytes
  for (int t = threadIdx.x; t < N; t += NT2) tw[t] = tw_g[t];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = threadIdx.x % LX, j = threadIdx.x / LX;
  const int nrows = cnt * cnt, nblk = (nrows + 2 * LX - 1) / (2 * LX);
  const unsigned sR = smem_u32(Rw);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  // warp 0: fetch the rows of (blk, comp); for comp 0 also publish the block's output offsets in drow[slot]
  auto issue = [&](int blk, int comp, int slot) {
    const int ridx = blk * (2 * LX) + lane;
    const bool valid = ridx < nrows;
    const int rr = valid ? ridx : nrows - 1;         // rows beyond the crop re-read the last row; their results are never stored
    const int zc = rr / cnt, yc = rr - zc * cnt;
    if (comp == 0) drow[slot * 2 * LX + lane] = valid ? (zc * cnt + yc) * cnt : -1;
    const float2* src = in + (long long)comp * in_bstride + ((zc + lo) * N + (yc + lo)) * cp;
    if (lane == 0) mbar_expect_tx(bar, 2 * LX * RP * 8);
    __syncwarp();
    bulk_g2s(sR + lane * RP * 8, src, RP * 8, bar);
  };
  int blk = blockIdx.x, slot = 0;
  unsigned parity = 0;
  if (warp == 0 && blk < nblk) issue(blk, 0, 0);
  for (; blk < nblk; blk += gridDim.x, slot ^= 1) {
    float2 fsq[SU];
#pragma unroll
    for (int s = 0; s < SU; ++s) fsq[s] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int comp = 0; comp < 3; ++comp) {
      mbar_wait(bar, parity);
      parity ^= 1;
      // Z[k] = A[k] + i B[k];  Z[N-k] = conj(A[k]) + i conj(B[k]);  imaginary parts of the k=0 and k=N/2 bins are dropped (c2r)
#pragma unroll
      for (int s = 0; s < (LX * KCH + NW - 1) / NW; ++s) {
        const int unit = warp + s * NW;
        if (unit < LX * KCH) {
          const int c = unit / KCH, k = (unit - c * KCH) * 32 + lane;
          if (k < HC) {
            float2 a = Rw[(2 * c) * RP + k], bq = Rw[(2 * c + 1) * RP + k];
            const bool edge = (k == 0 || 2 * k == N);
            if (edge) { a.y = 0.f; bq.y = 0.f; }
            Z[k * XP + c] = add_irot(a, bq);                               // (ar - bi, ai + br)
            if (!edge) Z[(N - k) * XP + c] = make_float2(a.x + bq.y, bq.x - a.y);
          }
        }
      }
      fence_proxy_async();                           // generic reads of Rw are ordered before the async-proxy writes of the next fetch
      __syncthreads();                               // Rw consumed, Z complete
      if (warp == 0) {
        if (comp < 2) issue(blk, comp + 1, slot);
        else if (blk + (int)gridDim.x < nblk) issue(blk + gridDim.x, 0, slot ^ 1);
      }
      {
        float2 v[R0];
        if (threadIdx.x < P::NA) { stageA_load<N, true, XP>(Z, j, col, v); PRadix<R0, true>::run(v); }
        __syncthreads();
        if (threadIdx.x < P::NA) stageA_store<N, XP>(Z, j, col, v);
        __syncthreads();
        if (threadIdx.x < P::NB) {
          float2* sp = Z + j * XP + col;
          stageB<N, true, XP>(Z, tw, j, col, [&](int r, float2 val) { sp[r * R0 * XP] = val; });
        }
        __syncthreads();
      }
      // unit = (column, chunk of 32 x): a column holds two output rows (real part -> even row, imaginary part -> odd row)
      float* o = out + (long long)comp * out_bstride;
      const int* dr = drow + slot * 2 * LX;
#pragma unroll
      for (int s = 0; s < SU; ++s) {
        const int unit = warp + s * NW;
        if (unit < LX * XCH) {
          const int c = unit / XCH, xc = (unit - c * XCH) * 32 + lane;
          const int dofA = dr[2 * c], dofB = dr[2 * c + 1];
          if (dofA >= 0 && xc < cnt) {
            float2 zz = pmul_s(scale, Z[(xc + lo) * XP + c]);
            o[dofA + xc] = zz.x;
            if (dofB >= 0) o[dofB + xc] = zz.y; else zz.y = 0.f;
            fsq[s] = pfma_v(zz, zz, fsq[s]);
          }
        }
      }
      __syncthreads();                               // Z is free for the next item
    }
    float mx = 0.f;
#pragma unroll
    for (int s = 0; s < SU; ++s) mx = fmaxf(mx, fmaxf(fsq[s].x, fsq[s].y));   // max |F|^2 (:208-223)
    mx = warp_max(mx);
    if (lane == 0 && mx > 0.f) atomic_max_float_nonneg(fmax_bits, mx);
  }
}

// ---------------------------------------------------------------------------------------------- x pass backward, 3 components, v4
// As v3 (persistent CTAs, rows fetched by the TMA engine as bulk copies signalled on an mbarrier), but with the lanes of a warp along the
// SEQUENCE instead of along the 16 columns (fft2.cuh, XRow): the Hermitian tangle happens in the registers of stage A, read straight from
// the staged rows, and stage B's results go from registers to global memory (each half-warp writes 64 contiguous bytes of one output row).
// Per item: 2 block barriers and ~5 N shared-memory accesses instead of 5 barriers and ~8 N (v3: tangle -> A -> B in place -> store pass).
// |F|^2 of a point needs its three components, which are three consecutive items: each stage-B thread accumulates them for its own R1 x 2
// outputs. Holding all of them in registers (38 for R1 = 19, next to the 38 of the radix-19 butterfly) spilled 264 bytes per thread and made
// every accumulation a local-memory round trip (ncu: long-scoreboard 37 %); only the first FSR stay in registers, the rest live in shared
// memory as fs[(r - FSR) * NB + tid] (each thread touches only its own entries: no synchronisation).
#ifndef FFTK_C2R4_FSR
#define FFTK_C2R4_FSR 6
#endif
// CW = complex columns (pairs of real rows) per item. 16 columns per 320-thread CTA left a 13 % tail at n = 304 (2097 row blocks over
// 296 persistent CTAs: 8 vs 7.08 blocks per CTA) and two CTAs per SM at the same phase boundaries; with 8 columns an item costs half, the CTAs are
// 160 threads and four fit an SM (same warps, same registers, 55 KB of shared memory each).
#ifndef FFTK_C2R4_CW
#define FFTK_C2R4_CW 8
#endif
template <int N> struct C2R4 {
  using X = XRow<N>;
  static constexpr int CW = FFTK_C2R4_CW;
  static constexpr int NA = CW * X::R1, NB = CW * 16;
  static constexpr int NT = ((NA > NB ? NA : NB) + 31) / 32 * 32;
  static constexpr int CPS = (N == 304 && CW == 8) ? 4 : 2;     // CTAs per SM the launch bounds ask for
  static constexpr int FSR = X::R1 < FFTK_C2R4_FSR ? X::R1 : FFTK_C2R4_FSR;
  static constexpr int FSS = X::R1 - FSR;            // accumulators per thread kept in shared memory
  static constexpr size_t smem = (size_t)2 * CW * X::RP * sizeof(float2) + (size_t)CW * X::YP * sizeof(float2) + (size_t)X::NTW * sizeof(float2) +
                                 (size_t)FSS * NB * sizeof(float2) + 2 * 2 * CW * sizeof(int) + 16;
};

template <int N>
__global__ void __launch_bounds__(C2R4<N>::NT, C2R4<N>::CPS) fft_x_c2r3_v4(const float2* __restrict__ in, int cp, float* __restrict__ out, int lo, int cnt, int in_bstride,
                                                                        int out_bstride, float scale, unsigned int* __restrict__ fmax_bits,
                                                                        const float2* __restrict__ tw_g) {
  using X = XRow<N>;
  using C = C2R4<N>;
  constexpr int NT2 = C::NT, R1 = X::R1, RP = X::RP, YP = X::YP, CW = C::CW, NR = 2 * CW;
  static_assert(NR <= 32, "one warp issues the row copies of an item");
  extern __shared__ __align__(16) unsigned char raw[];
  float2* Rw = reinterpret_cast<float2*>(raw);
  float2* Y = Rw + NR * RP;
  float2* twT = Y + CW * YP;
  constexpr int FSR = C::FSR, FSS = C::FSS;
  float2* fs = twT + X::NTW;                         // [FSS][NB]
  int* drow = reinterpret_cast<int*>(fs + FSS * C::NB);  // [2][NR]
  const unsigned bar = smem_u32(drow + 2 * NR);      // 8-byte aligned: every array before it is a multiple of 8 bytes
  for (int t = threadIdx.x; t < X::NTW; t += NT2) twT[t] = tw_g[((t >> 4) + 1) * (t & 15)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cA = threadIdx.x / R1, jA = threadIdx.x - cA * R1;     // stage A: lanes along k
  const int cB = threadIdx.x >> 4, jB = threadIdx.x & 15;          // stage B: lanes along x
  const bool actA = threadIdx.x < C::NA, actB = threadIdx.x < C::NB;
  const unsigned mask = crop_mask<N>(jB, lo, lo + cnt - 1);
  const int nrows = cnt * cnt, nblk = (nrows + NR - 1) / NR;
  const unsigned sR = smem_u32(Rw);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  // warp 0: fetch the rows of (blk, comp); for comp 0 also publish the block's output offsets in drow[slot]
  auto issue = [&](int blk, int comp, int slot) {
    if (lane == 0) mbar_expect_tx(bar, NR * RP * 8);
    __syncwarp();
    if (lane < NR) {
      const int ridx = blk * NR + lane;
      const bool valid = ridx < nrows;
      const int rr = valid ? ridx : nrows - 1;       // rows beyond the crop re-read the last row; their results are never stored
      const int zc = rr / cnt, yc = rr - zc * cnt;
      if (comp == 0) drow[slot * NR + lane] = valid ? (zc * cnt + yc) * cnt : -1;
      const float2* src = in + (long long)comp * in_bstride + ((zc + lo) * N + (yc + lo)) * cp;
      bulk_g2s(sR + lane * RP * 8, src, RP * 8, bar);
    }
  };
  int blk = blockIdx.x, slot = 0;
  unsigned parity = 0;
  if (warp == 0 && blk < nblk) issue(blk, 0, 0);
  const float2* rowA = Rw + (2 * cA) * RP;
  float2* ycolA = Y + cA * YP;
  const float2* ycolB = Y + cB * YP;
  float2* fsme = fs + threadIdx.x;
  for (; blk < nblk; blk += gridDim.x, slot ^= 1) {
    float2 fsq[FSR > 0 ? FSR : 1];
#pragma unroll
    for (int r = 0; r < FSR; ++r) fsq[r] = make_float2(0.f, 0.f);
    if (actB) {
#pragma unroll
      for (int r = 0; r < FSS; ++r) fsme[r * C::NB] = make_float2(0.f, 0.f);
    }
#pragma unroll 1
    for (int comp = 0; comp < 3; ++comp) {
      mbar_wait(bar, parity);
      parity ^= 1;
      if (actA) {
        float2 v[16];
        c2r_stageA_load<N>(rowA, rowA + RP, jA, v);
        PRadix<16, true>::run(v);
        c2r_stageA_store<N>(ycolA, jA, v);           // Y is free: every stage-B load of the previous item precedes its second barrier
      }
      fence_proxy_async();                           // generic reads of Rw are ordered before the async-proxy writes of the next fetch
      __syncthreads();                               // Rw consumed, Y complete
      if (warp == 0) {
        if (comp < 2) issue(blk, comp + 1, slot);
        else if (blk + (int)gridDim.x < nblk) issue(blk + gridDim.x, 0, slot ^ 1);
      }
      float2 u[R1];
      if (actB) c2r_stageB_load<N>(ycolB, jB, u);
      __syncthreads();                               // Y is free for the next item
      if (actB) {
        const int* dr = drow + slot * NR;
        const int dofA = dr[2 * cB], dofB = dr[2 * cB + 1];
        const unsigned m = dofA >= 0 ? mask : 0u;
        float* oA = out + (long long)comp * out_bstride + dofA + (jB - lo);
        float* oB = out + (long long)comp * out_bstride + dofB + (jB - lo);
        const bool hasB = dofB >= 0;
        c2r_stageB_finish<N>(u, twT, jB, [&](int r, float2 val) {
          if (m & (1u << r)) {
            float2 zz = pmul_s(scale, val);
            oA[16 * r] = zz.x;                       // real part -> even row, imaginary part -> odd row
            if (hasB) oB[16 * r] = zz.y; else zz.y = 0.f;
            if (r < FSR) fsq[r < FSR ? r : 0] = pfma_v(zz, zz, fsq[r < FSR ? r : 0]);
            else fsme[(r - FSR) * C::NB] = pfma_v(zz, zz, fsme[(r - FSR) * C::NB]);
          }
        });
      }
    }
    float mx = 0.f;
#pragma unroll
    for (int r = 0; r < FSR; ++r) mx = fmaxf(mx, fmaxf(fsq[r].x, fsq[r].y));   // max |F|^2 (:208-223)
    if (actB) {
#pragma unroll
      for (int r = 0; r < FSS; ++r) { const float2 q = fsme[r * C::NB]; mx = fmaxf(mx, fmaxf(q.x, q.y)); }
    }
    mx = warp_max(mx);
    if (lane == 0 && mx > 0.f) atomic_max_float_nonneg(fmax_bits, mx);
  }
}

}  // namespace fftk
