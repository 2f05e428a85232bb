// Short-range particle-particle forces on the cell-sorted particle array.
//   PPINT  (particle_mesh_threaded.f90:274-285, 324-361): all pairs inside one fine cell.
//   PP_EXT (particle_mesh_threaded.f90:378-624): all pairs between a fine cell and the cells within pp_range,
//           with the cut-off polynomial of :558-564.
// The reference walks unordered pairs and applies +-f to both partners; here every target particle gathers
// the force of all its partners (same sum, no write conflicts, deterministic given the array order).
#pragma once
#include "common.cuh"
#include "fine.cuh"

namespace pp {

constexpr int TPB = 128;   // 4 warps, one work item (fine cell) per warp

struct PPParams {
  float mass_p, rsoft, pp_bias, a_mid, G, dt, cutoff;
  int apply;             // pp_force_flag / pp_ext_force_flag
};

__device__ __forceinline__ void pair_force(const float3 pi, const float3 pj, const PPParams& P, bool ext, float3& acc) {
  const float sx = pi.x - pj.x, sy = pi.y - pj.y, sz = pi.z - pj.z;
  const float r = sqrtf(sx * sx + sy * sy + sz * sz);
  if (r > P.rsoft) {
    const float rb = r * P.pp_bias;
    float w = P.mass_p / (rb * rb * rb);
    if (ext && !(r > P.cutoff + 1.7320508f)) {
      const float u = rb / P.cutoff, u2 = u * u, u3 = u2 * u;
      w *= (1.0f - 1.75f * u3 + 0.75f * u3 * u2);
    }
    acc.x -= sx * w; acc.y -= sy * w; acc.z -= sz * w;   // pp_force_accum(:,ip) -= force_pp   (:346, :571)
  }
}

// one warp per fine cell with >= 2 particles
__global__ void __launch_bounds__(TPB) ppint_kernel(float* __restrict__ xv, const int* __restrict__ fstart, const int* __restrict__ list,
                                                    const int* __restrict__ n_list_ptr, int list_cap, PPParams P, int max_llf,
                                                    DevCounters* __restrict__ cnt, const int* __restrict__ n_items, int item_cap) {
  if (item_cap > 0 && n_items[0] + n_items[1] <= item_cap) return;    // the (cell, chunk) items were listed in full: ppint_cell_kernel does the work
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (TPB / 32);
  const int n_list = min(*n_list_ptr, list_cap);
  float fmax = 0.f;
  unsigned long long npair = 0;            // ordered pairs evaluated (bench.py: pairs/s and the FP32 fraction of this stage)
  for (int w = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); w < n_list; w += nwarps) {
    const int k = list[w];
    const int s0 = fstart[k], s1 = fstart[k + 1];
    if (s1 - s0 > max_llf) { if (lane == 0) atomicOr(&cnt->overflow, 4); continue; }   // 'exceeded max_llf' :280-283
    if (lane == 0) npair += (unsigned long long)(s1 - s0) * (unsigned long long)(s1 - s0 - 1);
    for (int ib = s0; ib < s1; ib += 32) {
      const int i = ib + lane;
      const bool vi = i < s1;
      float3 pi = make_float3(0.f, 0.f, 0.f);
      if (vi) { const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i; const float2 a = p[0]; pi = make_float3(a.x, a.y, p[1].x); }
      float3 acc = make_float3(0.f, 0.f, 0.f);
      for (int jb = s0; jb < s1; jb += 32) {
        const int j = jb + lane;
        float3 pj = make_float3(0.f, 0.f, 0.f);
        if (j < s1) { const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * j; const float2 a = p[0]; pj = make_float3(a.x, a.y, p[1].x); }
        const int nj = min(32, s1 - jb);
        for (int l = 0; l < nj; ++l) {
          const float3 q = make_float3(__shfl_sync(0xffffffffu, pj.x, l), __shfl_sync(0xffffffffu, pj.y, l), __shfl_sync(0xffffffffu, pj.z, l));
          if (vi && (jb + l) != i) pair_force(pi, q, P, false, acc);
        }
      }
      if (vi) {
        fmax = fmaxf(fmax, sqrtf(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z));   // :355-358
        if (P.apply) {
          float2* p = reinterpret_cast<float2*>(xv) + 3LL * i;
          const float s = P.a_mid * P.G * P.dt;
          float2 b = p[1], c = p[2];
          b.y += acc.x * s; c.x += acc.y * s; c.y += acc.z * s;
          p[1] = b; p[2] = c;
        }
      }
      __syncwarp();
    }
  }
  fmax = warp_max(fmax);
  if (lane == 0 && fmax > 0.f) atomic_max_float_nonneg(&cnt->pp_force_max_bits, fmax);
  if (lane == 0 && npair) atomicAdd(&cnt->pairs_ppint, npair);
}

// PP_EXT, one THREAD per target particle of the cell-sorted array (ghosts and non-physical cells are skipped). At the mean density
// of 1/8 particle per fine cell a warp-per-cell mapping leaves 31 of 32 lanes idle; here every lane owns a target and walks the
// (2*pr+1)^2 neighbour fine-cell rows. The cells gx-pr..gx+pr of one row lie in at most two coarse cells, and inside one coarse
// cell consecutive fx are consecutive keys, so each row is at most two contiguous particle ranges found with two fstart reads each.
// Neighbouring lanes hold neighbouring particles, so their range lookups and source loads hit the same L1 lines.
// The same-cell pairs belong to PPINT (:496-523 excludes the cell itself): the centre row is walked as [gx-pr,gx-1] and [gx+1,gx+pr].
// one contiguous range of sources
__device__ __forceinline__ void ppext_sources(const float* __restrict__ xv, int s, int e, const float3 pi, const PPParams& P, float3& acc, unsigned* npair = nullptr) {
  if (npair) *npair += (unsigned)max(e - s, 0);
#pragma unroll 1
  for (int j = s; j < e; ++j) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * j;
    const float2 a = p[0];
    pair_force(pi, make_float3(a.x, a.y, p[1].x), P, true, acc);
  }
}

constexpr int EXT_TPB = 128;
constexpr int EXT_MAXR = 2;    // largest pp_range (cubepm.par:92 uses 2)

// force on the target particle i (a physical-cell particle at position pi, fine cell (gx,gy,gz) of the hoc range) from every particle of the
// cells within pr, looked up in the global fine-cell table
__device__ __forceinline__ float3 ppext_direct(const float* __restrict__ xv, const int* __restrict__ fstart, int H, int pr, const float3 pi, int gx, int gy,
                                               int gz, const PPParams& P, unsigned* npair) {
  float3 acc = make_float3(0.f, 0.f, 0.f);
  // x cells gx-pr..gx+pr lie in coarse cells ca (fine cells fa0..fa1) and, if the window straddles a coarse boundary, cb (0..fb1)
  const int xa = gx - pr, xb = gx + pr;
  const int ca = xa >> 2, cb = xb >> 2;
  const bool two = cb != ca;
  const int fa0 = xa & 3, fa1 = two ? 3 : (xb & 3), fb1 = xb & 3;
#pragma unroll 1
  for (int dz = -pr; dz <= pr; ++dz) {
    const int nz = gz + dz;
    // all table look-ups of this z plane first (up to 4 per row, independent loads), then the pair loops
    int rs[2 * EXT_MAXR + 1][2], re[2 * EXT_MAXR + 1][2];
#pragma unroll
    for (int q = 0; q < 2 * EXT_MAXR + 1; ++q) {
      const int dy = q - EXT_MAXR;
      rs[q][0] = re[q][0] = rs[q][1] = re[q][1] = 0;
      if (dy >= -pr && dy <= pr) {
        const int ny = gy + dy;
        const long long rowkey = ((long long)((nz >> 2) * H + (ny >> 2)) * H) * 64 + (((nz & 3) << 4) | ((ny & 3) << 2));
        const long long ka = rowkey + (long long)ca * 64;
        rs[q][0] = fstart[ka + fa0]; re[q][0] = fstart[ka + fa1 + 1];
        if (two) { const long long kb = rowkey + (long long)cb * 64; rs[q][1] = fstart[kb]; re[q][1] = fstart[kb + fb1 + 1]; }
      }
    }
#pragma unroll
    for (int q = 0; q < 2 * EXT_MAXR + 1; ++q) {
      if (dz == 0 && q == EXT_MAXR) continue;       // the centre row is handled below (own cell excluded)
      ppext_sources(xv, rs[q][0], re[q][0], pi, P, acc, npair);
      ppext_sources(xv, rs[q][1], re[q][1], pi, P, acc, npair);
    }
  }
  {   // centre row: cells [gx-pr, gx-1] and [gx+1, gx+pr]; the pairs inside the own cell belong to PPINT (:496-523)
    const long long rowkey = ((long long)((gz >> 2) * H + (gy >> 2)) * H) * 64 + (((gz & 3) << 4) | ((gy & 3) << 2));
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const int x0 = side ? gx + 1 : gx - pr, x1 = side ? gx + pr : gx - 1;
      if (x0 > x1) continue;
      const int c0 = x0 >> 2, c1 = x1 >> 2;
      for (int cc = c0; cc <= c1; ++cc) {
        const int f0 = (cc == c0) ? (x0 & 3) : 0, f1 = (cc == c1) ? (x1 & 3) : 3;
        const long long k0 = rowkey + (long long)cc * 64;
        ppext_sources(xv, fstart[k0 + f0], fstart[k0 + f1 + 1], pi, P, acc, npair);
      }
    }
  }
  return acc;
}

// kick of pp_ext_force_accum (:576-590) on the record at p; returns |F| for pp_ext_force_max (:617)
__device__ __forceinline__ float ppext_apply(float2* __restrict__ p, const float3 acc, const PPParams& P) {
  if (P.apply) {
    const float s = P.a_mid * P.G * P.dt;
    float2 bq = p[1], c = p[2];
    bq.y += acc.x * s; c.x += acc.y * s; c.y += acc.z * s;
    p[1] = bq; p[2] = c;
  }
  return sqrtf(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z);
}

__global__ void __launch_bounds__(EXT_TPB) ppext_kernel(float* __restrict__ xv, const int* __restrict__ fstart, int np_all, int H, int b, int nc_buf,
                                                        int nc_node, int pr, PPParams P, DevCounters* __restrict__ cnt) {
  const int i = blockIdx.x * EXT_TPB + threadIdx.x;
  float fm = 0.f;
  unsigned npair = 0;
  if (i < np_all) {
    float2* p = reinterpret_cast<float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    const float z = p[1].x;
    const int gx = (int)floorf(a.x) + b, gy = (int)floorf(a.y) + b, gz = (int)floorf(z) + b;   // as part::make_key
    const int lo = nc_buf * 4, hi = (nc_buf + nc_node) * 4;
    if (gx >= lo && gx < hi && gy >= lo && gy < hi && gz >= lo && gz < hi)                        // kick only particles of the physical cells (:576-590)
      fm = ppext_apply(p, ppext_direct(xv, fstart, H, pr, make_float3(a.x, a.y, z), gx, gy, gz, P, &npair), P);
  }
  fm = warp_max(fm);
  if ((threadIdx.x & 31) == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fm);
  const int np_w = warp_sum_i((int)npair);
  if ((threadIdx.x & 31) == 0 && np_w) atomicAdd(&cnt->pairs_ppext, (unsigned long long)np_w);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// PP_EXT, tiled: one CTA per block of TB_X x TB_Y x TB_Z coarse cells (the targets) and the TB_HALO fine cells around it (the sources).
// The direct kernel above spends its time on table look-ups (up to 100 global fstart reads per target: the 4^3 fine cells of a coarse
// cell are stored (fz,fy,fx)-major, so a 5-cell x window of one (y,z) row is split over two coarse cells) and on divergence. Here the
// block's source particles are RE-SORTED IN SHARED MEMORY by the dense region cell index (z,y,x) with x running over the whole region
// width: every neighbour row of a target is then exactly ONE contiguous range between two adjacent-table entries, both in shared memory.
//   1. count : the (TB_Y+2) x (TB_Z+2) coarse x-rows that cover the region are contiguous ranges of the sorted particle array
//              (key = (cz*H+cy)*H+cx major): one warp per row, lanes along the range, shared-memory histogram over the region's fine cells
//   2. scan  : exclusive scan of the histogram (one chunk per thread, warp shuffles)
//   3. fill  : second read of the same ranges (L1/L2 hits), positions + global index scattered to their sorted place as float4
//   4. list  : the targets (particles of the block's own physical cells) are one contiguous range of the re-sorted array per interior
//              (z,y) row: a 64-entry row table built by one warp replaces a per-particle list
//   5. walk  : one thread per target; (2 pr + 1)^2 rows, two shared-memory table reads per row, pair loop over float4 sources
// A block whose region holds more than TB_CAP particles (density contrast > ~3 over the region) is appended to an overflow list and walked
// by ppext_blocklist_kernel through the global table: there the per-cell ranges are long, neighbouring lanes share them, and the direct
// walk is efficient. Keeping that path out of this kernel also keeps its register count (and so its occupancy) low.
// The pair weight uses the fast reciprocal (MUFU.RCP, 2 ulp) instead of two IEEE divisions: |error| ~ 3e-7 relative, far inside the 1e-4 gate.
constexpr int TB_X = 4, TB_Y = 2, TB_Z = 2, TB_HALO = EXT_MAXR;
constexpr int TB_RX = 4 * TB_X + 2 * TB_HALO, TB_RY = 4 * TB_Y + 2 * TB_HALO, TB_RZ = 4 * TB_Z + 2 * TB_HALO;   // 20 x 12 x 12 fine cells
constexpr int TB_NCELL = TB_RX * TB_RY * TB_RZ;
constexpr int TB_NT = 128;
constexpr int TB_CAP = 1024;                    // source particles per block held in shared memory (mean: 360 at 1/8 particle per fine cell);
                                                // 30 KB per CTA -> 7 CTAs (28 warps) per SM: the walk is latency-bound, occupancy is what it needs
constexpr int TB_NROW = (TB_Y + 2) * (TB_Z + 2);   // coarse x-rows read per block
constexpr int TB_CHUNK = (TB_NCELL + TB_NT - 1) / TB_NT;
constexpr int TB_TAB = (TB_NCELL + 1 + 3) / 4 * 4;   // table entries, padded for 16-byte zeroing
constexpr int TB_NTROW = 16 * TB_Y * TB_Z;           // interior (z,y) fine rows of a block
static_assert(TB_NTROW == 64 && TB_NROW == 16 && TB_NT == 128, "the row tables assume 64 target rows (2 per lane of one warp) and 4 warps");
constexpr size_t TB_SMEM = (size_t)TB_CAP * sizeof(float4) + (size_t)TB_TAB * sizeof(int);

__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// shared-memory loads through 32-bit window addresses computed once (the generic-pointer form made the compiler rebuild the
// window base from SR_CgaCtaId in front of every access inside the walk: 4-5 issue slots per load)
__device__ __forceinline__ int lds_i32(unsigned a) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 lds_f4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
// The PP_EXT pair weight (particle_mesh_threaded.f90:558-564), algebraically folded. With rb = r pp_bias and u = rb / nf_cutoff the reference's
//   w = mass_p / rb^3 * (1 - 7/4 u^3 + 3/4 u^5)        (always this branch: r <= 3 sqrt(3) < nf_cutoff + sqrt(3) at pp_range <= 2)
// is  w = c1 / r^3 + c4 r^2 - c5  with c1 = mass_p / pp_bias^3, c5 = 7/4 mass_p / nf_cutoff^3, c4 = 3/4 mass_p pp_bias^2 / nf_cutoff^5
// (mass_p / rb^3 * u^3 is the constant mass_p / nf_cutoff^3): one MUFU.RSQ, 2 FMUL, 2 FFMA instead of a square root, two reciprocals and the
// polynomial — 16 instead of 28 instructions per pair, in kernels that are bound by instruction issue. The soft-core test r > rsoft is r^2 > rsoft^2.
// No cancellation anywhere in the range (c1 / r^3 >= 20 c5 at r = 5.2). Differences to the reference's evaluation order are ~1e-6 relative.
struct PairConst { float c1, c4, c5, rs2; };
__device__ __forceinline__ PairConst make_pair_const(const PPParams& P) {
  PairConst K;
  const float ic = 1.0f / P.cutoff, ib = 1.0f / P.pp_bias;
  K.c1 = P.mass_p * ib * ib * ib;
  K.c5 = 1.75f * P.mass_p * ic * ic * ic;
  K.c4 = 0.75f * P.mass_p * P.pp_bias * P.pp_bias * ic * ic * ic * ic * ic;
  K.rs2 = P.rsoft * P.rsoft;
  return K;
}
__device__ __forceinline__ void pair_force_fast(const float3 pi, const float4 pj, const PairConst& K, float3& acc) {
  const float sx = pi.x - pj.x, sy = pi.y - pj.y, sz = pi.z - pj.z;
  const float r2 = sx * sx + sy * sy + sz * sz;
  const float ir = rsqrt_ftz(r2);                    // r2 = 0 gives +inf, which the select below discards
  const float ir3 = ir * ir * ir;
  float w = fmaf(K.c1, ir3, fmaf(K.c4, r2, -K.c5));
  w = r2 > K.rs2 ? w : 0.0f;
  acc.x = fmaf(-sx, w, acc.x); acc.y = fmaf(-sy, w, acc.y); acc.z = fmaf(-sz, w, acc.z);
}

__device__ __forceinline__ void ppext_sources_fast(const float* __restrict__ xv, int s, int e, const float3 pi, const PairConst& K, float3& acc) {
#pragma unroll 1
  for (int j = s; j < e; ++j) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * j;
    const float2 a = p[0];
    pair_force_fast(pi, make_float4(a.x, a.y, p[1].x, 0.f), K, acc);
  }
}

// ---- pp_ext_force_max of the MARGIN particles (particle_mesh_threaded.f90:617).
// The reference evaluates PP_EXT per tile over the tile's physical fine cells plus pp_range cells around them (:397-402) and takes the
// maximum of |pp_ext_force_accum| over EVERY particle of that region (:617), not only over the kicked (physical) ones. A particle in the
// margin of a tile has, in that tile, only the partial sum over the partner cells that lie inside the tile's region — and on a near-uniform
// particle load such a one-sided sum is larger than any complete sum, so it is what sets dt_pp_ext_acc (:692). The kicks never see these
// sums; they are recomputed here for the limiter only. A (particle, tile) pair with the particle in the tile's margin is a ROLE (up to 7 per
// particle, ~5 % of the particles have one): ppext_margin_list_kernel compacts the roles (one thread per particle, warp-aggregated append),
// ppext_margin_roles_kernel evaluates one role per thread with all lanes busy (a thread-per-particle version left 95 % of the lanes idle
// next to a lane walking 25 neighbour rows: 8.7 ms at 512^3 particles). Cell pairs whose two cells both lie in the tile's upper z margin
// are never visited by the reference's half stencil ("we never loop towards smaller z", k = 1..nf_physical_tile_dim+pp_range at :496).
using part::MarginGeom;
using part::margin_tiles;
using part::margin_roles;

// |partial sum| of the particle at pi (hoc-frame fine cell g) over the partner cells inside the region of tile t3
__device__ __forceinline__ float margin_role_sum(const float* __restrict__ xv, const int* __restrict__ fstart, const float3 pi, const int g[3], const int t3[3],
                                                 const MarginGeom& G, const PairConst& KC) {
  const int pr = G.pr, H = G.H;
  int lo[3], hi[3];                                   // the tile's region in the hoc-range frame (inclusive)
  for (int ax = 0; ax < 3; ++ax) { lo[ax] = t3[ax] * G.m - pr + G.b; hi[ax] = (t3[ax] + 1) * G.m + pr - 1 + G.b; }
  const int ztop = (t3[2] + 1) * G.m + G.b;           // first cell of the upper z margin
  float3 acc = make_float3(0.f, 0.f, 0.f);
  const int x0 = max(g[0] - pr, lo[0]), x1 = min(g[0] + pr, hi[0]);
#pragma unroll 1
  for (int nz = max(g[2] - pr, lo[2]); nz <= min(g[2] + pr, hi[2]); ++nz) {
    if (g[2] >= ztop && nz >= ztop) continue;
#pragma unroll 1
    for (int ny = max(g[1] - pr, lo[1]); ny <= min(g[1] + pr, hi[1]); ++ny) {
      const int rowkey = (((nz >> 2) * H + (ny >> 2)) * H) * 64 + (((nz & 3) << 4) | ((ny & 3) << 2));
      const bool centre = nz == g[2] && ny == g[1];
      for (int side = 0; side < (centre ? 2 : 1); ++side) {
        const int xa = centre ? (side ? g[0] + 1 : x0) : x0, xb = centre ? (side ? x1 : g[0] - 1) : x1;
        if (xa > xb) continue;
        const int ca = xa >> 2, cb = xb >> 2;
        // at most two coarse cells (the window is <= 5 cells wide): both ranges looked up before either is walked
        const int ka = rowkey + ca * 64, kb = rowkey + cb * 64;
        const int sa = fstart[ka + (xa & 3)], ea = fstart[ka + (ca == cb ? (xb & 3) : 3) + 1];
        int sb = 0, eb = 0;
        if (cb != ca) { sb = fstart[kb]; eb = fstart[kb + (xb & 3) + 1]; }
        ppext_sources_fast(xv, sa, ea, pi, KC, acc);
        ppext_sources_fast(xv, sb, eb, pi, KC, acc);
      }
    }
  }
  return sqrtf(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z);
}

// (A one-thread-per-face-adjacent-coarse-cell listing was tried: no gain on the lattice, 3.6x slower on a clustered box, where one thread walks a halo core.)
__global__ void __launch_bounds__(EXT_TPB) ppext_margin_list_kernel(const float* __restrict__ xv, int np_all, MarginGeom G, int2* __restrict__ roles, int cap,
                                                                    int* __restrict__ n_roles) {
  const int i = blockIdx.x * EXT_TPB + threadIdx.x;
  unsigned mask = 0;
  int tl[3] = {0, 0, 0};
  if (i < np_all) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    const int g[3] = {(int)floorf(a.x) + G.b, (int)floorf(a.y) + G.b, (int)floorf(p[1].x) + G.b};     // fine cell in the hoc-range frame (part::make_key)
    mask = margin_roles(g, G, tl);
  }
  const int n = __popc(mask), lane = threadIdx.x & 31;
  int inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
  const int total = __shfl_sync(0xffffffffu, inc, 31);
  if (total == 0) return;
  int base = 0;
  if (lane == 31) base = atomicAdd(n_roles, total);
  base = __shfl_sync(0xffffffffu, base, 31) + inc - n;
  while (mask) {
    const int bit = __ffs(mask) - 1;
    mask &= mask - 1;
    if (base < cap) roles[base] = make_int2(i, ((tl[2] + (bit >> 2)) * G.T + (tl[1] + ((bit >> 1) & 1))) * G.T + (tl[0] + (bit & 1)));
    ++base;
  }
}
// one role per thread (persistent grid). If the list overflowed (n_roles > cap: only with extreme clustering on the tile faces) the list is
// ignored and every particle evaluates its own roles, which is slow but complete.
__global__ void __launch_bounds__(EXT_TPB) ppext_margin_roles_kernel(const float* __restrict__ xv, const int* __restrict__ fstart, int np_all, MarginGeom G,
                                                                     const int2* __restrict__ roles, int cap, const int* __restrict__ n_roles, PPParams P,
                                                                     DevCounters* __restrict__ cnt) {
  const int n = *n_roles;
  const PairConst KC = make_pair_const(P);
  float fm = 0.f;
  if (n <= cap) {
    for (int r = blockIdx.x * EXT_TPB + threadIdx.x; r < n; r += gridDim.x * EXT_TPB) {
      const int2 role = roles[r];
      const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * role.x;
      const float2 a = p[0];
      const float z = p[1].x;
      const int g[3] = {(int)floorf(a.x) + G.b, (int)floorf(a.y) + G.b, (int)floorf(z) + G.b};
      const int t3[3] = {role.y % G.T, (role.y / G.T) % G.T, role.y / (G.T * G.T)};
      fm = fmaxf(fm, margin_role_sum(xv, fstart, make_float3(a.x, a.y, z), g, t3, G, KC));
    }
  } else {
    for (int i = blockIdx.x * EXT_TPB + threadIdx.x; i < np_all; i += gridDim.x * EXT_TPB) {
      const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i;
      const float2 a = p[0];
      const float z = p[1].x;
      const int g[3] = {(int)floorf(a.x) + G.b, (int)floorf(a.y) + G.b, (int)floorf(z) + G.b};
      int tl[3];
      unsigned mask = margin_roles(g, G, tl);
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        const int t3[3] = {tl[0] + (bit & 1), tl[1] + ((bit >> 1) & 1), tl[2] + (bit >> 2)};
        fm = fmaxf(fm, margin_role_sum(xv, fstart, make_float3(a.x, a.y, z), g, t3, G, KC));
      }
    }
  }
  fm = warp_max(fm);
  if ((threadIdx.x & 31) == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fm);
}

// PRT: compile-time pp_range (2 = cubepm.par:92, every loop of the walk unrolls and the row decode folds to constants) or -1 = run-time pr_rt
template <int PRT>
__global__ void __launch_bounds__(TB_NT) ppext_tiled_kernel(float* __restrict__ xv, const int* __restrict__ fstart, int H, int b, int nc_buf, int nc_node, int nbx,
                                                            int nby, int pr_rt, PPParams P, DevCounters* __restrict__ cnt, int* __restrict__ n_fallback, int* __restrict__ ovf_list) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* src = reinterpret_cast<float4*>(raw);
  int* tab = reinterpret_cast<int*>(src + TB_CAP);                       // [TB_NCELL + 1] (padded to TB_TAB): counts -> starts
  __shared__ int wsum[TB_NT / 32], trow_pre[TB_NTROW + 1], trow_s[TB_NTROW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pr = PRT >= 0 ? PRT : pr_rt;
  unsigned a_src = (unsigned)__cvta_generic_to_shared(src), a_tab = (unsigned)__cvta_generic_to_shared(tab);
  asm volatile("mov.u32 %0, %0;" : "+r"(a_src));   // opaque copies: otherwise the window base (SR_CgaCtaId arithmetic, 4 issue slots) is
  asm volatile("mov.u32 %0, %0;" : "+r"(a_tab));   // rematerialised in front of every shared-memory access of the walk
  const int bx = blockIdx.x % nbx, by = (blockIdx.x / nbx) % nby, bz = blockIdx.x / (nbx * nby);
  const int cx0 = nc_buf + bx * TB_X, cy0 = nc_buf + by * TB_Y, cz0 = nc_buf + bz * TB_Z;      // first coarse cell of the block (hoc-range coordinates)
  const int phys_hi = nc_buf + nc_node;                                                        // first non-physical coarse cell
  const int ox = 4 * cx0 - TB_HALO, oy = 4 * cy0 - TB_HALO, oz = 4 * cz0 - TB_HALO;            // fine cell (0,0,0) of the region
  for (int t = tid; t < TB_TAB / 4; t += TB_NT) reinterpret_cast<int4*>(tab)[t] = make_int4(0, 0, 0, 0);
  __syncthreads();
  // The region is covered by TB_NROW coarse x-rows, each ONE contiguous range of the sorted array. Warp w owns rows w, w + 4, w + 8, w + 12:
  // their four ranges are looked up first (8 independent loads), then walked as one flattened range with the lanes along it, two records per
  // iteration with both loads issued before either is used. visit(cell, x, y, z, global index) is called for every particle inside the region.
  int rd[4], ro[5];                                   // rd[q] = g0[q] - ro[q]: global index = flattened index + rd[row]
  {
    int g0[4], len[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = warp + (TB_NT / 32) * q;
      const int cy = cy0 - 1 + r % (TB_Y + 2), cz = cz0 - 1 + r / (TB_Y + 2);
      const bool ok = cy <= H - 1 && cz <= H - 1;
      const long long rk = (long long)(min(cz, H - 1) * H + min(cy, H - 1)) * H;
      g0[q] = fstart[(rk + cx0 - 1) * 64];
      len[q] = ok ? fstart[(rk + min(cx0 + TB_X, H - 1)) * 64 + 64] - g0[q] : 0;
    }
    ro[0] = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { ro[q + 1] = ro[q] + len[q]; rd[q] = g0[q] - ro[q]; }
  }
  const float2* xv2 = reinterpret_cast<const float2*>(xv);
  auto for_region = [&](auto visit) {
    const int n = ro[4];
    auto gidx = [&](int f) { return f + (f < ro[1] ? rd[0] : f < ro[2] ? rd[1] : f < ro[3] ? rd[2] : rd[3]); };
    auto one = [&](float2 a, float z, int gi) {
      const int lx = (int)floorf(a.x) + b - ox, ly = (int)floorf(a.y) + b - oy, lz = (int)floorf(z) + b - oz;
      if ((unsigned)lx < (unsigned)TB_RX && (unsigned)ly < (unsigned)TB_RY && (unsigned)lz < (unsigned)TB_RZ)
        visit((lz * TB_RY + ly) * TB_RX + lx, a.x, a.y, z, gi);
    };
#pragma unroll 1
    for (int f = lane; f < n; f += 64) {
      const bool two = f + 32 < n;
      const int ga = gidx(f), gb = two ? gidx(f + 32) : ga;
      const float2* pa = xv2 + 3LL * ga;
      const float2* pb = xv2 + 3LL * gb;
      const float2 a = pa[0], c = pb[0];
      const float za = pa[1].x, zc = pb[1].x;
      one(a, za, ga);
      if (two) one(c, zc, gb);
    }
  };
  // ---- 1. count
  for_region([&](int c, float, float, float, int) { atomicAdd(&tab[c + 1], 1); });
  __syncthreads();
  // ---- 2. exclusive scan of tab[1..NCELL] in place: tab[c + 1] = start of cell c (tab[0] = 0 = start of cell 0 after the fill)
  {
    const int c0 = 1 + tid * TB_CHUNK, c1 = min(c0 + TB_CHUNK, TB_NCELL + 1);
    int s = 0;
    for (int c = c0; c < c1; ++c) s += tab[c];
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int base = inc - s;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    for (int c = c0; c < c1; ++c) { const int v = tab[c]; tab[c] = base; base += v; }
  }
  int total = 0;
#pragma unroll
  for (int w = 0; w < TB_NT / 32; ++w) total += wsum[w];
  __syncthreads();
  float fm = 0.f;
  int npair = 0;                                     // ordered pairs evaluated by this thread (bench.py)
  const int plo = 4 * nc_buf, phi = 4 * phys_hi;
  if (total > TB_CAP) {
    if (tid == 0) ovf_list[atomicAdd(n_fallback, 1)] = blockIdx.x;     // walked by ppext_blocklist_kernel
  } else {
    // ---- 3. fill
    for_region([&](int c, float x, float y, float z, int gi) { src[atomicAdd(&tab[c + 1], 1)] = make_float4(x, y, z, __int_as_float(gi)); });
    __syncthreads();                                   // now tab[c] = start of cell c, c = 0..NCELL
    // ---- 4. targets = the particles of the block's own physical cells: in every interior (z,y) row they are ONE contiguous range of the
    //         re-sorted array, so the target list is a 64-entry table (first target, targets before the row), built by warp 0
    if (warp == 0) {
      const int xl = max(TB_HALO, plo - ox), xh = min(TB_HALO + 4 * TB_X, phi - ox);
      int cnt2[2], st2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = 2 * lane + h, ly = TB_HALO + row % (4 * TB_Y), lz = TB_HALO + row / (4 * TB_Y);
        const bool live = xh > xl && ly + oy >= plo && ly + oy < phi && lz + oz >= plo && lz + oz < phi;
        const int rb = (lz * TB_RY + ly) * TB_RX;
        st2[h] = live ? tab[rb + xl] : 0;
        cnt2[h] = live ? tab[rb + xh] - st2[h] : 0;
      }
      const int s2 = cnt2[0] + cnt2[1];
      int inc = s2;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
      trow_pre[2 * lane] = inc - s2; trow_pre[2 * lane + 1] = inc - s2 + cnt2[0];
      trow_s[2 * lane] = st2[0]; trow_s[2 * lane + 1] = st2[1];
      if (lane == 31) trow_pre[TB_NTROW] = inc;
    }
    __syncthreads();
    // ---- 5. walk
    const int nt = trow_pre[TB_NTROW];
    const PairConst KC = make_pair_const(P);
    for (int t = tid; t < nt; t += TB_NT) {
      int r = 0;
#pragma unroll
      for (int h = TB_NTROW / 2; h >= 1; h >>= 1) if (trow_pre[r + h] <= t) r += h;       // row of target t (largest r with pre[r] <= t)
      const float4 me = src[trow_s[r] + (t - trow_pre[r])];
      const float3 pi = make_float3(me.x, me.y, me.z);
      const int lx = (int)floorf(me.x) + b - ox, ly = (int)floorf(me.y) + b - oy, lz = (int)floorf(me.z) + b - oz;
      const int own = (lz * TB_RY + ly) * TB_RX + lx;
      const unsigned a_own = a_tab + 4u * (unsigned)own;          // address of tab[own]
      const int own_s = lds_i32(a_own), own_e = lds_i32(a_own + 4);   // the own cell's pairs belong to PPINT (:496-523)
      float3 acc = make_float3(0.f, 0.f, 0.f);
      // pass 1 (uniform over the rows, no divergence): bit q of `rows` = neighbour row q = (dz+pr)*w + (dy+pr) holds a source outside the own cell
      const int w = 2 * pr + 1, qc = (w * w - 1) >> 1;
      unsigned rows = 0;
      if constexpr (PRT >= 0) {
#pragma unroll
        for (int qz = 0; qz < 2 * PRT + 1; ++qz)
#pragma unroll
          for (int qy = 0; qy < 2 * PRT + 1; ++qy) {
            constexpr int W = 2 * PRT + 1;
            const int q = qz * W + qy, off = 4 * (((qz - PRT) * TB_RY + (qy - PRT)) * TB_RX);
            const int n = lds_i32(a_own + off + 4 * (PRT + 1)) - lds_i32(a_own + off - 4 * PRT) - (q == (W * W - 1) / 2 ? own_e - own_s : 0);
            npair += n;
            rows |= (n > 0 ? 1u : 0u) << q;
          }
      } else {
        int rb = own - pr * (TB_RY + 1) * TB_RX;
        for (int qz = 0, q = 0; qz < w; ++qz, rb += (TB_RY - w) * TB_RX)
          for (int qy = 0; qy < w; ++qy, ++q, rb += TB_RX) {
            const int n = tab[rb + pr + 1] - tab[rb - pr] - (q == qc ? own_e - own_s : 0);
            npair += n;
            rows |= (n > 0 ? 1u : 0u) << q;
          }
      }
      // pass 2: one pair per iteration; a lane whose row is exhausted takes its next non-empty row (a single predicated block, never a loop)
      int s = 0, e = 0;
      for (;;) {
        if (s == own_s) s = own_e;
        if (s >= e) {
          if (!rows) break;
          const int q = __ffs(rows) - 1;
          rows &= rows - 1;
          const int qz = (w == 5) ? (q * 52) >> 8 : (w == 3) ? (q * 86) >> 8 : 0, qy = q - qz * w;
          const unsigned a_rb = a_own + 4 * (((qz - pr) * TB_RY + (qy - pr)) * TB_RX);
          s = lds_i32(a_rb - 4 * pr); e = lds_i32(a_rb + 4 * (pr + 1));
          if (s == own_s) s = own_e;
        }
        pair_force_fast(pi, lds_f4(a_src + 16u * (unsigned)s), KC, acc);
        ++s;
      }
      fm = fmaxf(fm, ppext_apply(reinterpret_cast<float2*>(xv) + 3LL * __float_as_int(me.w), acc, P));
    }
  }
  fm = warp_max(fm);
  if (lane == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fm);
  npair = warp_sum_i(npair);
  if (lane == 0 && npair) atomicAdd(&cnt->pairs_ppext, (unsigned long long)npair);
}

// direct walk for the targets of the blocks the tiled kernel could not hold (same block decode)
__global__ void __launch_bounds__(TB_NT) ppext_blocklist_kernel(float* __restrict__ xv, const int* __restrict__ fstart, int H, int b, int nc_buf, int nc_node, int nbx,
                                                                int nby, int pr, PPParams P, DevCounters* __restrict__ cnt, const int* __restrict__ n_list,
                                                                const int* __restrict__ list, const int* __restrict__ n_items, int item_cap) {
  if (item_cap > 0 && *n_items <= item_cap) return;    // the (cell, chunk) items were listed in full: ppext_cell_kernel does the dense blocks
  const int n = *n_list, tid = threadIdx.x;
  const int phys_hi = nc_buf + nc_node;
  float fm = 0.f;
  unsigned npair = 0;
  for (int k = blockIdx.x; k < n; k += gridDim.x) {
    const int blk = list[k];
    const int bx = blk % nbx, by = (blk / nbx) % nby, bz = blk / (nbx * nby);
    const int cx0 = nc_buf + bx * TB_X, cy0 = nc_buf + by * TB_Y, cz0 = nc_buf + bz * TB_Z;
    for (int rr = 0; rr < TB_Y * TB_Z; ++rr) {      // TB_Y x TB_Z coarse x-rows, each a contiguous range of the sorted array
      const int cy = cy0 + rr % TB_Y, cz = cz0 + rr / TB_Y;
      if (cy >= phys_hi || cz >= phys_hi) continue;
      const long long rk = (long long)(cz * H + cy) * H;
      const int g0 = fstart[(rk + cx0) * 64], g1 = fstart[(rk + min(cx0 + TB_X, phys_hi) - 1) * 64 + 64];
      for (int i = g0 + tid; i < g1; i += TB_NT) {
        float2* p = reinterpret_cast<float2*>(xv) + 3LL * i;
        const float2 a = p[0];
        const float z = p[1].x;
        const int gx = (int)floorf(a.x) + b, gy = (int)floorf(a.y) + b, gz = (int)floorf(z) + b;
        fm = fmaxf(fm, ppext_apply(p, ppext_direct(xv, fstart, H, pr, make_float3(a.x, a.y, z), gx, gy, gz, P, &npair), P));
      }
    }
  }
  fm = warp_max(fm);
  if ((tid & 31) == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fm);
  const unsigned long long np_w = (unsigned long long)__reduce_add_sync(0xffffffffu, npair);
  if ((tid & 31) == 0 && np_w) atomicAdd(&cnt->pairs_ppext, np_w);
}


// ---------------------------------------------------------------------------------------------------------------------------------
// PP_EXT for the DENSE blocks (the overflow list of the tiled kernel): cell-pair tiling with one WARP per (target fine cell, chunk of 32 targets).
// In a clustered box most pair interactions live in a few fine cells holding tens to thousands of particles; every target of such a cell meets the
// SAME sources (the particles of the 5^3 - 1 cells around it: 24 neighbour rows + the two halves of the centre row = at most 52 contiguous ranges of
// the sorted array), so the warp looks the ranges up once (one row per lane, all look-ups in flight together) and then streams the sources through
// all lanes. Lanes = (target, source slice): a chunk of n <= 32 targets uses T = the next power of two >= n target slots and 32/T slices of the
// source stream, so a cell of 4 particles still keeps all 32 lanes on pair evaluations; the slices are reduced with shuffles at the end, the slice-0
// lane applies the kick and enters |F| into pp_ext_force_max. The one-thread-per-target direct walk this replaces ran one divergent 125-cell walk per
// lane and left the densest block to a single CTA: 39 G pairs/s on a z = 2 box (1.2 % of the FP32 peak).
// ppext_items_kernel lists the (cell, chunk) items of the overflow blocks; ppext_cell_kernel takes them from an atomic ticket (heavy cells' chunks
// are adjacent in the list, so they start together). If the item list overflows its capacity the direct walk (ppext_blocklist_kernel) is used
// instead, decided on the device.
__global__ void __launch_bounds__(TB_NT) ppext_items_kernel(const int* __restrict__ fstart, int H, int nc_buf, int nc_node, int nbx, int nby, const int* __restrict__ n_list,
                                                            const int* __restrict__ list, int2* __restrict__ items, int cap, int* __restrict__ n_items) {
  const int n = *n_list, tid = threadIdx.x;
  const int phys_hi = nc_buf + nc_node;
  for (int k = blockIdx.x; k < n; k += gridDim.x) {
    const int blk = list[k];
    const int bx = blk % nbx, by = (blk / nbx) % nby, bz = blk / (nbx * nby);
    const int cx0 = nc_buf + bx * TB_X, cy0 = nc_buf + by * TB_Y, cz0 = nc_buf + bz * TB_Z;
    for (int c = tid; c < TB_X * TB_Y * TB_Z * 64; c += TB_NT) {
      const int cc = c >> 6, f = c & 63;
      const int cx = cx0 + cc % TB_X, cy = cy0 + (cc / TB_X) % TB_Y, cz = cz0 + cc / (TB_X * TB_Y);
      if (cx >= phys_hi || cy >= phys_hi || cz >= phys_hi) continue;
      const int key = ((cz * H + cy) * H + cx) * 64 + f;
      const int cnt = fstart[key + 1] - fstart[key];
      if (cnt <= 0) continue;
      const int nch = (cnt + 31) >> 5;
      const int base = atomicAdd(n_items, nch);
      for (int q = 0; q < nch; ++q) if (base + q < cap) items[base + q] = make_int2(key, q);
    }
  }
}

__global__ void __launch_bounds__(TB_NT) ppext_cell_kernel(float* __restrict__ xv, const int* __restrict__ fstart, int H, int pr, PPParams P, DevCounters* __restrict__ cnt,
                                                           const int2* __restrict__ items, int cap, const int* __restrict__ n_items, int* __restrict__ ticket) {
  const int n = *n_items;
  if (n > cap) return;                                  // list overflow: ppext_blocklist_kernel does the work
  const int lane = threadIdx.x & 31;
  const float2* xv2 = reinterpret_cast<const float2*>(xv);
  const PairConst KC = make_pair_const(P);
  float fm = 0.f;
  unsigned long long npair = 0;
  // dynamic distribution (item costs span three orders of magnitude: a static round-robin deal was 15 % slower on the clustered boxes); the ticket of the
  // NEXT item is drawn before the current one is processed, so its L2 round trip is off the critical path
  int nxt = 0;
  if (lane == 0) nxt = atomicAdd(ticket, 1);
  for (;;) {
    const int it = __shfl_sync(0xffffffffu, nxt, 0);
    if (it >= n) break;
    if (lane == 0) nxt = atomicAdd(ticket, 1);
    const int2 item = items[it];
    const int key = item.x;
    const int f = key & 63, cc = key >> 6;
    const int cx = cc % H, cy = (cc / H) % H, cz = cc / (H * H);
    const int gx = 4 * cx + (f & 3), gy = 4 * cy + ((f >> 2) & 3), gz = 4 * cz + (f >> 4);
    const int t0 = fstart[key] + 32 * item.y, t1 = min(fstart[key + 1], t0 + 32);
    const int nt = t1 - t0;
    // ---- source ranges: lane r < 25 owns neighbour row r (dz = r/5 - 2, dy = r%5 - 2 for pr = 2; generally w = 2 pr + 1); the centre row's lane holds
    //      the cells left of the own cell, lane 25..: the cells right of it. Each row is at most two ranges (two coarse cells).
    const int w = 2 * pr + 1, nrow = w * w, qc = (nrow - 1) >> 1;
    int sa = 0, ea = 0, sb = 0, eb = 0;
    if (lane <= nrow) {
      const int r = lane < nrow ? lane : qc;
      const int nz = gz + r / w - pr, ny = gy + r % w - pr;
      int xa = gx - pr, xb = gx + pr;
      if (r == qc) { if (lane < nrow) xb = gx - 1; else xa = gx + 1; }
      if (xa <= xb) {
        const int rowkey = (((nz >> 2) * H + (ny >> 2)) * H) * 64 + (((nz & 3) << 4) | ((ny & 3) << 2));
        const int ca = xa >> 2, cb = xb >> 2;
        const int ka = rowkey + ca * 64, kb = rowkey + cb * 64;
        sa = fstart[ka + (xa & 3)]; ea = fstart[ka + (ca == cb ? (xb & 3) : 3) + 1];
        if (cb != ca) { sb = fstart[kb]; eb = fstart[kb + (xb & 3) + 1]; }
      }
    }
    // ---- lanes = (target slot, source slice)
    int T = 1;
    while (T < nt) T <<= 1;
    const int sf = 32 / T, slot = lane & (T - 1), slice = lane / T;
    const bool live = slot < nt;
    float3 pi = make_float3(0.f, 0.f, 0.f);
    if (live) { const float2* p = xv2 + 3LL * (t0 + slot); const float2 a = p[0]; pi = make_float3(a.x, a.y, p[1].x); }
    float3 acc = make_float3(0.f, 0.f, 0.f);
    int nsrc = 0;
#pragma unroll 1
    for (int r = 0; r <= nrow; ++r) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int s = __shfl_sync(0xffffffffu, h ? sb : sa, r), e = __shfl_sync(0xffffffffu, h ? eb : ea, r);
        nsrc += max(e - s, 0);
        // four sources per trip, all eight loads issued before the first pair is evaluated (the sources come from L2: the loop is bound by load
        // latency, not by the 16-instruction pair body, unless several loads are in flight per lane)
        int j = s + slice;
#pragma unroll 1
        for (; j + 3 * sf < e; j += 4 * sf) {
          float2 a[4]; float z[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { const float2* q = xv2 + 3LL * (j + u * sf); a[u] = q[0]; z[u] = q[1].x; }
#pragma unroll
          for (int u = 0; u < 4; ++u) pair_force_fast(pi, make_float4(a[u].x, a[u].y, z[u], 0.f), KC, acc);
        }
#pragma unroll 1
        for (; j < e; j += sf) {
          const float2* q = xv2 + 3LL * j;
          const float2 a = q[0];
          const float z = q[1].x;
          pair_force_fast(pi, make_float4(a.x, a.y, z, 0.f), KC, acc);
        }
      }
    }
    for (int o = T; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    }
    if (live && slice == 0) fm = fmaxf(fm, ppext_apply(reinterpret_cast<float2*>(xv) + 3LL * (t0 + slot), acc, P));
    if (lane == 0) npair += (unsigned long long)nsrc * (unsigned long long)nt;
  }
  fm = warp_max(fm);
  if (lane == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fm);
  if (lane == 0 && npair) atomicAdd(&cnt->pairs_ppext, npair);
}


// ---- the same kernel with TMA staging (CUBEP3M_B200_PPEXT_DENSE=tma).
// Source staging: ranges of at least CT_MIN records (the clumps, where the pairs are) are streamed through a per-warp shared-memory ring by the TMA engine —
// cp.async.bulk of up to CT_CH 24-byte records per chunk, completion on an mbarrier, the next chunk in flight while the current one is evaluated — so the
// pair loop reads its sources with two LDS instead of two L2 loads (ncu on the clustered profile box before: long-scoreboard 51 % of the stalls with the
// FMA pipe at 40 %). Records start 8-byte aligned (24 j), bulk copies need 16: an odd first record is fetched from 8 bytes earlier. Short ranges (the sparse
// cells around a clump) keep the direct loads, four in flight per lane.
// MEASURED (B200): no gain — z = 2 evolved box: dense blocks 1.78 ms with direct loads, 2.60 ms staged (2.75 with a two-slot ring); clump-dominated profile
// box: 4.3 vs 4.3 ms. The direct version already keeps 4 x 32 records in flight per warp out of L2 (96 % L2 hit rate), the staged one pays two warp barriers,
// an mbarrier wait and 25 KB of shared memory per CTA for every 64 records. Kept as an A/B knob; ppext_cell_kernel (direct loads) is the default.
constexpr int CT_CH = 64, CT_MIN = 24, CT_BUF = CT_CH * 24 + 16, CT_NS = 4;   // ring: CT_NS slots per warp, up to CT_NS chunks (256 records) in flight
__global__ void __launch_bounds__(TB_NT) ppext_cell_tma_kernel(float* __restrict__ xv, const int* __restrict__ fstart, int H, int pr, PPParams P, DevCounters* __restrict__ cnt,
                                                               const int2* __restrict__ items, int cap, const int* __restrict__ n_items, int* __restrict__ ticket) {
  __shared__ __align__(16) unsigned char ring[TB_NT / 32][CT_NS][CT_BUF];
  __shared__ __align__(8) unsigned long long bars[TB_NT / 32][CT_NS];
  __shared__ int meta[TB_NT / 32][CT_NS][2];            // (first record is odd, records) of the chunk in each slot
  const int n = *n_items;
  if (n > cap) return;                                  // list overflow: ppext_blocklist_kernel does the work
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float2* xv2 = reinterpret_cast<const float2*>(xv);
  const PairConst KC = make_pair_const(P);
  const unsigned bar_base = fftk::smem_u32(&bars[warp][0]), buf_base = fftk::smem_u32(&ring[warp][0][0]);
  if (lane == 0) { for (int q = 0; q < CT_NS; ++q) fftk::mbar_init(bar_base + 8u * q, 1); fftk::mbar_fence_init(); }
  __syncwarp();
  unsigned phase = 0;                                   // bit q: mbarrier phase parity of ring slot q (uniform over the warp)
  float fm = 0.f;
  unsigned long long npair = 0;
  // dynamic distribution (item costs span three orders of magnitude: a static round-robin deal was 15 % slower on the clustered boxes); the ticket of the
  // NEXT item is drawn before the current one is processed, so its L2 round trip is off the critical path
  int nxt = 0;
  if (lane == 0) nxt = atomicAdd(ticket, 1);
  for (;;) {
    const int it = __shfl_sync(0xffffffffu, nxt, 0);
    if (it >= n) break;
    if (lane == 0) nxt = atomicAdd(ticket, 1);
    const int2 item = items[it];
    const int key = item.x;
    const int f = key & 63, cc = key >> 6;
    const int cx = cc % H, cy = (cc / H) % H, cz = cc / (H * H);
    const int gx = 4 * cx + (f & 3), gy = 4 * cy + ((f >> 2) & 3), gz = 4 * cz + (f >> 4);
    const int t0 = fstart[key] + 32 * item.y, t1 = min(fstart[key + 1], t0 + 32);
    const int nt = t1 - t0;
    // ---- source ranges: lane r < 25 owns neighbour row r (dz = r/5 - 2, dy = r%5 - 2 for pr = 2; generally w = 2 pr + 1); the centre row's lane holds
    //      the cells left of the own cell, lane 25..: the cells right of it. Each row is at most two ranges (two coarse cells).
    const int w = 2 * pr + 1, nrow = w * w, qc = (nrow - 1) >> 1;
    int sa = 0, ea = 0, sb = 0, eb = 0;
    if (lane <= nrow) {
      const int r = lane < nrow ? lane : qc;
      const int nz = gz + r / w - pr, ny = gy + r % w - pr;
      int xa = gx - pr, xb = gx + pr;
      if (r == qc) { if (lane < nrow) xb = gx - 1; else xa = gx + 1; }
      if (xa <= xb) {
        const int rowkey = (((nz >> 2) * H + (ny >> 2)) * H) * 64 + (((nz & 3) << 4) | ((ny & 3) << 2));
        const int ca = xa >> 2, cb = xb >> 2;
        const int ka = rowkey + ca * 64, kb = rowkey + cb * 64;
        sa = fstart[ka + (xa & 3)]; ea = fstart[ka + (ca == cb ? (xb & 3) : 3) + 1];
        if (cb != ca) { sb = fstart[kb]; eb = fstart[kb + (xb & 3) + 1]; }
      }
    }
    // ---- lanes = (target slot, source slice)
    int T = 1;
    while (T < nt) T <<= 1;
    const int sf = 32 / T, slot = lane & (T - 1), slice = lane / T;
    const bool live = slot < nt;
    float3 pi = make_float3(0.f, 0.f, 0.f);
    if (live) { const float2* p = xv2 + 3LL * (t0 + slot); const float2 a = p[0]; pi = make_float3(a.x, a.y, p[1].x); }
    float3 acc = make_float3(0.f, 0.f, 0.f);
    int nsrc = 0;
    const int nrange = 2 * (nrow + 1);
    // ---- short ranges: direct loads
#pragma unroll 1
    for (int rr = 0; rr < nrange; ++rr) {
      const int s = __shfl_sync(0xffffffffu, (rr & 1) ? sb : sa, rr >> 1), e = __shfl_sync(0xffffffffu, (rr & 1) ? eb : ea, rr >> 1);
      const int len = e - s;
      nsrc += max(len, 0);
      if (len <= 0 || len >= CT_MIN) continue;
      int j = s + slice;
#pragma unroll 1
      for (; j + 3 * sf < e; j += 4 * sf) {
        float2 a[4]; float z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const float2* q = xv2 + 3LL * (j + u * sf); a[u] = q[0]; z[u] = q[1].x; }
#pragma unroll
        for (int u = 0; u < 4; ++u) pair_force_fast(pi, make_float4(a[u].x, a[u].y, z[u], 0.f), KC, acc);
      }
#pragma unroll 1
      for (; j < e; j += sf) {
        const float2* q = xv2 + 3LL * j;
        const float2 a = q[0];
        pair_force_fast(pi, make_float4(a.x, a.y, q[1].x, 0.f), KC, acc);
      }
    }
    // ---- long ranges: chunks of <= CT_CH records through the shared-memory ring
    {
      int rr = 0, cs = 0, ce = 0;                       // chunk iterator, uniform over the warp
      auto next_chunk = [&](int& j0, int& cntc) -> bool {
        while (cs >= ce) {
          if (rr >= nrange) return false;
          const int s = __shfl_sync(0xffffffffu, (rr & 1) ? sb : sa, rr >> 1), e = __shfl_sync(0xffffffffu, (rr & 1) ? eb : ea, rr >> 1);
          ++rr;
          if (e - s >= CT_MIN) { cs = s; ce = e; }
        }
        j0 = cs; cntc = min(CT_CH, ce - cs); cs += cntc;
        return true;
      };
      int issued = 0, consumed = 0;
      auto top_up = [&]() {
        while (issued - consumed < CT_NS) {
          int j0, cntc;
          if (!next_chunk(j0, cntc)) break;
          const int q = issued % CT_NS;
          if (lane == 0) {
            const unsigned odd = (unsigned)(j0 & 1);
            const unsigned bytes = (24u * (unsigned)cntc + 8u * odd + 15u) & ~15u;
            meta[warp][q][0] = (int)odd; meta[warp][q][1] = cntc;
            fftk::fence_proxy_async();                // the warp's generic reads of this slot are ordered before the async-proxy writes
            fftk::mbar_expect_tx(bar_base + 8u * q, bytes);
            fftk::bulk_g2s(buf_base + (unsigned)(CT_BUF * q), reinterpret_cast<const char*>(xv) + 24LL * j0 - 8 * (int)odd, bytes, bar_base + 8u * q);
          }
          ++issued;
        }
      };
      top_up();
      while (consumed < issued) {
        const int q = consumed % CT_NS;
        __syncwarp();                                   // lane 0's meta[] entry is visible
        fftk::mbar_wait(bar_base + 8u * q, (phase >> q) & 1u);
        phase ^= 1u << q;
        const int c0 = meta[warp][q][1];
        const unsigned base = buf_base + (unsigned)(CT_BUF * q) + 8u * (unsigned)meta[warp][q][0];
#pragma unroll 2
        for (int jj = slice; jj < c0; jj += sf) {
          float x, y, z;
          const unsigned ad = base + 24u * (unsigned)jj;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%3];\n\tld.shared.f32 %2, [%3+8];" : "=f"(x), "=f"(y), "=f"(z) : "r"(ad) : "memory");
          pair_force_fast(pi, make_float4(x, y, z, 0.f), KC, acc);
        }
        ++consumed;
        __syncwarp();                                   // every lane is done with slot q before it is refilled
        top_up();
      }
    }
    for (int o = T; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    }
    if (live && slice == 0) fm = fmaxf(fm, ppext_apply(reinterpret_cast<float2*>(xv) + 3LL * (t0 + slot), acc, P));
    if (lane == 0) npair += (unsigned long long)nsrc * (unsigned long long)nt;
  }
  fm = warp_max(fm);
  if (lane == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fm);
  if (lane == 0 && npair) atomicAdd(&cnt->pairs_ppext, npair);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// PPINT as (fine cell, 32-target chunk) items, the same lane mapping as ppext_cell_kernel with the cell's own range as the only source range
// (self-pair excluded). One warp per CELL (ppint_kernel) leaves a halo core of a few thousand particles to a single warp: 19 ms for 3.8e8 pairs on the
// clustered profile box, 20 G pairs/s.
// items: chunks of cells with more than PPINT_HEAVY particles fill the list from the front (n_items[0], drawn by ticket: their costs vary by orders of
// magnitude), all others from the back (n_items[1], dealt round-robin: there are millions of them and a ticket per 2-particle cell is pure latency)
constexpr int PPINT_HEAVY = 64;
__global__ void __launch_bounds__(TB_NT) ppint_items_kernel(const int* __restrict__ fstart, const int* __restrict__ list, const int* __restrict__ n_list_ptr, int list_cap,
                                                            int max_llf, int2* __restrict__ items, int cap, int* __restrict__ n_items, DevCounters* __restrict__ cnt) {
  const int n_list = min(*n_list_ptr, list_cap);
  for (int w = blockIdx.x * TB_NT + threadIdx.x; w < n_list; w += gridDim.x * TB_NT) {
    const int k = list[w];
    const int c = fstart[k + 1] - fstart[k];
    if (c > max_llf) { atomicOr(&cnt->overflow, 4); continue; }   // 'exceeded max_llf' :280-283
    const int nch = (c + 31) >> 5;
    if (c > PPINT_HEAVY) {
      const int base = atomicAdd(&n_items[0], nch);
      for (int q = 0; q < nch; ++q) if (base + q < cap) items[base + q] = make_int2(k, q);
    } else {
      const int base = atomicAdd(&n_items[1], nch);
      for (int q = 0; q < nch; ++q) if (base + q < cap) items[cap - 1 - (base + q)] = make_int2(k, q);
    }
  }
}

__global__ void __launch_bounds__(TB_NT) ppint_cell_kernel(float* __restrict__ xv, const int* __restrict__ fstart, PPParams P, DevCounters* __restrict__ cnt,
                                                           const int2* __restrict__ items, int cap, const int* __restrict__ n_items, int* __restrict__ ticket) {
  const int nh = n_items[0], nl = n_items[1];
  if (nh + nl > cap) return;                            // list overflow: ppint_kernel does the work
  const int lane = threadIdx.x & 31;
  const float2* xv2 = reinterpret_cast<const float2*>(xv);
  const PairConst KC = make_pair_const(P);
  float fm = 0.f;
  unsigned long long npair = 0;
  auto run_item = [&](const int2 item) {
    const int s = fstart[item.x], e = fstart[item.x + 1];
    const int t0 = s + 32 * item.y, nt = min(e, t0 + 32) - t0;
    int T = 1;
    while (T < nt) T <<= 1;
    const int sf = 32 / T, slot = lane & (T - 1), slice = lane / T;
    const bool live = slot < nt;
    const int me = t0 + slot;
    float3 pi = make_float3(0.f, 0.f, 0.f);
    if (live) { const float2* p = xv2 + 3LL * me; const float2 a = p[0]; pi = make_float3(a.x, a.y, p[1].x); }
    float3 acc = make_float3(0.f, 0.f, 0.f);
#pragma unroll 2
    for (int j = s + slice; j < e; j += sf) {
      const float2* q = xv2 + 3LL * j;
      const float2 a = q[0];
      const float z = q[1].x;
      // w = mass_p / (r pp_bias)^3 for r > rsoft (:344); the pair of a particle with itself has r = 0 and drops out with the soft core
      const float sx = pi.x - a.x, sy = pi.y - a.y, sz = pi.z - z;
      const float r2 = sx * sx + sy * sy + sz * sz;
      const float ir = rsqrt_ftz(r2);
      float w = KC.c1 * (ir * ir * ir);
      w = (r2 > KC.rs2 && j != me) ? w : 0.0f;
      acc.x = fmaf(-sx, w, acc.x); acc.y = fmaf(-sy, w, acc.y); acc.z = fmaf(-sz, w, acc.z);
    }
    for (int o = T; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    }
    if (live && slice == 0) fm = fmaxf(fm, ppext_apply(reinterpret_cast<float2*>(xv) + 3LL * me, acc, P));     // :349-358
    if (lane == 0) npair += (unsigned long long)(e - s - 1) * (unsigned long long)nt;
  };
  // heavy items: dynamic, the next ticket always one draw ahead (see ppext_cell_kernel)
  int nxt = 0;
  if (lane == 0) nxt = atomicAdd(ticket, 1);
  for (;;) {
    const int it = __shfl_sync(0xffffffffu, nxt, 0);
    if (it >= nh) break;
    if (lane == 0) nxt = atomicAdd(ticket, 1);
    run_item(items[it]);
  }
  // light items: round-robin over the warps of the grid
  const int nwarps = gridDim.x * (TB_NT / 32);
  for (int it = blockIdx.x * (TB_NT / 32) + (threadIdx.x >> 5); it < nl; it += nwarps) run_item(items[cap - 1 - it]);
  fm = warp_max(fm);
  if (lane == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_force_max_bits, fm);
  if (lane == 0 && npair) atomicAdd(&cnt->pairs_ppint, npair);
}

}  // namespace pp
