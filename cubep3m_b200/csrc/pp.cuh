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
                                                    DevCounters* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (TPB / 32);
  const int n_list = min(*n_list_ptr, list_cap);
  float fmax = 0.f;
  for (int w = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); w < n_list; w += nwarps) {
    const int k = list[w];
    const int s0 = fstart[k], s1 = fstart[k + 1];
    if (s1 - s0 > max_llf) { if (lane == 0) atomicOr(&cnt->overflow, 4); continue; }   // 'exceeded max_llf' :280-283
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
}

// PP_EXT, one THREAD per target particle of the cell-sorted array (ghosts and non-physical cells are skipped). At the mean density
// of 1/8 particle per fine cell a warp-per-cell mapping leaves 31 of 32 lanes idle; here every lane owns a target and walks the
// (2*pr+1)^2 neighbour fine-cell rows. The cells gx-pr..gx+pr of one row lie in at most two coarse cells, and inside one coarse
// cell consecutive fx are consecutive keys, so each row is at most two contiguous particle ranges found with two fstart reads each.
// Neighbouring lanes hold neighbouring particles, so their range lookups and source loads hit the same L1 lines.
// The same-cell pairs belong to PPINT (:496-523 excludes the cell itself): the centre row is walked as [gx-pr,gx-1] and [gx+1,gx+pr].
// one contiguous range of sources
__device__ __forceinline__ void ppext_sources(const float* __restrict__ xv, int s, int e, const float3 pi, const PPParams& P, float3& acc) {
#pragma unroll 1
  for (int j = s; j < e; ++j) {
    const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * j;
    const float2 a = p[0];
    pair_force(pi, make_float3(a.x, a.y, p[1].x), P, true, acc);
  }
}

constexpr int EXT_TPB = 128;
constexpr int EXT_MAXR = 2;    // largest pp_range (cubepm.par:92 uses 2)
__global__ void __launch_bounds__(EXT_TPB) ppext_kernel(float* __restrict__ xv, const int* __restrict__ fstart, int np_all, int H, int b, int nc_buf,
                                                        int nc_node, int pr, PPParams P, DevCounters* __restrict__ cnt) {
  const int i = blockIdx.x * EXT_TPB + threadIdx.x;
  float fm = 0.f;
  if (i < np_all) {
    float2* p = reinterpret_cast<float2*>(xv) + 3LL * i;
    const float2 a = p[0];
    const float z = p[1].x;
    const int gx = (int)floorf(a.x) + b, gy = (int)floorf(a.y) + b, gz = (int)floorf(z) + b;   // as part::make_key
    const int lo = nc_buf * 4, hi = (nc_buf + nc_node) * 4;
    if (gx >= lo && gx < hi && gy >= lo && gy < hi && gz >= lo && gz < hi) {                      // kick only particles of the physical cells (:576-590)
      const float3 pi = make_float3(a.x, a.y, z);
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
          ppext_sources(xv, rs[q][0], re[q][0], pi, P, acc);
          ppext_sources(xv, rs[q][1], re[q][1], pi, P, acc);
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
            ppext_sources(xv, fstart[k0 + f0], fstart[k0 + f1 + 1], pi, P, acc);
          }
        }
      }
      fm = sqrtf(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z);   // :617
      if (P.apply) {
        const float s = P.a_mid * P.G * P.dt;
        float2 bq = p[1], c = p[2];
        bq.y += acc.x * s; c.x += acc.y * s; c.y += acc.z * s;
        p[1] = bq; p[2] = c;
      }
    }
  }
  fm = warp_max(fm);
  if ((threadIdx.x & 31) == 0 && fm > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fm);
}

}  // namespace pp
