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

// one warp per occupied physical fine cell; neighbours = the (2*pr+1)^3 - 1 surrounding fine cells
__global__ void __launch_bounds__(TPB) ppext_kernel(float* __restrict__ xv, const int* __restrict__ fstart, const int* __restrict__ list,
                                                    const int* __restrict__ n_list_ptr, int list_cap, int H, int pr, PPParams P,
                                                    DevCounters* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (TPB / 32);
  const int n_list = min(*n_list_ptr, list_cap);
  const int side = 2 * pr + 1, ncell = side * side * side;
  float fmax = 0.f;
  for (int w = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); w < n_list; w += nwarps) {
    const int k = list[w];
    const int s0 = fstart[k], s1 = fstart[k + 1];
    // decode the fine cell
    const int sub = k & 63;
    const unsigned int cc = (unsigned int)k >> 6;
    const int gx = (int)(cc % H) * 4 + (sub & 3), gy = (int)((cc / H) % H) * 4 + ((sub >> 2) & 3), gz = (int)(cc / (H * H)) * 4 + (sub >> 4);
    // each lane looks up ceil(ncell/32) neighbour cells
    int nst[4], nen[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c = r * 32 + lane;
      nst[r] = 0; nen[r] = 0;
      if (c < ncell) {
        const int dx = c % side - pr, dy = (c / side) % side - pr, dz = c / (side * side) - pr;
        if (dx != 0 || dy != 0 || dz != 0) {
          const long long kk = fine::cell_key(gx + dx, gy + dy, gz + dz, H);
          nst[r] = fstart[kk]; nen[r] = fstart[kk + 1];
        }
      }
    }
    for (int ib = s0; ib < s1; ib += 32) {
      const int i = ib + lane;
      const bool vi = i < s1;
      float3 pi = make_float3(0.f, 0.f, 0.f);
      if (vi) { const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * i; const float2 a = p[0]; pi = make_float3(a.x, a.y, p[1].x); }
      float3 acc = make_float3(0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        unsigned m = __ballot_sync(0xffffffffu, nen[r] > nst[r]);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int t0 = __shfl_sync(0xffffffffu, nst[r], src), t1 = __shfl_sync(0xffffffffu, nen[r], src);
          for (int jb = t0; jb < t1; jb += 32) {
            const int j = jb + lane;
            float3 pj = make_float3(0.f, 0.f, 0.f);
            if (j < t1) { const float2* p = reinterpret_cast<const float2*>(xv) + 3LL * j; const float2 a = p[0]; pj = make_float3(a.x, a.y, p[1].x); }
            const int nj = min(32, t1 - jb);
            for (int l = 0; l < nj; ++l) {
              const float3 q = make_float3(__shfl_sync(0xffffffffu, pj.x, l), __shfl_sync(0xffffffffu, pj.y, l), __shfl_sync(0xffffffffu, pj.z, l));
              if (vi) pair_force(pi, q, P, true, acc);
            }
          }
        }
      }
      if (vi) {
        fmax = fmaxf(fmax, sqrtf(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z));   // :617
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
  if (lane == 0 && fmax > 0.f) atomic_max_float_nonneg(&cnt->pp_ext_force_max_bits, fmax);
}

}  // namespace pp
