// Particle-side stages: drift, ghost exchange (particle_pass), cell sort (the sorted-array replacement of
// link_list's ll/hoc chains), ghost deletion.
//
// Particle records stay in the reference's layout: xv(6,i) = 24-byte AoS (cubep3m.fh:75), accessed as 3 x float2.
// Positions are moved with explicit __fadd_rn/__fmul_rn so nvcc cannot contract them into FMAs: cell
// membership must be bit-identical to the reference's unfused evaluation.
#pragma once
#include "common.cuh"

namespace part {

constexpr int TPB = 256;
constexpr unsigned int KEY_DEAD = 0xffffffffu;

__device__ __forceinline__ void load_xv(const float* __restrict__ xv, long long i, float2& a, float2& b, float2& c) {
  const float2* p = reinterpret_cast<const float2*>(xv) + 3 * i;
  a = p[0]; b = p[1]; c = p[2];
}
__device__ __forceinline__ void store_xv(float* __restrict__ xv, long long i, float2 a, float2 b, float2 c) {
  float2* p = reinterpret_cast<float2*>(xv) + 3 * i;
  p[0] = a; p[1] = b; p[2] = c;
}

// update_position.f90:71  x = ((x + ((v*0.5)*(dt+dt_old))) + offset), evaluated unfused
__global__ void __launch_bounds__(TPB) drift_kernel(float* __restrict__ xv, int np, float hdt, float ox, float oy, float oz) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= np) return;
  float2 a, b, c;
  load_xv(xv, i, a, b, c);   // a=(x,y) b=(z,vx) c=(vy,vz)
  a.x = __fadd_rn(__fadd_rn(a.x, __fmul_rn(__fmul_rn(b.y, 0.5f), hdt)), ox);
  a.y = __fadd_rn(__fadd_rn(a.y, __fmul_rn(__fmul_rn(c.x, 0.5f), hdt)), oy);
  b.x = __fadd_rn(__fadd_rn(b.x, __fmul_rn(__fmul_rn(c.y, 0.5f), hdt)), oz);
  float2* p = reinterpret_cast<float2*>(xv) + 3 * i;
  p[0] = a; p[1] = b;
}

// move_grid_back.f90:20-23
__global__ void __launch_bounds__(TPB) shift_kernel(float* __restrict__ xv, int np, float sx, float sy, float sz) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= np) return;
  float2* p = reinterpret_cast<float2*>(xv) + 3 * i;
  float2 a = p[0], b = p[1];
  a.x = __fsub_rn(a.x, sx); a.y = __fsub_rn(a.y, sy); b.x = __fsub_rn(b.x, sz);
  p[0] = a; p[1] = b;
}

// checkpoint.f90:92: write(12) xv(1:3,j) - shake_offset, xv(4:6,j)  — a block of records staged for the writer
__global__ void __launch_bounds__(TPB) checkpoint_pack_kernel(const float* __restrict__ xv, long long first, int count, float sx, float sy, float sz,
                                                              float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= count) return;
  float2 a, b, c;
  load_xv(xv, first + i, a, b, c);
  a.x = __fsub_rn(a.x, sx); a.y = __fsub_rn(a.y, sy); b.x = __fsub_rn(b.x, sz);
  store_xv(out, i, a, b, c);
}

// link_list.f90:26-31: a particle is chained iff floor(x/4)+1 lies in [hoc_nc_l, hoc_nc_h] on all axes,
// i.e. -nf_buf <= x < mT + nf_buf.
__device__ __forceinline__ bool in_hoc_range(float x, float y, float z, float lo, float hi) {
  return x >= lo && x < hi && y >= lo && y < hi && z >= lo && z < hi;
}

// ---- cell keys.  Fine cell of a particle in the extended node frame: g = floor(x) + nf_buf in [0, mT + 2 nf_buf);
// coarse cell (0-based inside the hoc range) = g >> 2 == floor(x/4)+1 - hoc_nc_l  (link_list.f90:26-28);
// key = ((cz*H + cy)*H + cx)*64 + (fz*16 + fy*4 + fx): x-rows of coarse cells are contiguous, and so are the
// 64 fine cells of one coarse cell (the llf(.,4,4,4) binning of particle_mesh_threaded.f90:276-284).
__device__ __forceinline__ unsigned int make_key(float x, float y, float z, int b, int H) {
  const int gx = (int)floorf(x) + b, gy = (int)floorf(y) + b, gz = (int)floorf(z) + b;
  const unsigned int cx = gx >> 2, cy = gy >> 2, cz = gz >> 2;
  return ((cz * H + cy) * H + cx) * 64u + (unsigned)(((gz & 3) << 4) | ((gy & 3) << 2) | (gx & 3));
}

// key + histogram + boundary-candidate list for ONE particle at (x, y, z) (link_list.f90:26-47 as a counting sort): shared by key_hist_kernel and by the
// particle_pass kernels, which already hold every record in registers inside particle_mesh (the stand-alone pass over all records is saved).
// Lists the particles that sit within 2^-15 below an integer coordinate on any axis: only for those can the reference's tile-local cell
// floor(fl(x + offset)) (particle_mesh_threaded.f90:139-143) differ from floor(x)+offset.
// The value the histogram atomic returns is the particle's RANK inside its cell: it is stored (2 bytes per particle) and the scatter places the record at
// fstart[key] + rank without touching the table again (round 1 / early round 2: the scatter drew its slot with a second atomic that counted the cell back to
// zero — one more DRAM sector read-modify-write per particle on the sparse table; the table is now cleared by a memset that particle_mesh hides behind PP_EXT).
// The histogram holds two 16-bit counters per 32-bit word (cell k -> word k>>1, half k&1): the table is the largest array the sort touches
// (H^3*64 cells, ~8 per particle) and every atomic on it is a DRAM sector read-modify-write, so halving its footprint halves that traffic.
// A fine cell with >= 65535 particles raises the 'exceeded max_llf' flag (max_llf = 100000 in cubepm.par:183 is the same kind of bound).
struct KeyArgs { float lo, hi; int b, H; unsigned int* key; unsigned int* hist; float* cand; int cand_cap; unsigned short* rank; };
__device__ __forceinline__ void key_one(const KeyArgs& A, long long i, float x, float y, float z, DevCounters* __restrict__ cnt) {
  unsigned int k = KEY_DEAD;
  if (in_hoc_range(x, y, z, A.lo, A.hi)) {
    k = make_key(x, y, z, A.b, A.H);
    const unsigned sh = (k & 1u) << 4;
    const unsigned old = atomicAdd(&A.hist[k >> 1], 1u << sh);
    const unsigned before = (old >> sh) & 0xffffu;       // particles counted into this cell before this one = its rank inside the cell
    if (before >= 0xfffeu) atomicOr(&cnt->overflow, 4);
    A.rank[i] = (unsigned short)before;
    const float th = 3.0517578125e-05f;   // 2^-15 >= half an ulp of any |x + offset| < 1024
    // exact integers are not candidates: x + offset is exact for them, so both binnings agree
    const float ux = ceilf(x) - x, uy = ceilf(y) - y, uz = ceilf(z) - z;
    if ((ux > 0.f && ux <= th) || (uy > 0.f && uy <= th) || (uz > 0.f && uz <= th)) {
      const int slot = atomicAdd(&cnt->n_cand, 1);
      if (slot < A.cand_cap) { A.cand[3 * slot] = x; A.cand[3 * slot + 1] = y; A.cand[3 * slot + 2] = z; }
      else atomicOr(&cnt->overflow, 8);     // never silently truncated: the step returns ECAPACITY
    }
  } else {
    atomicAdd(&cnt->np_deleted, 1);   // 'PARTICLE DELETED' link_list.f90:32
  }
  A.key[i] = k;
}

// particle_pass.f90:73-94 (+ pass) and :173-192 (- pass) for one axis: every chained particle with
// x >= mT - nf_buf goes to the + neighbour, every one with x < nf_buf to the - neighbour.
// Both directions read the same pre-axis particle list (the reference relinks only after both).
// DRIFT: the first axis' pack also performs update_position (update_position.f90:71) on the record it has just loaded and writes the
// new position back, which saves the drift kernel's own pass over the particle array inside particle_mesh.
// The first axis' kernel (LIST = false) scans all np particles and also lists (blist) the chained ones that lie within nf_buf of a y or z
// face: only those, plus the ghosts received so far (indices >= np_first), can be sent along the later axes, whose kernels (LIST = true) then
// visit nlist + (np - np_first) records instead of all np.
// KEYS (first axis only): also the sort's key + histogram for every visited record (key_one).
template <bool DRIFT, bool LIST, bool KEYS = false>
__global__ void __launch_bounds__(TPB) pass_pack_kernel(float* __restrict__ xv, const int64_t* __restrict__ pid, int np, int axis,
                                                        float lo, float hi, float cut_hi, float cut_lo,
                                                        float* __restrict__ send_plus, float* __restrict__ send_minus,
                                                        int64_t* __restrict__ pid_plus, int64_t* __restrict__ pid_minus,
                                                        int cap, DevCounters* __restrict__ cnt, float hdt, float ox, float oy, float oz,
                                                        int* __restrict__ blist, int nlist, int np_first, KeyArgs KA) {
  const long long t = (long long)blockIdx.x * TPB + threadIdx.x;
  long long i = t;
  bool act = t < np;
  if (LIST) {
    act = t < (long long)nlist + (np - np_first);
    if (act) i = (t < nlist) ? blist[t] : (long long)np_first + (t - nlist);
  }
  bool gp = false, gm = false;
  float2 a, b, c;
  if (act) {
    load_xv(xv, i, a, b, c);
    if (DRIFT) {
      a.x = __fadd_rn(__fadd_rn(a.x, __fmul_rn(__fmul_rn(b.y, 0.5f), hdt)), ox);
      a.y = __fadd_rn(__fadd_rn(a.y, __fmul_rn(__fmul_rn(c.x, 0.5f), hdt)), oy);
      b.x = __fadd_rn(__fadd_rn(b.x, __fmul_rn(__fmul_rn(c.y, 0.5f), hdt)), oz);
      float2* p = reinterpret_cast<float2*>(xv) + 3 * i;
      p[0] = a; p[1] = b;
    }
    if (KEYS) key_one(KA, i, a.x, a.y, b.x, cnt);
    bool face = false;
    if (in_hoc_range(a.x, a.y, b.x, lo, hi)) {
      const float q = axis == 0 ? a.x : (axis == 1 ? a.y : b.x);
      gp = q >= cut_hi;
      gm = q < cut_lo;
      if (!LIST) face = a.y >= cut_hi || a.y < cut_lo || b.x >= cut_hi || b.x < cut_lo;
    }
    if (!LIST) {                                    // warp-aggregated append to the boundary list
      const unsigned mf = __ballot_sync(__activemask(), face);
      if (face) {
        const int lane_ = threadIdx.x & 31;
        const int leader = __ffs(mf) - 1;
        int base = 0;
        if (lane_ == leader) base = atomicAdd(&cnt->n_blist, __popc(mf));
        base = __shfl_sync(mf, base, leader);
        blist[base + __popc(mf & ((1u << lane_) - 1))] = (int)i;
      }
    }
  }
  // warp-aggregated slot allocation
  const unsigned mp = __ballot_sync(0xffffffffu, gp), mm = __ballot_sync(0xffffffffu, gm);
  const int lane = threadIdx.x & 31;
  int basep = 0, basem = 0;
  if (lane == 0) {
    if (mp) basep = atomicAdd(&cnt->n_send[0], __popc(mp));
    if (mm) basem = atomicAdd(&cnt->n_send[1], __popc(mm));
  }
  basep = __shfl_sync(0xffffffffu, basep, 0);
  basem = __shfl_sync(0xffffffffu, basem, 0);
  if (gp) {
    const int slot = basep + __popc(mp & ((1u << lane) - 1));
    if (slot < cap) { store_xv(send_plus, slot, a, b, c); if (pid) pid_plus[slot] = pid[i]; }
    else atomicOr(&cnt->overflow, 1);
  }
  if (gm) {
    const int slot = basem + __popc(mm & ((1u << lane) - 1));
    if (slot < cap) { store_xv(send_minus, slot, a, b, c); if (pid) pid_minus[slot] = pid[i]; }
    else atomicOr(&cnt->overflow, 1);
  }
}

// ---- peer-memory exchange (multi-GPU, NVLink): after the pack kernel has stored this rank's outgoing particles straight into the two
// neighbours' receive regions, one thread publishes the counts and raises the neighbours' flags; the receiver spins on its own two flags.
__global__ void pass_publish_kernel(const DevCounters* __restrict__ cnt, int* box_plus, int* box_minus, int epoch) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(box_plus) = cnt->n_send[0];
    *reinterpret_cast<volatile int*>(box_minus) = cnt->n_send[1];
    __threadfence_system();
    *reinterpret_cast<volatile int*>(box_plus + 1) = epoch;
    *reinterpret_cast<volatile int*>(box_minus + 1) = epoch;
  }
}
// mybox: [2 directions][count, flag]; out: {count from the - neighbour ("+" going), count from the + neighbour, time-out}
__global__ void pass_wait_kernel(const int* mybox, int epoch, int* __restrict__ out, long long timeout_cycles) {
  const int d = threadIdx.x;
  if (d < 2) {
    const volatile int* flag = mybox + 2 * d + 1;
    const long long t0 = clock64();
    bool ok = true;
    while (*flag != epoch) {
      if (clock64() - t0 > timeout_cycles) { ok = false; break; }
      __nanosleep(200);
    }
    __threadfence_system();
    out[d] = ok ? *reinterpret_cast<const volatile int*>(mybox + 2 * d) : 0;
    if (!ok) out[2] = 1;
  }
}

// receive side: shift + clamp, append at xv[np0 ...]
//   from the - neighbour's "+" buffer:  x = max(x - mT, -nf_buf)                          particle_pass.f90:162
//   from the + neighbour's "-" buffer:  |x|<eps -> +-eps ; x = min(x + mT, mT+nf_buf-eps)   particle_pass.f90:257-265
__global__ void __launch_bounds__(TPB) pass_unpack_kernel(float* __restrict__ xv, int64_t* __restrict__ pid, int np0, int axis,
                                                          const float* __restrict__ recv_plus, int n_plus,
                                                          const float* __restrict__ recv_minus, int n_minus,
                                                          const int64_t* __restrict__ rpid_plus, const int64_t* __restrict__ rpid_minus,
                                                          float fmT, float rnf_buf, float eps, float hi_clamp, KeyArgs KA, DevCounters* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= n_plus + n_minus) return;
  float2 a, b, c;
  const bool plus = i < n_plus;
  if (plus) load_xv(recv_plus, i, a, b, c); else load_xv(recv_minus, i - n_plus, a, b, c);
  float q = axis == 0 ? a.x : (axis == 1 ? a.y : b.x);
  if (plus) {
    q = fmaxf(__fsub_rn(q, fmT), -rnf_buf);
  } else {
    if (fabsf(q) < eps) q = (q < 0.0f) ? -eps : eps;
    q = fminf(__fadd_rn(q, fmT), hi_clamp);
  }
  if (axis == 0) a.x = q; else if (axis == 1) a.y = q; else b.x = q;
  store_xv(xv, (long long)np0 + i, a, b, c);
  if (KA.key) key_one(KA, (long long)np0 + i, a.x, a.y, b.x, cnt);   // inside particle_mesh: the ghost's sort key, while its record is in registers
  if (pid) pid[np0 + i] = plus ? rpid_plus[i] : rpid_minus[i - n_plus];
}

__global__ void __launch_bounds__(TPB) key_hist_kernel(const float* __restrict__ xv, int np, KeyArgs A, DevCounters* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= np) return;
  const float2* p = reinterpret_cast<const float2*>(xv) + 3 * i;
  const float2 a = p[0];
  key_one(A, i, a.x, a.y, p[1].x, cnt);
}

// ---- exclusive scan over NF fine-cell counts: block reduce -> scan of block sums -> block scan.
constexpr int SCAN_ITEMS = 16;                       // ints per thread
constexpr int SCAN_BLOCK = TPB * SCAN_ITEMS;         // 4096 ints per block

__device__ __forceinline__ int pair_sum(unsigned int w) { return (int)(w & 0xffffu) + (int)(w >> 16); }
// n (cells) is a multiple of 16: a scan block's SCAN_BLOCK cells are SCAN_BLOCK/2 words. One CTA reduces SCAN_RB consecutive scan blocks with
// all of its 2*SCAN_RB 16-byte loads in flight before the first add (two loads per thread left the kernel at half the DRAM bandwidth).
constexpr int SCAN_RB = 4;
__global__ void __launch_bounds__(TPB) scan_reduce_kernel(const unsigned int* __restrict__ hist, long long n, int* __restrict__ blocksum, int nb) {
  const int b0 = blockIdx.x * SCAN_RB;
  uint4 v[SCAN_RB][SCAN_ITEMS / 8];
#pragma unroll
  for (int q = 0; q < SCAN_RB; ++q) {
    const long long base = (long long)(b0 + q) * SCAN_BLOCK;
    const uint4* h4 = reinterpret_cast<const uint4*>(hist + (base >> 1));
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS / 8; ++it) {
      const long long e = base + ((long long)it * TPB + threadIdx.x) * 8;      // first cell of this thread's 4 words
      v[q][it] = (b0 + q < nb && e + 7 < n) ? h4[it * TPB + threadIdx.x] : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __shared__ int ws[SCAN_RB][TPB / 32];
#pragma unroll
  for (int q = 0; q < SCAN_RB; ++q) {
    int s = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS / 8; ++it) s += pair_sum(v[q][it].x) + pair_sum(v[q][it].y) + pair_sum(v[q][it].z) + pair_sum(v[q][it].w);
    s = warp_sum_i(s);
    if ((threadIdx.x & 31) == 0) ws[q][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < SCAN_RB && b0 + threadIdx.x < nb) {
    int t = 0;
    for (int w = 0; w < TPB / 32; ++w) t += ws[threadIdx.x][w];
    blocksum[b0 + threadIdx.x] = t;
  }
}

// single block: exclusive scan of the block sums in place; 8 consecutive entries per thread and pass (a 512^3-particle box has 300 k block sums:
// one entry per thread took 294 dependent passes = 0.27 ms)
constexpr int SBS_ITEMS = 8;
__global__ void __launch_bounds__(1024) scan_blocksums_kernel(int* __restrict__ blocksum, int nb) {
  __shared__ int ws[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024 * SBS_ITEMS) {
    const int i0 = base + threadIdx.x * SBS_ITEMS;
    int v[SBS_ITEMS];
    int s = 0;
#pragma unroll
    for (int q = 0; q < SBS_ITEMS; ++q) { v[q] = (i0 + q < nb) ? blocksum[i0 + q] : 0; s += v[q]; }
    int x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = ws[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
      ws[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    const int incl = x + warp_off;
    const int c = carry;
    int run = c + incl - s;
#pragma unroll
    for (int q = 0; q < SBS_ITEMS; ++q) { if (i0 + q < nb) blocksum[i0 + q] = run; run += v[q]; }
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
}

// block scan + offset; writes fstart (exclusive); also emits the PP work lists:
// physical fine cells (coarse cell in 1..nc_node on all axes) with >= 2 particles (PPINT) / >= 1 (PP_EXT).
// ONEPASS: single-pass scan with decoupled look-back (one read of the histogram instead of two, one launch instead of three). A CTA takes
// its tile from an atomic ticket (so every predecessor has started: spinning on them cannot deadlock), publishes its tile's AGGREGATE as soon as
// the local sums are known, looks back over the predecessors' status words (32 per step, one per lane) until it meets a PREFIX, and publishes
// its own inclusive PREFIX. status[t] = flag << 32 | value, flag 0 = empty (the array is cleared before the launch), 1 = aggregate, 2 = prefix;
// status[ntiles] is the ticket counter.
template <bool ONEPASS>
__global__ void __launch_bounds__(TPB) scan_apply_kernel(const unsigned int* __restrict__ hist_cur, long long n, const int* __restrict__ blocksum,
                                                         int* __restrict__ fstart, int H, int nc_buf, int nc_node,
                                                         int* __restrict__ multi_list, int* __restrict__ occ_list, int list_cap,
                                                         int want_multi, int want_occ, DevCounters* __restrict__ cnt,
                                                         unsigned long long* __restrict__ status, int ntiles) {
  __shared__ int s_tile, s_prefix;
  int tile = blockIdx.x;
  if (ONEPASS) {
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(status + ntiles, 1ULL);
    __syncthreads();
    tile = s_tile;
  }
  const long long base = (long long)tile * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q += 8) {       // 8 cells = 4 packed words (n is a multiple of 16)
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (base + q + 7 < n) t = *reinterpret_cast<const uint4*>(hist_cur + ((base + q) >> 1));
    v[q] = t.x & 0xffffu; v[q + 1] = t.x >> 16; v[q + 2] = t.y & 0xffffu; v[q + 3] = t.y >> 16;
    v[q + 4] = t.z & 0xffffu; v[q + 5] = t.z >> 16; v[q + 6] = t.w & 0xffffu; v[q + 7] = t.w >> 16;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[q + u];
  }
  // exclusive scan of s over the block
  __shared__ int ws[TPB / 32];
  int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = threadIdx.x < TPB / 32 ? ws[threadIdx.x] : 0;
#pragma unroll
    for (int o = 1; o < TPB / 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
    if (threadIdx.x < TPB / 32) ws[threadIdx.x] = w;
  }
  __syncthreads();
  int tile_prefix;
  if (ONEPASS) {
    if (threadIdx.x < 32) {
      const unsigned lane = threadIdx.x;
      const unsigned long long total = (unsigned long long)(unsigned)ws[TPB / 32 - 1];
      volatile unsigned long long* st = status;
      if (lane == 0) { __threadfence(); st[tile] = ((tile == 0 ? 2ULL : 1ULL) << 32) | total; }
      int excl = 0;
      if (tile > 0) {
        int j0 = tile - 1;                            // lane l inspects tile j0 - l
        for (;;) {
          const int j = j0 - (int)lane;
          unsigned long long w = 2ULL << 32;          // tiles before 0 count as a zero prefix
          if (j >= 0) { do { w = st[j]; } while ((w >> 32) == 0ULL); }
          const unsigned pm = __ballot_sync(0xffffffffu, (w >> 32) == 2ULL);   // lanes that found a prefix
          const int first = pm ? __ffs(pm) - 1 : 32;                             // nearest one
          int v = (int)lane <= first ? (int)(unsigned)(w & 0xffffffffULL) : 0;  // aggregates up to and including the nearest prefix
          v = warp_sum_i(v);
          excl += v;
          if (pm) break;
          j0 -= 32;
        }
        if (lane == 0) { __threadfence(); st[tile] = (2ULL << 32) | (unsigned long long)(unsigned)(excl + (int)total); }
      }
      if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    tile_prefix = s_prefix;
  } else {
    tile_prefix = blocksum[tile];
  }
  int run = tile_prefix + (x - s) + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0);
  // the 16 entries of one thread are a quarter of ONE coarse cell (64 fine cells): decode it once
  bool phys = false;
  if (base < n) {
    const unsigned int cc = (unsigned int)(base >> 6);
    const int cx = cc % H, cy = (cc / H) % H, cz = cc / (H * H);
    phys = cx >= nc_buf && cx < nc_buf + nc_node && cy >= nc_buf && cy < nc_buf + nc_node && cz >= nc_buf && cz < nc_buf + nc_node;
  }
  int o[SCAN_ITEMS];
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; ++q) { o[q] = run; run += v[q]; }
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q += 4) {
    if (base + q + 3 < n) {
      const int4 t = make_int4(o[q], o[q + 1], o[q + 2], o[q + 3]);
      *reinterpret_cast<int4*>(fstart + base + q) = t;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) if (base + q + u < n) fstart[base + q + u] = o[q + u];
    }
  }
  if (base + SCAN_ITEMS >= n && base < n) fstart[n] = run;   // total
  if (phys && s > 0) {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
      if (want_multi && v[q] >= 2) { const int slot = atomicAdd(&cnt->n_multi, 1); if (slot < list_cap) multi_list[slot] = (int)(base + q); }
      if (want_occ && v[q] >= 1) { const int slot = atomicAdd(&cnt->n_occ, 1); if (slot < list_cap) occ_list[slot] = (int)(base + q); }
    }
  }
}

// ---- PP_EXT margin roles (particle_mesh_threaded.f90:617; pp.cuh: ppext_margin_roles_kernel evaluates them). A (particle, tile) pair with the particle inside
// the tile's region [t*m - pr, (t+1)*m + pr) but outside the tile itself is a ROLE (up to 7 per particle, ~5 % of the particles have one).
struct MarginGeom { int H, b, m, T, pr; };

// tiles whose region [t*m - pr, (t+1)*m + pr) holds node-frame fine cell q: [tl, th] (empty if tl > th)
__device__ __forceinline__ void margin_tiles(int q, const MarginGeom& G, int& tl, int& th) {
  auto fdiv = [&](int v) { return v >= 0 ? v / G.m : -((-v + G.m - 1) / G.m); };
  tl = max(0, fdiv(q - G.pr)); th = min(G.T - 1, fdiv(q + G.pr));
}
// bit (dz*4 + dy*2 + dx) of the result = tile (tl + d) is a ROLE of the particle in hoc-frame fine cell g (in the region, not in the interior)
__device__ __forceinline__ unsigned margin_roles(const int g[3], const MarginGeom& G, int tl[3]) {
  int th[3];
  for (int ax = 0; ax < 3; ++ax) { margin_tiles(g[ax] - G.b, G, tl[ax], th[ax]); if (tl[ax] > th[ax]) return 0u; }
  unsigned mask = 0;
  for (int dz = 0; dz <= th[2] - tl[2]; ++dz)
    for (int dy = 0; dy <= th[1] - tl[1]; ++dy)
      for (int dx = 0; dx <= th[0] - tl[0]; ++dx) {
        const int t3[3] = {tl[0] + dx, tl[1] + dy, tl[2] + dz};
        bool interior = true;
        for (int ax = 0; ax < 3; ++ax) interior &= (g[ax] - G.b >= t3[ax] * G.m && g[ax] - G.b < (t3[ax] + 1) * G.m);
        if (!interior) mask |= 1u << (dz * 4 + dy * 2 + dx);
      }
  return mask;
}

// ROLES: the scatter already knows every particle's fine cell (its key) and its place in the sorted array, so it also appends the particle's margin roles
// (sorted index, tile) to the list the PP_EXT limiter works from — a separate listing kernel re-read all records (1.26 ms at 512^3).
// Clears the cell histogram for the next step. One warp per CTA and a grid of a few CTAs per SM: the kernel is meant to run on a side stream UNDER the
// issue-bound PP_EXT kernels (which leave the DRAM idle) without taking SM slots from them — cudaMemsetAsync's own kernel fills the SMs and cost PP_EXT
// the 0.5 ms it was supposed to hide.
__global__ void __launch_bounds__(32) zero_words_kernel(uint4* __restrict__ p, long long n16) {
  for (long long i = (long long)blockIdx.x * 32 + threadIdx.x; i < n16; i += (long long)gridDim.x * 32) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

template <bool ROLES>
__global__ void __launch_bounds__(TPB) scatter_kernel(const float* __restrict__ xv_in, const int64_t* __restrict__ pid_in,
                                                      const unsigned int* __restrict__ key, int np, const unsigned short* __restrict__ rank, unsigned int* __restrict__ hist,
                                                      const int* __restrict__ fstart, float* __restrict__ xv_out, int64_t* __restrict__ pid_out, int np_cap,
                                                      MarginGeom G, int2* __restrict__ roles, int role_cap, int* __restrict__ n_roles, int hist_mode) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  unsigned int k = KEY_DEAD;
  if (i < np) k = key[i];
  int dst = -1;
  if (k != KEY_DEAD) {
    float2 a, b, c;
    load_xv(xv_in, i, a, b, c);
    // slot = cell start + rank inside the cell (the count the histogram atomic of key_one returned): no atomics, the cell table is read-only here
    if (hist_mode == 2) {             // A/B: round 1's slot draw — a second atomic that counts the cell back to zero (no rank array read, no clearing pass)
      const unsigned sh = (k & 1u) << 4;
      dst = fstart[k] + (int)((atomicSub(&hist[k >> 1], 1u << sh) >> sh) & 0xffffu) - 1;
    } else {
      dst = fstart[k] + (int)rank[i];
      if (hist_mode == 1) hist[k >> 1] = 0u;     // A/B: clears the cell's word with a plain store (partial-sector writes: slow)
    }
    if ((unsigned)dst >= (unsigned)np_cap) dst = -1;   // only reachable after a 16-bit cell counter overflowed (flagged by key_hist_kernel, the step fails with EMAXLLF)
    else {
      store_xv(xv_out, dst, a, b, c);
      if (pid_in) pid_out[dst] = pid_in[i];
    }
  }
  if (ROLES) {
    unsigned mask = 0;
    int tl[3] = {0, 0, 0};
    if (dst >= 0) {
      const unsigned cc = k >> 6, f = k & 63u;
      const int H = G.H;
      const int g[3] = {(int)(4 * (cc % H) + (f & 3u)), (int)(4 * ((cc / H) % H) + ((f >> 2) & 3u)), (int)(4 * (cc / (H * H)) + (f >> 4))};
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
      if (base < role_cap) roles[base] = make_int2(dst, ((tl[2] + (bit >> 2)) * G.T + (tl[1] + ((bit >> 1) & 1))) * G.T + (tl[0] + (bit & 1)));
      ++base;
    }
  }
}

// ---- delete_particles.f90:14-50: keep particles with 0 <= x,y,z < mT, i.e. exactly those chained in coarse cells
// 1..nc_node. In the sorted array each physical (cy,cz) row is one contiguous range.
__global__ void __launch_bounds__(TPB) row_count_kernel(const int* __restrict__ fstart, int H, int nc_buf, int nc_node, int* __restrict__ rowlen) {
  const int r = blockIdx.x * TPB + threadIdx.x;
  if (r >= nc_node * nc_node) return;
  const int cy = r % nc_node + nc_buf, cz = r / nc_node + nc_buf;
  const long long k0 = ((long long)(cz * H + cy) * H + nc_buf) * 64;
  const long long k1 = k0 + (long long)nc_node * 64;
  rowlen[r] = fstart[k1] - fstart[k0];
}
// one CTA per physical row copies it to its compacted place; rowoff = exclusive scan of rowlen
__global__ void __launch_bounds__(TPB) compact_rows_kernel(const float* __restrict__ xv_in, const int64_t* __restrict__ pid_in,
                                                           const int* __restrict__ fstart, const int* __restrict__ rowoff, int H, int nc_buf,
                                                           int nc_node, float* __restrict__ xv_out, int64_t* __restrict__ pid_out) {
  const int r = blockIdx.x;
  const int cy = r % nc_node + nc_buf, cz = r / nc_node + nc_buf;
  const long long k0 = ((long long)(cz * H + cy) * H + nc_buf) * 64;
  const int s0 = fstart[k0], s1 = fstart[k0 + (long long)nc_node * 64];
  const int dst = rowoff[r];
  const float2* src2 = reinterpret_cast<const float2*>(xv_in) + 3LL * s0;
  float2* dst2 = reinterpret_cast<float2*>(xv_out) + 3LL * dst;
  const int nf2 = (s1 - s0) * 3;
  for (int q = threadIdx.x; q < nf2; q += TPB) dst2[q] = src2[q];
  if (pid_in) for (int q = threadIdx.x; q < s1 - s0; q += TPB) pid_out[dst + q] = pid_in[s0 + q];
}

// per coarse cell counts of the hoc range (parity getter): counts[c] = fstart[(c+1)*64] - fstart[c*64]
__global__ void __launch_bounds__(TPB) coarse_counts_kernel(const int* __restrict__ fstart, long long ncoarse, int* __restrict__ counts) {
  const long long c = (long long)blockIdx.x * TPB + threadIdx.x;
  if (c >= ncoarse) return;
  counts[c] = fstart[(c + 1) * 64] - fstart[c * 64];
}

}  // namespace part
