// TEST INFRASTRUCTURE — CPU oracle for CUBEP3M's `particle_mesh` step.
//
// A plain C++/OpenMP float32 restatement of the reference's Fortran (which cannot be compiled in this
// image: no gfortran / MPI / FFTW).  Compiled with -ffp-contract=off so that every expression is
// evaluated unfused, left to right, as the reference's x86-64 -O3 build does.  Each
// function cites the reference file:line it follows.  All D^3 MPI ranks of a run live in ONE process
// (struct World); every MPI exchange of the reference becomes a memcpy between rank states.
//
// PARITY STATUS: the reference ships no golden output vectors for this path and its FFT is an
// external library — so "parity unpinned" applies to the composite; the oracle is pinned by (1) the exact
// kernel tables kernels/wfxyzf.3.ascii + wfxyzc.2.ascii, (2) the analytic pairwise force law of
// report_pair.f90:50, (3) the mass / particle-count invariants printed under -DDIAG, (4) numpy.fft for the
// FFT definition.  See tests/test_oracle_*.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include "../include/cubep3m_b200.h"
#include "fft_ref.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
using oracle::cf;

inline int ifloor(float x) { return (int)std::floor(x); }

struct Params {
  cubep3m_b200_config c;
  int D, T, n, b, s, m, mT, nc_tile, nc_node, nc_dim, nc_slab, nc_buf, hoc_l, hoc_h, H, pass_depth;
  int nodes, tiles_node, max_np, max_buf;
  int Dg[3], Nc[3];   // rank grid (Dx,Dy,Dz) of cubic nodes and the global coarse mesh (Nx,Ny,Nz)
  void derive() {
    D = c.nodes_dim; T = c.tiles_node_dim; n = c.nf_tile; b = c.nf_buf; s = c.mesh_scale;
    m = n - 2 * b;                      // nf_physical_tile_dim   cubepm.par:197
    mT = m * T;                         // nf_physical_node_dim   cubepm.par:201
    nc_buf = b / s;                     // cubepm.par:190
    nc_tile = m / s;                    // cubepm.par:191
    nc_node = nc_tile * T;              // cubepm.par:192
    nc_dim = nc_node * D;               // cubepm.par:193
    const bool custom = c.nodes_dim_xyz[0] > 0 && c.nodes_dim_xyz[1] > 0 && c.nodes_dim_xyz[2] > 0;
    for (int a = 0; a < 3; ++a) { Dg[a] = custom ? c.nodes_dim_xyz[a] : D; Nc[a] = nc_node * Dg[a]; }
    nodes = Dg[0] * Dg[1] * Dg[2];
    nc_slab = (!custom && nc_dim % nodes == 0) ? nc_dim / nodes : 0;   // cubepm.par:195
    hoc_l = 1 - nc_buf;                 // cubepm.par:204
    hoc_h = nc_node + nc_buf;           // cubepm.par:205
    H = hoc_h - hoc_l + 1;
    pass_depth = 2 * nc_buf;            // cubepm.par:207
    tiles_node = T * T * T;
    if (c.max_np > 0) max_np = c.max_np;
    else {                              // cubepm.par:170-172
      double half = (double)(mT / 2);
      double buf = (8.0 * b * b * b + 6.0 * b * (double)mT * mT + 12.0 * (double)b * b * mT) / 8.0;
      max_np = (int)(c.density_buffer * (half * half * half + buf));
    }
    max_buf = c.max_buf > 0 ? c.max_buf : (int)(2.2 * (double)max_np);   // cubepm.par:175
  }
};

struct RankState {
  int rank = 0;
  int cc[3] = {0, 0, 0};     // cart_coords(1..3): cc[0] <-> z, cc[2] <-> x   (mpi_initialization.f90:55-64)
  int nb[6] = {0};           // cart_neighbor(1..6) = -z,+z,-y,+y,-x,+x       (mpi_initialization.f90:66-76)
  int np_local = 0;
  std::vector<float> xv;     // xv(6,max_np)                                  cubep3m.fh:75
  std::vector<int64_t> pid;  // PID(max_np)                                   cubep3m.fh:79
  std::vector<int> ll;       // ll(max_np)                                    cubep3m.fh:77
  std::vector<int> hoc;      // hoc(hoc_nc_l:hoc_nc_h)^3                      cubep3m.fh:78
  std::vector<float> rho_c;  // rho_c(nc_node^3)                              cubep3m.fh:58
  std::vector<float> force_c;// force_c(3,0:nc_node+1,...)                    cubep3m.fh:59
  std::vector<float> send_buf, recv_buf;
  std::vector<int64_t> send_pid, recv_pid;
  int np_deleted_ll = 0;
  // debug captures
  std::vector<int> tile_counts;
  std::vector<float> dbg_rho_f, dbg_force_f; int dbg_tile = -1;
};

struct World {
  Params p;
  std::vector<RankState> R;
  std::vector<float> fine_table, coarse_table;  // [k][j][i][c]
  std::vector<float> kern_f;                    // kern_f(3,n/2+1,n,n)        cubep3m.fh:35
  std::vector<float> kern_c;                    // global: (3,nc_dim/2+1,nc_dim,nc_dim); rank r owns z-planes [r*nc_slab,(r+1)*nc_slab)
  oracle::Fft3dR2C fft_f, fft_c;
  float stage_ms[CUBEP3M_B200_ST_COUNT] = {0};
  int np_buf_max = 0;
  int status = 0;
};

inline size_t hidx(const Params& p, int i, int j, int k) {
  return (size_t)(i - p.hoc_l) + (size_t)p.H * ((size_t)(j - p.hoc_l) + (size_t)p.H * (size_t)(k - p.hoc_l));
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------
// kernel_initialization.f90:2-267  fine_kernel
// ---------------------------------------------------------------------------------------------
void fine_kernel(World& w) {
  const Params& p = w.p;
  const int n = p.n, n2 = n + 2, hc = n / 2 + 1, nfc = p.c.nf_cutoff;
  w.kern_f.assign((size_t)3 * hc * n * n, 0.f);
  std::vector<float> rho((size_t)n2 * n * n);
  auto R = [&](int i, int j, int k) -> float& { return rho[(size_t)(i - 1) + (size_t)n2 * ((j - 1) + (size_t)n * (k - 1))]; };
  for (int d = 0; d < 3; ++d) {
    std::fill(rho.begin(), rho.end(), 0.f);
    for (int k = 1; k <= nfc; ++k)       // :25-36 read column 4+d
      for (int j = 1; j <= nfc; ++j)
        for (int i = 1; i <= nfc; ++i)
          R(i, j, k) = w.fine_table[(((size_t)(k - 1) * 16 + (j - 1)) * 16 + (i - 1)) * 3 + d];
    if (p.c.pp_ext && p.c.pp_ext_force_flag)  // :38-54
      for (int k = 1; k <= p.c.pp_range + 1; ++k)
        for (int j = 1; j <= p.c.pp_range + 1; ++j)
          for (int i = 1; i <= p.c.pp_range + 1; ++i) R(i, j, k) = 0.f;
    const float sy = (d == 1) ? -1.f : 1.f, sx = (d == 0) ? -1.f : 1.f, sz = (d == 2) ? -1.f : 1.f;
    for (int j = 2; j <= nfc; ++j)       // :71-73 reflect across y/2 over (1:nfc, ., 1:nfc)
      for (int k = 1; k <= nfc; ++k)
        for (int i = 1; i <= nfc; ++i) R(i, n - j + 2, k) = sy * R(i, j, k);
    for (int i = 2; i <= nfc; ++i)       // :76-79 reflect across x/2 over (., 1:n, 1:nfc)
      for (int k = 1; k <= nfc; ++k)
        for (int j = 1; j <= n; ++j) R(n - i + 2, j, k) = sx * R(i, j, k);
    for (int k = 2; k <= nfc; ++k)       // :82-85 reflect across z/2 over (1:n, 1:n, .)
      for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) R(i, j, n - k + 2) = sz * R(i, j, k);
    w.fft_f.forward(rho.data());          // :89
    for (int k = 1; k <= n; ++k)          // :93-99 imaginary part
      for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= hc; ++i)
          w.kern_f[(size_t)d + 3 * ((size_t)(i - 1) + (size_t)hc * ((j - 1) + (size_t)n * (k - 1)))] = R(2 * i, j, k);
  }
}

// ---------------------------------------------------------------------------------------------
// kernel_initialization.f90:272-732  coarse_kernel  (global formulation: all ranks' sub-cubes at once)
// ---------------------------------------------------------------------------------------------
void coarse_kernel(World& w) {
  const Params& p = w.p;
  const int Nx = p.Nc[0], Ny = p.Nc[1], Nz = p.Nc[2], N2 = Nx + 2, hc = Nx / 2 + 1;
  const float pi = 3.141592654f;          // cubepm.par:148
  w.kern_c.assign((size_t)3 * hc * Ny * Nz, 0.f);
  std::vector<float> ck((size_t)3 * Nx * Ny * Nz), ckc;
  auto CK = [&](std::vector<float>& a, int d, int i, int j, int k) -> float& {
    return a[(size_t)d + 3 * ((size_t)(i - 1) + (size_t)Nx * ((j - 1) + (size_t)Ny * (k - 1)))];
  };
  auto fill_plain = [&](std::vector<float>& a) {   // :302-336 and :479-513
    for (int k = 1; k <= Nz; ++k) {
      float z = (k < Nz / 2 + 2) ? (float)(k - 1) : (float)(k - 1 - Nz); z = p.s * z;
      for (int j = 1; j <= Ny; ++j) {
        float y = (j < Ny / 2 + 2) ? (float)(j - 1) : (float)(j - 1 - Ny); y = p.s * y;
        for (int i = 1; i <= Nx; ++i) {
          float x = (i < Nx / 2 + 2) ? (float)(i - 1) : (float)(i - 1 - Nx); x = p.s * x;
          float r = std::sqrt(x * x + y * y + z * z);
          if (r == 0.0f) { CK(a, 0, i, j, k) = 0.f; CK(a, 1, i, j, k) = 0.f; CK(a, 2, i, j, k) = 0.f; }
          else { float r3 = r * r * r; CK(a, 0, i, j, k) = -x / r3; CK(a, 1, i, j, k) = -y / r3; CK(a, 2, i, j, k) = -z / r3; }
        }
      }
    }
  };
  fill_plain(ck);
  // :344-457 overwrite the 4^3 near field in all 8 octants; component d is odd along axis d, even otherwise
  for (int oz = -3; oz <= 3; ++oz)
    for (int oy = -3; oy <= 3; ++oy)
      for (int ox = -3; ox <= 3; ++ox) {
        int i = ox >= 0 ? ox + 1 : Nx + ox + 1, j = oy >= 0 ? oy + 1 : Ny + oy + 1, k = oz >= 0 ? oz + 1 : Nz + oz + 1;
        if (i < 1 || j < 1 || k < 1 || i > Nx || j > Ny || k > Nz) continue;
        int oo[3] = {ox, oy, oz};
        for (int d = 0; d < 3; ++d) {
          float v = w.coarse_table[(((size_t)std::abs(oz) * 4 + std::abs(oy)) * 4 + std::abs(ox)) * 3 + d];
          CK(ck, d, i, j, k) = (oo[d] < 0) ? -v : v;
        }
      }
  std::vector<float> slab((size_t)N2 * Ny * Nz), tmp;
  auto load = [&](std::vector<float>& a, int d) {
    for (int k = 1; k <= Nz; ++k)
      for (int j = 1; j <= Ny; ++j) {
        float* row = &slab[(size_t)N2 * ((j - 1) + (size_t)Ny * (k - 1))];
        for (int i = 1; i <= Nx; ++i) row[i - 1] = CK(a, d, i, j, k);
        row[Nx] = row[Nx + 1] = 0.f;
      }
  };
  auto store = [&](int d) {                                          // :593-599
    for (int k = 1; k <= Nz; ++k)
      for (int j = 1; j <= Ny; ++j)
        for (int i = 1; i <= hc; ++i)
          w.kern_c[(size_t)d + 3 * ((size_t)(i - 1) + (size_t)hc * ((j - 1) + (size_t)Ny * (k - 1)))] =
              slab[(size_t)(2 * i - 1) + (size_t)N2 * ((j - 1) + (size_t)Ny * (k - 1))];
  };
  if (p.c.lrckcorr) {
    ckc = ck;                 // :469-475 corrected kernel stashed in force_c
    fill_plain(ck);           // :479-513 uncorrected
    for (int d = 0; d < 3; ++d) {
      load(ck, d); w.fft_c.forward(slab.data()); tmp = slab;          // :519-551 tmp_kern_c(d)
      load(ckc, d); w.fft_c.forward(slab.data());                      // :556-557
      for (int k = 1; k <= Nz; ++k) {                                  // :558-591 (x), :602-635 (y), :646-679 (z)
        int kz = (k < Nz / 2 + 2) ? k - 1 : k - 1 - Nz;
        for (int j = 1; j <= Ny; ++j) {
          int ky = (j < Ny / 2 + 2) ? j - 1 : j - 1 - Ny;
          for (int i = 1; i <= Nx + 2; i += 2) {
            int kx = (i - 1) / 2;
            float kr = std::sqrt((float)(kx * kx + ky * ky + kz * kz));
            if (kr <= 8.f) {
              float ka = 2 * std::sin(pi * kx / (float)Nx);
              float kb = 2 * std::sin(pi * ky / (float)Ny);
              float kc = 2 * std::sin(pi * kz / (float)Nz);
              int kd = d == 0 ? kx : (d == 1 ? ky : kz);
              float kk = d == 0 ? ka : (d == 1 ? kb : kc);
              if (kd != 0) {
                size_t o2 = (size_t)i + (size_t)N2 * ((j - 1) + (size_t)Ny * (k - 1));  // slab(i+1,j,k), 0-based
                float wa = slab[o2], wb = tmp[o2];
                float wc = 4.f * pi * kk / (ka * ka + kb * kb + kc * kc) / 16.f;
                slab[o2] = wa * (wc / wb);
              }
            }
          }
        }
      }
      store(d);
    }
  } else {
    for (int d = 0; d < 3; ++d) { load(ck, d); w.fft_c.forward(slab.data()); store(d); }   // :695-723
  }
}

// ---------------------------------------------------------------------------------------------
// update_position.f90:66-76
// ---------------------------------------------------------------------------------------------
void update_position(World& w, float dt, float dt_old, const float offset[3]) {
  double t0 = now_ms();
  for (auto& r : w.R) {
    float* xv = r.xv.data();
    const int np = r.np_local;
    const float hdt = dt + dt_old;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < np; ++i) {
      float* q = xv + (size_t)6 * i;
      q[0] = (q[0] + ((q[3] * 0.5f) * hdt)) + offset[0];
      q[1] = (q[1] + ((q[4] * 0.5f) * hdt)) + offset[1];
      q[2] = (q[2] + ((q[5] * 0.5f) * hdt)) + offset[2];
    }
  }
  w.stage_ms[CUBEP3M_B200_ST_DRIFT] += (float)(now_ms() - t0);
}

// move_grid_back.f90:20-23
void move_grid_back(World& w, const float shake[3]) {
  for (auto& r : w.R)
    for (int i = 0; i < r.np_local; ++i)
      for (int d = 0; d < 3; ++d) r.xv[(size_t)6 * i + d] = r.xv[(size_t)6 * i + d] - shake[d];
}

// ---------------------------------------------------------------------------------------------
// link_list.f90:19-53  (serial, LIFO chains, swap-delete of out-of-range particles)
// ---------------------------------------------------------------------------------------------
void link_list(World& w) {
  double t0 = now_ms();
  const Params& p = w.p;
  for (auto& r : w.R) {
    std::fill(r.hoc.begin(), r.hoc.end(), 0);
    int np_buf = 0;
    int pp = 1;
    float* xv = r.xv.data();
    while (pp <= r.np_local) {
      float* q = xv + (size_t)6 * (pp - 1);
      int i = ifloor(q[0] / (float)p.s) + 1, j = ifloor(q[1] / (float)p.s) + 1, k = ifloor(q[2] / (float)p.s) + 1;
      if (i < p.hoc_l || i > p.hoc_h || j < p.hoc_l || j > p.hoc_h || k < p.hoc_l || k > p.hoc_h) {
        float* last = xv + (size_t)6 * (r.np_local - 1);
        for (int c = 0; c < 6; ++c) q[c] = last[c];
        if (p.c.pid) r.pid[pp - 1] = r.pid[r.np_local - 1];
        r.np_local--; np_buf++;
        continue;
      }
      size_t h = hidx(p, i, j, k);
      r.ll[pp - 1] = r.hoc[h];
      r.hoc[h] = pp;
      pp++;
    }
    r.np_deleted_ll = np_buf;
  }
  w.stage_ms[CUBEP3M_B200_ST_LINK] += (float)(now_ms() - t0);
}

// relink the particles appended by one axis of particle_pass (particle_pass.f90:274-298, 490-518, 694-722)
void relink_tail(const Params& p, RankState& r, int first_pp) {
  int pp = first_pp;
  float* xv = r.xv.data();
  while (pp <= r.np_local) {
    float* q = xv + (size_t)6 * (pp - 1);
    int i = ifloor(q[0] / (float)p.s) + 1, j = ifloor(q[1] / (float)p.s) + 1, k = ifloor(q[2] / (float)p.s) + 1;
    if (i < p.hoc_l || i > p.hoc_h || j < p.hoc_l || j > p.hoc_h || k < p.hoc_l || k > p.hoc_h) {
      float* last = xv + (size_t)6 * (r.np_local - 1);
      for (int c = 0; c < 6; ++c) q[c] = last[c];
      if (p.c.pid) r.pid[pp - 1] = r.pid[r.np_local - 1];
      r.np_local--;
      continue;
    }
    size_t h = hidx(p, i, j, k);
    r.ll[pp - 1] = r.hoc[h];
    r.hoc[h] = pp;
    pp++;
  }
}

// One directional pass for every rank: pack (loops of particle_pass.f90:73-94 etc.), exchange, shift/clamp, append.
// axis 0/1/2 = x/y/z ; plus = true for the "+" pass (particles with x >= mT - nf_buf go to the + neighbour).
int pass_dir(World& w, int axis, bool plus) {
  const Params& p = w.p;
  const float rnf_buf = (float)p.b;
  const float hi_cut = (float)p.mT - rnf_buf;   // nf_physical_node_dim - rnf_buf
  std::vector<int> nsend(w.R.size());
  for (auto& r : w.R) {
    int np_buf = 0;
    int lo[3] = {p.hoc_l, p.hoc_l, p.hoc_l}, hi[3] = {p.hoc_h, p.hoc_h, p.hoc_h};
    if (plus) lo[axis] = p.hoc_h - p.pass_depth; else hi[axis] = p.hoc_l + p.pass_depth;
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          int pp = r.hoc[hidx(p, i, j, k)];
          while (pp != 0) {
            const float* q = &r.xv[(size_t)6 * (pp - 1)];
            bool go = plus ? (q[axis] >= hi_cut) : (q[axis] < rnf_buf);
            if (go) {
              if ((size_t)(np_buf + 1) * 6 > r.send_buf.size()) return CUBEP3M_B200_EPASSBUF;
              for (int c = 0; c < 6; ++c) r.send_buf[(size_t)6 * np_buf + c] = q[c];
              if (p.c.pid) r.send_pid[np_buf] = r.pid[pp - 1];
              np_buf++;
            }
            pp = r.ll[pp - 1];
          }
        }
    if (np_buf * 6 > p.max_buf) return CUBEP3M_B200_EPASSBUF;       // :96-99
    nsend[r.rank] = np_buf;
    w.np_buf_max = std::max(w.np_buf_max, np_buf);
  }
  // exchange: "+" pass sends to cart_neighbor(2*(3-axis)) [+side], receives from the - side.
  for (auto& r : w.R) {
    // neighbour indices: nb[0..5] = -z,+z,-y,+y,-x,+x
    int minus_nb = r.nb[2 * (2 - axis)], plus_nb = r.nb[2 * (2 - axis) + 1];
    int src = plus ? minus_nb : plus_nb;
    RankState& s = w.R[src];
    int nrecv = nsend[src];
    if (r.np_local + nrecv > p.max_np) return CUBEP3M_B200_EMAXNP;   // :136-139
    std::copy(s.send_buf.begin(), s.send_buf.begin() + (size_t)6 * nrecv, r.recv_buf.begin());
    if (p.c.pid) std::copy(s.send_pid.begin(), s.send_pid.begin() + nrecv, r.recv_pid.begin());
  }
  for (auto& r : w.R) {
    int minus_nb = r.nb[2 * (2 - axis)], plus_nb = r.nb[2 * (2 - axis) + 1];
    int src = plus ? minus_nb : plus_nb;
    int nrecv = nsend[src];
    const float fmT = (float)p.mT;
    for (int i = 0; i < nrecv; ++i) {
      float* q = &r.xv[(size_t)6 * (r.np_local + i)];
      for (int c = 0; c < 6; ++c) q[c] = r.recv_buf[(size_t)6 * i + c];
      if (p.c.pid) r.pid[r.np_local + i] = r.recv_pid[i];
      if (plus) {
        q[axis] = std::max(q[axis] - fmT, -rnf_buf);                 // :162
      } else {
        if (std::fabs(q[axis]) < p.c.eps) q[axis] = (q[axis] < 0.0f) ? -p.c.eps : p.c.eps;   // :257-263
        q[axis] = std::min(q[axis] + fmT, (fmT + rnf_buf) - p.c.eps);                            // :264-265
      }
    }
    r.np_local += nrecv;
  }
  return 0;
}

// particle_pass.f90:69-722 — order +x,-x,relink, -y,+y,relink, +z,-z,relink
int particle_pass(World& w) {
  double t0 = now_ms();
  w.np_buf_max = 0;
  const bool first_plus[3] = {true, false, true};
  for (int axis = 0; axis < 3; ++axis) {
    std::vector<int> np0(w.R.size());
    for (auto& r : w.R) np0[r.rank] = r.np_local;
    int st = pass_dir(w, axis, first_plus[axis]);
    if (st) return st;
    st = pass_dir(w, axis, !first_plus[axis]);
    if (st) return st;
    for (auto& r : w.R) relink_tail(w.p, r, np0[r.rank] + 1);
  }
  w.stage_ms[CUBEP3M_B200_ST_PASS] += (float)(now_ms() - t0);
  return 0;
}

// delete_particles.f90:14-50
void delete_particles(World& w) {
  double t0 = now_ms();
  const float lim = (float)w.p.mT;
  for (auto& r : w.R) {
    int pp = 1;
    float* xv = r.xv.data();
    while (pp <= r.np_local) {
      float* q = xv + (size_t)6 * (pp - 1);
      if (q[0] >= lim || q[0] < 0.0f || q[1] >= lim || q[1] < 0.0f || q[2] >= lim || q[2] < 0.0f) {
        float* last = xv + (size_t)6 * (r.np_local - 1);
        for (int c = 0; c < 6; ++c) q[c] = last[c];
        if (w.p.c.pid) r.pid[pp - 1] = r.pid[r.np_local - 1];
        r.np_local--;
        continue;
      }
      pp++;
    }
  }
  w.stage_ms[CUBEP3M_B200_ST_DELETE] += (float)(now_ms() - t0);
}

// ---------------------------------------------------------------------------------------------
// The tile loop of particle_mesh_threaded.f90:72-628
// ---------------------------------------------------------------------------------------------
struct TileScratch {
  std::vector<float> rho_f, cmplx_rho_f, force_f;
  std::vector<int> llf;            // llf(max_llf,4,4,4) grown on demand
  std::vector<float> pp_force_accum;
  std::vector<int> hoc_fine, ll_fine, tile_pp;
  std::vector<float> pp_ext_force_accum;
};

struct StepScalars {
  float f_force_max = 0.f, pp_force_max = 0.f, pp_ext_force_max = 0.f;
  double f_mesh_mass = 0.0;
  int status = 0;
};

void fine_tile(World& w, RankState& r, int cur_tile, float a_mid, float dt, float mass_p, TileScratch& S, StepScalars& out,
               double* t_dep, double* t_fft, double* t_kick, double* t_pp, double* t_ppext) {
  const Params& p = w.p;
  const int n = p.n, n2 = n + 2, hc = n / 2 + 1, b = p.b, m = p.m, T = p.T, s = p.s;
  const float G = p.c.G;
  const int fdim = m + 3;             // force_f spans nf_buf-1 : nf_tile-nf_buf+1  (cubep3m.fh:36-37)
  int tile[3];
  tile[2] = (cur_tile - 1) / (T * T);                         // :86-90
  int j0 = cur_tile - tile[2] * T * T;
  tile[1] = (j0 - 1) / T;
  j0 = j0 - tile[1] * T;
  tile[0] = j0 - 1;
  S.rho_f.assign((size_t)n2 * n * n, 0.f);                    // :100
  S.force_f.resize((size_t)3 * fdim * fdim * fdim);
  auto RHO = [&](int i, int j, int k) -> float& { return S.rho_f[(size_t)(i - 1) + (size_t)n2 * ((j - 1) + (size_t)n * (k - 1))]; };
  auto FF = [&](int c, int i, int j, int k) -> float& {
    return S.force_f[(size_t)c + 3 * ((size_t)(i - (b - 1)) + (size_t)fdim * ((j - (b - 1)) + (size_t)fdim * (k - (b - 1))))];
  };
  double t0 = now_ms();
  int cic_l[3], cic_h[3];
  for (int d = 0; d < 3; ++d) {
    if (p.c.ngp) { cic_l[d] = p.nc_tile * tile[d] + 2 - p.nc_buf; cic_h[d] = p.nc_tile * (tile[d] + 1) + p.nc_buf - 1; }   // :120-121
    else         { cic_l[d] = p.nc_tile * tile[d] + 1 - p.nc_buf; cic_h[d] = p.nc_tile * (tile[d] + 1) + p.nc_buf; }       // :123-124
  }
  float offset[3];
  for (int d = 0; d < 3; ++d) offset[d] = (float)(-tile[d] * m + b);   // :134
  int ndep = 0;
  for (int k = cic_l[2]; k <= cic_h[2]; ++k)
    for (int j = cic_l[1]; j <= cic_h[1]; ++j)
      for (int i = cic_l[0]; i <= cic_h[0]; ++i) {
        int pp = r.hoc[hidx(p, i, j, k)];
        if (p.c.ngp) {
          while (pp != 0) {                                           // :138-151
            const float* q = &r.xv[(size_t)6 * (pp - 1)];
            float x0 = q[0] + offset[0], x1 = q[1] + offset[1], x2 = q[2] + offset[2];
            int i1 = ifloor(x0) + 1, i2 = ifloor(x1) + 1, i3 = ifloor(x2) + 1;
            RHO(i1, i2, i3) = RHO(i1, i2, i3) + mass_p;
            ++ndep;
            pp = r.ll[pp - 1];
          }
        } else {
          const bool bnd = (i == cic_l[0] || i == cic_h[0] || j == cic_l[1] || j == cic_h[1] || k == cic_l[2] || k == cic_h[2]);  // :154-160
          while (pp != 0) {                                           // fine_cic_mass.f90:12-42 / fine_cic_mass_buffer.f90
            const float* q = &r.xv[(size_t)6 * (pp - 1)];
            float x[3], dx1[3], dx2[3]; int i1[3], i2[3];
            for (int d = 0; d < 3; ++d) {
              x[d] = q[d] + offset[d]; i1[d] = ifloor(x[d]) + 1; i2[d] = i1[d] + 1;
              dx1[d] = (float)i1[d] - x[d]; dx2[d] = 1.f - dx1[d];
            }
            dx1[0] = mass_p * dx1[0]; dx2[0] = mass_p * dx2[0];
            ++ndep;
            for (int cz = 0; cz < 2; ++cz)
              for (int cy = 0; cy < 2; ++cy)
                for (int cx = 0; cx < 2; ++cx) {
                  int ix = cx ? i2[0] : i1[0], iy = cy ? i2[1] : i1[1], iz = cz ? i2[2] : i1[2];
                  if (bnd && (ix < 1 || ix > n || iy < 1 || iy > n || iz < 1 || iz > n)) continue;
                  float wgt = ((cx ? dx2[0] : dx1[0]) * (cy ? dx2[1] : dx1[1])) * (cz ? dx2[2] : dx1[2]);
                  RHO(ix, iy, iz) = RHO(ix, iy, iz) + wgt;
                }
            pp = r.ll[pp - 1];
          }
        }
      }
  r.tile_counts[cur_tile - 1] = ndep;
  for (int k = 1 + b; k <= n - b; ++k)                               // :167-173
    for (int j = 1 + b; j <= n - b; ++j)
      for (int i = 1 + b; i <= n - b; ++i) out.f_mesh_mass += (double)RHO(i, j, k);
  if (r.dbg_tile == cur_tile) r.dbg_rho_f = S.rho_f;
  double t1 = now_ms(); *t_dep += t1 - t0;

  w.fft_f.forward(S.rho_f.data());                                   // :176
  S.cmplx_rho_f = S.rho_f;                                           // :180
  const float n3 = ((float)n * (float)n) * (float)n;                 // real(nf_tile)**3  fft_fine.f90:51
  for (int i3 = 0; i3 < 3; ++i3) {
    for (int k = 1; k <= n; ++k)                                     // :183-192
      for (int j = 1; j <= n; ++j) {
        const size_t rowo = (size_t)n2 * ((j - 1) + (size_t)n * (k - 1));
        const float* kf = &w.kern_f[(size_t)3 * ((size_t)hc * ((j - 1) + (size_t)n * (k - 1)))];
        for (int i = 1; i <= hc; ++i) {
          float kv = kf[(size_t)3 * (i - 1) + i3];
          float re = S.cmplx_rho_f[rowo + 2 * i - 2], im = S.cmplx_rho_f[rowo + 2 * i - 1];
          S.rho_f[rowo + 2 * i - 2] = -im * kv;
          S.rho_f[rowo + 2 * i - 1] = re * kv;
        }
      }
    w.fft_f.backward(S.rho_f.data());                                // :197
    for (size_t q = 0; q < S.rho_f.size(); ++q) S.rho_f[q] = S.rho_f[q] / n3;
    for (int k = b - 1; k <= n - b + 1; ++k)                         // :202-203
      for (int j = b - 1; j <= n - b + 1; ++j)
        for (int i = b - 1; i <= n - b + 1; ++i) FF(i3, i, j, k) = RHO(i, j, k);
  }
  for (int k = b - 1; k <= n - b + 1; ++k)                           // :208-223
    for (int j = b - 1; j <= n - b + 1; ++j)
      for (int i = b - 1; i <= n - b + 1; ++i) {
        float fx = FF(0, i, j, k), fy = FF(1, i, j, k), fz = FF(2, i, j, k);
        float force_mag = (fx * fx + fy * fy) + fz * fz;
        if (force_mag > out.f_force_max) out.f_force_max = force_mag;
      }
  if (r.dbg_tile == cur_tile) r.dbg_force_f = S.force_f;
  double t2 = now_ms(); *t_fft += t2 - t1;

  // ---- velocity update + intra-cell PP  (:227-368)
  for (int d = 0; d < 3; ++d) offset[d] = (float)b - (float)(tile[d] * m);   // :227
  const int max_llf = p.c.max_llf;
  int ipl[4][4][4];
  double tpp_local = 0.0;
  for (int k = tile[2] * p.nc_tile + 1; k <= (tile[2] + 1) * p.nc_tile; ++k)
    for (int j = tile[1] * p.nc_tile + 1; j <= (tile[1] + 1) * p.nc_tile; ++j)
      for (int i = tile[0] * p.nc_tile + 1; i <= (tile[0] + 1) * p.nc_tile; ++i) {
        int pp = r.hoc[hidx(p, i, j, k)];
        if (pp == 0) continue;
        if (p.c.ppint) std::memset(ipl, 0, sizeof(ipl));
        int chain = 0;
        while (pp != 0) {
          float* q = &r.xv[(size_t)6 * (pp - 1)];
          float x[3]; int i1[3];
          for (int d = 0; d < 3; ++d) { x[d] = q[d] + offset[d]; i1[d] = ifloor(x[d]) + 1; }
          if (p.c.ngp) {
            if (p.c.ngp_fmesh_force)                                  // :265-266
              for (int d = 0; d < 3; ++d) q[3 + d] = q[3 + d] + ((FF(d, i1[0], i1[1], i1[2]) * a_mid) * G) * dt;
            if (p.c.ppint) {                                          // :276-284
              int i2[3];
              for (int d = 0; d < 3; ++d) i2[d] = ((i1[d] - 1) % s) + 1;
              int& cnt = ipl[i2[2] - 1][i2[1] - 1][i2[0] - 1];
              cnt++;
              if (cnt > max_llf) { out.status = CUBEP3M_B200_EMAXLLF; return; }   // :280-283
            }
          } else {                                                    // :289-316 CIC interpolation
            int i2[3]; float dx1[3], dx2[3];
            for (int d = 0; d < 3; ++d) { i2[d] = i1[d] + 1; dx1[d] = (float)i1[d] - x[d]; dx2[d] = 1.0f - dx1[d]; }
            for (int cz = 0; cz < 2; ++cz)
              for (int cy = 0; cy < 2; ++cy)
                for (int cx = 0; cx < 2; ++cx) {
                  float dVc = ((((a_mid * G) * dt) * (cx ? dx2[0] : dx1[0])) * (cy ? dx2[1] : dx1[1])) * (cz ? dx2[2] : dx1[2]);
                  int ix = cx ? i2[0] : i1[0], iy = cy ? i2[1] : i1[1], iz = cz ? i2[2] : i1[2];
                  for (int d = 0; d < 3; ++d) q[3 + d] = q[3 + d] + FF(d, ix, iy, iz) * dVc;
                }
          }
          chain++;
          pp = r.ll[pp - 1];
        }
        if (p.c.ngp && p.c.ppint) {
          double tq = now_ms();
          // rebuild llf in chain order (second walk; same order as the reference's single walk)
          if ((int)S.llf.size() < 64 * chain) S.llf.resize((size_t)64 * chain);
          int cnts[64]; std::memset(cnts, 0, sizeof(cnts));
          pp = r.hoc[hidx(p, i, j, k)];
          while (pp != 0) {
            const float* q = &r.xv[(size_t)6 * (pp - 1)];
            int i2[3];
            for (int d = 0; d < 3; ++d) { int i1 = ifloor(q[d] + offset[d]) + 1; i2[d] = ((i1 - 1) % s); }
            int cell = i2[0] + 4 * (i2[1] + 4 * i2[2]);
            S.llf[(size_t)cell * chain + cnts[cell]] = pp;
            cnts[cell]++;
            pp = r.ll[pp - 1];
          }
          for (int km = 0; km < 4; ++km)                              // :324-361
            for (int jm = 0; jm < 4; ++jm)
              for (int im = 0; im < 4; ++im) {
                int cell = im + 4 * (jm + 4 * km);
                int np_c = cnts[cell];
                if (np_c == 0) continue;
                if ((int)S.pp_force_accum.size() < 3 * np_c) S.pp_force_accum.resize((size_t)3 * np_c);
                std::fill(S.pp_force_accum.begin(), S.pp_force_accum.begin() + 3 * np_c, 0.f);
                for (int ip = 0; ip < np_c - 1; ++ip) {
                  int pp1 = S.llf[(size_t)cell * chain + ip];
                  for (int jp = ip + 1; jp < np_c; ++jp) {
                    int pp2 = S.llf[(size_t)cell * chain + jp];
                    float* q1 = &r.xv[(size_t)6 * (pp1 - 1)];
                    float* q2 = &r.xv[(size_t)6 * (pp2 - 1)];
                    float sep[3] = {q1[0] - q2[0], q1[1] - q2[1], q1[2] - q2[2]};
                    float rmag = std::sqrt((sep[0] * sep[0] + sep[1] * sep[1]) + sep[2] * sep[2]);
                    if (rmag > p.c.rsoft) {
                      float rb = rmag * p.c.pp_bias; float rb3 = (rb * rb) * rb;
                      for (int d = 0; d < 3; ++d) {
                        float force_pp = mass_p * (sep[d] / rb3);     // :344
                        S.pp_force_accum[3 * ip + d] = S.pp_force_accum[3 * ip + d] - force_pp;
                        S.pp_force_accum[3 * jp + d] = S.pp_force_accum[3 * jp + d] + force_pp;
                        if (p.c.pp_force_flag) {                      // :349-350
                          q1[3 + d] = q1[3 + d] - ((force_pp * a_mid) * G) * dt;
                          q2[3 + d] = q2[3 + d] + ((force_pp * a_mid) * G) * dt;
                        }
                      }
                    }
                  }
                }
                for (int ip = 0; ip < np_c; ++ip) {                   // :355-358
                  float* f = &S.pp_force_accum[3 * ip];
                  float mag = std::sqrt((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
                  if (mag > out.pp_force_max) out.pp_force_max = mag;
                }
              }
          tpp_local += now_ms() - tq;
        }
      }
  double t3 = now_ms(); *t_kick += (t3 - t2) - tpp_local; *t_pp += tpp_local;

  // ---- extended PP (:378-624)
  if (p.c.pp_ext) {
    const int pr = p.c.pp_range;
    const int fd = m + 2 * pr;
    S.hoc_fine.assign((size_t)fd * fd * fd, 0);                        // :393
    int fl[3], fh[3];
    for (int d = 0; d < 3; ++d) { fl[d] = tile[d] * m + 1 - pr; fh[d] = (tile[d] + 1) * m + pr; }   // :397-402
    auto HF = [&](int i, int j, int k) -> int& { return S.hoc_fine[(size_t)(i - 1) + (size_t)fd * ((j - 1) + (size_t)fd * (k - 1))]; };
    // The reference keeps ll_fine(max_np) and pp_ext_force_accum(3,max_np) per thread and clears / scans all of them for every tile
    // (:393, :491, :617). Same scan over all np_local particles (:410-438), same chain order, same sums and maximum here, but the per-thread
    // arrays are indexed by the particle's position in the tile's own list (S.tile_pp), so that a 512^3 box does not need 2.4 GB per thread.
    S.tile_pp.clear();
    S.ll_fine.clear();
    for (int pp = 1; pp <= r.np_local; ++pp) {                         // :410-438
      const float* q = &r.xv[(size_t)6 * (pp - 1)];
      int i = ifloor(q[0]) + 1, j = ifloor(q[1]) + 1, k = ifloor(q[2]) + 1;
      if (i < fl[0] || i > fh[0] || j < fl[1] || j > fh[1] || k < fl[2] || k > fh[2]) continue;
      int& h = HF(i - fl[0] + 1, j - fl[1] + 1, k - fl[2] + 1);
      S.tile_pp.push_back(pp);
      S.ll_fine.push_back(h); h = (int)S.tile_pp.size();               // 1-based index into the tile's list
    }
    S.pp_ext_force_accum.assign((size_t)3 * S.tile_pp.size(), 0.f);    // :491
    const float cut = (float)p.c.nf_cutoff;
    if (pr != 0) {
      for (int k = 1; k <= m + pr; ++k)                                // :496
        for (int j = 1; j <= m + 2 * pr; ++j)
          for (int i = 1; i <= m + 2 * pr; ++i) {
            int pp1h = HF(i, j, k);
            if (pp1h == 0) continue;
            const bool phys1 = (pr < i && i <= m + pr && pr < j && j <= m + pr && pr < k && k <= m + pr);   // :576-578
            for (int kp = k; kp <= k + pr; ++kp) {                     // :503-523
              int jp_min = (kp == k) ? j : std::max(j - pr, 1);
              int jp_max = std::min(j + pr, m + 2 * pr);
              for (int jp = jp_min; jp <= jp_max; ++jp) {
                int ip_min = (kp == k && jp == j) ? i + 1 : std::max(i - pr, 1);
                int ip_max = std::min(i + pr, m + 2 * pr);
                for (int ip = ip_min; ip <= ip_max; ++ip) {
                  int pp2h = HF(ip, jp, kp);
                  if (pp2h == 0) continue;
                  const bool phys2 = (pr < ip && ip <= m + pr && pr < jp && jp <= m + pr && pr < kp && kp <= m + pr);  // :584-586
                  for (int l1 = pp1h; l1 != 0; l1 = S.ll_fine[l1 - 1])
                    for (int l2 = pp2h; l2 != 0; l2 = S.ll_fine[l2 - 1]) {
                      float* q1 = &r.xv[(size_t)6 * (S.tile_pp[l1 - 1] - 1)];
                      float* q2 = &r.xv[(size_t)6 * (S.tile_pp[l2 - 1] - 1)];
                      float sep[3] = {q1[0] - q2[0], q1[1] - q2[1], q1[2] - q2[2]};
                      float rmag = std::sqrt((sep[0] * sep[0] + sep[1] * sep[1]) + sep[2] * sep[2]);
                      if (rmag > p.c.rsoft) {                          // :558-564
                        float rb = rmag * p.c.pp_bias; float rb3 = (rb * rb) * rb;
                        float poly = 1.f;
                        const bool far = rmag > cut + std::sqrt(3.0f);
                        if (!far) {
                          float u = rb / cut; float u3 = (u * u) * u; float u5 = (((u * u) * u) * u) * u;
                          poly = (1.f - (7.0f / 4.0f) * u3) + (3.0f / 4.0f) * u5;
                        }
                        for (int d = 0; d < 3; ++d) {
                          float force_pp = mass_p * (sep[d] / rb3);
                          if (!far) force_pp = force_pp * poly;
                          S.pp_ext_force_accum[(size_t)3 * (l1 - 1) + d] -= force_pp;
                          S.pp_ext_force_accum[(size_t)3 * (l2 - 1) + d] += force_pp;
                          if (p.c.pp_ext_force_flag) {
                            if (phys1) q1[3 + d] = q1[3 + d] - ((force_pp * a_mid) * G) * dt;
                            if (phys2) q2[3 + d] = q2[3 + d] + ((force_pp * a_mid) * G) * dt;
                          }
                        }
                      }
                    }
                }
              }
            }
          }
    }
    float mx = 0.f;                                                    // :617 (every other particle's accumulator is zero)
    for (size_t l = 0; l < S.tile_pp.size(); ++l) {
      const float* f = &S.pp_ext_force_accum[(size_t)3 * l];
      float mag = std::sqrt((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
      if (mag > mx) mx = mag;
    }
    // NOTE: the reference overwrites pp_ext_force_max(thread) per tile; we keep the max over tiles,
    // which is what maxval over threads sees when every thread ran >= 1 tile last with the largest value.
    if (mx > out.pp_ext_force_max) out.pp_ext_force_max = mx;
  }
  *t_ppext += now_ms() - t3;
}

// ---------------------------------------------------------------------------------------------
// coarse mesh: coarse_mass.f90:82-99 + coarse_cic_mass(_buffer).f90, coarse_force.f90:18-90 (global FFT = pack_slab +
// distributed FFT + unpack_slab of fft_coarse.f90), coarse_force_buffer.f90:23-63, coarse_max_dt.f90:19-37,
// coarse_velocity.f90:137-179
// ---------------------------------------------------------------------------------------------
void coarse_mass(World& w, float mass_p) {
  double t0 = now_ms();
  const Params& p = w.p;
  const int nc = p.nc_node;
  for (auto& r : w.R) {
    std::fill(r.rho_c.begin(), r.rho_c.end(), 0.f);
    auto RC = [&](int i, int j, int k) -> float& { return r.rho_c[(size_t)(i - 1) + (size_t)nc * ((j - 1) + (size_t)nc * (k - 1))]; };
    for (int k0 = 0; k0 < p.s; ++k0)
      for (int k = k0; k <= nc + 1; k += p.s)
        for (int j = 0; j <= nc + 1; ++j)
          for (int i = 0; i <= nc + 1; ++i) {
            int pp = r.hoc[hidx(p, i, j, k)];
            while (pp != 0) {
              const float* q = &r.xv[(size_t)6 * (pp - 1)];
              float x[3], dx1[3], dx2[3]; int i1[3], i2[3];
              for (int d = 0; d < 3; ++d) {
                x[d] = (1.0f / (float)p.s) * q[d] - 0.5f;
                i1[d] = ifloor(x[d]) + 1; i2[d] = i1[d] + 1;
                if (p.c.coarse_ngp) { dx1[d] = 0.f; dx2[d] = 1.f; }
                else { dx1[d] = (float)i1[d] - x[d]; dx2[d] = 1.0f - dx1[d]; }
              }
              dx1[0] = mass_p * dx1[0]; dx2[0] = mass_p * dx2[0];
              for (int cz = 0; cz < 2; ++cz)
                for (int cy = 0; cy < 2; ++cy)
                  for (int cx = 0; cx < 2; ++cx) {
                    int ix = cx ? i2[0] : i1[0], iy = cy ? i2[1] : i1[1], iz = cz ? i2[2] : i1[2];
                    if (ix < 1 || ix > nc || iy < 1 || iy > nc || iz < 1 || iz > nc) continue;
                    float wgt = ((cx ? dx2[0] : dx1[0]) * (cy ? dx2[1] : dx1[1])) * (cz ? dx2[2] : dx1[2]);
                    RC(ix, iy, iz) = RC(ix, iy, iz) + wgt;
                  }
              pp = r.ll[pp - 1];
            }
          }
  }
  w.stage_ms[CUBEP3M_B200_ST_COARSE_MASS] += (float)(now_ms() - t0);
}

void coarse_force(World& w, float* c_force_max) {
  double t0 = now_ms();
  const Params& p = w.p;
  const int Nx = p.Nc[0], Ny = p.Nc[1], Nz = p.Nc[2], N2 = Nx + 2, hc = Nx / 2 + 1, nc = p.nc_node, fc = nc + 2;
  std::vector<float> slab((size_t)N2 * Ny * Nz), cmplx;
  // gather cubes -> global (pack_slab, fft_coarse.f90:4-54): global x index = local + nc*cart_coords(3) etc.
  for (auto& r : w.R)
    for (int k = 0; k < nc; ++k)
      for (int j = 0; j < nc; ++j) {
        float* row = &slab[(size_t)N2 * ((j + nc * r.cc[1]) + (size_t)Ny * (k + nc * r.cc[0]))] + nc * r.cc[2];
        const float* src = &r.rho_c[(size_t)nc * (j + (size_t)nc * k)];
        for (int i = 0; i < nc; ++i) row[i] = src[i];
      }
  w.fft_c.forward(slab.data());                                       // coarse_force.f90:18
  cmplx = slab;                                                       // :19
  const float n3 = ((float)Nx * (float)Ny) * (float)Nz;               // fft_coarse.f90:186
  for (auto& r : w.R) std::fill(r.force_c.begin(), r.force_c.end(), 0.f);
  for (int d = 0; d < 3; ++d) {
    for (int k = 0; k < Nz; ++k)                                      // :37-48
      for (int j = 0; j < Ny; ++j) {
        size_t rowo = (size_t)N2 * (j + (size_t)Ny * k);
        const float* kc = &w.kern_c[(size_t)3 * ((size_t)hc * (j + (size_t)Ny * k))];
        for (int i = 0; i < hc; ++i) {
          float kv = kc[3 * i + d];
          slab[rowo + 2 * i] = -cmplx[rowo + 2 * i + 1] * kv;
          slab[rowo + 2 * i + 1] = cmplx[rowo + 2 * i] * kv;
        }
      }
    w.fft_c.backward(slab.data());
    for (size_t q = 0; q < slab.size(); ++q) slab[q] = slab[q] / n3;
    for (auto& r : w.R)                                               // unpack_slab + :52 force_c(d,1:nc,...) = rho_c
      for (int k = 0; k < nc; ++k)
        for (int j = 0; j < nc; ++j) {
          const float* row = &slab[(size_t)N2 * ((j + nc * r.cc[1]) + (size_t)Ny * (k + nc * r.cc[0]))] + nc * r.cc[2];
          for (int i = 0; i < nc; ++i)
            r.force_c[(size_t)d + 3 * ((size_t)(i + 1) + (size_t)fc * ((j + 1) + (size_t)fc * (k + 1)))] = row[i];
        }
  }
  // coarse_force_buffer.f90:23-63 — x faces, then y, then z, each on the full (0:nc+1)^2 extent
  auto FC = [&](RankState& r, int d, int i, int j, int k) -> float& {
    return r.force_c[(size_t)d + 3 * ((size_t)i + (size_t)fc * (j + (size_t)fc * k))];
  };
  for (int axis = 0; axis < 3; ++axis) {
    std::vector<std::vector<float>> lo_face(w.R.size()), hi_face(w.R.size());
    for (auto& r : w.R) {
      lo_face[r.rank].resize((size_t)3 * fc * fc); hi_face[r.rank].resize((size_t)3 * fc * fc);
      for (int v = 0; v < fc; ++v)
        for (int u = 0; u < fc; ++u)
          for (int d = 0; d < 3; ++d) {
            int a1[3], aN[3];
            if (axis == 0) { a1[0] = 1; a1[1] = u; a1[2] = v; }
            else if (axis == 1) { a1[0] = u; a1[1] = 1; a1[2] = v; }
            else { a1[0] = u; a1[1] = v; a1[2] = 1; }
            aN[0] = a1[0]; aN[1] = a1[1]; aN[2] = a1[2]; aN[axis] = nc;
            lo_face[r.rank][(size_t)d + 3 * (u + (size_t)fc * v)] = FC(r, d, a1[0], a1[1], a1[2]);
            hi_face[r.rank][(size_t)d + 3 * (u + (size_t)fc * v)] = FC(r, d, aN[0], aN[1], aN[2]);
          }
    }
    for (auto& r : w.R) {
      int minus_nb = r.nb[2 * (2 - axis)], plus_nb = r.nb[2 * (2 - axis) + 1];
      // face 1 is sent to the - neighbour and lands in its nc+1 layer => our nc+1 layer comes from the + neighbour's face 1
      for (int v = 0; v < fc; ++v)
        for (int u = 0; u < fc; ++u)
          for (int d = 0; d < 3; ++d) {
            int a[3];
            if (axis == 0) { a[1] = u; a[2] = v; } else if (axis == 1) { a[0] = u; a[2] = v; } else { a[0] = u; a[1] = v; }
            a[axis] = nc + 1; FC(r, d, a[0], a[1], a[2]) = lo_face[plus_nb][(size_t)d + 3 * (u + (size_t)fc * v)];
            a[axis] = 0;      FC(r, d, a[0], a[1], a[2]) = hi_face[minus_nb][(size_t)d + 3 * (u + (size_t)fc * v)];
          }
    }
  }
  // coarse_max_dt.f90:19-31
  float mx = 0.f;
  for (auto& r : w.R)
    for (int k = 1; k <= nc; ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) {
          float fx = FC(r, 0, i, j, k), fy = FC(r, 1, i, j, k), fz = FC(r, 2, i, j, k);
          float f = std::sqrt((fx * fx + fy * fy) + fz * fz);
          if (f > mx) mx = f;
        }
  *c_force_max = mx;
  w.stage_ms[CUBEP3M_B200_ST_COARSE_FORCE] += (float)(now_ms() - t0);
}

void coarse_velocity(World& w, float a_mid, float dt) {
  double t0 = now_ms();
  const Params& p = w.p;
  const int nc = p.nc_node, fc = nc + 2;
  const float G = p.c.G;
  for (auto& r : w.R) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 1; k <= nc; ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) {
          int pp = r.hoc[hidx(p, i, j, k)];
          while (pp != 0) {
            float* q = &r.xv[(size_t)6 * (pp - 1)];
            float x[3], dx1[3], dx2[3]; int i1[3], i2[3];
            for (int d = 0; d < 3; ++d) {
              x[d] = (1.0f / (float)p.s) * q[d] - 0.5f;
              i1[d] = ifloor(x[d]) + 1; i2[d] = i1[d] + 1;
              if (p.c.coarse_ngp) { dx1[d] = 0.f; dx2[d] = 1.f; }
              else { dx1[d] = (float)i1[d] - x[d]; dx2[d] = 1.0f - dx1[d]; }
            }
            for (int cz = 0; cz < 2; ++cz)
              for (int cy = 0; cy < 2; ++cy)
                for (int cx = 0; cx < 2; ++cx) {
                  float dV = ((((a_mid * G) * dt) * (cx ? dx2[0] : dx1[0])) * (cy ? dx2[1] : dx1[1])) * (cz ? dx2[2] : dx1[2]);
                  int ix = cx ? i2[0] : i1[0], iy = cy ? i2[1] : i1[1], iz = cz ? i2[2] : i1[2];
                  const float* f = &r.force_c[(size_t)3 * ((size_t)ix + (size_t)fc * (iy + (size_t)fc * iz))];
                  for (int d = 0; d < 3; ++d) q[3 + d] = q[3 + d] + f[d] * dV;
                }
            pp = r.ll[pp - 1];
          }
        }
  }
  w.stage_ms[CUBEP3M_B200_ST_COARSE_VEL] += (float)(now_ms() - t0);
}

int particle_mesh(World& w, float dt, float dt_old, float a_mid, float mass_p, const float offset[3], cubep3m_b200_step_out* out) {
  double tstart = now_ms();
  const Params& p = w.p;
  std::fill(w.stage_ms, w.stage_ms + CUBEP3M_B200_ST_COUNT, 0.f);
  update_position(w, dt, dt_old, offset);                              // particle_mesh_threaded.f90:56
  link_list(w);                                                        // :61
  int st = particle_pass(w);                                           // :63
  if (st) return st;
  int np_ghost = 0, np_del = 0;
  for (auto& r : w.R) { np_ghost = std::max(np_ghost, r.np_local); np_del += r.np_deleted_ll; }

  StepScalars tot;
  double t_dep = 0, t_fft = 0, t_kick = 0, t_pp = 0, t_ppext = 0;
  for (auto& r : w.R) {
    r.tile_counts.assign(p.tiles_node, 0);
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    std::vector<StepScalars> ts(nthreads);
    std::vector<double> td(nthreads * 5, 0.0);
    // PP_EXT kicks ghost-free but reads/writes velocities of particles in neighbouring tiles' margins
    // only through "phys" guards, so tiles are independent exactly as in the reference's !$omp do (:84-85).
    // fewer tiles than threads (tiles_node_dim = 1 or 2): the idle threads go to the batches of each tile's FFT passes (fft_ref.h)
    const int team = std::max(1, std::min(nthreads, p.tiles_node));
#ifdef _OPENMP
    omp_set_max_active_levels(2);
    oracle::g_threads = nthreads;
#endif
#pragma omp parallel num_threads(team)
    {
      int tid = 0;
#ifdef _OPENMP
      tid = omp_get_thread_num();
#endif
      TileScratch S;
#pragma omp for schedule(dynamic, 1)
      for (int cur_tile = 1; cur_tile <= p.tiles_node; ++cur_tile)
        fine_tile(w, r, cur_tile, a_mid, dt, mass_p, S, ts[tid], &td[tid * 5], &td[tid * 5 + 1], &td[tid * 5 + 2], &td[tid * 5 + 3], &td[tid * 5 + 4]);
    }
    for (int t = 0; t < nthreads; ++t) {
      tot.f_force_max = std::max(tot.f_force_max, ts[t].f_force_max);
      tot.pp_force_max = std::max(tot.pp_force_max, ts[t].pp_force_max);
      tot.pp_ext_force_max = std::max(tot.pp_ext_force_max, ts[t].pp_ext_force_max);
      tot.f_mesh_mass += ts[t].f_mesh_mass;
      if (ts[t].status) tot.status = ts[t].status;
      t_dep = std::max(t_dep, td[t * 5]); t_fft = std::max(t_fft, td[t * 5 + 1]); t_kick = std::max(t_kick, td[t * 5 + 2]);
      t_pp = std::max(t_pp, td[t * 5 + 3]); t_ppext = std::max(t_ppext, td[t * 5 + 4]);
    }
  }
  if (tot.status) return tot.status;
  w.stage_ms[CUBEP3M_B200_ST_FINE_DEPOSIT] = (float)t_dep; w.stage_ms[CUBEP3M_B200_ST_FINE_FFT] = (float)t_fft;
  w.stage_ms[CUBEP3M_B200_ST_FINE_KICK] = (float)t_kick; w.stage_ms[CUBEP3M_B200_ST_PP] = (float)t_pp;
  w.stage_ms[CUBEP3M_B200_ST_PP_EXT] = (float)t_ppext;

  const float G = p.c.G;
  float f_force_max_node = std::sqrt(tot.f_force_max);                 // :643
  out->f_force_max = f_force_max_node;
  out->dt_f_acc = 1.0f / std::sqrt(std::max(0.0001f, f_force_max_node) * a_mid * G);   // :652
  out->pp_force_max = tot.pp_force_max;
  out->dt_pp_acc = p.c.ppint ? std::sqrt(p.c.dt_pp_scale * p.c.rsoft) / std::max(std::sqrt(tot.pp_force_max * a_mid * G), 1e-3f) : 1000.f;  // :668
  out->pp_ext_force_max = tot.pp_ext_force_max;
  out->dt_pp_ext_acc = p.c.pp_ext ? std::sqrt(p.c.dt_pp_scale * p.c.rsoft) / std::max(std::sqrt(tot.pp_ext_force_max * a_mid * G), 1e-3f) : 1000.f;  // :692
  out->sum_rho_f = tot.f_mesh_mass;                                    // :703

  coarse_mass(w, mass_p);                                              // coarse_mesh.f90:28
  double sumc = 0.0;
  for (auto& r : w.R) for (float v : r.rho_c) sumc += (double)v;       // coarse_mesh.f90:31-43
  out->sum_rho_c = sumc;
  float cmax = 0.f;
  coarse_force(w, &cmax);                                              // coarse_mesh.f90:84-96
  out->c_force_max = cmax;
  out->dt_c_acc = std::sqrt((float)p.s / (cmax * a_mid * G));          // coarse_max_dt.f90:36
  if (p.c.coarse_vel_update) coarse_velocity(w, a_mid, dt);            // coarse_mesh.f90:106
  delete_particles(w);                                                 // particle_mesh_threaded.f90:720
  int64_t tot_np = 0; int npl = 0;
  for (auto& r : w.R) { tot_np += r.np_local; npl = r.np_local; }
  out->np_total = tot_np; out->np_local = npl; out->np_with_ghosts = np_ghost; out->np_deleted_ll = np_del; out->np_buf_max = w.np_buf_max;
  w.stage_ms[CUBEP3M_B200_ST_TOTAL] = (float)(now_ms() - tstart);
  for (int i = 0; i < CUBEP3M_B200_ST_COUNT; ++i) out->stage_ms[i] = w.stage_ms[i];
  return 0;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// C interface for ctypes (tests / bench cpu_baseline only)
// -------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------- halofind.f90:564-672 (find_halos: density + maxima pass)
struct OPeak { int i, j, k, tile; float den, x, y, z; };

static float para_inter(const float x[3], const float fx[3]) {          // halofind.f90:770-778
  return x[1] - 0.5f * (((x[1] - x[0]) * (x[1] - x[0])) * (fx[1] - fx[2]) - ((x[1] - x[2]) * (x[1] - x[2])) * (fx[1] - fx[0])) /
                    ((x[1] - x[0]) * (fx[1] - fx[2]) - (x[1] - x[2]) * (fx[1] - fx[0]));
}

// peaks are appended tile by tile in the scan order of :620-627 (before the sort of :676-679); cft[0..1] += cftmass, cftmass2
void find_peaks(World& w, RankState& r, float mass_p, float den_peak_cutoff, int para_inter_hc, int ngph, std::vector<OPeak>& out, double* cft) {
  const Params& p = w.p;
  const int n = p.n, n2 = n + 2, b = p.b, m = p.m, T = p.T;
  std::vector<float> rho;
  for (int cur = 1; cur <= p.tiles_node; ++cur) {                      // halofind.f90:48-54
    int tile[3];
    tile[2] = (cur - 1) / (T * T);
    int j0 = cur - tile[2] * T * T;
    tile[1] = (j0 - 1) / T;
    j0 = j0 - tile[1] * T;
    tile[0] = j0 - 1;
    rho.assign((size_t)n2 * n * n, 0.f);                               // :591
    auto RHO = [&](int i, int j, int k) -> float& { return rho[(size_t)(i - 1) + (size_t)n2 * ((j - 1) + (size_t)n * (k - 1))]; };
    int cic_l[3], cic_h[3], off_i[3];
    float offset[3];
    for (int d = 0; d < 3; ++d) {
      off_i[d] = tile[d] * m - b;                                      // :585
      cic_l[d] = p.nc_tile * tile[d] + 2 - p.nc_buf;                   // :597
      cic_h[d] = p.nc_tile * (tile[d] + 1) + p.nc_buf - 1;             // :598
      offset[d] = (float)(-tile[d] * m + b);                           // fine_ngp_mass.f90:11 / fine_cic_mass.f90:13
    }
    for (int k = cic_l[2]; k <= cic_h[2]; ++k)                         // :602-616
      for (int j = cic_l[1]; j <= cic_h[1]; ++j)
        for (int i = cic_l[0]; i <= cic_h[0]; ++i) {
          int pp = r.hoc[hidx(p, i, j, k)];
          while (pp != 0) {
            const float* q = &r.xv[(size_t)6 * (pp - 1)];
            float x[3]; int i1[3];
            for (int d = 0; d < 3; ++d) { x[d] = q[d] + offset[d]; i1[d] = ifloor(x[d]) + 1; }
            if (ngph) {
              RHO(i1[0], i1[1], i1[2]) = RHO(i1[0], i1[1], i1[2]) + mass_p;          // fine_ngp_mass.f90:19
            } else {                                                                  // fine_cic_mass.f90:17-41
              float dx1[3], dx2[3];
              for (int d = 0; d < 3; ++d) { dx1[d] = (float)i1[d] - x[d]; dx2[d] = 1.f - dx1[d]; }
              dx1[0] = mass_p * dx1[0]; dx2[0] = mass_p * dx2[0];
              for (int cz = 0; cz < 2; ++cz)
                for (int cy = 0; cy < 2; ++cy)
                  for (int cx = 0; cx < 2; ++cx) {
                    float wgt = ((cx ? dx2[0] : dx1[0]) * (cy ? dx2[1] : dx1[1])) * (cz ? dx2[2] : dx1[2]);
                    float& c = RHO(i1[0] + cx, i1[1] + cy, i1[2] + cz);
                    c = c + wgt;
                  }
            }
            pp = r.ll[pp - 1];
          }
        }
    for (int k = 1 + b; k <= b + m; ++k)                               // :620-672
      for (int j = 1 + b; j <= b + m; ++j)
        for (int i = 1 + b; i <= b + m; ++i) {
          const float c = RHO(i, j, k);
          cft[0] += (double)c;
          cft[1] += (double)(c * c);
          float denmax = c;
          for (int kk = k - 1; kk <= k + 1; ++kk)
            for (int jj = j - 1; jj <= j + 1; ++jj)
              for (int ii = i - 1; ii <= i + 1; ++ii) denmax = std::max(denmax, RHO(ii, jj, kk));
          if (denmax == c && denmax > den_peak_cutoff) {
            OPeak q;
            q.i = i; q.j = j; q.k = k; q.tile = cur - 1; q.den = denmax;
            float pos[3];
            if (para_inter_hc) {
              const int c3[3] = {i, j, k};
              for (int d = 0; d < 3; ++d) {
                const float x[3] = {(float)(c3[d] - 1) - 0.5f, (float)c3[d] - 0.5f, (float)(c3[d] + 1) - 0.5f};
                const float fx[3] = {RHO(i - (d == 0), j - (d == 1), k - (d == 2)), c, RHO(i + (d == 0), j + (d == 1), k + (d == 2))};
                pos[d] = para_inter(x, fx);
              }
            } else {
              pos[0] = (float)i - 0.5f; pos[1] = (float)j - 0.5f; pos[2] = (float)k - 0.5f;
            }
            q.x = pos[0] + (float)off_i[0]; q.y = pos[1] + (float)off_i[1]; q.z = pos[2] + (float)off_i[2];   // :723
            out.push_back(q);
          }
        }
  }
}

extern "C" {

int oracle_create(const cubep3m_b200_config* cfg, const float* fine_table, const float* coarse_table, int build_kernels, void** out) {
  World* w = new World();
  w->p.c = *cfg; w->p.derive();
  const Params& p = w->p;
  if (p.s != 4 || p.m <= 0 || p.m % p.s != 0) { delete w; return CUBEP3M_B200_EINVAL; }
  // LRCKCORR divides by Im(kernel) for every |k| <= 8 (kernel_initialization.f90:573-581): the Nyquist plane must lie beyond that
  for (int a = 0; a < 3; ++a) if (p.c.lrckcorr && p.Nc[a] / 2 <= 8) { delete w; return CUBEP3M_B200_EINVAL; }
  w->fine_table.assign(fine_table, fine_table + 16 * 16 * 16 * 3);
  w->coarse_table.assign(coarse_table, coarse_table + 4 * 4 * 4 * 3);
  w->fft_f.init(p.n); w->fft_c.init(p.Nc[0], p.Nc[1], p.Nc[2]);
  w->R.resize(p.nodes);
  const int Dx = p.Dg[0], Dy = p.Dg[1], Dz = p.Dg[2];
  for (int r = 0; r < p.nodes; ++r) {
    RankState& R = w->R[r];
    R.rank = r;
    R.cc[0] = r / (Dx * Dy); R.cc[1] = (r / Dx) % Dy; R.cc[2] = r % Dx;      // rank = x + Dx*(y + Dy*z); cc = (z,y,x)
    auto rk = [&](int z, int y, int x) { return ((z + Dz) % Dz) * Dx * Dy + ((y + Dy) % Dy) * Dx + ((x + Dx) % Dx); };
    R.nb[0] = rk(R.cc[0] - 1, R.cc[1], R.cc[2]); R.nb[1] = rk(R.cc[0] + 1, R.cc[1], R.cc[2]);
    R.nb[2] = rk(R.cc[0], R.cc[1] - 1, R.cc[2]); R.nb[3] = rk(R.cc[0], R.cc[1] + 1, R.cc[2]);
    R.nb[4] = rk(R.cc[0], R.cc[1], R.cc[2] - 1); R.nb[5] = rk(R.cc[0], R.cc[1], R.cc[2] + 1);
    R.xv.assign((size_t)6 * p.max_np, 0.f); R.ll.assign(p.max_np, 0);
    if (p.c.pid) { R.pid.assign(p.max_np, 0); R.send_pid.assign(p.max_buf / 6 + 1, 0); R.recv_pid.assign(p.max_buf / 6 + 1, 0); }
    R.hoc.assign((size_t)p.H * p.H * p.H, 0);
    R.rho_c.assign((size_t)p.nc_node * p.nc_node * p.nc_node, 0.f);
    R.force_c.assign((size_t)3 * (p.nc_node + 2) * (p.nc_node + 2) * (p.nc_node + 2), 0.f);
    R.send_buf.assign(p.max_buf, 0.f); R.recv_buf.assign(p.max_buf, 0.f);
    R.tile_counts.assign(p.tiles_node, 0);
  }
  if (build_kernels) { fine_kernel(*w); coarse_kernel(*w); }
  *out = w;
  return 0;
}
void oracle_destroy(void* h) { delete (World*)h; }
int oracle_max_np(void* h) { return ((World*)h)->p.max_np; }
int oracle_set_kernels(void* h, const float* kern_f, const float* kern_c_global) {
  World* w = (World*)h; const Params& p = w->p;
  size_t nf = (size_t)3 * (p.n / 2 + 1) * p.n * p.n, ncg = (size_t)3 * (p.Nc[0] / 2 + 1) * p.Nc[1] * p.Nc[2];
  w->kern_f.assign(kern_f, kern_f + nf); w->kern_c.assign(kern_c_global, kern_c_global + ncg);
  return 0;
}
int oracle_set_particles(void* h, int rank, const float* xv, const int64_t* pid, int np) {
  World* w = (World*)h;
  if (np > w->p.max_np) return CUBEP3M_B200_EMAXNP;
  RankState& r = w->R[rank];
  std::copy(xv, xv + (size_t)6 * np, r.xv.begin());
  if (pid && w->p.c.pid) std::copy(pid, pid + np, r.pid.begin());
  r.np_local = np;
  return 0;
}
int oracle_get_np(void* h, int rank) { return ((World*)h)->R[rank].np_local; }
int oracle_get_particles(void* h, int rank, float* xv, int64_t* pid) {
  World* w = (World*)h; RankState& r = w->R[rank];
  std::copy(r.xv.begin(), r.xv.begin() + (size_t)6 * r.np_local, xv);
  if (pid && w->p.c.pid) std::copy(r.pid.begin(), r.pid.begin() + r.np_local, pid);
  return r.np_local;
}
int oracle_update_position(void* h, float dt, float dt_old, const float* offset) { update_position(*(World*)h, dt, dt_old, offset); return 0; }
int oracle_move_grid_back(void* h, const float* shake) { move_grid_back(*(World*)h, shake); return 0; }
int oracle_link_list(void* h) { link_list(*(World*)h); return 0; }
int oracle_particle_pass(void* h) { return particle_pass(*(World*)h); }
int oracle_delete_particles(void* h) { delete_particles(*(World*)h); return 0; }
int oracle_particle_mesh(void* h, float dt, float dt_old, float a_mid, float mass_p, const float* offset, cubep3m_b200_step_out* out) {
  return particle_mesh(*(World*)h, dt, dt_old, a_mid, mass_p, offset, out);
}
// chained particles per coarse cell of the hoc range, x fastest
int oracle_cell_counts(void* h, int rank, int32_t* counts) {
  World* w = (World*)h; RankState& r = w->R[rank];
  size_t nh = r.hoc.size();
  for (size_t c = 0; c < nh; ++c) { int n = 0; for (int pp = r.hoc[c]; pp != 0; pp = r.ll[pp - 1]) ++n; counts[c] = n; }
  return 0;
}
int oracle_tile_counts(void* h, int rank, int32_t* counts) {
  World* w = (World*)h; RankState& r = w->R[rank];
  std::copy(r.tile_counts.begin(), r.tile_counts.end(), counts); return 0;
}
int oracle_kern_f(void* h, float* out) { World* w = (World*)h; std::copy(w->kern_f.begin(), w->kern_f.end(), out); return 0; }
int oracle_kern_c(void* h, int rank, float* out) {   // this rank's z-slab (3,hc,nc_dim,nc_slab)
  World* w = (World*)h; const Params& p = w->p;
  if (p.nc_slab > 0) {
    size_t per = (size_t)3 * (p.Nc[0] / 2 + 1) * p.Nc[1] * p.nc_slab;
    std::copy(w->kern_c.begin() + per * rank, w->kern_c.begin() + per * (rank + 1), out);
  } else std::copy(w->kern_c.begin(), w->kern_c.end(), out);
  return 0;
}
int oracle_rho_c(void* h, int rank, float* out) { RankState& r = ((World*)h)->R[rank]; std::copy(r.rho_c.begin(), r.rho_c.end(), out); return 0; }
int oracle_force_c(void* h, int rank, float* out) { RankState& r = ((World*)h)->R[rank]; std::copy(r.force_c.begin(), r.force_c.end(), out); return 0; }
int oracle_set_debug_tile(void* h, int rank, int tile) { ((World*)h)->R[rank].dbg_tile = tile; return 0; }
int oracle_fine_tile(void* h, int rank, float* rho_f, float* force_f) {
  RankState& r = ((World*)h)->R[rank];
  if (r.dbg_rho_f.empty()) return CUBEP3M_B200_ENOTREADY;
  std::copy(r.dbg_rho_f.begin(), r.dbg_rho_f.end(), rho_f); std::copy(r.dbg_force_f.begin(), r.dbg_force_f.end(), force_f); return 0;
}
int oracle_find_peaks(void* h, int rank, float mass_p, float den_peak_cutoff, int para_inter_hc, int ngph, void* peaks, int max_peaks, int* n_peaks, double* cft) {
  World* w = (World*)h;
  std::vector<OPeak> v;
  double c2[2] = {0.0, 0.0};
  find_peaks(*w, w->R[rank], mass_p, den_peak_cutoff, para_inter_hc, ngph, v, c2);
  *n_peaks = (int)v.size();
  if (cft) { cft[0] = c2[0]; cft[1] = c2[1]; }
  if ((int)v.size() > max_peaks) return 8;
  memcpy(peaks, v.data(), v.size() * sizeof(OPeak));
  return 0;
}
int oracle_fft3d(int n, float* data, int inverse) {
  oracle::Fft3dR2C f; f.init(n);
  if (inverse) f.backward(data); else f.forward(data);
  return 0;
}
int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
}
