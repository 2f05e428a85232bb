/* replay_driver.c — plain-C replay of the CUBEP3M driver's call order through the C ABI of include/cubep3m_b200.h.
 *
 * The Fortran driver (and the shim fortran/particle_mesh_b200.f90) cannot be compiled in this image, so this program stands in for it:
 * it walks the main loop of cubepm.f90:103-236 — timestep, particle_mesh, and on a checkpoint step dt_old = 0, update_position (which
 * draws a SECOND shake offset, update_position.f90:56-63), [move_grid_back], checkpoint, link_list, particle_pass, delete_particles,
 * dt = 0 — exactly as the shim does, in
 *     strict   mode (upload before / download after every call that touches xv: -DB200_STRICT) and
 *     resident mode (the device copy is authoritative; download only at the checkpoint step),
 * and checks both against the CPU oracle driven through the same sequence:
 *   - step 1 (identical input on both sides): positions bit-exact;
 *   - every later comparison point: same particle count, every particle (matched by PID) within 2e-3 fine cells and its velocity
 *     within 2e-3 relative of the oracle's (fp32 summation order; positions after a kick are no longer bit-comparable);
 *   - limiters within 1e-3;
 *   - the two modes agree with each other on the particle count at every step.
 * Usage: replay_driver <fine_table.npy> <coarse_table.npy> [steps] [checkpoint_step]     (exit code 0 = all checks passed)
 * TEST INFRASTRUCTURE: links liboracle; not part of the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/cubep3m_b200.h"

/* oracle C interface (oracle/cubep3m_oracle.cpp) */
int oracle_create(const cubep3m_b200_config* cfg, const float* fine_table, const float* coarse_table, int build_kernels, void** out);
void oracle_destroy(void* h);
int oracle_set_particles(void* h, int rank, const float* xv, const int64_t* pid, int np);
int oracle_get_np(void* h, int rank);
int oracle_get_particles(void* h, int rank, float* xv, int64_t* pid);
int oracle_update_position(void* h, float dt, float dt_old, const float* offset);
int oracle_move_grid_back(void* h, const float* shake);
int oracle_link_list(void* h);
int oracle_particle_pass(void* h);
int oracle_delete_particles(void* h);
int oracle_particle_mesh(void* h, float dt, float dt_old, float a_mid, float mass_p, const float* offset, cubep3m_b200_step_out* out);

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++failures; fprintf(stderr, "CHECK FAILED %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)
#define ST(call) do { int st__ = (call); if (st__) { fprintf(stderr, "%s -> status %d (%s)\n", #call, st__, cubep3m_b200_strerror(st__)); exit(2); } } while (0)

static float* read_npy_f32(const char* path, size_t count) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  unsigned char hdr[10];
  if (fread(hdr, 1, 10, f) != 10 || memcmp(hdr, "\x93NUMPY", 6)) { fprintf(stderr, "%s: not an npy file\n", path); exit(2); }
  const size_t hlen = hdr[8] | ((size_t)hdr[9] << 8);
  fseek(f, (long)(10 + hlen), SEEK_SET);
  float* a = (float*)malloc(count * sizeof(float));
  if (fread(a, sizeof(float), count, f) != count) { fprintf(stderr, "%s: short read\n", path); exit(2); }
  fclose(f);
  return a;
}

/* the driver's random_number stand-in: both sides get the same offsets */
static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static float rnd01(void) { lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull; return (float)((lcg_state >> 40) & 0xFFFFFF) / 16777216.0f; }
static void draw_offset(float shake[3], float off[3]) {      /* update_position.f90:56-58 */
  for (int a = 0; a < 3; ++a) { off[a] = (rnd01() - 0.5f) * 4.0f * 4.0f - shake[a]; shake[a] = shake[a] + off[a]; }
}

typedef struct { int64_t id; const float* p; } rec_t;
static int cmp_rec(const void* a, const void* b) { const int64_t x = ((const rec_t*)a)->id, y = ((const rec_t*)b)->id; return (x > y) - (x < y); }

/* compare two particle lists matched by PID */
static void compare_sets(const char* what, int step, const float* g, const int64_t* gp, int ng, const float* r, const int64_t* rp, int nr, int bit_exact, float period) {
  CHECK(ng == nr, "%s step %d: particle count %d vs oracle %d", what, step, ng, nr);
  if (ng != nr) return;
  rec_t* a = (rec_t*)malloc(sizeof(rec_t) * ng); rec_t* b = (rec_t*)malloc(sizeof(rec_t) * nr);
  for (int i = 0; i < ng; ++i) { a[i].id = gp[i]; a[i].p = g + 6 * (size_t)i; b[i].id = rp[i]; b[i].p = r + 6 * (size_t)i; }
  qsort(a, ng, sizeof(rec_t), cmp_rec); qsort(b, nr, sizeof(rec_t), cmp_rec);
  double max_dx = 0, sum2 = 0; long nbad = 0, nv = 0;
  for (int i = 0; i < ng; ++i) {
    if (a[i].id != b[i].id) { ++nbad; continue; }
    double dv2 = 0, v2 = 0;
    for (int c = 0; c < 3; ++c) {
      double dx = fabs((double)a[i].p[c] - (double)b[i].p[c]);
      if (dx > 0.5 * period) dx = period - dx;                      /* the same particle may sit on either side of the periodic seam */
      if (dx > max_dx) max_dx = dx;
      if (bit_exact && a[i].p[c] != b[i].p[c]) ++nbad;
      const double dv = (double)a[i].p[3 + c] - (double)b[i].p[3 + c];
      dv2 += dv * dv; v2 += (double)b[i].p[3 + c] * (double)b[i].p[3 + c];
    }
    if (v2 > 0) { sum2 += dv2 / v2; ++nv; }
  }
  const double rms = nv ? sqrt(sum2 / nv) : 0.0;
  CHECK(nbad == 0, "%s step %d: %ld mismatching records%s", what, step, nbad, bit_exact ? " (bit-exact positions required)" : "");
  CHECK(max_dx <= 2e-3, "%s step %d: max position difference %.3e fine cells", what, step, max_dx);
  CHECK(rms <= 2e-3, "%s step %d: rms relative velocity difference %.3e", what, step, rms);
  printf("  %-28s step %d: n=%d max|dx|=%.2e rms(dv/v)=%.2e%s\n", what, step, ng, max_dx, rms, bit_exact ? " [bit-exact x]" : "");
  free(a); free(b);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s fine_table.npy coarse_table.npy [steps] [checkpoint_step]\n", argv[0]); return 2; }
  float* fine_table = read_npy_f32(argv[1], 16 * 16 * 16 * 3);
  float* coarse_table = read_npy_f32(argv[2], 4 * 4 * 4 * 3);
  const int steps = argc > 3 ? atoi(argv[3]) : 5, ckpt = argc > 4 ? atoi(argv[4]) : 3;

  cubep3m_b200_config cfg;
  cubep3m_b200_default_config(&cfg);
  cfg.nf_tile = 112; cfg.tiles_node_dim = 2; cfg.pp_ext = 1; cfg.pid = 1;       /* 128^3 fine cells, 64^3 particles, PPINT + PP_EXT, -DPID_FLAG */
  const int mT = (cfg.nf_tile - 2 * cfg.nf_buf) * cfg.tiles_node_dim, nside = mT / 2, np0 = nside * nside * nside;
  /* perturbed lattice: one particle per 2 fine cells (dist_init_dm.f90:1019-1036 layout) with smooth displacements and velocities */
  float* xv0 = (float*)malloc(sizeof(float) * 6 * (size_t)np0);
  int64_t* pid0 = (int64_t*)malloc(sizeof(int64_t) * (size_t)np0);
  for (int k = 0, i = 0; k < nside; ++k)
    for (int j = 0; j < nside; ++j)
      for (int ii = 0; ii < nside; ++ii, ++i) {
        const float x = 2.0f * ii + 0.5f, y = 2.0f * j + 0.5f, z = 2.0f * k + 0.5f;
        const float w = 6.2831853f / (float)mT;
        const float dx = 0.9f * sinf(w * x) * cosf(2 * w * y), dy = 0.9f * sinf(w * y + 1.0f) * cosf(w * z), dz = 0.9f * sinf(2 * w * z) * cosf(w * x + 0.5f);
        float* p = xv0 + 6 * (size_t)i;
        p[0] = fmodf(x + dx + 0.37f * rnd01() + (float)mT, (float)mT); p[1] = fmodf(y + dy + 0.37f * rnd01() + (float)mT, (float)mT);
        p[2] = fmodf(z + dz + 0.37f * rnd01() + (float)mT, (float)mT);
        p[3] = 0.8f * dx; p[4] = 0.8f * dy; p[5] = 0.8f * dz;
        pid0[i] = 1000 + 3 * (int64_t)i;
      }
  const float mass_p = ((float)mT * (float)mT * (float)mT) / (float)np0;
  const int max_np = 2 * np0 + 400000;
  cfg.max_np = max_np;
  float* host = (float*)malloc(sizeof(float) * 6 * (size_t)max_np);      /* the driver's xv(6,max_np) */
  int64_t* hpid = (int64_t*)malloc(sizeof(int64_t) * (size_t)max_np);
  float* oxv = (float*)malloc(sizeof(float) * 6 * (size_t)max_np);
  int64_t* opid = (int64_t*)malloc(sizeof(int64_t) * (size_t)max_np);
  int np_mode[2][64];

  for (int mode = 0; mode < 2; ++mode) {                                   /* 0 = strict, 1 = resident */
    printf("== %s mode\n", mode == 0 ? "strict" : "resident");
    lcg_state = 0x2545F4914F6CDD1Dull;
    cubep3m_b200_ctx* ctx = NULL;
    void* orc = NULL;
    ST(cubep3m_b200_init(&cfg, fine_table, coarse_table, NULL, NULL, NULL, 1, &ctx));
    ST(oracle_create(&cfg, fine_table, coarse_table, 1, &orc));
    memcpy(host, xv0, sizeof(float) * 6 * (size_t)np0); memcpy(hpid, pid0, sizeof(int64_t) * (size_t)np0);
    int32_t np_local = np0;
    ST(cubep3m_b200_upload_particles(ctx, host, hpid, np_local));          /* after particle_initialize, cubepm.f90:52 */
    ST(oracle_set_particles(orc, 0, xv0, pid0, np0));
    cubep3m_b200_clock cg, co;
    cubep3m_b200_clock_init(&cg, 20.0f, 0.24f, 0.76f); cubep3m_b200_clock_init(&co, 20.0f, 0.24f, 0.76f);
    cg.ppint = co.ppint = 1; cg.pp_ext = co.pp_ext = 1;
    float shake[3] = {0, 0, 0};
    for (int step = 1; step <= steps; ++step) {
      /* a checkpoint at step `ckpt`: a_checkpoint is whatever the scale factor would reach with a 40 % shorter step (timestep.f90:128-137) */
      if (step == ckpt) { cubep3m_b200_clock t = cg; t.a_target = 1.0f; cubep3m_b200_timestep(&t); cg.a_target = co.a_target = cg.a + 0.6f * (t.a - cg.a); }
      else cg.a_target = co.a_target = 1.0f;
      cubep3m_b200_timestep(&cg); cubep3m_b200_timestep(&co);              /* cubepm.f90:104 */
      CHECK(cg.checkpoint_step == (step == ckpt), "step %d: checkpoint_step = %d", step, cg.checkpoint_step);
      float off[3];
      draw_offset(shake, off);                                              /* the shim's b200_draw_offset, once per particle_mesh */
      cubep3m_b200_step_out og, oo;
      if (mode == 0) ST(cubep3m_b200_upload_particles(ctx, host, hpid, np_local));
      ST(cubep3m_b200_particle_mesh(ctx, cg.dt, cg.dt_old, cg.a_mid, mass_p, off, &og));      /* cubepm.f90:143 */
      np_local = og.np_local;
      if (mode == 0) ST(cubep3m_b200_download_particles(ctx, host, hpid, &np_local));
      ST(oracle_particle_mesh(orc, co.dt, co.dt_old, co.a_mid, mass_p, off, &oo));
      cg.dt_f_acc = og.dt_f_acc; cg.dt_pp_acc = og.dt_pp_acc; cg.dt_pp_ext_acc = og.dt_pp_ext_acc; cg.dt_c_acc = og.dt_c_acc;
      co.dt_f_acc = oo.dt_f_acc; co.dt_pp_acc = oo.dt_pp_acc; co.dt_pp_ext_acc = oo.dt_pp_ext_acc; co.dt_c_acc = oo.dt_c_acc;
      CHECK(og.np_total == oo.np_total && og.np_local == oo.np_local, "step %d: np %d/%lld vs oracle %d/%lld", step, og.np_local, (long long)og.np_total, oo.np_local, (long long)oo.np_total);
      CHECK(fabsf(og.dt_f_acc / oo.dt_f_acc - 1.f) < 1e-3f && fabsf(og.dt_c_acc / oo.dt_c_acc - 1.f) < 1e-3f && fabsf(og.dt_pp_acc / oo.dt_pp_acc - 1.f) < 1e-2f &&
                fabsf(og.dt_pp_ext_acc / oo.dt_pp_ext_acc - 1.f) < 1e-2f,
            "step %d: limiters f %g/%g c %g/%g pp %g/%g ppext %g/%g", step, og.dt_f_acc, oo.dt_f_acc, og.dt_c_acc, oo.dt_c_acc, og.dt_pp_acc, oo.dt_pp_acc, og.dt_pp_ext_acc, oo.dt_pp_ext_acc);
      np_mode[mode][step] = og.np_local;
      if (mode == 0 || step == 1) {                                         /* strict mode has xv on the host after every step */
        if (mode == 1) ST(cubep3m_b200_download_particles(ctx, host, hpid, &np_local));
        const int nr = oracle_get_particles(orc, 0, oxv, opid);
        compare_sets("after particle_mesh", step, host, hpid, np_local, oxv, opid, nr, step == 1, (float)mT);
      }
      if (cg.checkpoint_step) {                                             /* cubepm.f90:171-233 */
        cg.dt_old = co.dt_old = 0.0f;                                       /* :175 */
        float off2[3];
        draw_offset(shake, off2);                                           /* update_position draws a new offset (update_position.f90:56-63) */
        if (mode == 0) ST(cubep3m_b200_upload_particles(ctx, host, hpid, np_local));
        ST(cubep3m_b200_update_position(ctx, cg.dt, cg.dt_old, off2));     /* :176 */
        ST(cubep3m_b200_move_grid_back(ctx, shake));                        /* :179 (-DMOVE_GRID_BACK), what checkpoint.f90:92 otherwise subtracts on the fly */
        ST(cubep3m_b200_download_particles(ctx, host, hpid, &np_local));    /* the checkpoint needs xv on the host in both modes */
        ST(oracle_update_position(orc, co.dt, co.dt_old, off2));
        ST(oracle_move_grid_back(orc, shake));
        shake[0] = shake[1] = shake[2] = 0.0f;                              /* move_grid_back.f90:24 */
        int nr = oracle_get_particles(orc, 0, oxv, opid);
        compare_sets("checkpoint (half drift)", step, host, hpid, np_local, oxv, opid, nr, 0, (float)mT);
        /* projection / halofind block: link_list, particle_pass, [consumers], delete_particles (:193-228) */
        int32_t ndel = 0, npg = 0, npl = 0;
        ST(cubep3m_b200_link_list(ctx, &ndel));
        ST(cubep3m_b200_particle_pass(ctx, &npg));
        ST(oracle_link_list(orc)); ST(oracle_particle_pass(orc));
        const int npg_o = oracle_get_np(orc, 0);
        CHECK(abs(npg - npg_o) <= 2, "step %d: particles incl. ghosts after the checkpoint pass %d vs oracle %d", step, npg, npg_o);
        ST(cubep3m_b200_delete_particles(ctx, &npl));
        ST(oracle_delete_particles(orc));
        CHECK(npl == oracle_get_np(orc, 0) && npl == np0, "step %d: np_local after delete_particles %d vs oracle %d", step, npl, oracle_get_np(orc, 0));
        np_local = npl;
        if (mode == 0) ST(cubep3m_b200_download_particles(ctx, host, hpid, &np_local));
        cg.dt = co.dt = 0.0f;                                               /* :231 */
      }
    }
    ST(cubep3m_b200_download_particles(ctx, host, hpid, &np_local));
    const int nr = oracle_get_particles(orc, 0, oxv, opid);
    compare_sets("final state", steps, host, hpid, np_local, oxv, opid, nr, 0, (float)mT);
    cubep3m_b200_finalize(ctx);
    oracle_destroy(orc);
  }
  for (int step = 1; step <= steps; ++step) CHECK(np_mode[0][step] == np_mode[1][step], "step %d: strict np %d != resident np %d", step, np_mode[0][step], np_mode[1][step]);
  printf(failures ? "REPLAY FAILED: %d check(s)\n" : "REPLAY OK (%d failures)\n", failures);
  return failures ? 1 : 0;
}
