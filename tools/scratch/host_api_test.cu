// Pure-C harness for fot_plan_batch_host (no Python/torch in the process): checks upload/compute overlap.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <vector>
#include <cuda_runtime.h>
#include "fot.h"
int main() {
  const int nT = 11, nd = 19, nB = 7, nx = 7, NQ = 4096, P = 50, TOBS = 51;
  fot_config_t cfg{}; cfg.dt = 0.1; cfg.max_speed = 10; cfg.max_road_width = 2.7;
  cfg.k_j = cfg.k_t = cfg.k_d = cfg.k_s_dot = cfg.k_lat = cfg.k_lon = 1.0;
  cfg.collide_r2 = cfg.collide_r2_single = 1.44; cfg.n_T = nT; cfg.n_d = nd; cfg.n_B = nB; cfg.n_total = 51; cfg.nx = nx;
  std::vector<double> T(nT), inv4(4 * nT), inv5(9 * nT), Tb(nB), inv4b(4 * nB), inv5b(9 * nB), dg(nd), knots(nx), xa(nx), z(nx, 0.0), xb(nx - 1, 1.0);
  std::vector<int32_t> ns(nT), nsb(nB);
  auto fill = [](double t, double* i4, double* i5) {   // closed-form inverses are fine for a timing harness
    double a = 3 * t * t, b = 4 * t * t * t, c = 6 * t, d = 12 * t * t, det = a * d - b * c;
    i4[0] = d / det; i4[1] = -b / det; i4[2] = -c / det; i4[3] = a / det;
    for (int k = 0; k < 9; ++k) i5[k] = (k % 4 == 0) ? 1.0 / (t * t * t) : 0.0;
  };
  for (int j = 0; j < nT; ++j) { T[j] = 4.0 + 0.1 * j; ns[j] = 40 + j; fill(T[j], &inv4[4 * j], &inv5[9 * j]); }
  for (int j = 0; j < nB; ++j) { Tb[j] = 0.5 * (j + 1); nsb[j] = 5 * (j + 1); fill(Tb[j], &inv4b[4 * j], &inv5b[9 * j]); }
  for (int i = 0; i < nd; ++i) dg[i] = (i - 9) * 0.3;
  for (int i = 0; i < nx; ++i) { knots[i] = 10.0 * i; xa[i] = 10.0 * i; }
  fot_tables_t tb{T.data(), ns.data(), inv4.data(), inv5.data(), Tb.data(), nsb.data(), inv4b.data(), inv5b.data(), dg.data(),
                  knots.data(), xa.data(), xb.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data()};
  fot_handle_t* h = nullptr;
  if (fot_create(&cfg, &tb, 0, &h)) { printf("create: %s\n", fot_last_error()); return 1; }
  const int NT = fot_n_t_max(h);
  double *fr, *tg, *lm, *sd, *vg, *dyn; int32_t* nv;
  cudaMallocHost(&fr, NQ * 48); cudaMallocHost(&tg, NQ * 8); cudaMallocHost(&lm, NQ * 32); cudaMallocHost(&sd, NQ * 8);
  cudaMallocHost(&vg, NQ * 6 * 8); cudaMallocHost(&nv, NQ * 4); cudaMallocHost(&dyn, (size_t)NQ * P * TOBS * 16);
  srand(1);
  for (int q = 0; q < NQ; ++q) {
    double f[6] = {5.0 + q % 10, 5.0, 0.1, 0.2, 0.0, 0.0}; memcpy(fr + 6 * q, f, 48);
    tg[q] = 6.0; double l[4] = {10, 2, 0.2, 3}; memcpy(lm + 4 * q, l, 32); sd[q] = NAN; nv[q] = 6;
    for (int k = 0; k < 6; ++k) vg[6 * q + k] = k < 5 ? 6.0 - k * 5 / 3.6 : 0.0;
    for (int p = 0; p < P; ++p) { double x0 = 5 + 40.0 * rand() / RAND_MAX, y0 = -10 + 20.0 * rand() / RAND_MAX;
      for (int k = 0; k < TOBS; ++k) { dyn[(((size_t)q * P + p) * TOBS + k) * 2] = x0; dyn[(((size_t)q * P + p) * TOBS + k) * 2 + 1] = y0 + 0.05 * k; } }
  }
  int32_t *bi, *st, *wl; double *bc, *w;
  cudaMallocHost(&bi, NQ * 4); cudaMallocHost(&bc, NQ * 8); cudaMallocHost(&st, NQ * 32); cudaMallocHost(&wl, NQ * 4); cudaMallocHost(&w, (size_t)NQ * 15 * NT * 8);
  fot_batch_t b{}; b.n_q = NQ; b.n_v_max = 6; b.frenet = fr; b.target_speed = tg; b.limits = lm; b.stop_dist = sd; b.v_grid = vg; b.n_v = nv;
  b.dyn = dyn; b.S = 1; b.P = P; b.T_obs = TOBS; b.dyn_mode = FOT_DYN_SINGLE;
  fot_result_t r{}; r.best_idx = bi; r.best_cost = bc; r.stats = st; r.winner_len = wl; r.winner = w;
  for (int rep = 0; rep < 6; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    int rc = fot_plan_batch_host(h, &b, &r);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("rc=%d host call %.2f ms  (ok=%d of q0)\n", rc, ms, st[0]);
  }
  fot_destroy(h);
  return 0;
}
