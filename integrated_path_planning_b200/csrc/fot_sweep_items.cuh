// fot_sweep_items.cuh -- the sample-major sweep kernel (sm_100a, fp64).
//
// Same contract as fot_sweep (fot_kernels.cuh): one fused kernel for the whole candidate sweep of
// every query, per-block arg-min partials out, no trajectory written to HBM.  The work is laid out
// the other way round:
//
//   block  = (query, horizon T_j, chunk of terminal speeds)  -- or the query's brake ladder
//   thread = one (terminal speed v_k, sample t_n) "item": everything that depends on the
//            longitudinal profile only -- s(t_n), the spline reference point, heading, curvature,
//            1/s_dot -- is computed ONCE per item and lives in registers;
//   loop   = the lateral targets d_i.  The quintic solve is linear in the target, so
//            d_i(t_n) = A(t_n) + d_i * B(t_n)  (and likewise the two time derivatives): three FMAs
//            per candidate sample from six per-thread constants, no table, no per-candidate Horner.
//
// A candidate's validity flags (priority chain fp.py:964-991, silent drops :933-956) are an OR over
// its samples, i.e. over the threads of one speed: warp `redux.or` over the lanes of that speed,
// then one shared-memory atomic per 4 candidates.  The NaN-prefix truncation (fp.py:851-875) depends
// on the reference point only, so it is a per-speed constant, not per-candidate state.
//
// Collision (fp.py:1035-1233): a clean candidate's sample lies exactly on the normal of the
// reference line through the item's reference point at lateral offset |d| <= road half-width, so an
// obstacle sample can only touch it when it is within r along the tangent and W + r across it.
// Each item tests the obstacles of its own time step against that window (the obstacle tensor is
// read in the reference's own [S][P][T][2] layout: consecutive lanes are consecutive time steps,
// so the 16-byte loads coalesce; small fields are staged in shared memory by ONE bulk copy,
// cp.async.bulk + mbarrier) and pushes the few survivors into a block-wide queue; the queue is then
// drained by all threads, one entry each, with the reference's exact un-fused
// `dx*dx + dy*dy <= r^2` against every still-alive clean candidate of that speed.
//
// Numerics: costs are bit-identical to the reference (same arithmetic order, NumPy pairwise
// sums); the validity chain uses algebraically equal forms that differ by a few ulp (see
// DESIGN.md section 7); the winner's sequences are regenerated in reference order by fot_winner.
#pragma once
#include "fot_kernels.cuh"

namespace fot {

constexpr int kItemThreads = 320;   // largest block of fot_sweep_items
constexpr int kRefW = 8;            // reference row: rx ry cos sin | kappa s 1/s_dot kappa'
constexpr int kLabW = 6;            // lateral basis row: A B A' B' A'' B''
constexpr int kLabC = 10;           // lateral basis coefficients: a0 a1 a2 c3 c4 c5 | b3 b4 b5 | pad
constexpr unsigned F_DROP = 32u;    // silent drop (singular / non-finite / teleport), fp.py:826-833, :944-956

struct ItemGeom {
  int32_t ppc;               // terminal speeds (pairs) per grid block
  int32_t chunks;            // grid blocks per horizon
  int32_t grid_blocks;       // n_T * chunks
  int32_t ppb;               // brake horizons per brake block
  int32_t brake_blocks;
  int32_t blocks_per_query;
  int32_t threads;           // block size (multiple of 32, >= items of the largest block)
  int32_t pcap;              // max(ppc, ppb): per-pair table slots
  int32_t nlat_cap;          // lateral-basis slots: 1 for grid blocks, ppb for brake blocks
  int32_t jcap;              // max(n_d, ppb): lateral jerk-sum slots
  int32_t qcap;              // collision queue capacity (entries)
  int32_t ochunk;            // obstacles culled per queue round
  int32_t stage_dyn;         // 1: the query's obstacle block is staged in shared memory by one bulk copy
  int32_t spline_smem;       // 1: spline tables copied to shared memory
  int32_t vwords;            // u32 words per candidate of the chance-constraint violation bitmap (0: no budget)
  int32_t nw4, nwc;          // flag words (4 candidates each) / bit-mask words (32 candidates each) per pair
  int32_t n_bad;             // words of the NaN-trajectory bitmap
  // byte offsets into dynamic shared memory
  int32_t o_tt, o_ref, o_lab, o_labc, o_lonc, o_js, o_jp, o_dend, o_dgrid, o_vlast, o_spl, o_dyn;
  int32_t o_pi, o_flags, o_clean, o_hit, o_viol, o_queue, o_kobs, o_bad;
};

struct SplineView {
  const double *knots, *xa, *xb, *xc, *xd, *ya, *yb, *yc, *yd;
  int nx;
};
struct RefFast {
  double rx, ry, cth, sth, rk, rdk;
};
// Reference point at arc length s (cs.py:47-166, :215-288) with heading as a unit vector:
// cos(atan2(y', x')) = x'/|r'| (cc.py:128-129 takes cos/sin of the yaw).  NaN outside the knots.
__device__ __forceinline__ RefFast spline_ref_fast(const SplineView& V, double s) {
  RefFast o;
  const int nx = V.nx;
  if (!(s >= V.knots[0] && s <= V.knots[nx - 1])) {     // cs.py:62
    o.rx = o.ry = o.cth = o.sth = o.rk = o.rdk = qnan();
    return o;
  }
  int lo = 0, hi = nx;                                   // searchsorted(side='right') (cs.py:162)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (V.knots[mid] <= s) lo = mid + 1; else hi = mid;
  }
  int seg = lo - 1;
  seg = seg < 0 ? 0 : (seg > nx - 2 ? nx - 2 : seg);     // cs.py:165
  const double dx = s - V.knots[seg];
  const double dx2 = dx * dx, dx3 = dx2 * dx;
  const double xa = V.xa[seg], xb = V.xb[seg], xc = V.xc[seg], xd = V.xd[seg];
  const double ya = V.ya[seg], yb = V.yb[seg], yc = V.yc[seg], yd = V.yd[seg];
  o.rx = xa + xb * dx + xc * dx2 + xd * dx3;             // cs.py:73-74
  o.ry = ya + yb * dx + yc * dx2 + yd * dx3;
  const double x1 = xb + 2.0 * xc * dx + 3.0 * xd * dx2; // cs.py:100
  const double y1 = yb + 2.0 * yc * dx + 3.0 * yd * dx2;
  const double x2 = 2.0 * xc + 6.0 * xd * dx;            // cs.py:125
  const double y2 = 2.0 * yc + 6.0 * yd * dx;
  const double x3 = 6.0 * xd, y3 = 6.0 * yd;             // cs.py:149
  const double D = x1 * x1 + y1 * y1;
  const double rD = 1.0 / sqrt(D);
  o.cth = x1 * rD;
  o.sth = y1 * rD;
  const double iD15 = rD * rD * rD;                      // D ** -1.5
  o.rk = (y2 * x1 - x2 * y1) * iD15;                     // cs.py:246
  const double a = x1 * y2 - y1 * x2;
  const double b = x1 * y3 - y1 * x3;
  const double c = x1 * x2 + y1 * y2;
  o.rdk = b * iD15 - 3.0 * a * c * (iD15 * rD * rD);     // cs.py:273
  return o;
}

// 1/sqrt(x) to fp64 accuracy from the fp32 MUFU seed and two Newton steps (branch-free; the
// argument is q^2 + d'^2 of a sample that is either well inside the normal range or belongs to a
// candidate that is dropped / speed-rejected anyway).
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y = (double)rsqrtf((float)x);
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);
  y = fma(y, fma(-hx * y, y, 0.5), y);
  return y;
}

// NumPy pairwise sum (see np_pairwise_sum) of f(0..n-1) computed by the 8 lanes of an aligned lane
// group: lane a owns accumulator r_a, the combine tree and the sequential tail are NumPy's.
template <class F>
__device__ __forceinline__ double np_sum_8lanes(const F& f, int n, int sub, unsigned gmask) {
  double res;
  if (n > 128) {
    res = np_pairwise_sum(f, 0, n);          // rare: long time grids, every lane does the serial sum
  } else if (n < 8) {
    res = 0.0;
    for (int i = 0; i < n; ++i) res += f(i);
  } else {
    const int stop = n - (n % 8);
    double r = f(sub);
    for (int i = 8 + sub; i < stop; i += 8) r += f(i);
    r = r + __shfl_xor_sync(gmask, r, 1);    // (r0+r1), (r2+r3), ...
    r = r + __shfl_xor_sync(gmask, r, 2);    // (r0+r1)+(r2+r3), (r4+r5)+(r6+r7)
    r = r + __shfl_xor_sync(gmask, r, 4);
    res = r;
    for (int i = stop; i < n; ++i) res += f(i);
  }
  return res;
}

__global__ void __launch_bounds__(kItemThreads, 2)
fot_sweep_items(const Plan P, const Batch B, const Out O, const ItemGeom G) {
  extern __shared__ __align__(16) unsigned char smb[];
  double* tt = reinterpret_cast<double*>(smb + G.o_tt);        // [NT][kTT]
  double* ref = reinterpret_cast<double*>(smb + G.o_ref);      // [pcap][NT][kRefW]
  double* lab = reinterpret_cast<double*>(smb + G.o_lab);      // [nlat_cap][NT][kLabW]
  double* labc = reinterpret_cast<double*>(smb + G.o_labc);    // [nlat_cap][kLabC]
  double* lonc = reinterpret_cast<double*>(smb + G.o_lonc);    // [pcap][6]  a0..a4, s_dot at the last sample
  double* js = reinterpret_cast<double*>(smb + G.o_js);        // [pcap]
  double* jp = reinterpret_cast<double*>(smb + G.o_jp);        // [jcap]
  double* dend = reinterpret_cast<double*>(smb + G.o_dend);    // [jcap]
  double* dgrid = reinterpret_cast<double*>(smb + G.o_dgrid);  // [n_d]
  double* vlast = reinterpret_cast<double*>(smb + G.o_vlast);  // [pcap][n_d]  v^2 at the last kept sample
  double* spl = reinterpret_cast<double*>(smb + G.o_spl);      // [9][nx] when spline_smem
  const double2* dynst = reinterpret_cast<const double2*>(smb + G.o_dyn);   // [SP][T_obs] when stage_dyn
  int* pi_hold = reinterpret_cast<int*>(smb + G.o_pi);         // [pcap]
  int* pi_fn = pi_hold + G.pcap;                               // [pcap] first NaN sample (INT_MAX: none)
  unsigned* flags = reinterpret_cast<unsigned*>(smb + G.o_flags);   // [pcap][nw4]
  unsigned* cleanw = reinterpret_cast<unsigned*>(smb + G.o_clean);  // [pcap][nwc] kinematically clean
  unsigned* hitw = reinterpret_cast<unsigned*>(smb + G.o_hit);      // [pcap][nwc] decisive collision
  unsigned* viol = reinterpret_cast<unsigned*>(smb + G.o_viol);     // [pcap][n_d][vwords]
  unsigned* queue = reinterpret_cast<unsigned*>(smb + G.o_queue);   // [qcap]
  int* kobs = reinterpret_cast<int*>(smb + G.o_kobs);               // [NT]
  unsigned* bad = reinterpret_cast<unsigned*>(smb + G.o_bad);       // [n_bad]
  __shared__ int s_qcount[2];
  __shared__ int s_anybad;
  __shared__ int s_stats[FOT_N_STATS];
  __shared__ double s_cost[kItemThreads / 32];
  __shared__ int s_idx[kItemThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;

  const int NT = P.n_t_max;
  const int q = blockIdx.x / G.blocks_per_query;
  const int b = blockIdx.x % G.blocks_per_query;
  const int tid = threadIdx.x, lane = tid & 31, bd = blockDim.x;
  const double* fs = B.frenet + 6 * (size_t)q;
  const int n_v = B.n_v[q];
  const int n_d = P.cfg.n_d;
  const bool brake_blk = b >= G.grid_blocks;
  const size_t part = (size_t)q * G.blocks_per_query + b;

  // ---- which pairs does this block own? -----------------------------------------------------
  int jT = 0, k_lo = 0, n_k = 0, N = 0, n_dl = 0;
  if (!brake_blk) {
    jT = b / G.chunks;
    k_lo = (b - jT * G.chunks) * G.ppc;
    n_k = min(G.ppc, n_v - k_lo);
    N = P.n_steps[jT] + 1;
    n_dl = n_d;
  } else {
    const int b0 = (b - G.grid_blocks) * G.ppb;
    if (fs[1] > 0.1 && b0 < P.cfg.n_B) {                 // fp.py:469 BRAKE_MIN_SPEED
      k_lo = b0;
      n_k = min(G.ppb, P.cfg.n_B - b0);
    }
    N = P.cfg.n_total;
    n_dl = 1;
  }
  if (n_k <= 0) {                                        // uniform per block
    if (tid == 0) { O.part_cost[part] = INFINITY; O.part_idx[part] = -1; }
    return;
  }
  const int n_cand = n_k * n_dl;
  const int cand0 = brake_blk ? P.cfg.n_T * n_v * n_d + k_lo : (jT * n_v + k_lo) * n_d;   // generation order (fp.py:398-449)
  // A non-finite Frenet state makes every sample of every candidate non-finite: the reference drops
  // them all silently (empty / non-finite guards fp.py:933-946).
  const bool state_ok = fabs(fs[0]) + fabs(fs[1]) + fabs(fs[2]) + fabs(fs[3]) + fabs(fs[4]) + fabs(fs[5]) < INFINITY;
  if (!state_ok) {
    for (int c = tid; c < n_cand; c += bd) {
      if (O.cand_cat) O.cand_cat[(size_t)q * O.cand_stride + cand0 + c] = (uint8_t)FOT_CAT_DROP;
      if (O.cand_cost) O.cand_cost[(size_t)q * O.cand_stride + cand0 + c] = qnan();
    }
    if (tid == 0) { O.part_cost[part] = INFINITY; O.part_idx[part] = -1; }
    return;
  }

  const bool has_dyn = B.dyn_raw != nullptr;
  const int SP = has_dyn ? B.S * B.P : 0;
  const double2* dyn_q = has_dyn ? reinterpret_cast<const double2*>(B.dyn_raw) + (size_t)q * SP * B.T_obs : nullptr;

  // ---- phase 0: tables ------------------------------------------------------------------------
  if (tid == 0) {
    s_qcount[0] = 0; s_qcount[1] = 0; s_anybad = 0;
    if (G.stage_dyn) {
      mbar_init(&s_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t bytes = (uint32_t)SP * (uint32_t)B.T_obs * 16u;
      mbar_expect_tx(&s_bar, bytes);
      tma_bulk_g2s(smb + G.o_dyn, dyn_q, bytes, &s_bar);
    }
  }
  if (tid < FOT_N_STATS) s_stats[tid] = 0;
  for (int n = tid; n < NT; n += bd) {
    tt_fill(tt, n, P.cfg.dt);
    const double kf = rint(((double)n * P.cfg.dt) / P.cfg.dt);                 // fp.py:1226-1227
    const int kmax_i = B.T_obs > 0 ? B.T_obs - 1 : 0;
    kobs[n] = kf < 0.0 ? 0 : (kf > (double)kmax_i ? kmax_i : (int)kf);
  }
  for (int i = tid; i < n_d; i += bd) dgrid[i] = P.d_grid[i];
  for (int i = tid; i < G.pcap; i += bd) pi_fn[i] = 0x7fffffff;
  for (int i = tid; i < G.pcap * G.nw4; i += bd) flags[i] = 0u;
  for (int i = tid; i < G.pcap * G.nwc; i += bd) { cleanw[i] = 0u; hitw[i] = 0u; }
  for (int i = tid; i < G.pcap * n_d * G.vwords; i += bd) viol[i] = 0u;
  for (int i = tid; i < G.n_bad; i += bd) bad[i] = (!G.stage_dyn && B.dyn_bad) ? B.dyn_bad[(size_t)q * G.n_bad + i] : 0u;
  SplineView V;
  V.nx = P.cfg.nx;
  if (G.spline_smem) {
    const int nx = V.nx;
    for (int i = tid; i < nx; i += bd) {
      spl[i] = P.knots[i];
      spl[nx + i] = P.xa[i];      spl[3 * nx + i] = P.xc[i];
      spl[5 * nx + i] = P.ya[i];  spl[7 * nx + i] = P.yc[i];
      if (i < nx - 1) {
        spl[2 * nx + i] = P.xb[i]; spl[4 * nx + i] = P.xd[i];
        spl[6 * nx + i] = P.yb[i]; spl[8 * nx + i] = P.yd[i];
      }
    }
    V.knots = spl; V.xa = spl + nx; V.xb = spl + 2 * nx; V.xc = spl + 3 * nx; V.xd = spl + 4 * nx;
    V.ya = spl + 5 * nx; V.yb = spl + 6 * nx; V.yc = spl + 7 * nx; V.yd = spl + 8 * nx;
  } else {
    V.knots = P.knots; V.xa = P.xa; V.xb = P.xb; V.xc = P.xc; V.xd = P.xd;
    V.ya = P.ya; V.yb = P.yb; V.yc = P.yc; V.yd = P.yd;
  }
  __syncthreads();

  // quartic solve per pair (fp.py:619-647); lateral basis coefficients
  if (tid < n_k) {
    Lon L;
    if (!brake_blk)
      L = lon_solve(fs, B.v_grid[(size_t)q * B.n_v_max + k_lo + tid], P.T[jT], P.inv4 + 4 * jT, n_v == 1, N - 1);
    else
      L = lon_solve(fs, 0.0, P.Tb[k_lo + tid], P.inv4b + 4 * (k_lo + tid), true, P.n_steps_b[k_lo + tid]);
    double* lc = lonc + 6 * tid;
    lc[0] = L.a0; lc[1] = L.a1; lc[2] = L.a2; lc[3] = L.a3; lc[4] = L.a4;
    lc[5] = lon_p1(L, tt, N - 1);                                              // terminal speed of the cost (fp.py:724)
    pi_hold[tid] = L.hold;
    if (brake_blk) {                                                           // one lateral profile per brake horizon (fp.py:480-482)
      const Lat Lb = lat_solve(fs, fs[3], P.Tb[k_lo + tid], P.inv5b + 9 * (k_lo + tid), true, P.n_steps_b[k_lo + tid]);
      double* c = labc + kLabC * tid;
      c[0] = Lb.a0; c[1] = Lb.a1; c[2] = Lb.a2; c[3] = Lb.a3; c[4] = Lb.a4; c[5] = Lb.a5;
      c[6] = 0.0; c[7] = 0.0; c[8] = 0.0;
    }
  }
  if (!brake_blk && tid == bd - 1) {
    // d_i(t) = A(t) + d_i * B(t): the quintic's right-hand side is linear in the target (fp.py:676-683)
    const double T = P.T[jT];
    const double* Ai = P.inv5 + 9 * jT;
    const double a0 = fs[3], a1 = fs[4], a2 = fs[5] / 2.0;
    const double r0 = -a0 - a1 * T - a2 * T * T, r1 = -a1 - 2.0 * a2 * T, r2 = -2.0 * a2;
    double* c = labc;
    c[0] = a0; c[1] = a1; c[2] = a2;
    c[3] = fma(r2, Ai[2], fma(r1, Ai[1], r0 * Ai[0]));
    c[4] = fma(r2, Ai[5], fma(r1, Ai[4], r0 * Ai[3]));
    c[5] = fma(r2, Ai[8], fma(r1, Ai[7], r0 * Ai[6]));
    c[6] = Ai[0]; c[7] = Ai[3]; c[8] = Ai[6];
  }
  __syncthreads();

  // lateral basis table and the items' reference points
  const int n_lat = brake_blk ? n_k : 1;
  for (int idx = tid; idx < n_lat * N; idx += bd) {
    const int lt = idx / N, n = idx - lt * N;
    const int hold = brake_blk ? P.n_steps_b[k_lo + lt] : N - 1;
    const double t = tt[kTT * min(n, hold)];
    const double* c = labc + kLabC * lt;
    // Horner with running derivatives, A: a0..c5, B: b3..b5 (B has no terms below t^3)
    double pA = fma(c[5], t, c[4]), dA = c[5], ddA;
    ddA = dA;               dA = fma(dA, t, pA);  pA = fma(pA, t, c[3]);
    ddA = fma(ddA, t, dA);  dA = fma(dA, t, pA);  pA = fma(pA, t, c[2]);
    ddA = fma(ddA, t, dA);  dA = fma(dA, t, pA);  pA = fma(pA, t, c[1]);
    ddA = fma(ddA, t, dA);  dA = fma(dA, t, pA);  pA = fma(pA, t, c[0]);
    double pB = fma(c[8], t, c[7]), dB = c[8], ddB;
    ddB = dB;               dB = fma(dB, t, pB);  pB = fma(pB, t, c[6]);
    ddB = fma(ddB, t, dB);  dB = fma(dB, t, pB);  pB = pB * t;
    ddB = fma(ddB, t, dB);  dB = fma(dB, t, pB);  pB = pB * t;
    ddB = fma(ddB, t, dB);  dB = fma(dB, t, pB);  pB = pB * t;
    const bool held = n > hold;                                                // fp.py:487-499 brake padding
    double* r = lab + ((size_t)lt * NT + n) * kLabW;
    r[0] = pA; r[1] = pB;
    r[2] = held ? 0.0 : dA;        r[3] = held ? 0.0 : dB;
    r[4] = held ? 0.0 : 2.0 * ddA; r[5] = held ? 0.0 : 2.0 * ddB;
  }
  const int n_items = n_k * N;
  const bool active = tid < n_items;
  const int p = active ? tid / N : -1;                                         // pair slot of this item
  const int n = active ? tid - p * N : 0;                                      // its sample
  double i_rx = 0, i_ry = 0, i_cth = 0, i_sth = 0, i_rk = 0, i_rdk = 0, i_sd = 0, i_sdd = 0, i_isd = 0;
  if (active) {
    Lon L;
    const double* lc = lonc + 6 * p;
    L.a0 = lc[0]; L.a1 = lc[1]; L.a2 = lc[2]; L.a3 = lc[3]; L.a4 = lc[4]; L.hold = pi_hold[p];
    const double s = lon_p0(L, tt, n);
    i_sd = lon_p1(L, tt, n);
    i_sdd = lon_p2(L, tt, n);
    const RefFast rp = spline_ref_fast(V, s);
    i_rx = rp.rx; i_ry = rp.ry; i_cth = rp.cth; i_sth = rp.sth; i_rk = rp.rk; i_rdk = rp.rdk;
    i_isd = fabs(i_sd) > 1e-3 ? 1.0 / i_sd : 0.0;                              // fp.py:792 EPS_S_DOT
    double* r = ref + ((size_t)p * NT + n) * kRefW;
    r[0] = i_rx; r[1] = i_ry; r[2] = i_cth; r[3] = i_sth; r[4] = i_rk; r[5] = s; r[6] = i_isd; r[7] = i_rdk;
    if (i_rx != i_rx || i_ry != i_ry) atomicMin(&pi_fn[p], n);                 // fp.py:851-866
  }
  __syncthreads();

  // ---- phase 1: validity chain, thread = item, loop = lateral targets --------------------------
  const int fn = active ? pi_fn[p] : 0;
  const int keep = fn == 0x7fffffff ? N : (fn >= 2 ? fn : 0);                  // fp.py:866
  const bool valid = active && n < keep;
  const bool chk = valid && n >= 1;                                            // limits skip index 0 (fp.py:964-983)
  const double* lim = B.limits + 4 * (size_t)q;
  const double vmax = lim[0], amax = lim[1], kmax = lim[2], latmax = lim[3];
  const double vmax2 = vmax * vmax;
  const double road_thr = P.cfg.max_road_width + 1e-9;                         // fp.py:982
  const double tele_thr = fmax(vmax, P.cfg.max_speed) * P.cfg.dt * 3.0;        // fp.py:955
  const double tele2 = tele_thr * tele_thr;
  const double stop_dist = B.stop_dist[q];
  const bool want_vlast = (stop_dist == stop_dist) && valid && n == keep - 1;
  const int lt_own = brake_blk ? (active ? p : 0) : 0;
  const double* labr = lab + ((size_t)lt_own * NT + n) * kLabW;
  const double A0 = labr[0], B0 = labr[1], A1 = labr[2], B1 = labr[3], A2 = labr[4], B2 = labr[5];
  const double* labp = n >= 1 ? labr - kLabW : labr;
  const double Ap = labp[0], Bp = labp[1];
  const double* refp = ref + ((size_t)(active ? p : 0) * NT + (n >= 1 ? n - 1 : n)) * kRefW;
  const double rxp = chk ? refp[0] : i_rx, ryp = chk ? refp[1] : i_ry, cthp = chk ? refp[2] : i_cth, sthp = chk ? refp[3] : i_sth;
  const double sd2 = i_sd * i_sd, isd2 = i_isd * i_isd;
  const unsigned keepmask = chk ? 0xffu : F_DROP;                              // at n = 0 only the drop guards apply
  const unsigned segmask = __match_any_sync(0xffffffffu, p);
  const bool seg_leader = (__ffs(segmask) - 1) == lane;
  const double kTan01Sq = 0.010067046422495888;                                // tan(0.1)^2
  unsigned* flags_p = flags + (size_t)(active ? p : 0) * G.nw4;

  for (int i0 = 0; i0 < n_dl; i0 += 4) {
    unsigned acc = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u;
      if (i < n_dl && valid) {
        const double di = brake_blk ? 0.0 : dgrid[i];
        const double d = fma(di, B0, A0), d1 = fma(di, B1, A1), d2 = fma(di, B2, A2);
        const double dprev = fma(di, Bp, Ap);
        const double qq = fma(-i_rk, d, 1.0);                                  // 1 - kappa_r d
        const double x = fma(-i_sth, d, i_rx), y = fma(i_cth, d, i_ry);        // cc.py:131-132
        const double xp = fma(-sthp, dprev, rxp), yp = fma(cthp, dprev, ryp);
        const double ex = x - xp, ey = y - yp;
        const double step2 = fma(ex, ex, ey * ey);                             // fp.py:954 (squared)
        const double dpr = d1 * i_isd;                                         // d' (fp.py:792-799)
        const double dpp = (d2 - dpr * i_sdd) * isd2;                          // d''
        const double h2 = fma(qq, qq, dpr * dpr);
        const double rh = rsqrt_nr(h2);
        const double h = h2 * rh;                                              // hypot(q, d') = q / cos(delta)
        const double m = fma(i_rdk, d, i_rk * dpr);                            // kappa_r' d + kappa_r d'
        const double kap = fma(fma(dpp, qq, m * dpr) * rh, rh, i_rk) * rh;     // cc.py:144-147
        const double v2 = sd2 * h2;                                            // v^2 (cc.py:150-152)
        const double t2 = fma(dpr, fma(h, kap, -i_rk), -m);
        const double aq = h * fma(sd2, t2, i_sdd * qq);                        // a * q (cc.py:155-157)
        const bool fast = v2 > 0.25;                                           // v > 0.5 (fp.py:1019)
        const double akap = fabs(kap);
        unsigned f = 0u;
        f |= ((qq <= 0.05) & (fabs(qq) < INFINITY)) ? F_DROP : 0u;             // fp.py:826-833
        f |= !(fabs(v2) + fabs(aq) + akap < INFINITY) ? F_DROP : 0u;           // fp.py:944-946
        f |= (step2 > tele2) ? F_DROP : 0u;                                    // fp.py:953-956
        f |= (v2 > vmax2) ? F_SPEED : 0u;                                      // fp.py:964
        f |= (fabs(aq) > amax * qq) ? F_ACCEL : 0u;                            // fp.py:966
        f |= (fast & (akap > kmax)) ? F_CURV : 0u;                             // fp.py:1020
        f |= (v2 * akap > latmax) ? F_LAT : 0u;                                // fp.py:975
        f |= (fabs(d) > road_thr) ? F_ROAD : 0u;                               // fp.py:982
        if (chk & !fast & !(f & F_CURV)) {                                     // low-speed regime fp.py:1022-1032
          const double* rp1 = ref + ((size_t)p * NT + n - 1) * kRefW;
          const double s_now = rp1[kRefW + 5], s_prev = rp1[5];
          if (fabs(d - dprev) > fmax(1.5 * fabs(s_now - s_prev), 0.02)) {
            f |= F_CURV;
          } else {
            // |wrap(yaw_n - yaw_{n-1})| is the angle between the heading vectors u = R(theta_r)(q, d')
            const double q_prev = fma(-rp1[4], dprev, 1.0);
            const double dp_prev = fma(di, labr[3 - kLabW], labr[2 - kLabW]) * rp1[6];
            const double ux = i_cth * qq - i_sth * dpr, uy = i_sth * qq + i_cth * dpr;
            const double uxp = cthp * q_prev - sthp * dp_prev, uyp = sthp * q_prev + cthp * dp_prev;
            const double cr = uxp * uy - uyp * ux, dt_ = uxp * ux + uyp * uy;
            if (kmax * kmax * step2 <= 0.01) {
              // the threshold is the 0.1 rad floor: angle > 0.1 <=> dot <= 0 or cross^2 > tan(0.1)^2 dot^2
              if (dt_ <= 0.0 || cr * cr > kTan01Sq * dt_ * dt_) f |= F_CURV;
            } else if (fabs(atan2(cr, dt_)) > kmax * sqrt(step2)) {
              f |= F_CURV;
            }
          }
        }
        acc |= (f & keepmask) << (8 * u);
        if (want_vlast) vlast[(size_t)p * n_d + i] = v2;
      }
    }
    const unsigned red = __reduce_or_sync(segmask, acc);
    if (seg_leader && active && red) atomicOr(&flags_p[i0 >> 2], red);
  }
  // Samples beyond the NaN prefix that are inside the spline domain again still count for the
  // candidate-wide singularity guard (fp.py:826-833 runs before the truncation).  Essentially never.
  if (active && !valid && i_rx == i_rx && keep > 0) {
    for (int i = 0; i < n_dl; ++i) {
      const double qq = fma(-i_rk, fma(brake_blk ? 0.0 : dgrid[i], B0, A0), 1.0);
      if ((qq <= 0.05) & (fabs(qq) < INFINITY)) atomicOr(&flags_p[i >> 2], F_DROP << (8 * (i & 3)));
    }
  }
  __syncthreads();

  // ---- phase 2: collision (fp.py:1035-1233) -------------------------------------------------
  for (int c = tid; c < n_cand; c += bd) {
    const int cp = c / n_dl, ci = c - cp * n_dl;
    const unsigned byte = (flags[(size_t)cp * G.nw4 + (ci >> 2)] >> (8 * (ci & 3))) & 0xffu;
    const int cfn = pi_fn[cp];
    if (byte == 0u && (cfn == 0x7fffffff || cfn >= 2)) atomicOr(&cleanw[(size_t)cp * G.nwc + (ci >> 5)], 1u << (ci & 31));
  }
  __syncthreads();
  const bool dist_mode = (B.dyn_mode == FOT_DYN_DISTRIBUTION);
  const int max_viol = dist_mode ? (int)floor(P.cfg.chance_epsilon * (double)B.S) : 0;   // fp.py:1114
  const int n_circ = P.cfg.n_circles;
  double max_off = 0.0;                                  // footprint circles sit within max|offset| of the path point
  for (int i = 0; i < n_circ; ++i) max_off = fmax(max_off, fabs(P.cfg.circle_offsets[i]));
  bool pair_clean = false;
  if (valid)
    for (int w = 0; w < G.nwc; ++w) pair_clean |= cleanw[(size_t)p * G.nwc + w] != 0u;
  int qsel = 0;

  // exact test of queue entry (item, obstacle) against every live clean candidate of the item's pair
  auto process = [&](int it, double ox, double oy, int oj, double r2, bool budget) {
    const int ep = it / N, en = it - ep * N;
    const double* r = ref + ((size_t)ep * NT + en) * kRefW;
    const double* l = lab + ((size_t)(brake_blk ? ep : 0) * NT + en) * kLabW;
    const double rx = r[0], ry = r[1], cth = r[2], sth = r[3];
    const double a0 = l[0], b0 = l[1];
    for (int w = 0; w < G.nwc; ++w) {
      unsigned mbits = cleanw[(size_t)ep * G.nwc + w];
      if (!budget) mbits &= ~hitw[(size_t)ep * G.nwc + w];
      while (mbits) {
        const int bit = __ffs(mbits) - 1;
        mbits &= mbits - 1u;
        const int i = w * 32 + bit;
        const double di = brake_blk ? 0.0 : dgrid[i];
        const double d = fma(di, b0, a0);
        const double x = fma(-sth, d, rx), y = fma(cth, d, ry);
        bool hit = false;
        if (n_circ == 0) {
          const double dx = x - ox, dy = y - oy;
          hit = dx * dx + dy * dy <= r2;                                       // fp.py:1196-1198, :1231-1233
        } else {                                                               // fp.py:1158-1167
          const double dpr = fma(di, l[3], l[2]) * r[6];
          const double qq = fma(-r[4], d, 1.0);
          const double rh = 1.0 / sqrt(fma(qq, qq, dpr * dpr));
          const double hx = (cth * qq - sth * dpr) * rh, hy = (sth * qq + cth * dpr) * rh;   // (cos yaw, sin yaw)
          for (int ci = 0; ci < n_circ && !hit; ++ci) {
            const double dx = (x + P.cfg.circle_offsets[ci] * hx) - ox, dy = (y + P.cfg.circle_offsets[ci] * hy) - oy;
            hit = dx * dx + dy * dy <= r2;
          }
        }
        if (hit) {
          if (!budget) atomicOr(&hitw[(size_t)ep * G.nwc + w], 1u << bit);
          else { const int sidx = oj / B.P; atomicOr(&viol[((size_t)ep * n_d + i) * G.vwords + (sidx >> 5)], 1u << (sidx & 31)); }
        }
      }
    }
  };

  // one obstacle set: `n_obs` obstacles, obstacle j of item (pair, n) at src[j * stride + koff(n)]
  auto run_set = [&](const double2* src, int n_obs, int stride, bool timed, bool check_bad, double r2, bool budget) {
    const double rc = sqrt(r2) * (1.0 + 1e-9) + 1e-9 + max_off;
    const double wc = fmax(P.cfg.max_road_width + 1e-9, fabs(fs[3])) + rc;
    const int koff = timed ? kobs[n] : 0;
    const bool cull = valid && pair_clean;
    for (int j0 = 0; j0 < n_obs; j0 += G.ochunk) {
      const int j1 = min(n_obs, j0 + G.ochunk);
      if (cull) {
        const double2* sp = src + (size_t)j0 * stride + koff;
#pragma unroll 4
        for (int j = j0; j < j1; ++j, sp += stride) {
          const double2 o = *sp;
          const double ex = o.x - i_rx, ey = o.y - i_ry;
          const double al = fma(ex, i_cth, ey * i_sth), ac = fma(ey, i_cth, -(ex * i_sth));
          if ((fabs(al) <= rc) & (fabs(ac) <= wc)) {                           // NaN -> false
            if (check_bad && ((bad[j >> 5] >> (j & 31)) & 1u)) continue;       // fp.py:1211-1222 NaN trajectory
            const int slot = atomicAdd(&s_qcount[qsel], 1);
            if (slot < G.qcap) queue[slot] = ((unsigned)tid << 16) | (unsigned)(j - j0);
            else process(tid, o.x, o.y, j, r2, budget);                        // queue full: test right here
          }
        }
      }
      __syncthreads();
      const int cnt = min(s_qcount[qsel], G.qcap);
      if (tid == 0) s_qcount[qsel ^ 1] = 0;
      for (int e = tid; e < cnt; e += bd) {
        const unsigned ent = queue[e];
        const int it = (int)(ent >> 16), j = j0 + (int)(ent & 0xffffu);
        const int en = it % N;
        const double2 o = src[(size_t)j * stride + (timed ? kobs[en] : 0)];
        process(it, o.x, o.y, j, r2, budget);
      }
      qsel ^= 1;
      if (j1 < n_obs && !budget) {
        // leave early once every clean candidate of the block has its decisive hit
        bool alive = false;
        if (tid < n_k * G.nwc) alive = (cleanw[tid] & ~hitw[tid]) != 0u;
        if (!__syncthreads_or(alive ? 1 : 0)) break;
      } else {
        __syncthreads();
      }
    }
  };

  if (B.n_static > 0 && B.static_raw) {
    const int qs = B.static_per_query ? q : 0;
    run_set(reinterpret_cast<const double2*>(B.static_raw) + (size_t)qs * B.n_static, B.n_static, 1, false, false,
            P.cfg.collide_r2, false);
  }
  if (has_dyn) {
    if (G.stage_dyn) {
      mbar_wait(&s_bar, 0u);
      // NaN-trajectory rule (fp.py:1211-1222): np.min/np.max over a trajectory with a NaN anywhere is
      // NaN, the AABB overlap test is then False and that pedestrian never collides.
      for (int e = tid; e < SP * B.T_obs; e += bd) {
        const double2 o = dynst[e];
        if (o.x != o.x || o.y != o.y) { const int j = e / B.T_obs; atomicOr(&bad[j >> 5], 1u << (j & 31)); s_anybad = 1; }
      }
      __syncthreads();
    } else if (tid == 0) {
      int any = 0;
      for (int i = 0; i < G.n_bad; ++i) any |= bad[i] != 0u;
      s_anybad = any;
    }
    if (!G.stage_dyn) __syncthreads();
    const bool check_bad = s_anybad != 0;
    const bool budget = dist_mode && max_viol > 0;
    run_set(G.stage_dyn ? dynst : dyn_q, SP, B.T_obs, true, check_bad,
            dist_mode ? P.cfg.collide_r2 : P.cfg.collide_r2_single, budget);   // fp.py:1099-1104
  }

  // ---- cost pieces: jerk sums in NumPy's pairwise order, 8 lanes per profile (fp.py:718-722) ----
  {
    const int n_prof = n_k + (brake_blk ? n_k : n_d);
    const int sub = tid & 7;
    const unsigned gmask = 0xffu << (lane & 24);
    for (int pr0 = 0; pr0 < n_prof; pr0 += bd / 8) {
      const int pr = pr0 + (tid >> 3);
      if (pr < n_k) {                                    // longitudinal
        Lon L;
        const double* lc = lonc + 6 * pr;
        L.a0 = lc[0]; L.a1 = lc[1]; L.a2 = lc[2]; L.a3 = lc[3]; L.a4 = lc[4]; L.hold = pi_hold[pr];
        auto jerk2 = [&](int k) { const double j = lon_p3(L, tt, k); return j * j; };
        const double sres = np_sum_8lanes(jerk2, N, sub, gmask);
        if (sub == 0) js[pr] = sres;
      } else if (pr < n_prof) {                          // lateral
        const int li = pr - n_k;
        Lat L;
        if (brake_blk) L = lat_solve(fs, fs[3], P.Tb[k_lo + li], P.inv5b + 9 * (k_lo + li), true, P.n_steps_b[k_lo + li]);
        else L = lat_solve(fs, dgrid[li], P.T[jT], P.inv5 + 9 * jT, n_d == 1, N - 1);
        auto jerk2 = [&](int k) { const double j = lat_p3(L, tt, k); return j * j; };
        const double sres = np_sum_8lanes(jerk2, N, sub, gmask);
        if (sub == 0) { jp[li] = sres; dend[li] = lat_p0(L, tt, N - 1); }
      }
    }
  }
  __syncthreads();

  // ---- phase 3: category, cost, block arg-min, histogram ----------------------------------------
  double my_cost = INFINITY;
  int my_idx = 0x7fffffff;
  for (int c = tid; c < n_cand; c += bd) {
    const int cp = c / n_dl, ci = c - cp * n_dl;
    const int li = brake_blk ? cp : ci;
    // cost on the un-truncated profile (fp.py:703-734)
    const double Jp = jp[li], d_end = dend[li];
    const double Jd = d_end * d_end;
    const double Js = js[cp];
    const double dv = B.target[q] - lonc[6 * cp + 5];
    const double Jv = dv * dv;
    const double Jt = tt[kTT * (N - 1)];
    const double lat_cost = P.cfg.k_j * Jp + P.cfg.k_t * Jt + P.cfg.k_d * Jd;
    const double lon_cost = P.cfg.k_j * Js + P.cfg.k_t * Jt + P.cfg.k_s_dot * Jv;
    const double cost = P.cfg.k_lat * lat_cost + P.cfg.k_lon * lon_cost;
    const unsigned byte = (flags[(size_t)cp * G.nw4 + (ci >> 2)] >> (8 * (ci & 3))) & 0xffu;
    const int cfn = pi_fn[cp];
    const int ckeep = cfn == 0x7fffffff ? N : (cfn >= 2 ? cfn : 0);
    int cat;
    if (ckeep == 0 || (byte & F_DROP)) cat = FOT_CAT_DROP;                     // fp.py:831-833, :933, :944, :953
    else if (byte & F_SPEED) cat = FOT_CAT_SPEED;
    else if (byte & F_ACCEL) cat = FOT_CAT_ACCEL;
    else if (byte & F_CURV) cat = FOT_CAT_CURV;
    else if (byte & F_LAT) cat = FOT_CAT_LAT;
    else if (byte & F_ROAD) cat = FOT_CAT_ROAD;
    else {
      bool hit = (hitw[(size_t)cp * G.nwc + (ci >> 5)] >> (ci & 31)) & 1u;
      if (G.vwords > 0) {
        int nv = 0;
        for (int w = 0; w < G.vwords; ++w) nv += __popc(viol[((size_t)cp * n_d + ci) * G.vwords + w]);
        hit = hit || nv > max_viol;                                            // fp.py:1113-1124
      }
      if (hit) {
        cat = FOT_CAT_COLL;                                                    // fp.py:986-989
      } else {
        cat = FOT_CAT_OK;
        if (stop_dist == stop_dist) {                                          // fp.py:307-324
          const double v_last = sqrt(vlast[(size_t)cp * n_d + ci]);
          const double s_span = ref[((size_t)cp * NT + ckeep - 1) * kRefW + 5] - ref[(size_t)cp * NT * kRefW + 5];
          if (!(v_last <= 0.15 && s_span <= stop_dist + 1e-6)) cat = FOT_CAT_STOP;
        }
      }
    }
    const int cand_idx = cand0 + c;
    if (cat < FOT_N_STATS) atomicAdd(&s_stats[cat], 1);
    if (O.cand_cat) O.cand_cat[(size_t)q * O.cand_stride + cand_idx] = (uint8_t)cat;
    if (O.cand_cost) O.cand_cost[(size_t)q * O.cand_stride + cand_idx] = cost;
    if (cat == FOT_CAT_OK && cost < INFINITY) argmin_merge(my_cost, my_idx, cost, cand_idx);
  }
  for (int off = 16; off > 0; off >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, off);
    const int oi = __shfl_down_sync(0xffffffffu, my_idx, off);
    argmin_merge(my_cost, my_idx, oc, oi);
  }
  if (lane == 0) { s_cost[tid >> 5] = my_cost; s_idx[tid >> 5] = my_idx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (bd >> 5); ++w) argmin_merge(my_cost, my_idx, s_cost[w], s_idx[w]);
    O.part_cost[part] = my_cost;
    O.part_idx[part] = (my_idx == 0x7fffffff) ? -1 : my_idx;
  }
  if (tid < FOT_N_STATS && s_stats[tid] != 0) atomicAdd(&O.stats[(size_t)q * FOT_N_STATS + tid], s_stats[tid]);
}

// NaN-trajectory bitmap for obstacle fields too large to stage in shared memory: one warp per
// predicted trajectory, bit j of bad[q] set when any of its samples is NaN (fp.py:1211-1222).
__global__ void fot_bad_prepass(const double2* __restrict__ dyn, unsigned* __restrict__ bad, int n_q, int SP,
                                int T_obs, int n_bad) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)n_q * SP) return;
  const int q = (int)(warp / SP), j = (int)(warp % SP);
  const double2* src = dyn + (size_t)warp * T_obs;
  bool b = false;
  for (int k = lane; k < T_obs; k += 32) {
    const double2 o = src[k];
    b |= (o.x != o.x) || (o.y != o.y);
  }
  b = __any_sync(0xffffffffu, b);
  if (lane == 0 && b) atomicOr(&bad[(size_t)q * n_bad + (j >> 5)], 1u << (j & 31));
}

}  // namespace fot
