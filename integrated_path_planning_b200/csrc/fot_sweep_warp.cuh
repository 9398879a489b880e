// fot_sweep_warp.cuh -- the sample-major sweep with two barriers per block and a barrier-free collision queue
// (sm_100a, fp64).
//
// Same contract, same work decomposition and the same arithmetic as fot_sweep_items (fot_sweep_items.cuh: block =
// the pairs of one horizon, thread = one (pair, sample) item, loop = lateral targets, affine-in-d_i validity chain in
// squared form, tangent-frame cull, exact `dx*dx + dy*dy <= r^2` for the survivors).  What changes is the
// synchronisation.  fot_sweep_items runs every block through six block-wide barriers (reset | items | validity |
// cull | exact tests | category) and ncu shows the price: 3.1 barrier-stall cycles per issued instruction, 30 % of all
// warp time (profiles/r1) -- warps that skip the validity loop, or whose samples are far from every pedestrian, wait
// for the ones that are busy.  Here a block has TWO barriers, and the imbalanced part runs between them without one:
//
//   B   per item: quartic solve, reference point, lateral basis -> item row in shared memory
//   --- barrier (i): rows and the pairs' NaN prefixes are complete
//   CD  per WARP: validity screens and loop (flags by shared atomics, as before), low-speed units of its own items;
//       then the warp boxes its own 32 reference points, lists the obstacles whose trajectory box meets that box
//       (ballot compaction into a warp-private list -- tighter than a block-wide list) and culls them against its own
//       items.  Survivors go as (item, obstacle) entries into ONE block-wide queue, and every warp that has finished
//       producing turns consumer: 32 tickets at a time, exact tests against the clean candidates of the entry's pair,
//       until all warps have produced and the queue is empty.  A slot is valid when it is non-zero -- no barrier
//       between producing and consuming, so the warps that had little to do take over the tests of the others.
//       A consumer works with the flags as they are at that moment: a candidate that another warp flags later may
//       get an exact test it did not need -- harmless, the validity categories outrank the collision category
//       (fp.py:964-991) -- and the lateral window of the cull only ever narrows as flags arrive.
//   --- barrier (ii): flags, hit words, violation bitmaps are final
//   E   category, cost, arg-min, histogram (unchanged)
//
// and none between E of one block and B of the next: the per-block state (flags, hit words, first-NaN slots, staged
// cost entries, terminal speeds, queue counters) is double-buffered, and the buffer of block b+1 is cleared by the
// threads of block b right after barrier (i), when every thread is provably done with block b-1.
//
// Shapes: obstacle entries per query (static + S*P) up to kWarpListCap; larger fields run fot_sweep_items, whose
// block-wide lists and multi-round queue are made for them.
#pragma once
#include "fot_sweep_items.cuh"

namespace fot {

constexpr int kWarpListCap = 512;    // obstacle entries a warp-private list can hold
constexpr int kWarpQueue = 32;       // warp-private scratch words
constexpr int kBlockQueue = 2048;    // slots of the block-wide (item, obstacle) queue; beyond it entries are tested in place

struct WarpGeom {
  int32_t ppc, chunks, grid_blocks, ppb, brake_blocks, blocks_per_query, bpc, ctas_per_query, threads, pcap, ct_lcap;
  int32_t lcap;                // list capacity per warp (static + dynamic entries of one query)
  int32_t stage_dyn, spline_smem, vwords, nw4, nwc;
  // byte offsets into dynamic shared memory, once per CTA
  int32_t o_row, o_dgrid, o_spl, o_dyn, o_box, o_wlist, o_wq, o_bq;
  // per-block state: two copies, buf_bytes apart, starting at o_buf; offsets inside one copy
  int32_t o_buf, buf_bytes;
  int32_t b_fnr, b_flags, b_hit, b_viol, b_dirty, b_qctl, n_zero;   // zero-initialised region: fnr | flags | hit | viol | dirty | queue counters
  int32_t b_sdl, b_ct, b_vlast, b_span;
  int32_t fused_box, gate_q0, gate_per;
  uint32_t gate_epoch;
  unsigned* gate;
};

// nibble with bit k set when byte k of x is non-zero
__device__ __forceinline__ unsigned nonzero_bytes(unsigned x) {
  const unsigned t = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
  return (((t >> 7) * 0x00204081u) >> 21) & 0xfu;
}

template <bool kFused>
__global__ void __launch_bounds__(kItemThreads, FOT_ITEM_MIN_CTAS)
fot_sweep_warp(const Plan P, const Batch B, const Out O, const WarpGeom G) {
  extern __shared__ __align__(16) unsigned char smb[];
  double* row = reinterpret_cast<double*>(smb + G.o_row);      // [pcap][NT][kRowW]
  double* dgrid = reinterpret_cast<double*>(smb + G.o_dgrid);  // [n_d]
  double* spl = reinterpret_cast<double*>(smb + G.o_spl);      // [9][nx] when spline_smem
  const double2* dynst = reinterpret_cast<const double2*>(smb + G.o_dyn);   // [SP][T_obs] when stage_dyn
  float4* sbox = reinterpret_cast<float4*>(smb + G.o_box);     // [SP] trajectory boxes when fused_box
  __shared__ int s_stats[FOT_N_STATS];
  __shared__ double s_cost[kItemThreads / 32];
  __shared__ int s_idx[kItemThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_abort;

  const int NT = P.n_t_max;
  const int q = blockIdx.x / G.ctas_per_query;
  const int cta = blockIdx.x - q * G.ctas_per_query;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bd = blockDim.x;
  const unsigned lt_mask = (1u << lane) - 1u;
  const double* fs = B.frenet + 6 * (size_t)q;
  const int n_v = B.n_v[q];
  const int n_d = P.cfg.n_d;
  const double dt = P.cfg.dt;
  const size_t part = (size_t)q * G.ctas_per_query + cta;
  const int b_first = cta * G.bpc, b_last = min(G.blocks_per_query, b_first + G.bpc);
  const bool state_ok = fabs(fs[0]) + fabs(fs[1]) + fabs(fs[2]) + fabs(fs[3]) + fabs(fs[4]) + fabs(fs[5]) < INFINITY;
  unsigned* wlist = reinterpret_cast<unsigned*>(smb + G.o_wlist) + warp * G.lcap;     // this warp's obstacle list
  unsigned* wq = reinterpret_cast<unsigned*>(smb + G.o_wq) + warp * kWarpQueue;       // this warp's scratch (slow-item slots)
  unsigned* bq = reinterpret_cast<unsigned*>(smb + G.o_bq);                           // block-wide queue of (item, obstacle) entries

  const bool has_dyn = B.dyn_raw != nullptr;
  const int SP = has_dyn ? B.S * B.P : 0;
  const int M = B.static_raw ? B.n_static : 0;
  const double2* dyn_q = has_dyn ? reinterpret_cast<const double2*>(B.dyn_raw) + (size_t)q * SP * B.T_obs : nullptr;
  const double2* stat_q = M > 0 ? reinterpret_cast<const double2*>(B.static_raw) + (size_t)(B.static_per_query ? q : 0) * M : nullptr;

  // ---- once per CTA: obstacle block in flight, grids and spline tables, both state buffers cleared --------
  if (kFused && G.gate) {
    if (tid == 0) {
      const unsigned* flag = G.gate + (G.gate_q0 + q) / G.gate_per;
      int abort_ = 0;
      long long t0 = 0;
      for (unsigned spins = 0;; ++spins) {
        unsigned seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if (seen == G.gate_epoch) break;
        __nanosleep(spins < 64 ? 100 : FOT_GATE_SLEEP_NS);
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > kGateTimeoutNs) { abort_ = 1; G.gate[kGateSlices] = 1u; break; }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");
      s_abort = abort_;
    }
    __syncthreads();
    if (s_abort) return;
  }
  if (tid == 0 && G.stage_dyn && state_ok) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)SP * (uint32_t)B.T_obs * 16u;
    mbar_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(smb + G.o_dyn, dyn_q, bytes, &s_bar);
  }
  if (tid < FOT_N_STATS) s_stats[tid] = 0;
  for (int i = tid; i < n_d; i += bd) dgrid[i] = P.d_grid[i];
  if (G.spline_smem) {
    const int nx = P.cfg.nx;
    for (int i = tid; i < nx; i += bd) {
      spl[i] = P.knots[i];
      spl[nx + i] = P.xa[i];      spl[3 * nx + i] = P.xc[i];
      spl[5 * nx + i] = P.ya[i];  spl[7 * nx + i] = P.yc[i];
      if (i < nx - 1) {
        spl[2 * nx + i] = P.xb[i]; spl[4 * nx + i] = P.xd[i];
        spl[6 * nx + i] = P.yb[i]; spl[8 * nx + i] = P.yd[i];
      }
    }
  }
  for (int i = tid; i < 2 * (G.buf_bytes / 4); i += bd) {
    // only the zero regions need it, but the buffers are small
    reinterpret_cast<unsigned*>(smb + G.o_buf)[i] = 0u;
  }
  for (int i = tid; i < kBlockQueue; i += bd) bq[i] = 0u;          // an empty slot is zero; consumers hand slots back empty
  if (!kFused && state_ok)
    for (int j = tid; j < SP; j += bd) sbox[j] = B.dyn_box[(size_t)q * SP + j];      // fot_prepass boxes: read by every warp of every block
  if (kFused) {
    __syncthreads();                                     // the mbarrier thread 0 initialised is visible to every warp
    // box every predicted trajectory of the staged obstacle block once per CTA (what fot_prepass does for a resident
    // tensor): one warp per trajectory, fp32 rounded outward, NaN trajectory -> NaN box
    if (state_ok && G.stage_dyn) {
      mbar_wait(&s_bar, 0u);
      for (int j = warp; j < SP; j += bd >> 5) {
        const double2* src = dynst + (size_t)j * B.T_obs;
        double xlo = INFINITY, xhi = -INFINITY, ylo = INFINITY, yhi = -INFINITY;
        bool bad = false;
        for (int k = lane; k < B.T_obs; k += 32) {
          const double2 o = src[k];
          bad |= (o.x != o.x) || (o.y != o.y);
          xlo = fmin(xlo, o.x); xhi = fmax(xhi, o.x); ylo = fmin(ylo, o.y); yhi = fmax(yhi, o.y);
        }
        for (int off = 16; off > 0; off >>= 1) {
          xlo = fmin(xlo, __shfl_xor_sync(0xffffffffu, xlo, off)); xhi = fmax(xhi, __shfl_xor_sync(0xffffffffu, xhi, off));
          ylo = fmin(ylo, __shfl_xor_sync(0xffffffffu, ylo, off)); yhi = fmax(yhi, __shfl_xor_sync(0xffffffffu, yhi, off));
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
          const float nanf_ = __int_as_float(0x7fc00000);
          sbox[j] = bad ? make_float4(nanf_, nanf_, nanf_, nanf_)
                        : make_float4(__double2float_rd(xlo), __double2float_ru(xhi), __double2float_rd(ylo), __double2float_ru(yhi));
        }
      }
    }
  }
  __syncthreads();                                       // tables, cleared buffers (and fused boxes) visible

  double my_cost = INFINITY;                             // running arg-min over every block this CTA sweeps
  int my_idx = 0x7fffffff;
  int par = 0;                                           // state buffer of the current block
#ifdef FOT_PHASE_CLOCKS
  long long t_phase = clock64();
#endif

  // block-independent constants of the validity chain and the collision tests
  const int n_circ = P.cfg.n_circles;
  double max_off = 0.0;                                  // footprint circles sit within max|offset| of the path point
  for (int i = 0; i < n_circ; ++i) max_off = fmax(max_off, fabs(P.cfg.circle_offsets[i]));
  const bool dist_mode = (B.dyn_mode == FOT_DYN_DISTRIBUTION);
  const double r2_dyn = dist_mode ? P.cfg.collide_r2 : P.cfg.collide_r2_single;   // fp.py:1099-1104, :1173
  const double rc_s = sqrt(P.cfg.collide_r2) * (1.0 + 1e-9) + 1e-9 + max_off;
  const double rc_d = sqrt(r2_dyn) * (1.0 + 1e-9) + 1e-9 + max_off;
  const double wroad = fmax(P.cfg.max_road_width + 1e-9, fabs(fs[3]));
  const int max_viol = dist_mode ? (int)floor(P.cfg.chance_epsilon * (double)B.S) : 0;   // fp.py:1114
  const bool budget = dist_mode && max_viol > 0;
  const double* lim = B.limits + 4 * (size_t)q;
  const double inf = INFINITY;
  auto uni = [&](double x) {       // warp-uniform value through redux: lives in uniform registers
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)__double2loint(x));
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)__double2hiint(x));
    return __hiloint2double((int)hi, (int)lo);
  };
  const double vmax2 = uni(lim[0] < 0.0 ? -inf : lim[0] * lim[0]);
  const double amax2 = uni(lim[1] < 0.0 ? -inf : lim[1] * lim[1]);
  const double kmax2 = uni(lim[2] < 0.0 ? -inf : lim[2] * lim[2]);
  const double latmax2 = uni(lim[3] < 0.0 ? -inf : lim[3] * lim[3]);
  const double road_thr = P.cfg.max_road_width + 1e-9;                         // fp.py:982
  const double tele_thr = fmax(lim[0], P.cfg.max_speed) * dt * 3.0;            // fp.py:955
  const double tele2 = uni(tele_thr * tele_thr);
  const double fast2 = 0.25;                                                   // v > 0.5 (fp.py:1019)
  const double stop_dist = B.stop_dist[q];
  const double kTan01Sq = 0.010067046422495888;                                // tan(0.1)^2

  for (int b = b_first; b < b_last; ++b) {
  const bool brake_blk = b >= G.grid_blocks;
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
  if (n_k <= 0) continue;                                // uniform per block
  const int n_cand = n_k * n_dl;
  const int cand0 = brake_blk ? P.cfg.n_T * n_v * n_d + k_lo : (jT * n_v + k_lo) * n_d;   // generation order (fp.py:398-449)
  if (!state_ok) {
    for (int c = tid; c < n_cand; c += bd) {
      if (O.cand_cat) O.cand_cat[(size_t)q * O.cand_stride + cand0 + c] = (uint8_t)FOT_CAT_DROP;
      if (O.cand_cost) O.cand_cost[(size_t)q * O.cand_stride + cand0 + c] = qnan();
    }
    continue;
  }
  // this block's state buffer
  unsigned char* buf = smb + G.o_buf + par * G.buf_bytes;
  unsigned char* buf_next = smb + G.o_buf + (par ^ 1) * G.buf_bytes;
  int* pi_fnr = reinterpret_cast<int*>(buf + G.b_fnr);              // [pcap] 0x7fffffff - first NaN sample (0: none)
  unsigned* flags = reinterpret_cast<unsigned*>(buf + G.b_flags);   // [pcap][nw4]
  unsigned* hitw = reinterpret_cast<unsigned*>(buf + G.b_hit);      // [pcap][nwc] decisive collision
  unsigned* viol = reinterpret_cast<unsigned*>(buf + G.b_viol);     // [pcap][n_d][vwords]
  unsigned* dirty = reinterpret_cast<unsigned*>(buf + G.b_dirty);   // [pcap][nwc] candidates with a validity flag
  double* sdl = reinterpret_cast<double*>(buf + G.b_sdl);           // [pcap] s_dot at the last sample
  double* ctb = reinterpret_cast<double*>(buf + G.b_ct);            // [pcap + 2 max(n_d, ppb)] cost-table entries: Js | Jp | d_end
  double* vlast = reinterpret_cast<double*>(buf + G.b_vlast);       // [pcap][n_d] v^2 at the last kept sample
  double* sspan = reinterpret_cast<double*>(buf + G.b_span);        // [pcap] s[keep - 1] - s[0]
  unsigned* qctl = reinterpret_cast<unsigned*>(buf + G.b_qctl);     // block queue: [0] slots reserved, [1] tickets taken, [2] warps done producing
  const unsigned magicN = ((1u << 20) + (unsigned)N - 1u) / (unsigned)N;   // item / N = (item * magicN) >> 20, exact for item < 512, N <= 128

  // ---- phase B: one item per thread ---------------------------------------------------------------
  {
    // this block's jerk sums / terminal offsets (fot_prepass tables) -> shared memory, asynchronously; phase E
    // reads them two barriers from now
    const int nTv = P.cfg.n_T * B.n_v_max, nTd = P.cfg.n_T * n_d, nB = P.cfg.n_B;
    const double* ct = B.cost_tab + (size_t)q * (nTv + 2 * nTd + 3 * nB);
    const int n_lat = brake_blk ? n_k : n_d, lcap = G.ct_lcap;
    if (tid < n_k + 2 * n_lat) {
      const double* src;
      double* dst;
      if (tid < n_k) {
        src = brake_blk ? ct + nTv + 2 * nTd + k_lo + tid : ct + jT * B.n_v_max + k_lo + tid;
        dst = ctb + tid;
      } else if (tid < n_k + n_lat) {
        const int li = tid - n_k;
        src = brake_blk ? ct + nTv + 2 * nTd + nB + k_lo + li : ct + nTv + jT * n_d + li;
        dst = ctb + G.pcap + li;
      } else {
        const int li = tid - n_k - n_lat;
        src = brake_blk ? ct + nTv + 2 * nTd + 2 * nB + k_lo + li : ct + nTv + nTd + jT * n_d + li;
        dst = ctb + G.pcap + lcap + li;
      }
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int n_items = n_k * N;
  const bool active = tid < n_items;
  const int p = active ? tid / N : -1;                                         // pair slot of this item
  const int n = active ? tid - p * N : 0;                                      // its sample
  double i_rx = 0, i_ry = 0, i_cth = 0, i_sth = 0, i_rk = 0, i_rdk = 0, i_sd = 0, i_sdd = 0, i_isd = 0;
  double A0 = 0, B0 = 0, A1 = 0, B1 = 0, A2 = 0, B2 = 0;
  if (active) {
    Lon L;
    if (!brake_blk)
      L = lon_solve(fs, B.v_grid[(size_t)q * B.n_v_max + k_lo + p], P.T[jT], P.inv4 + 4 * jT, n_v == 1, N - 1);
    else
      L = lon_solve(fs, 0.0, P.Tb[k_lo + p], P.inv4b + 4 * (k_lo + p), true, P.n_steps_b[k_lo + p]);
    const bool held = n > L.hold;                                              // fp.py:487-499 brake padding
    const TPow tp = tpow(held ? L.hold : n, dt);
    const double s = L.a0 + L.a1 * tp.t + L.a2 * tp.t2 + L.a3 * tp.t3 + L.a4 * tp.t4;             // fp.py:644
    i_sd = held ? 0.0 : L.a1 + 2.0 * L.a2 * tp.t + 3.0 * L.a3 * tp.t2 + 4.0 * L.a4 * tp.t3;      // fp.py:645
    i_sdd = held ? 0.0 : 2.0 * L.a2 + 6.0 * L.a3 * tp.t + 12.0 * L.a4 * tp.t2;                    // fp.py:646
    SplineView V;
    V.nx = P.cfg.nx;
    if (G.spline_smem) {
      const int nx = V.nx;
      V.knots = spl; V.xa = spl + nx; V.xb = spl + 2 * nx; V.xc = spl + 3 * nx; V.xd = spl + 4 * nx;
      V.ya = spl + 5 * nx; V.yb = spl + 6 * nx; V.yc = spl + 7 * nx; V.yd = spl + 8 * nx;
    } else {
      V.knots = P.knots; V.xa = P.xa; V.xb = P.xb; V.xc = P.xc; V.xd = P.xd;
      V.ya = P.ya; V.yb = P.yb; V.yc = P.yc; V.yd = P.yd;
    }
    const RefFast rp = spline_ref_fast(V, s);
    i_rx = rp.rx; i_ry = rp.ry; i_cth = rp.cth; i_sth = rp.sth; i_rk = rp.rk; i_rdk = rp.rdk;
    i_isd = fabs(i_sd) > 1e-3 ? rcp_nr(i_sd) : 0.0;                            // fp.py:792 EPS_S_DOT
    // lateral basis at this sample: d_i(t) = A(t) + d_i * B(t) (fp.py:676-683); Horner with running derivatives
    double c0, c1, c2, c3, c4, c5, b3, b4, b5;
    if (!brake_blk) {
      const double T = P.T[jT];
      const double* Ai = P.inv5 + 9 * jT;
      c0 = fs[3]; c1 = fs[4]; c2 = fs[5] / 2.0;
      const double r0 = -c0 - c1 * T - c2 * T * T, r1 = -c1 - 2.0 * c2 * T, r2 = -2.0 * c2;
      c3 = fma(r2, Ai[2], fma(r1, Ai[1], r0 * Ai[0]));
      c4 = fma(r2, Ai[5], fma(r1, Ai[4], r0 * Ai[3]));
      c5 = fma(r2, Ai[8], fma(r1, Ai[7], r0 * Ai[6]));
      b3 = Ai[0]; b4 = Ai[3]; b5 = Ai[6];
    } else {                                                                   // one lateral profile per brake horizon (fp.py:480-482)
      const Lat Lb = lat_solve(fs, fs[3], P.Tb[k_lo + p], P.inv5b + 9 * (k_lo + p), true, P.n_steps_b[k_lo + p]);
      c0 = Lb.a0; c1 = Lb.a1; c2 = Lb.a2; c3 = Lb.a3; c4 = Lb.a4; c5 = Lb.a5;
      b3 = b4 = b5 = 0.0;
    }
    {
      const double t = tp.t;
      double pA = fma(c5, t, c4), dA = c5, ddA;
      ddA = dA;               dA = fma(dA, t, pA);  pA = fma(pA, t, c3);
      ddA = fma(ddA, t, dA);  dA = fma(dA, t, pA);  pA = fma(pA, t, c2);
      ddA = fma(ddA, t, dA);  dA = fma(dA, t, pA);  pA = fma(pA, t, c1);
      ddA = fma(ddA, t, dA);  dA = fma(dA, t, pA);  pA = fma(pA, t, c0);
      double pB = fma(b5, t, b4), dB = b5, ddB;
      ddB = dB;               dB = fma(dB, t, pB);  pB = fma(pB, t, b3);
      ddB = fma(ddB, t, dB);  dB = fma(dB, t, pB);  pB = pB * t;
      ddB = fma(ddB, t, dB);  dB = fma(dB, t, pB);  pB = pB * t;
      ddB = fma(ddB, t, dB);  dB = fma(dB, t, pB);  pB = pB * t;
      A0 = pA; B0 = pB;
      A1 = held ? 0.0 : dA;        B1 = held ? 0.0 : dB;
      A2 = held ? 0.0 : 2.0 * ddA; B2 = held ? 0.0 : 2.0 * ddB;
    }
    double* r = row + (p * NT + n) * kRowW;
    r[0] = i_rx; r[1] = i_ry; r[2] = i_cth; r[3] = i_sth; r[4] = i_rk; r[5] = s; r[6] = i_isd; r[7] = i_sd;
    r[8] = A0; r[9] = B0; r[10] = A1; r[11] = B1;
    if (i_rx != i_rx || i_ry != i_ry) atomicMax(&pi_fnr[p], 0x7fffffff - n);   // fp.py:851-866 (first NaN sample)
    if (n == N - 1) sdl[p] = i_sd;                                             // terminal speed of the cost (fp.py:724)
  }
  FOT_PHASE_MARK(0);
  __syncthreads();                                       // ---- barrier (i)
  FOT_PHASE_MARK(1);
  // every thread is past phase E of the previous block: clear the state the NEXT block accumulates into
  for (int i = tid; i < G.n_zero; i += bd) reinterpret_cast<unsigned*>(buf_next + G.b_fnr)[i] = 0u;

  // ---- phase CD, per warp ---------------------------------------------------------------------------
  const int fnr = active ? pi_fnr[p] : 0;
  const int fn = 0x7fffffff - fnr;                                             // 0x7fffffff: no NaN sample
  const int keep = fnr == 0 ? N : (fn >= 2 ? fn : 0);                          // fp.py:866
  const bool valid = active && n < keep;
  const bool chk = valid && n >= 1;                                            // limits skip index 0 (fp.py:964-983)
  if (tid < n_k) {
    // stop-distance span of pair `tid` (fp.py:319): s at its last kept sample minus s at its first
    const int f2 = 0x7fffffff - pi_fnr[tid];
    const int k2 = pi_fnr[tid] == 0 ? N : (f2 >= 2 ? f2 : 0);
    sspan[tid] = k2 > 0 ? row[(tid * NT + k2 - 1) * kRowW + 5] - row[tid * NT * kRowW + 5] : 0.0;
  }
  const unsigned keep4 = chk ? 0xffffffffu : F_DROP * 0x01010101u;             // n = 0: only the drop guards apply
  const double* rown = row + ((active ? p : 0) * NT + n) * kRowW;
  const double* rowp = chk ? rown - kRowW : rown;
  const double sd2 = i_sd * i_sd, isd2 = i_isd * i_isd;
  const double Q0 = fma(-i_rk, A0, 1.0), Q1 = -(i_rk * B0);                    // q = 1 - kappa_r d
  const double P0 = A1 * i_isd, P1 = B1 * i_isd;                               // d' (fp.py:792-799)
  const double R0 = (A2 - P0 * i_sdd) * isd2, R1 = (B2 - P1 * i_sdd) * isd2;   // d''
  const double M0 = fma(i_rdk, A0, i_rk * P0), M1 = fma(i_rdk, B0, i_rk * P1); // kappa_r' d + kappa_r d'
  const double S0 = i_sdd * Q0, S1 = i_sdd * Q1;                               // s_ddot q
  const double E0x = (i_rx - rowp[0]) - (i_sth * A0 - rowp[3] * rowp[8]), E1x = -(i_sth * B0 - rowp[3] * rowp[9]);
  const double E0y = (i_ry - rowp[1]) + (i_cth * A0 - rowp[2] * rowp[8]), E1y = i_cth * B0 - rowp[2] * rowp[9];
  const unsigned segmask = __match_any_sync(0xffffffffu, p);
  const bool seg_leader = (__ffs(segmask) - 1) == lane;
  unsigned* flags_p = flags + (active ? p : 0) * G.nw4;
  unsigned* dirty_p = dirty + (active ? p : 0) * G.nwc;
  const double sd4 = sd2 * sd2;
  unsigned anyslow = 0u;
  // per-item settlement of the tests that are affine / convex in d_i, interval screen of the rest: see
  // fot_sweep_items.cuh (identical arithmetic)
  bool lite, skip;
  {
    const double ga = brake_blk ? 0.0 : P.d_min, gb = brake_blk ? 0.0 : P.d_max, gabs = fmax(fabs(ga), fabs(gb));
    const double bx = fabs(E0x) + gabs * fabs(E1x), by = fabs(E0y) + gabs * fabs(E1y);
    const bool ok_tele = fma(bx, bx, by * by) <= 0.99 * tele2;
    const bool ok_road = fabs(fma(ga, B0, A0)) <= road_thr && fabs(fma(gb, B0, A0)) <= road_thr;
    const double qa = fma(ga, Q1, Q0), pa = fma(ga, P1, P0), qb = fma(gb, Q1, Q0), pb = fma(gb, P1, P0);
    const double vcap = vmax2 * (1.0 - 1e-12);
    const bool ok_speed = sd2 * fma(qa, qa, pa * pa) <= vcap && sd2 * fma(qb, qb, pb * pb) <= vcap;
    const bool ok_sing = fmin(qa, qb) > 0.05;
    const double mag = fabs(Q0) + fabs(P0) + fabs(R0) + fabs(M0) + fabs(S0) + sd2 + fabs(i_rk) +
                       gabs * (fabs(Q1) + fabs(P1) + fabs(R1) + fabs(M1) + fabs(S1));
    const bool ok_fin = mag <= 1e40;
    const bool lite_ok = ok_tele && ok_road && ok_speed && ok_sing && ok_fin;
    lite = __all_sync(0xffffffffu, !valid || lite_ok);                         // NaN anywhere: full chain
    const double qmin = fmin(qa, qb), qmax = fmax(qa, qb), Pm = fmax(fabs(pa), fabs(pb));
    const double Rm = fmax(fabs(fma(ga, R1, R0)), fabs(fma(gb, R1, R0)));
    const double Mm = fmax(fabs(fma(ga, M1, M0)), fabs(fma(gb, M1, M0)));
    const double Sm = fmax(fabs(fma(ga, S1, S0)), fabs(fma(gb, S1, S0)));
    const double h2lo = qmin * qmin, h2hi = fma(qmax, qmax, Pm * Pm), ark = fabs(i_rk);
    const double Wm = fma(ark, h2hi, fma(Rm, qmax, Mm * Pm));                  // |kappa h^3|
    const double Tm = fma(Pm, fma(ark, h2hi, Wm), Mm * h2hi);
    const double Zm = fma(sd2, Tm, Sm * h2hi);                                 // |a h q|
    const double slack = 1.0 + 1e-9, Wm2 = Wm * Wm * slack;
    const bool ok_rest = Wm2 <= kmax2 * (h2lo * h2lo * h2lo) && sd4 * Wm2 <= latmax2 * h2lo &&
                         Zm * Zm * slack <= amax2 * (h2lo * h2lo) && sd2 * h2lo > 0.25 * slack;
    skip = __all_sync(0xffffffffu, !valid || (lite_ok && (!chk || ok_rest)));
  }
  auto sample = [&](auto lite_tag, double di, unsigned& acc, unsigned sh) {
    constexpr bool kLite = decltype(lite_tag)::value;
    const double qq = fma(di, Q1, Q0), dpr = fma(di, P1, P0), dpp = fma(di, R1, R0);
    const double m = fma(di, M1, M0), sq = fma(di, S1, S0);
    const double h2 = fma(qq, qq, dpr * dpr);                                  // hypot(q, d')^2 = (q / cos delta)^2
    const double w = fma(i_rk, h2, fma(dpp, qq, m * dpr));                     // kappa h^3   (cc.py:144-147)
    const double h6 = h2 * h2 * h2;
    const double w2 = w * w;
    const double v2 = sd2 * h2;                                                // v^2         (cc.py:150-152)
    const double T = fma(dpr, fma(-i_rk, h2, w), -(m * h2));
    const double Z = fma(sd2, T, sq * h2);                                     // a h q       (cc.py:155-157)
    const double acc_rhs = amax2 * (qq * qq * h2), curv_rhs = kmax2 * h6, lat_lhs = sd4 * w2, lat_rhs = latmax2 * h2;
    const double Z2 = Z * Z;
    if constexpr (kLite) {
      asm("{\n .reg .pred p, f;\n"
          " setp.gt.f64 f, %2, %3;\n"                                             // v > 0.5 (fp.py:1019)
          " @!f or.b32 %1, %1, 1;\n"
          " setp.gt.and.f64 p, %4, %5, f;\n"                                      // |kappa| > k_max when fast (fp.py:1020)
          " @p or.b32 %0, %0, %6;\n"
          "}"
          : "+r"(acc), "+r"(anyslow)
          : "d"(v2), "d"(fast2), "d"(w2), "d"(curv_rhs), "r"(F_CURV << sh));
    } else {
      const double ex = fma(di, E1x, E0x), ey = fma(di, E1y, E0y);
      const double step2 = fma(ex, ex, ey * ey);                               // fp.py:954 (squared)
      const double fin = fabs(Z) + fabs(w) + h6;
      asm("{\n .reg .pred p, f;\n .reg .f64 t;\n"
          " abs.f64 t, %2;\n setp.lt.f64 p, t, 0d7FF0000000000000;\n setp.le.and.f64 p, %2, 0d3FA999999999999A, p;\n"   // q <= 0.05 and finite (fp.py:826-833)
          " setp.geu.or.f64 p, %3, 0d7FF0000000000000, p;\n"                      // non-finite v / a / kappa (fp.py:944-946)
          " setp.gt.or.f64 p, %4, %5, p;\n"                                       // teleport (fp.py:953-956)
          " @p or.b32 %0, %0, %6;\n"
          " setp.gt.f64 f, %7, %8;\n"                                             // v > 0.5 (fp.py:1019)
          " @!f or.b32 %1, %1, 1;\n"
          " setp.gt.and.f64 p, %9, %10, f;\n"                                     // |kappa| > k_max when fast (fp.py:1020)
          " @p or.b32 %0, %0, %11;\n"
          "}"
          : "+r"(acc), "+r"(anyslow)
          : "d"(qq), "d"(fin), "d"(step2), "d"(tele2), "r"(F_DROP << sh), "d"(v2), "d"(fast2), "d"(w2), "d"(curv_rhs), "r"(F_CURV << sh));
      flag_gt(acc, v2, vmax2, F_SPEED << sh);                                  // fp.py:964
      flag_abs_gt(acc, fma(di, B0, A0), road_thr, F_ROAD << sh);               // fp.py:982
    }
    flag_gt(acc, Z2, acc_rhs, F_ACCEL << sh);                                  // fp.py:966
    flag_gt(acc, lat_lhs, lat_rhs, F_LAT << sh);                               // fp.py:975  v^2 |kappa| > a_lat
  };
  auto sweep_targets = [&](auto lite_tag) {
    for (int i0 = 0; i0 < n_dl; i0 += 4) {
      unsigned acc = 0u;
      if (valid) {
        if (brake_blk) {
          sample(lite_tag, 0.0, acc, 0u);
        } else if (i0 + 4 <= n_dl) {
          const double g0 = dgrid[i0], g1 = dgrid[i0 + 1], g2 = dgrid[i0 + 2], g3 = dgrid[i0 + 3];
          sample(lite_tag, g0, acc, 0u); sample(lite_tag, g1, acc, 8u); sample(lite_tag, g2, acc, 16u); sample(lite_tag, g3, acc, 24u);
        } else {
          for (int u = 0; i0 + u < n_dl; ++u) sample(lite_tag, dgrid[i0 + u], acc, 8u * u);
        }
      }
      const unsigned red = __reduce_or_sync(segmask, acc & keep4);
      if (seg_leader && active && red) {
        atomicOr(&flags_p[i0 >> 2], red);
        atomicOr(&dirty_p[i0 >> 5], nonzero_bytes(red) << (i0 & 31));
      }
    }
  };
  FOT_PHASE_MARK(2);
  if (skip) { }
  else if (lite) sweep_targets(std::true_type{});
  else sweep_targets(std::false_type{});
  FOT_PHASE_MARK(3);
  // Samples beyond the NaN prefix that are inside the spline domain again still count for the candidate-wide
  // singularity guard (fp.py:826-833 runs before the truncation).  Essentially never.
  if (active && !valid && i_rx == i_rx && keep > 0) {
    for (int i = 0; i < n_dl; ++i) {
      const double qq = fma(brake_blk ? 0.0 : dgrid[i], Q1, Q0);
      if ((qq <= 0.05) & (fabs(qq) < inf)) {
        atomicOr(&flags_p[i >> 2], F_DROP << (8 * (i & 3)));
        atomicOr(&dirty_p[i >> 5], 1u << (i & 31));
      }
    }
  }
  // stop-distance directive (fp.py:307-324) needs v at the last kept sample
  if (stop_dist == stop_dist && valid && n == keep - 1) {
    for (int i = 0; i < n_dl; ++i) {
      const double di = brake_blk ? 0.0 : dgrid[i];
      const double qq = fma(di, Q1, Q0), dpr = fma(di, P1, P0);
      vlast[p * n_d + i] = sd2 * fma(qq, qq, dpr * dpr);
    }
  }
  __syncwarp();

  // Low-speed regime (fp.py:1022-1032): the warp redoes the two low-speed tests for its own items that saw a
  // candidate with v <= 0.5, one (item, candidate) unit per lane, candidate-major over the compacted slow items.
  {
    const unsigned slow_mask = __ballot_sync(0xffffffffu, anyslow && chk);
    if (slow_mask) {
      const int n_slow = __popc(slow_mask);
      if (anyslow && chk) wq[__popc(slow_mask & lt_mask)] = (unsigned)tid;
      __syncwarp();
      int k = lane % n_slow, i = lane / n_slow;            // unit u = i * n_slow + k for u = lane, lane + 32, ...
      const int dk = 32 % n_slow, di_step = 32 / n_slow;
      while (i < n_dl) {
        const int it = (int)wq[k];
        const int sp = (int)(((unsigned)it * magicN) >> 20), sn = it - sp * N;
        // a candidate that already carries a flag of curvature priority or higher cannot change category
        if (!((flags[sp * G.nw4 + (i >> 2)] >> (8 * (i & 3))) & (F_DROP | F_SPEED | F_ACCEL | F_CURV))) {
          const double* r1 = row + (sp * NT + sn) * kRowW;                     // sample n
          const double* r0 = r1 - kRowW;                                       // sample n - 1 (only checked samples queue)
          const double di = brake_blk ? 0.0 : dgrid[i];
          const double d = fma(di, r1[9], r1[8]), dprev = fma(di, r0[9], r0[8]);
          const double qq = fma(-r1[4], d, 1.0), dpr = fma(di, r1[11], r1[10]) * r1[6];
          const double ssd = r1[7];
          if (!(ssd * ssd * fma(qq, qq, dpr * dpr) > 0.25)) {                  // else: this candidate is in the fast regime here
            bool badc;
            if (fabs(d - dprev) > fmax(1.5 * fabs(r1[5] - r0[5]), 0.02)) {
              badc = true;
            } else {
              // |wrap(yaw_n - yaw_{n-1})| is the angle between the heading vectors u = R(theta_r)(q, d')
              const double kmax = lim[2];
              const double q_prev = fma(-r0[4], dprev, 1.0);
              const double dp_prev = fma(di, r0[11], r0[10]) * r0[6];
              const double ux = r1[2] * qq - r1[3] * dpr, uy = r1[3] * qq + r1[2] * dpr;
              const double uxp = r0[2] * q_prev - r0[3] * dp_prev, uyp = r0[3] * q_prev + r0[2] * dp_prev;
              const double cr = uxp * uy - uyp * ux, dt_ = uxp * ux + uyp * uy;
              const double ex = fma(-r1[3], d, r1[0]) - fma(-r0[3], dprev, r0[0]);
              const double ey = fma(r1[2], d, r1[1]) - fma(r0[2], dprev, r0[1]);
              const double step2 = fma(ex, ex, ey * ey);
              if (kmax * kmax * step2 <= 0.01)
                // the threshold is the 0.1 rad floor: angle > 0.1 <=> dot <= 0 or cross^2 > tan(0.1)^2 dot^2
                badc = dt_ <= 0.0 || cr * cr > kTan01Sq * dt_ * dt_;
              else
                badc = fabs(atan2(cr, dt_)) > kmax * sqrt(step2);
            }
            if (badc) {
              atomicOr(&flags[sp * G.nw4 + (i >> 2)], F_CURV << (8 * (i & 3)));
              atomicOr(&dirty[sp * G.nwc + (i >> 5)], 1u << (i & 31));
            }
          }
        }
        k += dk; i += di_step;
        if (k >= n_slow) { k -= n_slow; ++i; }
      }
      __syncwarp();
    }
  }

  FOT_PHASE_MARK(4);
  // ---- collision (fp.py:1035-1233) ------------------------------------------------------------------
  // Every warp PRODUCES (item, obstacle) entries from the cull of its own items into one block-wide queue and then
  // CONSUMES entries of the whole block, 32 tickets at a time, until every warp has finished producing and the queue
  // is drained: warps that skipped the validity loop or have no obstacle near their samples take over the exact
  // tests of the others.  No barrier separates the two halves: a slot is valid as soon as it is non-zero.
  if (has_dyn || M > 0) {
    // exact test of one entry against every live clean candidate of the item's pair
    auto process = [&](unsigned ent) {
      const int it = (int)((ent >> 21) & 0x1ffu);
      const bool is_dyn = (ent >> 20) & 1u;
      const unsigned j = ent & 0xfffffu;
      const int ep = (int)(((unsigned)it * magicN) >> 20), en = it - ep * N;
      const unsigned off = is_dyn ? j * (unsigned)B.T_obs : j;
      const unsigned ok_ = off + (unsigned)(B.T_obs > 0 ? min(en, B.T_obs - 1) : 0);
      const double2 o = is_dyn ? (G.stage_dyn ? dynst[ok_] : dyn_q[ok_]) : stat_q[off];
      const double r2 = is_dyn ? r2_dyn : P.cfg.collide_r2;
      const bool use_budget = budget && is_dyn;
      const double* r = row + (ep * NT + en) * kRowW;
      const double cth = r[2], sth = r[3];
      const double X0 = fma(-sth, r[8], r[0]) - o.x, X1 = -(sth * r[9]);       // x - ox = X0 + d_i X1
      const double Y0 = fma(cth, r[8], r[1]) - o.y, Y1 = cth * r[9];
      for (int w = 0; w < G.nwc; ++w) {
        const int rem = n_dl - 32 * w;
        unsigned mbits = ~dirty[ep * G.nwc + w] & (rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u));
        if (!use_budget) mbits &= ~hitw[ep * G.nwc + w];
        while (mbits) {
          const int bit = __ffs(mbits) - 1;
          mbits &= mbits - 1u;
          const int i = w * 32 + bit;
          const double di = brake_blk ? 0.0 : dgrid[i];
          bool hit = false;
          if (n_circ == 0) {
            const double dx = fma(di, X1, X0), dy = fma(di, Y1, Y0);
            hit = dx * dx + dy * dy <= r2;                                     // fp.py:1196-1198, :1231-1233
          } else {                                                             // fp.py:1158-1167
            const double d = fma(di, r[9], r[8]);
            const double dpr = fma(di, r[11], r[10]) * r[6];
            const double qq = fma(-r[4], d, 1.0);
            const double rh = 1.0 / sqrt(fma(qq, qq, dpr * dpr));
            const double hx = (cth * qq - sth * dpr) * rh, hy = (sth * qq + cth * dpr) * rh;   // (cos yaw, sin yaw)
            for (int ci = 0; ci < n_circ && !hit; ++ci) {
              const double dx = fma(di, X1, X0) + P.cfg.circle_offsets[ci] * hx, dy = fma(di, Y1, Y0) + P.cfg.circle_offsets[ci] * hy;
              hit = dx * dx + dy * dy <= r2;
            }
          }
          if (hit) {
            if (!use_budget) atomicOr(&hitw[ep * G.nwc + w], 1u << bit);
            else { const int sidx = (int)j / B.P; atomicOr(&viol[(ep * n_d + i) * G.vwords + (sidx >> 5)], 1u << (sidx & 31)); }
          }
        }
      }
    };

    // -- produce: box of this warp's reference points -> obstacle list -> tangent-frame cull of its own items
    const bool okb = valid && i_rx == i_rx && i_ry == i_ry;
    const unsigned uxlo = __reduce_min_sync(0xffffffffu, okb ? f2ord(__double2float_rd(i_rx)) : 0xffffffffu);
    if (uxlo != 0xffffffffu) {                           // warp-uniform
      const unsigned uxhi = __reduce_max_sync(0xffffffffu, okb ? f2ord(__double2float_ru(i_rx)) : 0u);
      const unsigned uylo = __reduce_min_sync(0xffffffffu, okb ? f2ord(__double2float_rd(i_ry)) : 0xffffffffu);
      const unsigned uyhi = __reduce_max_sync(0xffffffffu, okb ? f2ord(__double2float_ru(i_ry)) : 0u);
      // the obstacles whose (trajectory) box meets that box padded by the widest reach of a clean candidate; a NaN
      // box (fp.py:1211-1222) fails every comparison
      const float pad = __double2float_ru(wroad + fmax(rc_s, rc_d));
      const float bx0 = __fsub_rd(ord2f(uxlo), pad), bx1 = __fadd_ru(ord2f(uxhi), pad);
      const float by0 = __fsub_rd(ord2f(uylo), pad), by1 = __fadd_ru(ord2f(uyhi), pad);
      int n_ws = 0, n_wd = 0;                            // static / dynamic list lengths (warp-uniform)
      for (int j0 = 0; j0 < M; j0 += 32) {
        const int j = j0 + lane;
        bool in = false;
        if (j < M) {
          const double2 o = stat_q[j];
          in = o.x >= (double)bx0 && o.x <= (double)bx1 && o.y >= (double)by0 && o.y <= (double)by1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) wlist[n_ws + __popc(bal & lt_mask)] = (unsigned)j;
        n_ws += __popc(bal);
      }
      if (SP > 0) {
        if (G.stage_dyn) mbar_wait(&s_bar, 0u);          // the staged obstacle block has landed
        for (int j0 = 0; j0 < SP; j0 += 32) {
          const int j = j0 + lane;
          bool in = false;
          if (j < SP) {
            const float4 ob = sbox[j];                   // xmin xmax ymin ymax
            in = ob.x <= bx1 && ob.y >= bx0 && ob.z <= by1 && ob.w >= by0;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, in);
          if (in) wlist[n_ws + n_wd + __popc(bal & lt_mask)] = (unsigned)j;
          n_wd += __popc(bal);
        }
      }
      __syncwarp();
      FOT_PHASE_MARK(5);
      const int n_l = n_ws + n_wd;
      if (n_l > 0) {
        // this item's pair as it stands now: lowest / highest clean candidate -> lateral window of the cull
        int i_lo = -1, i_hi = -1;
        if (valid)
          for (int w = 0; w < G.nwc; ++w) {
            const int rem = n_dl - 32 * w;
            const unsigned cwd = ~dirty_p[w] & (rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u));
            if (cwd) { if (i_lo < 0) i_lo = 32 * w + __ffs(cwd) - 1; i_hi = 32 * w + 31 - __clz(cwd); }
          }
        const bool cull = valid && i_lo >= 0;
        const int kob = B.T_obs > 0 ? min(n, B.T_obs - 1) : 0;                 // clip(round(t/dt)) = n (fp.py:1226-1227)
        // tangent-frame window: along = (o - ref).t within the collision radius, across = (o - ref).n within the
        // radius of the lateral offsets the pair's clean candidates take at this sample
        const double ca = fma(i_rx, i_cth, i_ry * i_sth), cn = fma(i_ry, i_cth, -(i_rx * i_sth));
        double d_lo = A0, d_hi = A0;
        if (cull && !brake_blk) {
          const double ga = P.d_sorted ? dgrid[i_lo] : P.d_min, gb = P.d_sorted ? dgrid[i_hi] : P.d_max;
          d_lo = A0 + fmin(ga * B0, gb * B0) - 1e-9;
          d_hi = A0 + fmax(ga * B0, gb * B0) + 1e-9;
        }
        for (int e = 0; e < n_l; ++e) {
          bool surv = false;
          const bool is_dyn = e >= n_ws;
          const unsigned j = wlist[e];
          if (cull) {
            const double2 o = is_dyn ? (G.stage_dyn ? dynst[j * (unsigned)B.T_obs + kob] : dyn_q[j * (unsigned)B.T_obs + kob]) : stat_q[j];
            const double rc = is_dyn ? rc_d : rc_s;
            const double al = fma(o.x, i_cth, fma(o.y, i_sth, -ca));
            const double ac = fma(o.y, i_cth, fma(-o.x, i_sth, -cn));
            surv = (fabs(al) <= rc) & (ac >= d_lo - rc) & (ac <= d_hi + rc);     // NaN -> false
          }
          const unsigned bal = __ballot_sync(0xffffffffu, surv);
          if (bal) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&qctl[0], (unsigned)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (surv) {
              const unsigned slot = base + __popc(bal & lt_mask);
              const unsigned ent = 0x80000000u | ((unsigned)tid << 21) | ((unsigned)is_dyn << 20) | j;
              if (slot < (unsigned)kBlockQueue) *reinterpret_cast<volatile unsigned*>(bq + slot) = ent;
              else process(ent);                         // queue full: test right here
            }
          }
        }
      }
    }
    FOT_PHASE_MARK(6);
    // -- this warp has produced everything it will produce
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicAdd(&qctl[2], 1u); }
    // -- consume: 32 tickets at a time
    const unsigned n_warps = (unsigned)(bd >> 5);
    for (;;) {
      unsigned h = 0;
      if (lane == 0) h = atomicAdd(&qctl[1], 32u);
      h = __shfl_sync(0xffffffffu, h, 0);
      if (h >= (unsigned)kBlockQueue) break;             // beyond the queue: those entries were tested in place
      const unsigned idx = h + lane;
      unsigned ent = 0;
      if (idx < (unsigned)kBlockQueue) {
        volatile unsigned* slot = bq + idx;
        for (;;) {
          ent = *slot;
          if (ent) break;
          unsigned done;
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(done) : "r"(smem_u32(&qctl[2])) : "memory");
          if (done == n_warps) { ent = *slot; break; }   // nobody will write this slot any more
          __nanosleep(64);
        }
        if (ent) { process(ent); *slot = 0u; }           // (slots are handed back empty for the next block)
      }
      unsigned fin = 0;
      if (lane == 0) {
        unsigned done;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(done) : "r"(smem_u32(&qctl[2])) : "memory");
        fin = done == n_warps && h + 32u >= *reinterpret_cast<volatile unsigned*>(&qctl[0]);
      }
      if (__shfl_sync(0xffffffffu, fin, 0)) break;
    }
  }
  FOT_PHASE_MARK(7);
  asm volatile("cp.async.wait_all;" ::: "memory");   // this block's cost-table entries have landed (visible after the barrier)
  __syncthreads();                                       // ---- barrier (ii)
  FOT_PHASE_MARK(8);

  // ---- phase E: category, cost, block arg-min, histogram ----------------------------------------
  for (int c = tid; c < n_cand; c += bd) {
    const int cp = c / n_dl, ci = c - cp * n_dl;
    const int li = brake_blk ? cp : ci;
    const double Js = ctb[cp], Jp = ctb[G.pcap + li], d_end = ctb[G.pcap + G.ct_lcap + li];
    const double Jd = d_end * d_end;
    const double dv = B.target[q] - sdl[cp];
    const double Jv = dv * dv;
    const double Jt = (double)(N - 1) * dt;
    const double lat_cost = P.cfg.k_j * Jp + P.cfg.k_t * Jt + P.cfg.k_d * Jd;
    const double lon_cost = P.cfg.k_j * Js + P.cfg.k_t * Jt + P.cfg.k_s_dot * Jv;
    const double cost = P.cfg.k_lat * lat_cost + P.cfg.k_lon * lon_cost;
    const unsigned byte = (flags[cp * G.nw4 + (ci >> 2)] >> (8 * (ci & 3))) & 0xffu;
    const int cfnr = pi_fnr[cp];
    const int cfn = 0x7fffffff - cfnr;
    const int ckeep = cfnr == 0 ? N : (cfn >= 2 ? cfn : 0);
    int cat;
    if (ckeep == 0 || (byte & F_DROP)) cat = FOT_CAT_DROP;                     // fp.py:831-833, :933, :944, :953
    else if (byte & F_SPEED) cat = FOT_CAT_SPEED;
    else if (byte & F_ACCEL) cat = FOT_CAT_ACCEL;
    else if (byte & F_CURV) cat = FOT_CAT_CURV;
    else if (byte & F_LAT) cat = FOT_CAT_LAT;
    else if (byte & F_ROAD) cat = FOT_CAT_ROAD;
    else {
      bool hit = (hitw[cp * G.nwc + (ci >> 5)] >> (ci & 31)) & 1u;
      if (G.vwords > 0) {
        int nv = 0;
        for (int w = 0; w < G.vwords; ++w) nv += __popc(viol[(cp * n_d + ci) * G.vwords + w]);
        hit = hit || nv > max_viol;                                            // fp.py:1113-1124
      }
      if (hit) {
        cat = FOT_CAT_COLL;                                                    // fp.py:986-989
      } else {
        cat = FOT_CAT_OK;
        if (stop_dist == stop_dist) {                                          // fp.py:307-324
          const double v_last = sqrt(vlast[cp * n_d + ci]);
          if (!(v_last <= 0.15 && sspan[cp] <= stop_dist + 1e-6)) cat = FOT_CAT_STOP;
        }
      }
    }
    const int cand_idx = cand0 + c;
    if (cat < FOT_N_STATS) atomicAdd(&s_stats[cat], 1);
    if (O.cand_cat) O.cand_cat[(size_t)q * O.cand_stride + cand_idx] = (uint8_t)cat;
    if (O.cand_cost) O.cand_cost[(size_t)q * O.cand_stride + cand_idx] = cost;
    if (cat == FOT_CAT_OK && cost < INFINITY) argmin_merge(my_cost, my_idx, cost, cand_idx);
  }
  FOT_PHASE_MARK(9);
  par ^= 1;                                              // no barrier: the next block works in the other buffer
  }  // blocks of this CTA

  if (G.stage_dyn && state_ok) mbar_wait(&s_bar, 0u);    // the bulk copy must have landed before the CTA can exit
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

}  // namespace fot
