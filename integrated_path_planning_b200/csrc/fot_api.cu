// fot_api.cu -- C ABI (include/fot.h) over the sweep kernels.  Host side only: argument
// checks, device tables, scratch, launches.  There is no CPU fallback anywhere in this file:
// without a CUDA device every entry point that computes returns FOT_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <chrono>
#include <vector>
#include <mutex>
#include <set>

#include "fot_kernels.cuh"
#include "fot_sweep_items.cuh"
#include "fot_sweep_warp.cuh"
#include "fot_sweep_pairs.cuh"
#include "fot_predict.cuh"

using namespace fot;

namespace {

thread_local std::string g_err;

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess)
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  else
    snprintf(buf, sizeof buf, "%s", what);
  g_err = buf;
  return code;
}

#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return fail(FOT_ERR_CUDA, #call, e__);      \
  } while (0)

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Growable device / pinned-host buffer.
struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  bool host = false;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    release();
    size_t want = n + n / 4 + 256;
    cudaError_t e = host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want; else p = nullptr;
    return e;
  }
  void release() {
    if (p) { if (host) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
  }
};

// Tuning / test knobs (FOT_* environment variables).  They are resolved ONCE, when the handle is created -- no entry
// point reads the environment on the planning path -- and again only on an explicit fot_reload_options() (the tests
// and tuning scripts switch variants on a live handle that way).  Defaults are the product behaviour.
struct Options {
  int item_threads = 0;        // FOT_ITEM_THREADS   0: kItemThreads
  int bpc = 0;                 // FOT_BPC            blocks per CTA (0: rule of item_geometry)
  int qcap = 1024;             // FOT_QCAP           collision queue entries (tests force the queue-full path)
  int fused_box = 0;           // FOT_FUSED_BOX      1: trajectory boxes in the sweep even for a resident tensor
  int stage_dyn = 1;           // FOT_STAGE_DYN      0: never stage the obstacle block in shared memory
  int sweep = 0;               // FOT_SWEEP          0 auto (fot_sweep_pairs, else fot_sweep_items, else the candidate-major kernel), 1 "items" (fail if
                               //                    unsupported), 2 "generic" (candidate-major kernel), 3 "warp" (fot_sweep_warp: two
                               //                    barriers per block + barrier-free collision queue; measured equal to "items",
                               //                    DESIGN.md section 4c; fail if unsupported)
                               //                    4 "pairs" (fot_sweep_pairs: one longitudinal profile per warp, no block
                               //                    barriers; fail if unsupported)
  int pair_cpq = 0;            // FOT_PAIR_CPQ       CTAs per query of fot_sweep_pairs (0: rule of pair_geometry)
  int pair_simple = 1;         // FOT_PAIR_SIMPLE    0: always the instantiation of fot_sweep_pairs with every mode compiled in (tests)
  int pair_feat = 0;           // FOT_PAIR_FEAT      mode bits (PAIR_*) added to what the batch needs (tests, tuning)
  int host_chunks = 0;         // FOT_HOST_CHUNKS    equal chunks of the host-pointer call (0: rule)
  std::string chunk_waves;     // FOT_CHUNK_WAVES    chunk sizes in sweep waves, "1,2,3" (+ the rest)
  int host_streams = 2;        // FOT_HOST_STREAMS   1: chunks on one compute stream
  int gated = 1;               // FOT_GATED          0: chunked launches instead of gated ones
  int gate_uploads = 16;       // FOT_GATE_UPLOADS   upload slices of a gated call
  int gate_copy_streams = 1;   // FOT_GATE_COPY_STREAMS
  int gate_flag_stream = 0;    // FOT_GATE_FLAG_STREAM  1: slice flags written by a second stream
  int gate_tail_bpc = 0;       // FOT_GATE_TAIL_BPC  blocks per CTA of the last gated range
  int gate_memcpy = 0;         // FOT_GATE_MEMCPY    1: flags by 4-byte copies instead of stream writes
  int debug_timing = 0;        // FOT_DEBUG_TIMING   1: print the timeline of every host-pointer call
  void resolve() {
    *this = Options{};
    auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    item_threads = geti("FOT_ITEM_THREADS", 0);
    bpc = geti("FOT_BPC", 0);
    qcap = std::max(1, std::min(1 << 15, geti("FOT_QCAP", 1024)));
    fused_box = geti("FOT_FUSED_BOX", 0) != 0;
    stage_dyn = geti("FOT_STAGE_DYN", 1) != 0;
    if (const char* e = getenv("FOT_SWEEP")) sweep = !strcmp(e, "generic") ? 2 : !strcmp(e, "items") ? 1 : !strcmp(e, "warp") ? 3 : !strcmp(e, "pairs") ? 4 : 0;
    pair_cpq = geti("FOT_PAIR_CPQ", 0);
    pair_simple = geti("FOT_PAIR_SIMPLE", 1) != 0;
    pair_feat = geti("FOT_PAIR_FEAT", 0) & 31;
    host_chunks = geti("FOT_HOST_CHUNKS", 0);
    if (const char* e = getenv("FOT_CHUNK_WAVES")) chunk_waves = e;
    host_streams = geti("FOT_HOST_STREAMS", 2);
    gated = geti("FOT_GATED", 1) != 0;
    gate_uploads = geti("FOT_GATE_UPLOADS", 16);
    gate_copy_streams = geti("FOT_GATE_COPY_STREAMS", 1) >= 2 ? 2 : 1;
    gate_flag_stream = geti("FOT_GATE_FLAG_STREAM", 0) != 0;
    gate_tail_bpc = geti("FOT_GATE_TAIL_BPC", 0);
    gate_memcpy = getenv("FOT_GATE_MEMCPY") != nullptr;
    debug_timing = getenv("FOT_DEBUG_TIMING") != nullptr;
  }
};

std::mutex g_live_mu;
std::set<fot_handle*> g_live;     // handles alive in this process (fot_reload_options(NULL) walks them)

}  // namespace

// The instantiations of fot_sweep_pairs that exist (mode sets, PAIR_*), smallest first: the campaign shape, + static
// obstacles, + per-candidate outputs, + both, the unstaged obstacle block (large sample sets), the violation budget
// (chance-constrained planning), static obstacles + footprint (the corridor scenarios), everything.
#define FOT_PAIR_INSTANCES(X) X(0) X(PAIR_STATIC) X(PAIR_OUTPUTS) X(PAIR_STATIC | PAIR_OUTPUTS) X(PAIR_LOOSE) X(PAIR_BUDGET) \
  X(PAIR_STATIC | PAIR_FOOTPRINT) X(PAIR_ALL)
#define FOT_PAIR_LIST(f) (f),
static const int kPairInstances[] = {FOT_PAIR_INSTANCES(FOT_PAIR_LIST)};
static const int kPairInstanceCount = (int)(sizeof(kPairInstances) / sizeof(kPairInstances[0]));
typedef void (*PairKernel)(const Plan, const Batch, const Out, const PairGeom);
static PairKernel pair_kernel(bool fused, int feat) {
  switch (feat) {
#define FOT_PAIR_CASE(f) case (f): return fused ? fot_sweep_pairs<true, (f)> : fot_sweep_pairs<false, (f)>;
    FOT_PAIR_INSTANCES(FOT_PAIR_CASE)
    default: return fused ? fot_sweep_pairs<true, PAIR_ALL> : fot_sweep_pairs<false, PAIR_ALL>;
  }
}

struct fot_handle {
  Options opt;
  int device = 0;
  Plan plan{};
  void* tables_dev = nullptr;
  cudaStream_t stream = nullptr, stream2 = nullptr, copy_stream = nullptr, copy_stream2 = nullptr, d2h_stream = nullptr;
  static constexpr int kMaxChunks = 16;
  cudaEvent_t ev_copy[kMaxChunks] = {}, ev_done[kMaxChunks] = {};
  cudaEvent_t ev_join = nullptr, ev_blob = nullptr;
  cudaEvent_t ev_slice[64] = {};     // gated uploads: one event per slice (kGateSlices)
  cudaStream_t pstream[4] = {};      // gated launches: one stream per range of queries, earlier ranges at higher priority
  Buf gate_d, gate_h;                // gated launches: device progress / error words, pinned source words
  uint32_t gate_epoch = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  static constexpr int kRing = 256;
  std::vector<cudaEvent_t> ring;     // kRing x 4 events: start, after prepass, after sweep, after winner
  long long n_launch = 0;
  int smem_optin = 0;
  int sms = 148;                     // SM count of the device (block-per-CTA grouping, host chunk sizes)
  Buf obs_tm, obs_max2, stat_tm, stat_max2, part_cost, part_idx, dyn_box, cost_tab;   // device scratch
  int last_pair_feat = -1;           // mode set (PAIR_*) of the fot_sweep_pairs instantiation of the last launch
  int last_sweep_kind = 0;           // 4: fot_sweep_pairs, 1: fot_sweep_items, 3: fot_sweep_warp, 2: fot_sweep (generic)
  fot_result_t mirror{};             // fot_set_result_mirror: second destination of the winner block (all null: none)
  unsigned* mirror_flag = nullptr;   // word published behind each mirrored launch (fot_set_result_mirror), or null
  unsigned mirror_seq = 0;
  const double* last_winner_d = nullptr;   // full winner series of the last host-result call (device, in out_d)
  int last_winner_nq = 0;
  Buf stage_h, stage_d, out_d, dyn_d, stat_d;   // host-API staging
  fot_handle() { stage_h.host = true; gate_h.host = true; }
};

extern "C" int fot_abi_version(void) { return FOT_ABI_VERSION; }
extern "C" const char* fot_last_error(void) { return g_err.c_str(); }

extern "C" int fot_create(const fot_config_t* cfg, const fot_tables_t* tb, int device, fot_handle_t** out) {
  if (!cfg || !tb || !out) return fail(FOT_ERR_ARG, "fot_create: null argument");
  *out = nullptr;
  if (cfg->n_T < 1 || cfg->n_d < 1 || cfg->n_B < 0 || cfg->nx < 2 || cfg->n_total < 1 ||
      cfg->n_circles < 0 || cfg->n_circles > FOT_MAX_CIRCLES || !(cfg->dt > 0.0))
    return fail(FOT_ERR_ARG, "fot_create: bad grid sizes");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(FOT_ERR_NO_DEVICE, "fot_create: no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= n_dev) return fail(FOT_ERR_ARG, "fot_create: bad device ordinal");
  CK(cudaSetDevice(device));

  fot_handle* h = new fot_handle();
  // any failure below releases everything created so far (streams, events, tables) through fot_destroy
  struct Guard { fot_handle* h; ~Guard() { if (h) fot_destroy(h); } } guard{h};
  h->opt.resolve();
  h->device = device;
  h->plan.cfg = *cfg;
  const int nT = cfg->n_T, nB = cfg->n_B, nd = cfg->n_d, nx = cfg->nx;
  int n_t_max = nB > 0 ? cfg->n_total : 0;
  for (int j = 0; j < nT; ++j) {
    if (tb->n_steps[j] < 1) return fail(FOT_ERR_ARG, "fot_create: horizon shorter than 2 samples");
    n_t_max = std::max(n_t_max, tb->n_steps[j] + 1);
  }
  for (int j = 0; j < nB; ++j)
    if (tb->n_steps_b[j] + 1 > cfg->n_total) return fail(FOT_ERR_ARG, "fot_create: brake horizon longer than max_t");
  h->plan.n_t_max = n_t_max;
  h->plan.d_sorted = 1;
  h->plan.d_min = h->plan.d_max = tb->d_grid[0];
  for (int i = 0; i < nd; ++i) {
    if (i > 0 && !(tb->d_grid[i] >= tb->d_grid[i - 1])) h->plan.d_sorted = 0;
    h->plan.d_min = std::min(h->plan.d_min, tb->d_grid[i]);
    h->plan.d_max = std::max(h->plan.d_max, tb->d_grid[i]);
  }

  // one device blob: doubles first, then the int tables
  std::vector<double> dbl;
  auto push = [&](const double* p, size_t n) { size_t o = dbl.size(); dbl.insert(dbl.end(), p, p + n); return o; };
  const size_t oT = push(tb->T, nT), oI4 = push(tb->inv4, 4 * (size_t)nT), oI5 = push(tb->inv5, 9 * (size_t)nT);
  const size_t oTb = nB ? push(tb->Tb, nB) : 0, oI4b = nB ? push(tb->inv4b, 4 * (size_t)nB) : 0,
               oI5b = nB ? push(tb->inv5b, 9 * (size_t)nB) : 0;
  const size_t oD = push(tb->d_grid, nd), oK = push(tb->knots, nx);
  const size_t oxa = push(tb->xa, nx), oxb = push(tb->xb, nx - 1), oxc = push(tb->xc, nx), oxd = push(tb->xd, nx - 1);
  const size_t oya = push(tb->ya, nx), oyb = push(tb->yb, nx - 1), oyc = push(tb->yc, nx), oyd = push(tb->yd, nx - 1);
  std::vector<int32_t> ints(tb->n_steps, tb->n_steps + nT);
  if (nB) ints.insert(ints.end(), tb->n_steps_b, tb->n_steps_b + nB);
  const size_t dbytes = dbl.size() * sizeof(double), ibytes = ints.size() * sizeof(int32_t);
  cudaError_t e = cudaMalloc(&h->tables_dev, dbytes + ibytes);
  if (e != cudaSuccess) { h->tables_dev = nullptr; return fail(FOT_ERR_CUDA, "cudaMalloc(tables)", e); }
  e = cudaMemcpy(h->tables_dev, dbl.data(), dbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy((char*)h->tables_dev + dbytes, ints.data(), ibytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(FOT_ERR_CUDA, "cudaMemcpy(tables)", e);
  const double* D = (const double*)h->tables_dev;
  const int32_t* I = (const int32_t*)((char*)h->tables_dev + dbytes);
  Plan& P = h->plan;
  P.T = D + oT; P.inv4 = D + oI4; P.inv5 = D + oI5;
  P.Tb = D + oTb; P.inv4b = D + oI4b; P.inv5b = D + oI5b;
  P.d_grid = D + oD; P.knots = D + oK;
  P.xa = D + oxa; P.xb = D + oxb; P.xc = D + oxc; P.xd = D + oxd;
  P.ya = D + oya; P.yb = D + oyb; P.yc = D + oyc; P.yd = D + oyd;
  P.n_steps = I; P.n_steps_b = I + nT;

  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_blob, cudaEventDisableTiming));
  for (auto& ev : h->ev_slice) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  {
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));     // numerically lower = higher priority
    for (int i = 0; i < 4; ++i)
      CK(cudaStreamCreateWithPriority(&h->pstream[i], cudaStreamNonBlocking, std::min(least, greatest + i)));
  }
  CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->copy_stream2, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
  for (auto& ev : h->ev_copy) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto& ev : h->ev_done) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  h->ring.resize((size_t)fot_handle::kRing * 4);
  for (auto& ev : h->ring) CK(cudaEventCreate(&ev));
  CK(cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  CK(cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device));
  h->smem_optin -= 2048;   // leave room for the kernels' static shared memory
  CK(cudaFuncSetAttribute(fot_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
  CK(cudaFuncSetAttribute(fot_sweep_items<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
  CK(cudaFuncSetAttribute(fot_sweep_items<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
  CK(cudaFuncSetAttribute(fot_sweep_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
  CK(cudaFuncSetAttribute(fot_sweep_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
  for (int i = 0; i < kPairInstanceCount; ++i)
    for (int f = 0; f < 2; ++f)
      CK(cudaFuncSetAttribute(pair_kernel(f != 0, kPairInstances[i]), cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
  { std::lock_guard<std::mutex> lk(g_live_mu); g_live.insert(h); }
  guard.h = nullptr;
  *out = h;
  return FOT_OK;
}

extern "C" int fot_reload_options(fot_handle_t* h) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  if (h) { if (!g_live.count(h)) return fail(FOT_ERR_ARG, "fot_reload_options: unknown handle"); h->opt.resolve(); return FOT_OK; }
  for (fot_handle* x : g_live) x->opt.resolve();
  return FOT_OK;
}

extern "C" int fot_destroy(fot_handle_t* h) {
  if (!h) return FOT_OK;
  { std::lock_guard<std::mutex> lk(g_live_mu); g_live.erase(h); }
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();          // every stream of the handle (compute, priority, copy) is idle before its buffers go
  for (Buf* b : {&h->gate_d, &h->gate_h, &h->obs_tm, &h->obs_max2, &h->stat_tm, &h->stat_max2, &h->part_cost, &h->part_idx, &h->dyn_box, &h->cost_tab, &h->stage_h, &h->stage_d, &h->out_d,
                 &h->dyn_d, &h->stat_d})
    b->release();
  if (h->tables_dev) cudaFree(h->tables_dev);
  for (auto ev : h->ring) if (ev) cudaEventDestroy(ev);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto ev : h->ev_copy) if (ev) cudaEventDestroy(ev);
  for (auto ev : h->ev_done) if (ev) cudaEventDestroy(ev);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_blob) cudaEventDestroy(h->ev_blob);
  for (auto ev : h->ev_slice) if (ev) cudaEventDestroy(ev);
  for (auto ps : h->pstream) if (ps) cudaStreamDestroy(ps);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->copy_stream2) cudaStreamDestroy(h->copy_stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return FOT_OK;
}

extern "C" int fot_n_t_max(const fot_handle_t* h) { return h ? h->plan.n_t_max : FOT_ERR_ARG; }

extern "C" int fot_candidate_count(const fot_handle_t* h, int n_v, int has_brake) {
  if (!h || n_v < 0) return FOT_ERR_ARG;
  return h->plan.cfg.n_T * n_v * h->plan.cfg.n_d + (has_brake ? h->plan.cfg.n_B : 0);
}

// Block geometry: how many (speed, lateral) candidates one block takes, bounded by the shared
// memory its per-speed reference samples need.
static int sweep_geometry(const fot_handle* h, const fot_batch_t* b, SweepGeom* g, size_t* smem_bytes) {
  const int n_v_max = b->n_v_max;
  const int NT = h->plan.n_t_max, nd = h->plan.cfg.n_d, nB = h->plan.cfg.n_B;
  const size_t budget = std::min<size_t>((size_t)h->smem_optin, 200 * 1024);
  // obstacle ring: stage capacity in entries per plane.  Small fields: several whole time planes per
  // stage (about 4 KB); large fields: 512-entry chunks (12 KB).
  const int SPp = b->dyn_mode != FOT_DYN_NONE ? ((b->S * b->P + 3) & ~3) : 0;
  const int Mp = (b->n_static + 3) & ~3;
  int tile_cap = 0, n_stages = 0;
  if (SPp > 0 || Mp > 0) {
    if (SPp > 0 && SPp <= 512) tile_cap = std::min(kGmax, std::max(1, 160 / SPp)) * SPp;   // whole planes, about 4 KB per stage
    else if (SPp > 512) tile_cap = 512;
    else tile_cap = std::min(512, Mp);
    n_stages = 3;
  }
  const bool footprint = h->plan.cfg.n_circles > 0;
  // layout (doubles): tt | hot | js,lonc (6 kv) | jp,dend (2 jp_cap) | kin | [phase-2 region] | ints
  auto layout = [&](int kv, int* phase2_off, int* ints_off) {
    const size_t jp_cap = (size_t)kv * nd + (((size_t)kv * nd) & 1);
    const size_t head = (size_t)kTT * NT + (size_t)kHot * kv * NT + 6 * (size_t)kv + 2 * jp_cap;
    const size_t kin_sz = (size_t)kKin * kv * NT;
    const size_t p2 = (size_t)n_stages * 3 * tile_cap + (size_t)kSweepThreads * kRec;
    const bool alias = !footprint && p2 <= kin_sz;
    const size_t p2_off = alias ? head : head + kin_sz;
    const size_t ints = std::max(head + kin_sz, p2_off + p2);
    if (phase2_off) *phase2_off = (int)p2_off;
    if (ints_off) *ints_off = (int)ints;
    return ints * sizeof(double) + ((size_t)4 * kv + NT + kSweepThreads + (size_t)kv * kGmax) * sizeof(int32_t) +
           (size_t)kv * kGmax * kCullCap * sizeof(unsigned short) + 16;
  };
  auto bytes = [&](int kv) { return layout(kv, nullptr, nullptr); };
  const int nT = h->plan.cfg.n_T;
  const long long grid_total = (long long)nT * n_v_max * nd;
  // whole pairs per block when a pair fits (no pair straddles two blocks, so its reference samples and
  // cull lists are built once); otherwise plain 128-candidate chunks
  int ch = nd <= kSweepThreads ? (kSweepThreads / nd) * nd : kSweepThreads;
  ch = (int)std::min<long long>(ch, grid_total);
  ch = std::max(ch, 1);
  auto pairs_of = [&](int c) {
    const int p = (c % nd == 0) ? c / nd : (c - 1) / nd + 2;
    return (int)std::min<long long>((long long)nT * n_v_max, p);
  };
  int kv = std::max(pairs_of(ch), std::min(nB, 8));
  while (bytes(kv) > std::min<size_t>(budget, 72 * 1024) && ch > nd && ch > 1) {   // keep >= 3 blocks/SM when possible
    ch = nd <= ch / 2 ? (ch / 2 / nd) * nd : std::max(nd, ch / 2);
    kv = std::max(pairs_of(ch), std::min(nB, 8));
  }
  while (bytes(kv) > budget && ch > 1) { ch = std::max(1, ch / 2); kv = std::max(pairs_of(ch), 1); }
  if (bytes(kv) > budget) return fail(FOT_ERR_TOO_LARGE, "time grid too long for shared memory");
  g->ch_eff = ch;
  g->kv_cap = std::max(kv, 1);
  g->jp_cap = g->kv_cap * nd + ((g->kv_cap * nd) & 1);   // even: keeps the following tables 16-byte aligned
  int p2 = 0, io = 0;
  layout(g->kv_cap, &p2, &io);
  g->phase2_off = p2;
  g->ints_off = io;
  g->tile_cap = tile_cap;
  g->n_stages = n_stages;
  g->grid_blocks = (int)((grid_total + ch - 1) / ch);
  g->brake_blocks = nB > 0 ? (nB + g->kv_cap - 1) / g->kv_cap : 0;
  g->blocks_per_query = g->grid_blocks + g->brake_blocks;
  *smem_bytes = bytes(g->kv_cap);
  return FOT_OK;
}


// Geometry of the sample-major kernel (fot_sweep_items).  Returns false when the batch's shape is
// outside what that kernel covers (very long time grids, huge per-step obstacle counts); the
// candidate-major fot_sweep then runs.
static bool item_geometry(const fot_handle* h, const fot_batch_t* b, ItemGeom* g, size_t* smem_bytes,
                          bool want_fused_box = false, int bpc_override = 0) {
  const int NT = h->plan.n_t_max, nd = h->plan.cfg.n_d, nB = h->plan.cfg.n_B, nx = h->plan.cfg.nx;
  const bool has_dyn = b->dyn_mode != FOT_DYN_NONE;
  const long long SPl = has_dyn ? (long long)b->S * b->P : 0;
  if (NT > 128 || nd > 8192 || SPl > 32768 || b->n_static > 32768) return false;   // 128: one NumPy pairwise block
  const int SP = (int)SPl;
  int max_threads = kItemThreads;
  if (h->opt.item_threads > 0) max_threads = std::max(NT, std::min(kItemThreads, h->opt.item_threads));
  const int ppc_max = max_threads / NT;
  ItemGeom G{};
  G.chunks = (b->n_v_max + ppc_max - 1) / ppc_max;
  G.ppc = (b->n_v_max + G.chunks - 1) / G.chunks;
  G.grid_blocks = h->plan.cfg.n_T * G.chunks;
  G.ppb = nB > 0 ? std::min(nB, ppc_max) : 0;
  G.brake_blocks = nB > 0 ? (nB + G.ppb - 1) / G.ppb : 0;
  G.blocks_per_query = G.grid_blocks + G.brake_blocks;
  // blocks per CTA: a whole query per CTA once the batch alone fills the GPU several times over,
  // one block per CTA for single plan() calls
  {
    const int sms = h->sms;
    const long long total = (long long)b->n_q * G.blocks_per_query;
    long long bpc = total / ((long long)sms * 2 * 4);
    bpc = std::max<long long>(1, std::min<long long>(bpc, G.blocks_per_query));
    if (bpc_override > 0) bpc = std::max(1, std::min(bpc_override, (int)G.blocks_per_query));
    if (h->opt.bpc > 0) bpc = std::max(1, std::min(h->opt.bpc, (int)G.blocks_per_query));
    G.ctas_per_query = (int)((G.blocks_per_query + bpc - 1) / bpc);
    G.bpc = (G.blocks_per_query + G.ctas_per_query - 1) / G.ctas_per_query;
  }
  G.pcap = std::max(G.ppc, std::max(G.ppb, 1));
  G.threads = (G.pcap * NT + 31) / 32 * 32;
  G.ct_lcap = std::max(nd, std::max(G.ppb, 1));
  G.nw4 = (nd + 3) / 4;
  G.nwc = (nd + 31) / 32;
  const int max_viol = b->dyn_mode == FOT_DYN_DISTRIBUTION ? (int)std::floor(h->plan.cfg.chance_epsilon * (double)b->S) : 0;
  G.vwords = max_viol > 0 ? (b->S + 31) / 32 : 0;
  G.lcap = std::max(1, SP + b->n_static);
  G.ochunk = G.lcap <= 256 ? G.lcap : 128;
  G.qcap = h->opt.qcap;                   // (the tests shrink it to force the queue-full path)
  G.spline_smem = nx <= 128 ? 1 : 0;
  const size_t dyn_bytes = (size_t)SP * b->T_obs * 16;
  // option fused_box (tests, tuning): trajectory boxes in the sweep even for a resident tensor
  const bool fuse = want_fused_box || h->opt.fused_box;
  auto layout = [&](bool stage) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) / 16 * 16; return (int32_t)o; };
    G.o_row = take((size_t)G.pcap * NT * kRowW * 8);
    G.o_sdl = take((size_t)G.pcap * 8);
    G.o_ct = take((size_t)(G.pcap + 2 * G.ct_lcap) * 8);
    G.o_dgrid = take((size_t)nd * 8);
    G.o_vlast = take((size_t)G.pcap * nd * 8);
    G.o_spl = take(G.spline_smem ? (size_t)9 * nx * 8 : 0);
    G.o_dyn = take(stage ? dyn_bytes : 0);
    G.o_box = take(stage && fuse ? (size_t)SP * 16 : 0);
    G.o_fn = take((size_t)G.pcap * 4);
    G.o_flags = take((size_t)G.pcap * G.nw4 * 4);
    G.o_hit = take((size_t)G.pcap * G.nwc * 4);
    G.o_viol = take((size_t)G.pcap * nd * G.vwords * 4);
    G.n_zero = (int32_t)((off - (size_t)G.o_flags) / 4);
    G.o_queue = take((size_t)G.qcap * 4);
    G.o_list = take((size_t)G.lcap * 4);
    G.o_clean = take((size_t)G.pcap * G.nwc * 4);
    G.o_slow = take((size_t)G.threads * 2);
    G.stage_dyn = stage ? 1 : 0;
    G.fused_box = stage && fuse ? 1 : 0;
    return off;
  };
  // stage the query's obstacle block in shared memory when two blocks per SM still fit
  bool stage = has_dyn && dyn_bytes > 0 && dyn_bytes <= 64 * 1024 && ((uintptr_t)b->dyn & 15) == 0;
  stage = stage && h->opt.stage_dyn;
  size_t bytes = layout(stage);
  if (stage && bytes > 110 * 1024) { stage = false; bytes = layout(false); }
  if (bytes > (size_t)h->smem_optin) return false;
  *g = G;
  *smem_bytes = bytes;
  return true;
}

// Geometry of the warp-local kernel (fot_sweep_warp): the block decomposition of fot_sweep_items, per-warp obstacle
// lists and survivor queues, and the per-block state in two copies.  False when the shape is outside its range
// (more than kWarpListCap obstacle entries per query, long time grids): fot_sweep_items then runs.
static bool warp_geometry(const fot_handle* h, const fot_batch_t* b, WarpGeom* g, size_t* smem_bytes,
                          bool want_fused_box = false, int bpc_override = 0) {
  const int NT = h->plan.n_t_max, nd = h->plan.cfg.n_d, nB = h->plan.cfg.n_B, nx = h->plan.cfg.nx;
  const bool has_dyn = b->dyn_mode != FOT_DYN_NONE;
  const long long SPl = has_dyn ? (long long)b->S * b->P : 0;
  if (NT > 128 || nd > 8192 || SPl + b->n_static > kWarpListCap) return false;
  const int SP = (int)SPl;
  int max_threads = kItemThreads;
  if (h->opt.item_threads > 0) max_threads = std::max(NT, std::min(kItemThreads, h->opt.item_threads));
  const int ppc_max = max_threads / NT;
  WarpGeom G{};
  G.chunks = (b->n_v_max + ppc_max - 1) / ppc_max;
  G.ppc = (b->n_v_max + G.chunks - 1) / G.chunks;
  G.grid_blocks = h->plan.cfg.n_T * G.chunks;
  G.ppb = nB > 0 ? std::min(nB, ppc_max) : 0;
  G.brake_blocks = nB > 0 ? (nB + G.ppb - 1) / G.ppb : 0;
  G.blocks_per_query = G.grid_blocks + G.brake_blocks;
  {
    const long long total = (long long)b->n_q * G.blocks_per_query;
    long long bpc = total / ((long long)h->sms * 2 * 4);
    bpc = std::max<long long>(1, std::min<long long>(bpc, G.blocks_per_query));
    if (bpc_override > 0) bpc = std::max(1, std::min(bpc_override, (int)G.blocks_per_query));
    if (h->opt.bpc > 0) bpc = std::max(1, std::min(h->opt.bpc, (int)G.blocks_per_query));
    G.ctas_per_query = (int)((G.blocks_per_query + bpc - 1) / bpc);
    G.bpc = (G.blocks_per_query + G.ctas_per_query - 1) / G.ctas_per_query;
  }
  G.pcap = std::max(G.ppc, std::max(G.ppb, 1));
  G.threads = (G.pcap * NT + 31) / 32 * 32;
  G.ct_lcap = std::max(nd, std::max(G.ppb, 1));
  G.nw4 = (nd + 3) / 4;
  G.nwc = (nd + 31) / 32;
  const int max_viol = b->dyn_mode == FOT_DYN_DISTRIBUTION ? (int)std::floor(h->plan.cfg.chance_epsilon * (double)b->S) : 0;
  G.vwords = max_viol > 0 ? (b->S + 31) / 32 : 0;
  G.lcap = std::max(1, SP + b->n_static);
  G.spline_smem = nx <= 128 ? 1 : 0;
  const size_t dyn_bytes = (size_t)SP * b->T_obs * 16;
  const bool fuse = want_fused_box || h->opt.fused_box;
  auto layout = [&](bool stage) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) / 16 * 16; return (int32_t)o; };
    G.o_row = take((size_t)G.pcap * NT * kRowW * 8);
    G.o_dgrid = take((size_t)nd * 8);
    G.o_spl = take(G.spline_smem ? (size_t)9 * nx * 8 : 0);
    G.o_dyn = take(stage ? dyn_bytes : 0);
    G.o_box = take((size_t)SP * 16);                      // trajectory boxes: built here (fused) or copied from fot_prepass
    G.o_wlist = take((size_t)(G.threads / 32) * G.lcap * 4);
    G.o_wq = take((size_t)(G.threads / 32) * kWarpQueue * 4);
    G.o_bq = take((size_t)kBlockQueue * 4);
    G.o_buf = take(0);
    size_t bo = 0;
    auto btake = [&](size_t bytes) { size_t o = bo; bo = (bo + bytes + 15) / 16 * 16; return (int32_t)o; };
    G.b_fnr = btake((size_t)G.pcap * 4);
    G.b_flags = btake((size_t)G.pcap * G.nw4 * 4);
    G.b_hit = btake((size_t)G.pcap * G.nwc * 4);
    G.b_viol = btake((size_t)G.pcap * nd * G.vwords * 4);
    G.b_dirty = btake((size_t)G.pcap * G.nwc * 4);
    G.b_qctl = btake(16);
    G.n_zero = (int32_t)((bo - (size_t)G.b_fnr) / 4);
    G.b_sdl = btake((size_t)G.pcap * 8);
    G.b_ct = btake((size_t)(G.pcap + 2 * G.ct_lcap) * 8);
    G.b_vlast = btake((size_t)G.pcap * nd * 8);
    G.b_span = btake((size_t)G.pcap * 8);
    G.buf_bytes = (int32_t)bo;
    off += 2 * bo;
    G.stage_dyn = stage ? 1 : 0;
    G.fused_box = stage && fuse ? 1 : 0;
    return off;
  };
  bool stage = has_dyn && dyn_bytes > 0 && dyn_bytes <= 64 * 1024 && ((uintptr_t)b->dyn & 15) == 0;
  stage = stage && h->opt.stage_dyn;
  size_t bytes = layout(stage);
  if (stage && bytes > 110 * 1024) { stage = false; bytes = layout(false); }
  if (bytes > (size_t)h->smem_optin) return false;
  *g = G;
  *smem_bytes = bytes;
  return true;
}


// Geometry of the pair-per-warp kernel (fot_sweep_pairs): one CTA per query (several for small batches), every warp
// with a private slice of shared memory for the pair it is working on.  False when the shape is outside its range
// (time grids beyond 64 samples, lateral grids beyond 96 targets, tables that do not fit): fot_sweep_items then runs.
static bool pair_geometry(const fot_handle* h, const fot_batch_t* b, PairGeom* g, size_t* smem_bytes,
                          bool want_fused_box = false, int cpq_override = 0) {
  const int NT = h->plan.n_t_max, nd = h->plan.cfg.n_d, nB = h->plan.cfg.n_B, nx = h->plan.cfg.nx;
  const bool has_dyn = b->dyn_mode != FOT_DYN_NONE;
  const long long SPl = has_dyn ? (long long)b->S * b->P : 0;
  if (NT > kPairNT || nd > kPairND || b->n_v_max > kPairNV || nx > kPairNX || h->plan.cfg.n_T > kPairNH || SPl > (1 << 20) || b->n_static > (1 << 20)) return false;
  if (SPl * std::max(1, b->T_obs) >= (1ll << 27)) return false;                            // list entries: element offsets
  const int SP = (int)SPl;
  PairGeom G{};
  G.nw4 = (nd + 3) / 4;
  G.nwc = (nd + 31) / 32;
  const int max_viol = b->dyn_mode == FOT_DYN_DISTRIBUTION ? (int)std::floor(h->plan.cfg.chance_epsilon * (double)b->S) : 0;
  G.vwords = max_viol > 0 ? (b->S + 31) / 32 : 0;
  G.viol_bytes = (nd * G.vwords * 4 + 15) / 16 * 16;
  const size_t dyn_bytes = (size_t)SP * b->T_obs * 16;
  const bool fuse = want_fused_box || h->opt.fused_box;
  auto layout = [&](bool stage) {
    size_t off = sizeof(PairShared);
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) / 16 * 16; return (int32_t)o; };
    const bool box = has_dyn && ((stage && fuse) || (size_t)SP * 16 <= 8 * 1024);
    G.o_dyn = take(stage ? dyn_bytes : 0);
    G.o_box = take(box ? (size_t)SP * 16 : 0);
    G.o_viol = take((size_t)kPairWarps * G.viol_bytes);
    G.stage_dyn = stage ? 1 : 0;
    G.fused_box = stage && fuse ? 1 : 0;
    G.box_smem = box ? 1 : 0;
    return off;
  };
  // stage the query's obstacle block in shared memory when two CTAs per SM still fit
  bool stage = has_dyn && dyn_bytes > 0 && dyn_bytes <= 64 * 1024 && ((uintptr_t)b->dyn & 15) == 0;
  stage = stage && h->opt.stage_dyn;
  size_t bytes = layout(stage);
  if (stage && bytes > 113 * 1024 - 512) { stage = false; bytes = layout(false); }
  if (bytes > (size_t)h->smem_optin) return false;
  // CTAs per query: one once the batch alone fills the GPU several times over; for small batches the pairs of a
  // query are dealt to several CTAs (about one pair per warp for a single plan() call)
  {
    const int n_units = h->plan.cfg.n_T * b->n_v_max + nB;
    long long cpq = ((long long)h->sms * 2 * 4 + b->n_q - 1) / b->n_q;
    cpq = std::max<long long>(1, std::min<long long>(cpq, (n_units + kPairWarps - 1) / kPairWarps));
    if (cpq_override > 0) cpq = std::max(1, std::min(cpq_override, n_units));
    if (h->opt.pair_cpq > 0) cpq = std::max(1, std::min(h->opt.pair_cpq, n_units));
    G.ctas_per_query = (int)cpq;
  }
  *g = G;
  *smem_bytes = bytes;
  return true;
}

static int check_batch(const fot_handle* h, const fot_batch_t* b, const fot_result_t* r) {
  if (!h || !b || !r) return fail(FOT_ERR_ARG, "null argument");
  if (b->n_q < 1 || b->n_v_max < 1) return fail(FOT_ERR_ARG, "n_q and n_v_max must be >= 1");
  if (!b->frenet || !b->target_speed || !b->limits || !b->stop_dist || !b->v_grid || !b->n_v)
    return fail(FOT_ERR_ARG, "null query array");
  if (b->n_static < 0 || (b->n_static > 0 && !b->static_obs)) return fail(FOT_ERR_ARG, "static obstacles missing");
  if (b->dyn_mode != FOT_DYN_NONE) {
    if (b->dyn_mode != FOT_DYN_SINGLE && b->dyn_mode != FOT_DYN_DISTRIBUTION) return fail(FOT_ERR_ARG, "bad dyn_mode");
    if (!b->dyn || b->S < 1 || b->P < 1 || b->T_obs < 1) return fail(FOT_ERR_ARG, "dynamic obstacle shape");
    if (b->dyn_mode == FOT_DYN_SINGLE && b->S != 1) return fail(FOT_ERR_ARG, "single-sample mode needs S == 1");
    if (b->dyn_mode == FOT_DYN_DISTRIBUTION && b->S > 64 &&
        std::floor(h->plan.cfg.chance_epsilon * (double)b->S) > 0.0)
      return fail(FOT_ERR_ARG, "chance_epsilon > 0 supports at most 64 samples");
  }
  if (!r->best_idx || !r->best_cost || !r->stats || !r->winner_len || !r->winner)
    return fail(FOT_ERR_ARG, "null result array");
  if ((r->cand_cat || r->cand_cost) && r->cand_stride < fot_candidate_count(h, b->n_v_max, 1))
    return fail(FOT_ERR_ARG, "cand_stride too small");
  if (r->winner_samples < 0) return fail(FOT_ERR_ARG, "winner_samples must be >= 0");
  return FOT_OK;
}

// q_off / q_total: this launch handles queries [q_off, q_off + b->n_q) of a batch of q_total whose chunks may be
// in flight on two streams at once, so every scratch buffer is sized for the batch and addressed by q_off.
// cuStreamWriteValue32 through the runtime's driver entry-point lookup (no link-time dependency on libcuda): a
// stream-ordered 4-byte write with a memory fence before it, far cheaper than a 4-byte cudaMemcpyAsync.
typedef int (*StreamWrite32)(cudaStream_t, unsigned long long, unsigned, unsigned);
static StreamWrite32 stream_write32() {
  static StreamWrite32 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr{};
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (StreamWrite32)p;
  }();
  return fn;
}

struct Gate {
  unsigned* word = nullptr;      // device slice flags (nullptr: not gated)
  uint32_t epoch = 0;
  int per = 1;                   // queries per upload slice
};

static int launch_all(fot_handle* h, const fot_batch_t* b, const fot_result_t* r, cudaStream_t st, size_t q_off = 0,
                      size_t q_total = 0, Gate gate = Gate{}, int bpc_override = 0, bool record_span = true) {
  if (q_total == 0) q_total = (size_t)b->n_q;
  // kernel choice: the sample-major fot_sweep_items unless the shape is outside its range (or
  // FOT_SWEEP=generic asks for the candidate-major kernel, which the tests use as a cross-check)
  ItemGeom ig{};
  size_t ismem = 0;
  bool use_items = item_geometry(h, b, &ig, &ismem, gate.word != nullptr, bpc_override);
  if (h->opt.sweep == 2) use_items = false;
  // the two-barrier kernel on request (FOT_SWEEP=warp)
  WarpGeom wg{};
  size_t wsmem = 0;
  bool use_warp = h->opt.sweep == 3 && warp_geometry(h, b, &wg, &wsmem, gate.word != nullptr, bpc_override);
  if (use_warp && gate.word && !wg.fused_box) use_warp = false;
  if (h->opt.sweep == 3 && !use_warp && !gate.word) return fail(FOT_ERR_ARG, "FOT_SWEEP=warp: shape not supported by fot_sweep_warp");
  if (use_warp) {
    // same block decomposition and scratch as fot_sweep_items: the code below sizes everything from `ig`
    use_items = true;
    ig.blocks_per_query = wg.blocks_per_query; ig.ctas_per_query = wg.ctas_per_query; ig.threads = wg.threads;
    ig.fused_box = wg.fused_box;
  }
  // the pair-per-warp kernel: the default wherever its shape limits allow (FOT_SWEEP=pairs insists on it)
  PairGeom pg{};
  size_t psmem = 0;
  bool use_pairs = (h->opt.sweep == 0 || h->opt.sweep == 4) && use_items && pair_geometry(h, b, &pg, &psmem, gate.word != nullptr);
  if (use_pairs && gate.word && !pg.fused_box) use_pairs = false;
  if (h->opt.sweep == 4 && !use_pairs && !gate.word) return fail(FOT_ERR_ARG, "FOT_SWEEP=pairs: shape not supported by fot_sweep_pairs");
  if (use_pairs) {
    // same scratch (cost tables, trajectory boxes, one arg-min partial per CTA) as fot_sweep_items
    use_items = true;
    use_warp = false;
    ig.blocks_per_query = h->plan.cfg.n_T * b->n_v_max + h->plan.cfg.n_B;      // upper bound of ctas_per_query (scratch stride)
    ig.ctas_per_query = pg.ctas_per_query; ig.threads = kPairThreads;
    ig.fused_box = pg.fused_box;
    if (gate.word) { pg.gate = gate.word; pg.gate_epoch = gate.epoch; pg.gate_per = gate.per; pg.gate_q0 = (int32_t)q_off; }
  }
  if (gate.word && !(use_items && ig.fused_box)) return fail(FOT_ERR_ARG, "gated launch needs a sample-major sweep with a staged obstacle block");
  if (gate.word) {
    ig.gate = gate.word; ig.gate_epoch = gate.epoch; ig.gate_per = gate.per; ig.gate_q0 = (int32_t)q_off;
    wg.gate = gate.word; wg.gate_epoch = gate.epoch; wg.gate_per = gate.per; wg.gate_q0 = (int32_t)q_off;
  }
  if (h->opt.sweep == 1 && !use_items) return fail(FOT_ERR_ARG, "FOT_SWEEP=items: shape not supported by fot_sweep_items");
  SweepGeom g{};
  size_t smem = 0;
  if (!use_items) {
    int rc = sweep_geometry(h, b, &g, &smem);
    if (rc != FOT_OK) return rc;
  } else {
    g.blocks_per_query = ig.ctas_per_query;     // fot_winner reads one partial per CTA
  }
  const bool has_dyn = b->dyn_mode != FOT_DYN_NONE;
  const size_t n_part = (size_t)b->n_q * g.blocks_per_query;
  if ((size_t)b->n_q * (size_t)g.blocks_per_query > 0x7fffffffull) return fail(FOT_ERR_ARG, "batch too large for one launch");
  // partial winners: chunks of one batch may cut their queries into different numbers of CTAs; the stride that
  // addresses the scratch is the largest any chunk can use
  const size_t part_stride = use_items ? (size_t)std::max(ig.blocks_per_query, ig.ctas_per_query) : (size_t)g.blocks_per_query;
  if (q_off != 0 && !use_items) return fail(FOT_ERR_ARG, "chunk offsets need the fot_sweep_items path");
  CK(h->part_cost.reserve(q_total * part_stride * sizeof(double)));
  CK(h->part_idx.reserve(q_total * part_stride * sizeof(int32_t)));
  const int SP = has_dyn ? b->S * b->P : 0;
  const int SPp = (SP + 3) & ~3, Mp = (b->n_static + 3) & ~3;
  const int nq_s = b->static_per_query ? b->n_q : 1;
  const bool need_box = use_items && has_dyn && !ig.fused_box;
  const size_t cost_stride = (size_t)h->plan.cfg.n_T * (b->n_v_max + 2 * h->plan.cfg.n_d) + 3 * (size_t)h->plan.cfg.n_B;
  if (!use_items) {
    if (has_dyn) {
      CK(h->obs_tm.reserve((size_t)b->n_q * b->T_obs * 3 * SPp * sizeof(double)));
      CK(h->obs_max2.reserve((size_t)b->n_q * sizeof(double)));
    }
    if (b->n_static > 0) {
      CK(h->stat_tm.reserve((size_t)nq_s * 3 * Mp * sizeof(double)));
      CK(h->stat_max2.reserve((size_t)nq_s * sizeof(double)));
    }
  } else {
    if (need_box) CK(h->dyn_box.reserve(q_total * SP * sizeof(float4)));
    CK(h->cost_tab.reserve(q_total * cost_stride * sizeof(double)));
  }
  float4* dyn_box = need_box ? (float4*)h->dyn_box.p + q_off * SP : nullptr;
  double* cost_tab = use_items ? (double*)h->cost_tab.p + q_off * cost_stride : nullptr;

  Batch B{};
  B.n_q = b->n_q; B.n_v_max = b->n_v_max;
  B.frenet = b->frenet; B.target = b->target_speed; B.limits = b->limits; B.stop_dist = b->stop_dist;
  B.v_grid = b->v_grid; B.n_v = b->n_v;
  B.static_tm = (!use_items && b->n_static > 0) ? (const double*)h->stat_tm.p : nullptr;
  B.static_max2 = (!use_items && b->n_static > 0) ? (const double*)h->stat_max2.p : nullptr;
  B.n_static = b->n_static; B.static_per_query = b->static_per_query;
  B.obs_tm = (!use_items && has_dyn) ? (const double*)h->obs_tm.p : nullptr;
  B.obs_max2 = (!use_items && has_dyn) ? (const double*)h->obs_max2.p : nullptr;
  B.S = has_dyn ? b->S : 0; B.P = has_dyn ? b->P : 0; B.T_obs = has_dyn ? b->T_obs : 0; B.dyn_mode = b->dyn_mode;
  B.dyn_raw = has_dyn ? b->dyn : nullptr;
  B.static_raw = b->n_static > 0 ? b->static_obs : nullptr;
  B.dyn_box = dyn_box;
  B.cost_tab = cost_tab;
  Out O{};
  O.best_idx = r->best_idx; O.best_cost = r->best_cost; O.stats = r->stats; O.winner_len = r->winner_len;
  O.winner = r->winner; O.cand_cat = r->cand_cat; O.cand_cost = r->cand_cost; O.cand_stride = r->cand_stride;
  O.part_cost = (double*)h->part_cost.p + q_off * part_stride; O.part_idx = (int32_t*)h->part_idx.p + q_off * part_stride;
  if (h->mirror.winner && q_off == 0 && q_total == (size_t)b->n_q) {     // whole-batch launches only (fot_plan_batch_device)
    O.m_best_idx = h->mirror.best_idx; O.m_best_cost = h->mirror.best_cost; O.m_stats = h->mirror.stats;
    O.m_winner_len = h->mirror.winner_len; O.m_winner = h->mirror.winner;
  }

  cudaEvent_t* ring = h->ring.data() + (size_t)(h->n_launch % fot_handle::kRing) * 4;
  if (record_span) CK(cudaEventRecord(h->ev0, st));
  {
    const size_t n_cat = r->cand_cat ? (size_t)b->n_q * r->cand_stride : 0;
    const size_t work = std::max<size_t>((size_t)b->n_q * FOT_N_STATS, n_cat);
    const int blocks = (int)std::min<size_t>(1024, (work + 255) / 256);
    const bool planes = !use_items;
    fot_init_kernel<<<blocks, 256, 0, st>>>(r->stats, b->n_q * FOT_N_STATS,
                                            planes && has_dyn ? (double*)h->obs_max2.p : nullptr, planes && has_dyn ? b->n_q : 0,
                                            planes && b->n_static > 0 ? (double*)h->stat_max2.p : nullptr,
                                            planes && b->n_static > 0 ? nq_s : 0, r->cand_cat, n_cat);
  }
  CK(cudaEventRecord(ring[0], st));
  if (!use_items) {
    if (has_dyn) {
      const long long warps = (long long)b->n_q * SP;
      const int blocks = (int)((warps * 32 + 255) / 256);
      fot_obstacle_prepass<<<blocks, 256, 0, st>>>((const double2*)b->dyn, (double*)h->obs_tm.p,
                                                   (double*)h->obs_max2.p, b->n_q, SP, b->T_obs);
    }
    if (b->n_static > 0) {
      const int total = nq_s * b->n_static;
      fot_static_prepass<<<(total + 255) / 256, 256, 0, st>>>((const double2*)b->static_obs, (double*)h->stat_tm.p,
                                                              (double*)h->stat_max2.p, nq_s, b->n_static);
    }
  } else {
    // cost tables (one thread per polynomial profile) and trajectory boxes (one warp per trajectory), one launch
    const long long n_prof = (long long)h->plan.cfg.n_T * (b->n_v_max + h->plan.cfg.n_d) + 2 * h->plan.cfg.n_B;
    const long long cblocks = ((long long)b->n_q * n_prof + 255) / 256;
    const long long n_traj = need_box ? (long long)b->n_q * SP : 0;
    const long long ablocks = ((n_traj + kBoxPerWarp - 1) / kBoxPerWarp * 32 + 255) / 256;
    if (cblocks + ablocks > 0x7fffffffll) return fail(FOT_ERR_ARG, "batch too large for one launch");
    fot_prepass<<<(unsigned)(cblocks + ablocks), 256, 0, st>>>(h->plan, B, cost_tab, (unsigned)cblocks,
                                                               (const double2*)b->dyn, dyn_box, n_traj, b->T_obs);
  }
  CK(cudaEventRecord(ring[1], st));
  if (use_pairs) {
    // the modes this batch needs; the smallest instantiated superset runs (everything else is compiled out of it)
    int need = 0;
    if (b->n_static > 0) need |= PAIR_STATIC;
    if (h->plan.cfg.n_circles > 0) need |= PAIR_FOOTPRINT;
    if (pg.vwords > 0) need |= PAIR_BUDGET;
    if (r->cand_cat || r->cand_cost) need |= PAIR_OUTPUTS;
    if (has_dyn && !(pg.stage_dyn && pg.box_smem)) need |= PAIR_LOOSE;
    if (!h->plan.d_sorted) need |= PAIR_LOOSE;
    need |= h->opt.pair_feat;
    if (!h->opt.pair_simple) need = PAIR_ALL;
    int feat = PAIR_ALL;
    for (int i = 0; i < kPairInstanceCount; ++i)
      if ((kPairInstances[i] & need) == need) { feat = kPairInstances[i]; break; }
    pair_kernel(pg.fused_box != 0, feat)<<<(unsigned)n_part, kPairThreads, psmem, st>>>(h->plan, B, O, pg);
    h->last_pair_feat = feat;
  } else if (use_warp && wg.fused_box) fot_sweep_warp<true><<<(unsigned)n_part, wg.threads, wsmem, st>>>(h->plan, B, O, wg);
  else if (use_warp) fot_sweep_warp<false><<<(unsigned)n_part, wg.threads, wsmem, st>>>(h->plan, B, O, wg);
  else if (use_items && ig.fused_box) fot_sweep_items<true><<<(unsigned)n_part, ig.threads, ismem, st>>>(h->plan, B, O, ig);
  else if (use_items) fot_sweep_items<false><<<(unsigned)n_part, ig.threads, ismem, st>>>(h->plan, B, O, ig);
  else fot_sweep<<<(unsigned)n_part, kSweepThreads, smem, st>>>(h->plan, B, O, g);
  h->last_sweep_kind = use_pairs ? 4 : use_warp ? 3 : use_items ? 1 : 2;
  CK(cudaEventRecord(ring[2], st));
  fot_winner<<<b->n_q, 128, (size_t)kTT * h->plan.n_t_max * sizeof(double), st>>>(h->plan, B, O, g);
  if (O.m_winner && h->mirror_flag) fot_publish_kernel<<<1, 1, 0, st>>>(h->mirror_flag, ++h->mirror_seq);
  CK(cudaEventRecord(ring[3], st));
  if (record_span) CK(cudaEventRecord(h->ev1, st));
  h->n_launch++;
  CK(cudaGetLastError());
  h->timed = true;
  return FOT_OK;
}

extern "C" int fot_plan_batch_device(fot_handle_t* h, const fot_batch_t* b, const fot_result_t* r, void* stream) {
  int rc = check_batch(h, b, r);
  if (rc != FOT_OK) return rc;
  CK(cudaSetDevice(h->device));
  if (r->winner_samples != 0) return fail(FOT_ERR_ARG, "winner_samples applies to the host-result entry points only");
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  rc = launch_all(h, b, r, st);
  if (rc != FOT_OK) return rc;
  if (!stream) CK(cudaStreamSynchronize(st));
  return FOT_OK;
}

// First k samples of every winner series, packed [n][FOT_N_SERIES][k] (fot_result_t.winner_samples): a strided copy of
// 16-byte rows is slow on the copy engines, a contiguous one is not.
__global__ void fot_pack_heads(const double* __restrict__ winner, double* __restrict__ heads, long long n_rows, int NT, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * k) return;
  const long long row = i / k;
  heads[i] = winner[row * NT + (i - row * k)];
}

// Read-back of one chunk's winner block into the caller's arrays on stream `ds` (after `st`'s kernels).
static int read_back(fot_handle* h, const fot_result_t* r, const fot_result_t& dr, double* heads_d, int q0, int cq, int NT,
                     cudaStream_t st, cudaStream_t ds, cudaEvent_t done) {
  const int k = r->winner_samples > 0 ? std::min(r->winner_samples, NT) : 0;
  if (k > 0) {
    const long long n_rows = (long long)cq * FOT_N_SERIES;
    fot_pack_heads<<<(unsigned)((n_rows * k + 255) / 256), 256, 0, st>>>(dr.winner, heads_d, n_rows, NT, k);
  }
  CK(cudaEventRecord(done, st));
  CK(cudaStreamWaitEvent(ds, done, 0));
  CK(cudaMemcpyAsync(r->best_idx + q0, dr.best_idx, (size_t)cq * 4, cudaMemcpyDeviceToHost, ds));
  CK(cudaMemcpyAsync(r->best_cost + q0, dr.best_cost, (size_t)cq * 8, cudaMemcpyDeviceToHost, ds));
  CK(cudaMemcpyAsync(r->stats + (size_t)q0 * FOT_N_STATS, dr.stats, (size_t)cq * FOT_N_STATS * 4, cudaMemcpyDeviceToHost, ds));
  CK(cudaMemcpyAsync(r->winner_len + q0, dr.winner_len, (size_t)cq * 4, cudaMemcpyDeviceToHost, ds));
  if (k > 0)
    CK(cudaMemcpyAsync(r->winner + (size_t)q0 * FOT_N_SERIES * k, heads_d, (size_t)cq * FOT_N_SERIES * k * 8, cudaMemcpyDeviceToHost, ds));
  else
    CK(cudaMemcpyAsync(r->winner + (size_t)q0 * FOT_N_SERIES * NT, dr.winner, (size_t)cq * FOT_N_SERIES * NT * 8,
                       cudaMemcpyDeviceToHost, ds));
  if (r->cand_cat)
    CK(cudaMemcpyAsync(r->cand_cat + (size_t)q0 * r->cand_stride, dr.cand_cat, (size_t)cq * r->cand_stride,
                       cudaMemcpyDeviceToHost, ds));
  if (r->cand_cost)
    CK(cudaMemcpyAsync(r->cand_cost + (size_t)q0 * r->cand_stride, dr.cand_cost, (size_t)cq * r->cand_stride * 8,
                       cudaMemcpyDeviceToHost, ds));
  return FOT_OK;
}

// A failed multi-chunk call may leave earlier chunks in flight: kernels, upload slices and read-backs into the
// CALLER's result arrays, and (gated mode) CTAs spinning on slice flags that will never be written.  Nothing of that
// may still be running when the error is reported -- the caller is free to release its arrays, and the next call
// reuses the handle's scratch and gate flags.  Releases the waiting CTAs (their results are discarded), waits for
// every stream of the handle and clears the gate error word; keeps the error text of the failure.
static void drain_after_failure(fot_handle* h) {
  const std::string keep = g_err;
  if (h->gate_d.p && h->gate_epoch) {
    std::vector<uint32_t> w(kGateSlices, h->gate_epoch);
    cudaMemcpy(h->gate_d.p, w.data(), kGateSlices * sizeof(uint32_t), cudaMemcpyHostToDevice);
  }
  for (cudaStream_t s : {h->copy_stream, h->copy_stream2, h->stream, h->stream2, h->pstream[0], h->pstream[1], h->pstream[2],
                         h->pstream[3], h->d2h_stream})
    if (s) cudaStreamSynchronize(s);
  if (h->gate_d.p) cudaMemset((uint32_t*)h->gate_d.p + kGateSlices, 0, sizeof(uint32_t));
  cudaGetLastError();
  g_err = keep;
}

// device-side span of a call that ran on several streams (fot_last_kernel_ms): `first` carries ev0; ev1 is recorded
// on it after every other stream of the handle has been joined
static int close_span(fot_handle* h, cudaStream_t first) {
  for (cudaStream_t s : {h->stream, h->stream2, h->pstream[0], h->pstream[1], h->pstream[2], h->pstream[3]}) {
    if (s == first) continue;
    CK(cudaEventRecord(h->ev_join, s));
    CK(cudaStreamWaitEvent(first, h->ev_join, 0));
  }
  CK(cudaEventRecord(h->ev1, first));
  return FOT_OK;
}

static int plan_batch_host_impl(fot_handle_t* h, const fot_batch_t* b, const fot_result_t* r);
static int plan_batch_device_to_host_impl(fot_handle_t* h, const fot_batch_t* b, const fot_result_t* r, void* stream);

extern "C" int fot_plan_batch_host(fot_handle_t* h, const fot_batch_t* b, const fot_result_t* r) {
  int rc = check_batch(h, b, r);
  if (rc != FOT_OK) return rc;
  CK(cudaSetDevice(h->device));
  rc = plan_batch_host_impl(h, b, r);
  if (rc != FOT_OK) drain_after_failure(h);
  return rc;
}

extern "C" int fot_plan_batch_device_to_host(fot_handle_t* h, const fot_batch_t* b, const fot_result_t* r, void* stream) {
  int rc = check_batch(h, b, r);
  if (rc != FOT_OK) return rc;
  CK(cudaSetDevice(h->device));
  rc = plan_batch_device_to_host_impl(h, b, r, stream);
  if (rc != FOT_OK) drain_after_failure(h);
  return rc;
}

// Host-pointer entry point.  Small per-query arrays go through one pinned blob; the obstacle tensor
// is copied straight from the caller's memory (pinned memory makes that a true async DMA).  Large
// batches are cut into chunks of queries: chunk c+1's obstacle upload runs on the copy stream while
// chunk c's kernels run on the compute stream, and each chunk's winners go back as soon as its
// kernels finish, directly into the caller's result arrays.
static int plan_batch_host_impl(fot_handle_t* h, const fot_batch_t* b, const fot_result_t* r) {
  int rc = FOT_OK;
  cudaStream_t st = h->stream;
  const int nq = b->n_q, NT = h->plan.n_t_max;
  // ---- small per-query arrays: one pinned blob, one H2D -------------------------------
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  const size_t o_fr = take((size_t)nq * 6 * 8), o_tg = take((size_t)nq * 8), o_lm = take((size_t)nq * 4 * 8),
               o_sd = take((size_t)nq * 8), o_vg = take((size_t)nq * b->n_v_max * 8), o_nv = take((size_t)nq * 4);
  CK(h->stage_h.reserve(off));
  CK(h->stage_d.reserve(off));
  char* sh = (char*)h->stage_h.p;
  memcpy(sh + o_fr, b->frenet, (size_t)nq * 6 * 8);
  memcpy(sh + o_tg, b->target_speed, (size_t)nq * 8);
  memcpy(sh + o_lm, b->limits, (size_t)nq * 4 * 8);
  memcpy(sh + o_sd, b->stop_dist, (size_t)nq * 8);
  memcpy(sh + o_vg, b->v_grid, (size_t)nq * b->n_v_max * 8);
  memcpy(sh + o_nv, b->n_v, (size_t)nq * 4);
  // The compute stream carries kernels and event records only: every copy runs on the copy / d2h
  // streams.  (A copy on the compute stream makes the driver service that stream's later timed event
  // records on the copy engine, where they queue behind the uploads and serialise the pipeline.)
  CK(cudaMemcpyAsync(h->stage_d.p, sh, off, cudaMemcpyHostToDevice, h->copy_stream));
  char* sd = (char*)h->stage_d.p;
  const bool has_dyn = b->dyn_mode != FOT_DYN_NONE;
  const size_t dyn_q_bytes = has_dyn ? (size_t)b->S * b->P * b->T_obs * 16 : 0;
  if (has_dyn) CK(h->dyn_d.reserve((size_t)nq * dyn_q_bytes));
  if (b->n_static > 0) {
    const size_t bytes = (size_t)(b->static_per_query ? nq : 1) * b->n_static * 16;
    CK(h->stat_d.reserve(bytes));
    CK(cudaMemcpyAsync(h->stat_d.p, b->static_obs, bytes, cudaMemcpyHostToDevice, h->copy_stream));
  }
  // ---- device result blob ---------------------------------------------------------------
  size_t ro = 0;
  auto rtake = [&](size_t bytes) { size_t o = ro; ro = align_up(ro + bytes); return o; };
  const size_t r_bi = rtake((size_t)nq * 4), r_bc = rtake((size_t)nq * 8), r_st = rtake((size_t)nq * FOT_N_STATS * 4),
               r_wl = rtake((size_t)nq * 4), r_w = rtake((size_t)nq * FOT_N_SERIES * NT * 8);
  const size_t r_cc = r->cand_cat ? rtake((size_t)nq * r->cand_stride) : 0;
  const size_t r_cs = r->cand_cost ? rtake((size_t)nq * r->cand_stride * 8) : 0;
  const int k_head = r->winner_samples > 0 ? std::min(r->winner_samples, NT) : 0;
  const size_t r_wh = k_head ? rtake((size_t)nq * FOT_N_SERIES * k_head * 8) : 0;
  CK(h->out_d.reserve(ro));
  char* od = (char*)h->out_d.p;
  h->last_winner_d = (const double*)(od + r_w);     // fot_fetch_winners
  h->last_winner_nq = nq;

  // chunking: only worth it when the obstacle upload is large.  The kernels are the longer leg of the
  // pipeline (the upload runs at PCIe speed), so the first chunks are small -- the sweep starts after
  // 1/16 of the upload -- and the rest are large enough to keep the wave tail short.
  int n_chunks = 1;
  constexpr int kMaxChunks = fot_handle::kMaxChunks;
  int bounds[kMaxChunks + 1];
  bounds[0] = 0;
  for (int c = 1; c <= kMaxChunks; ++c) bounds[c] = nq;
  if ((size_t)nq * dyn_q_bytes > ((size_t)8 << 20) && nq >= 1024) {
    // chunk sizes in whole waves of the sweep (one CTA per query, two CTAs per SM) where possible
    const int wave = 2 * h->sms;
    auto waves = [&](int frac16) { const int w = std::max(1, (int)((long long)nq * frac16 / 16 / wave)); return w * wave; };
    const int cuts[] = {waves(1), waves(3), waves(6), waves(11), nq};  // about 1/16, 1/8, 3/16, 5/16, 5/16 of the queries
    n_chunks = 5;
    for (int c = 0; c < n_chunks; ++c) bounds[c + 1] = std::min(nq, std::max(cuts[c], bounds[c]));
  }
  if (h->opt.host_chunks > 0) {
    n_chunks = std::max(1, std::min(kMaxChunks, h->opt.host_chunks));
    const int per_ = (nq + n_chunks - 1) / n_chunks;
    for (int c = 0; c <= n_chunks; ++c) bounds[c] = std::min(nq, c * per_);
  }
  if (!h->opt.chunk_waves.empty()) {                       // tuning: chunk sizes in sweep waves, "1,2,3,4,3" (+ the rest)
    const char* env = h->opt.chunk_waves.c_str();
    const int wave = 2 * h->sms;
    n_chunks = 0;
    int at = 0;
    for (const char* p = env; *p && n_chunks < kMaxChunks - 1 && at < nq;) {
      at = std::min(nq, at + std::max(1, atoi(p)) * wave);
      bounds[++n_chunks] = at;
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
    if (at < nq) bounds[++n_chunks] = nq;
    for (int c = n_chunks + 1; c <= kMaxChunks; ++c) bounds[c] = nq;
  }
  // chunks alternate between two compute streams, so that the next chunk's CTAs fill the SMs the tail wave of
  // the previous chunk leaves idle (the chunks touch disjoint scratch, see launch_all)
  bool two_streams = n_chunks > 1;
  two_streams = two_streams && h->opt.host_streams >= 2;
  if (two_streams) {
    ItemGeom ig{};
    size_t ismem = 0;
    fot_batch_t probe = *b;
    probe.n_q = bounds[1] - bounds[0];
    if (!item_geometry(h, &probe, &ig, &ismem) || h->opt.sweep == 2) two_streams = false;
  }
  // Gated launches (large uploads through fot_sweep_items): the sweep does not wait for whole chunks.  The upload
  // is cut into fine slices, each followed by a 4-byte copy that publishes "queries uploaded so far"; the sweep
  // is launched over few, large ranges of queries and every CTA waits for its own query's slice (ItemGeom::gate).
  // The ranges only exist so that winners of the first queries go back while the last are still being swept.
  bool gated = false;
  int n_up = 0, ub[kGateSlices + 1] = {0};
  Gate gate{};
  {
    if (h->opt.gated && n_chunks > 1 && has_dyn && nq < (1 << 22) && h->opt.sweep != 2) {
      ItemGeom ig{};
      size_t ismem = 0;
      gated = item_geometry(h, b, &ig, &ismem, true) && ig.fused_box;
    }
  }
  if (gated) {
    const bool fresh = h->gate_d.p == nullptr;
    CK(h->gate_d.reserve((kGateSlices + 1) * sizeof(uint32_t)));
    CK(h->gate_h.reserve((kGateSlices + 1) * sizeof(uint32_t)));
    if (fresh) CK(cudaMemset(h->gate_d.p, 0, (kGateSlices + 1) * sizeof(uint32_t)));
    if (++h->gate_epoch == 0) h->gate_epoch = 1;             // 0 is the value of a fresh flag
    gate.word = (unsigned*)h->gate_d.p;
    gate.epoch = h->gate_epoch;
    n_up = std::max(1, std::min(kGateSlices, h->opt.gate_uploads));
    gate.per = (nq + n_up - 1) / n_up;
    n_up = (nq + gate.per - 1) / gate.per;
    for (int u = 0; u <= n_up; ++u) ub[u] = std::min(nq, u * gate.per);
    // launch ranges: about 1/2, 5/16, 3/16 of the queries, in whole waves
    const int wave = 2 * h->sms;
    auto waves = [&](int frac16) { const int w = std::max(1, (int)((long long)nq * frac16 / 16 / wave)); return w * wave; };
    if (h->opt.host_chunks <= 0 && h->opt.chunk_waves.empty()) {
      const int cuts[] = {waves(8), waves(13), nq};
      n_chunks = 3;
      for (int c = 0; c < n_chunks; ++c) bounds[c + 1] = std::min(nq, std::max(cuts[c], bounds[c]));
      for (int c = n_chunks + 1; c <= kMaxChunks; ++c) bounds[c] = nq;
    }
    two_streams = n_chunks > 1;
  }
  const auto t_wall0 = std::chrono::steady_clock::now();
  const bool dbg = h->opt.debug_timing != 0;
  cudaEvent_t d0 = nullptr, d1 = nullptr, d2 = nullptr, d3 = nullptr, d4 = nullptr;
  cudaEvent_t tl[kMaxChunks][4] = {};                    // debug timeline per chunk: upload done, kernels start / end, read-back done
  if (dbg) for (auto& row : tl) for (auto& ev : row) cudaEventCreate(&ev);
  if (dbg) { cudaEventCreate(&d0); cudaEventCreate(&d1); cudaEventCreate(&d2); cudaEventCreate(&d3); cudaEventCreate(&d4);
             cudaEventRecord(d0, st); cudaEventRecord(d4, h->copy_stream); }
  CK(cudaEventRecord(h->ev_blob, h->copy_stream));
  // every obstacle upload is queued first (they depend on nothing), one event per chunk
  if (gated) {
    // each slice is followed, in its stream, by the write of its flag (two alternating upload streams are a
    // tuning knob; measured slower than one)
    uint32_t* words = (uint32_t*)h->gate_h.p;
    const int n_cs = h->opt.gate_copy_streams;
    const bool flag_stream = h->opt.gate_flag_stream && n_cs == 1;
    CK(cudaStreamWaitEvent(h->copy_stream2, h->ev_blob, 0));      // nothing of this call before the previous call's flags are history
    for (int u = 0; u < n_up; ++u) {
      const int q0 = ub[u], cq = ub[u + 1] - q0;
      if (cq <= 0) continue;
      cudaStream_t cs = (n_cs == 2 && (u & 1)) ? h->copy_stream2 : h->copy_stream;
      char* dst = (char*)h->dyn_d.p + (size_t)q0 * dyn_q_bytes;
      const char* src = (const char*)b->dyn + (size_t)q0 * dyn_q_bytes;
      CK(cudaMemcpyAsync(dst, src, (size_t)cq * dyn_q_bytes, cudaMemcpyHostToDevice, cs));
      words[u] = gate.epoch;
      // FOT_GATE_FLAG_STREAM=1: the flag is written by a second stream behind an event of the slice, so that the
      // upload stream carries nothing but copies and event records (a stream write between two slices costs the
      // upload ~15 us of DMA idle time: uploads done at 3.3 ms instead of 3.1).  Faster in isolation (3.75 against
      // 3.86 ms per call), no better inside bench.py, and it depends on the two streams not sharing a hardware
      // queue; the default keeps the flag in line, where its order behind the slice needs nothing else.
      cudaStream_t fs = cs;
      if (flag_stream) {
        CK(cudaEventRecord(h->ev_slice[u], cs));
        CK(cudaStreamWaitEvent(h->copy_stream2, h->ev_slice[u], 0));
        fs = h->copy_stream2;
      }
      if (StreamWrite32 wr = h->opt.gate_memcpy ? nullptr : stream_write32()) {
        if (wr(fs, (unsigned long long)(uintptr_t)(gate.word + u), gate.epoch, 0u) != 0)
          return fail(FOT_ERR_CUDA, "cuStreamWriteValue32");
      } else {
        CK(cudaMemcpyAsync(gate.word + u, words + u, sizeof(uint32_t), cudaMemcpyHostToDevice, fs));
      }
    }
  } else if (has_dyn)
    for (int c = 0; c < n_chunks; ++c) {
      const int q0 = bounds[c], cq = bounds[c + 1] - q0;
      if (cq <= 0) continue;
      char* dst = (char*)h->dyn_d.p + (size_t)q0 * dyn_q_bytes;
      const char* src = (const char*)b->dyn + (size_t)q0 * dyn_q_bytes;
      CK(cudaMemcpyAsync(dst, src, (size_t)cq * dyn_q_bytes, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaEventRecord(h->ev_copy[c], h->copy_stream));
      if (dbg) cudaEventRecord(tl[c][0], h->copy_stream);
    }
  if (dbg) {
    if (gated) { cudaEventRecord(h->ev_join, h->copy_stream2); cudaStreamWaitEvent(h->copy_stream, h->ev_join, 0); }
    cudaEventRecord(d1, h->copy_stream);
    cudaPointerAttributes pa{};
    cudaError_t pe = cudaPointerGetAttributes(&pa, b->dyn);
    fprintf(stderr, "[fot] dyn pointer attr: err=%d type=%d (0 unregistered, 1 host, 2 device, 3 managed)\n", (int)pe, (int)pa.type);
  }
  int n_issued = 0;
  cudaStream_t span_first = nullptr;
  for (int c = 0; c < n_chunks; ++c) {
    const int q0 = bounds[c], cq = bounds[c + 1] - q0;
    if (cq <= 0) continue;
    // Gated ranges run on streams of descending priority: every range's kernels are eligible from the start, and a
    // CTA of a later range that got an SM slot early would only sit there waiting for its slice of the upload.
    cudaStream_t st = gated ? h->pstream[n_issued & 3] : (two_streams && (n_issued & 1)) ? h->stream2 : h->stream;
    ++n_issued;
    if (has_dyn && !gated) CK(cudaStreamWaitEvent(st, h->ev_copy[c], 0));
    else CK(cudaStreamWaitEvent(st, h->ev_blob, 0));
    if (!span_first) { span_first = st; CK(cudaEventRecord(h->ev0, st)); }
    if (dbg && c == 0) cudaEventRecord(d2, st);
    if (dbg) cudaEventRecord(tl[c][1], st);
    fot_batch_t db = *b;
    db.n_q = cq;
    db.frenet = (const double*)(sd + o_fr) + (size_t)q0 * 6;
    db.target_speed = (const double*)(sd + o_tg) + q0;
    db.limits = (const double*)(sd + o_lm) + (size_t)q0 * 4;
    db.stop_dist = (const double*)(sd + o_sd) + q0;
    db.v_grid = (const double*)(sd + o_vg) + (size_t)q0 * b->n_v_max;
    db.n_v = (const int32_t*)(sd + o_nv) + q0;
    if (has_dyn) db.dyn = (const double*)((char*)h->dyn_d.p + (size_t)q0 * dyn_q_bytes);
    if (b->n_static > 0)
      db.static_obs = (const double*)h->stat_d.p + (b->static_per_query ? (size_t)q0 * b->n_static * 2 : 0);
    fot_result_t dr = *r;
    dr.best_idx = (int32_t*)(od + r_bi) + q0;
    dr.best_cost = (double*)(od + r_bc) + q0;
    dr.stats = (int32_t*)(od + r_st) + (size_t)q0 * FOT_N_STATS;
    dr.winner_len = (int32_t*)(od + r_wl) + q0;
    dr.winner = (double*)(od + r_w) + (size_t)q0 * FOT_N_SERIES * NT;
    dr.cand_cat = r->cand_cat ? (uint8_t*)(od + r_cc) + (size_t)q0 * r->cand_stride : nullptr;
    dr.cand_cost = r->cand_cost ? (double*)(od + r_cs) + (size_t)q0 * r->cand_stride : nullptr;
    // (tuning knob: blocks per CTA of the last gated range; shorter CTAs there measured no better than the
    // rule of item_geometry)
    int tail_bpc = 0;
    if (gated && bounds[c + 1] >= nq) tail_bpc = h->opt.gate_tail_bpc;
    rc = launch_all(h, &db, &dr, st, (size_t)q0, (size_t)nq, gate, tail_bpc, /*record_span=*/false);
    if (rc != FOT_OK) return rc;
    // winners of this chunk straight into the caller's arrays, on their own stream so the next
    // chunk's kernels never queue behind a copy engine that is busy with the uploads
    cudaStream_t ds = h->d2h_stream;
    if (dbg) cudaEventRecord(tl[c][2], st);
    rc = read_back(h, r, dr, (double*)(od + r_wh) + (size_t)q0 * FOT_N_SERIES * k_head, q0, cq, NT, st, ds, h->ev_done[c]);
    if (rc != FOT_OK) return rc;
    if (dbg) cudaEventRecord(tl[c][3], ds);
  }
  if (dbg) {
    cudaEventRecord(h->ev_join, h->stream2); cudaStreamWaitEvent(st, h->ev_join, 0);
    for (auto ps : h->pstream) { cudaEventRecord(h->ev_join, ps); cudaStreamWaitEvent(st, h->ev_join, 0); }
    cudaEventRecord(d3, st);
  }
  if (span_first) { rc = close_span(h, span_first); if (rc != FOT_OK) return rc; }
  CK(cudaStreamSynchronize(h->d2h_stream));   // the last read-backs wait for the last kernels of every compute stream
  CK(cudaStreamSynchronize(st));
  CK(cudaStreamSynchronize(h->stream2));
  if (gated) {
    for (auto ps : h->pstream) CK(cudaStreamSynchronize(ps));
    CK(cudaStreamSynchronize(h->copy_stream));
    CK(cudaStreamSynchronize(h->copy_stream2));
  }
  if (gated && std::chrono::steady_clock::now() - t_wall0 > std::chrono::nanoseconds(kGateTimeoutNs)) {
    unsigned gave_up = 0;
    CK(cudaMemcpy(&gave_up, gate.word + kGateSlices, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (gave_up) {
      cudaMemset(gate.word + kGateSlices, 0, sizeof(unsigned));
      return fail(FOT_ERR_CUDA, "gated sweep: the obstacle upload did not arrive");
    }
  }
  if (dbg) {
    float a = 0, bb = 0, cc = 0, dd = 0;
    cudaEventElapsedTime(&a, d0, d1); cudaEventElapsedTime(&bb, d0, d2); cudaEventElapsedTime(&cc, d0, d3); cudaEventElapsedTime(&dd, d0, d4);
    fprintf(stderr, "[fot] chunks=%d copy-stream start %.3f ms, copies done %.3f ms, first kernel may start %.3f ms, kernels done %.3f ms\n", n_chunks, dd, a, bb, cc);
    for (int c = 0; c < n_chunks; ++c) {
      float t[4] = {-1, -1, -1, -1};
      for (int k = 0; k < 4; ++k) if (k > 0 || (has_dyn && !gated)) cudaEventElapsedTime(&t[k], d0, tl[c][k]);
      fprintf(stderr, "[fot]   chunk %2d queries %5d..%5d  upload done %.3f  kernels %.3f -> %.3f  read-back done %.3f\n", c,
              bounds[c], bounds[c + 1], t[0], t[1], t[2], t[3]);
    }
    for (auto& row : tl) for (auto& ev : row) cudaEventDestroy(ev);
    cudaEventDestroy(d0); cudaEventDestroy(d1); cudaEventDestroy(d2); cudaEventDestroy(d3); cudaEventDestroy(d4);
  }
  return FOT_OK;
}

// Device inputs, HOST results: the call of a caller whose obstacle tensor was produced on the GPU (the prediction
// post-processing of fot_predict.cuh) but who consumes the winners on the host.  The queries are swept in three
// ranges on the priority streams, and each range's winners go back over PCIe while the next range is swept, so
// only the last range's read-back is exposed.  `stream` is the stream on which the inputs become ready (NULL:
// they are ready now).  Returns when the results are in `res`.
static int plan_batch_device_to_host_impl(fot_handle_t* h, const fot_batch_t* b, const fot_result_t* r, void* stream) {
  int rc = FOT_OK;
  const int nq = b->n_q, NT = h->plan.n_t_max;
  size_t ro = 0;
  auto rtake = [&](size_t bytes) { size_t o = ro; ro = align_up(ro + bytes); return o; };
  const size_t r_bi = rtake((size_t)nq * 4), r_bc = rtake((size_t)nq * 8), r_st = rtake((size_t)nq * FOT_N_STATS * 4),
               r_wl = rtake((size_t)nq * 4), r_w = rtake((size_t)nq * FOT_N_SERIES * NT * 8);
  const size_t r_cc = r->cand_cat ? rtake((size_t)nq * r->cand_stride) : 0;
  const size_t r_cs = r->cand_cost ? rtake((size_t)nq * r->cand_stride * 8) : 0;
  const int k_head = r->winner_samples > 0 ? std::min(r->winner_samples, NT) : 0;
  const size_t r_wh = k_head ? rtake((size_t)nq * FOT_N_SERIES * k_head * 8) : 0;
  CK(h->out_d.reserve(ro));
  char* od = (char*)h->out_d.p;
  h->last_winner_d = (const double*)(od + r_w);     // fot_fetch_winners
  h->last_winner_nq = nq;
  // ranges: about 1/2, 5/16, 3/16 of the queries in whole waves; one range for small batches or the fallback kernel
  int n_ranges = 1, bounds[4] = {0, nq, nq, nq};
  {
    ItemGeom ig{};
    size_t ismem = 0;
    const bool items = item_geometry(h, b, &ig, &ismem) && h->opt.sweep != 2;
    if (items && nq >= 1024) {
      const int wave = 2 * h->sms;
      auto waves = [&](int frac16) { const int w = std::max(1, (int)((long long)nq * frac16 / 16 / wave)); return w * wave; };
      const int cuts[] = {waves(8), waves(13), nq};
      n_ranges = 3;
      for (int c = 0; c < n_ranges; ++c) bounds[c + 1] = std::min(nq, std::max(cuts[c], bounds[c]));
    }
  }
  if (stream) CK(cudaEventRecord(h->ev_blob, (cudaStream_t)stream));
  const bool has_dyn = b->dyn_mode != FOT_DYN_NONE;
  const size_t dyn_q_bytes = has_dyn ? (size_t)b->S * b->P * b->T_obs * 16 : 0;
  for (int c = 0; c < n_ranges; ++c) {
    const int q0 = bounds[c], cq = bounds[c + 1] - q0;
    if (cq <= 0) continue;
    cudaStream_t st = h->pstream[c & 3];
    if (stream) CK(cudaStreamWaitEvent(st, h->ev_blob, 0));
    fot_batch_t db = *b;
    db.n_q = cq;
    db.frenet = b->frenet + (size_t)q0 * 6;
    db.target_speed = b->target_speed + q0;
    db.limits = b->limits + (size_t)q0 * 4;
    db.stop_dist = b->stop_dist + q0;
    db.v_grid = b->v_grid + (size_t)q0 * b->n_v_max;
    db.n_v = b->n_v + q0;
    if (has_dyn) db.dyn = (const double*)((const char*)b->dyn + (size_t)q0 * dyn_q_bytes);
    if (b->n_static > 0 && b->static_per_query) db.static_obs = b->static_obs + (size_t)q0 * b->n_static * 2;
    fot_result_t dr = *r;
    dr.best_idx = (int32_t*)(od + r_bi) + q0;
    dr.best_cost = (double*)(od + r_bc) + q0;
    dr.stats = (int32_t*)(od + r_st) + (size_t)q0 * FOT_N_STATS;
    dr.winner_len = (int32_t*)(od + r_wl) + q0;
    dr.winner = (double*)(od + r_w) + (size_t)q0 * FOT_N_SERIES * NT;
    dr.cand_cat = r->cand_cat ? (uint8_t*)(od + r_cc) + (size_t)q0 * r->cand_stride : nullptr;
    dr.cand_cost = r->cand_cost ? (double*)(od + r_cs) + (size_t)q0 * r->cand_stride : nullptr;
    if (c == 0) CK(cudaEventRecord(h->ev0, st));
    rc = launch_all(h, &db, &dr, st, (size_t)q0, (size_t)nq, Gate{}, 0, /*record_span=*/false);
    if (rc != FOT_OK) return rc;
    rc = read_back(h, r, dr, (double*)(od + r_wh) + (size_t)q0 * FOT_N_SERIES * k_head, q0, cq, NT, st, h->d2h_stream, h->ev_done[c]);
    if (rc != FOT_OK) return rc;
  }
  rc = close_span(h, h->pstream[0]);
  if (rc != FOT_OK) return rc;
  CK(cudaStreamSynchronize(h->d2h_stream));
  for (int c = 0; c < n_ranges; ++c) CK(cudaStreamSynchronize(h->pstream[c & 3]));
  return FOT_OK;
}

// ---- the gather of a sharded sweep over peer memory ---------------------------------------------------------------
extern "C" int fot_set_result_mirror(fot_handle_t* h, const fot_result_t* mirror, void* flag) {
  if (!h) return fail(FOT_ERR_ARG, "fot_set_result_mirror: null handle");
  if (!mirror) { h->mirror = fot_result_t{}; h->mirror_flag = nullptr; h->mirror_seq = 0; return FOT_OK; }
  if (!mirror->best_idx || !mirror->best_cost || !mirror->stats || !mirror->winner_len || !mirror->winner)
    return fail(FOT_ERR_ARG, "fot_set_result_mirror: null mirror array");
  h->mirror = *mirror;
  if ((unsigned*)flag != h->mirror_flag) h->mirror_seq = 0;     // a new flag word starts a new sequence: 1, 2, 3, ...
  h->mirror_flag = (unsigned*)flag;
  return FOT_OK;
}

extern "C" int fot_peer_alloc(int device, size_t bytes, void** ptr, unsigned char handle[64]) {
  if (!ptr || !handle || bytes == 0) return fail(FOT_ERR_ARG, "fot_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaSetDevice(device));
  CK(cudaMalloc(ptr, bytes));
  CK(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, *ptr);
  if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return fail(FOT_ERR_CUDA, "cudaIpcGetMemHandle", e); }
  memcpy(handle, &hd, 64);
  return FOT_OK;
}

extern "C" int fot_peer_open(int device, const unsigned char handle[64], void** ptr) {
  if (!ptr || !handle) return fail(FOT_ERR_ARG, "fot_peer_open: bad argument");
  CK(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, 64);
  CK(cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return FOT_OK;
}

extern "C" int fot_peer_close(void* ptr) {
  if (ptr) CK(cudaIpcCloseMemHandle(ptr));
  return FOT_OK;
}

extern "C" int fot_peer_free(void* ptr) {
  if (ptr) CK(cudaFree(ptr));
  return FOT_OK;
}

extern "C" int fot_peer_await(int device, void* stream, const void* flags, int world, unsigned seq, void* err_word) {
  if (!flags || !err_word || world < 1 || world > 1024) return fail(FOT_ERR_ARG, "fot_peer_await: bad argument");
  CK(cudaSetDevice(device));
  fot_await_kernel<<<1, (world + 31) / 32 * 32, 0, (cudaStream_t)stream>>>((const unsigned*)flags, world, seq, 2000000000ll, (unsigned*)err_word);
  CK(cudaGetLastError());
  return FOT_OK;
}

extern "C" int fot_fetch_winners(fot_handle_t* h, int q0, int n, double* out) {
  if (!h || !out || q0 < 0 || n < 1) return fail(FOT_ERR_ARG, "fot_fetch_winners: bad argument");
  if (!h->last_winner_d || q0 + n > h->last_winner_nq) return fail(FOT_ERR_ARG, "fot_fetch_winners: no such queries in the last host-result call");
  CK(cudaSetDevice(h->device));
  const size_t row = (size_t)FOT_N_SERIES * h->plan.n_t_max;
  CK(cudaMemcpyAsync(out, h->last_winner_d + (size_t)q0 * row, (size_t)n * row * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
  CK(cudaStreamSynchronize(h->d2h_stream));
  return FOT_OK;
}

extern "C" float fot_last_kernel_ms(const fot_handle_t* h) {
  if (!h || !h->timed) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

extern "C" int fot_last_sweep_kind(const fot_handle_t* h) { return h ? h->last_sweep_kind : 0; }
extern "C" int fot_last_pair_features(const fot_handle_t* h) { return h && h->last_sweep_kind == 4 ? h->last_pair_feat : -1; }

extern "C" int fot_launch_stage_ms(const fot_handle_t* h, int back, float ms[3]) {
  if (!h || !ms || back < 0 || back >= fot_handle::kRing || back >= h->n_launch)
    return fail(FOT_ERR_ARG, "fot_launch_stage_ms: no such launch");
  const cudaEvent_t* ring = h->ring.data() + (size_t)((h->n_launch - 1 - back) % fot_handle::kRing) * 4;
  CK(cudaEventSynchronize(ring[3]));
  for (int i = 0; i < 3; ++i) CK(cudaEventElapsedTime(&ms[i], ring[i], ring[i + 1]));
  return FOT_OK;
}

extern "C" int fot_probe_fma_tflops(int device, int kind, double* tflops_out) {
  if (!tflops_out || kind < 0 || kind > 2) return fail(FOT_ERR_ARG, "fot_probe_fma_tflops: bad argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(FOT_ERR_NO_DEVICE, "no CUDA device");
  CK(cudaSetDevice(device));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  float* sink = nullptr;
  CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256, iters = kind == 0 ? 4096 : 8192;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    if (kind == 0) fot_probe_kernel<0><<<blocks, threads>>>(sink, iters);
    else if (kind == 1) fot_probe_kernel<1><<<blocks, threads>>>(sink, iters);
    else fot_probe_kernel<2><<<blocks, threads>>>(sink, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fmas = (double)blocks * threads * iters * 64.0 * (kind == 2 ? 2.0 : 1.0);
    const double tf = 2.0 * fmas / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  CK(cudaGetLastError());
  *tflops_out = best;
  return FOT_OK;
}

// ---- prediction post-processing ------------------------------------------
static int pred_begin(int device, const char* what) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(FOT_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= n_dev) return fail(FOT_ERR_ARG, what);
  CK(cudaSetDevice(device));
  return FOT_OK;
}
static int pred_end(void* stream) {
  CK(cudaGetLastError());
  if (!stream) CK(cudaStreamSynchronize(nullptr));
  return FOT_OK;
}

extern "C" int fot_predict_cv_device(int device, void* stream, int n_q, int P, const double* p_curr, const double* p_prev,
                                     const double* staleness, double sgan_dt, const double* time_target, int n_steps,
                                     const double* cur_pos, int obs_float32, double* out) {
  if (n_q < 1 || P < 1 || n_steps < 1 || !p_curr || !time_target || !out || !(sgan_dt > 0.0))
    return fail(FOT_ERR_ARG, "fot_predict_cv_device: bad argument");
  int rc = pred_begin(device, "fot_predict_cv_device: bad device ordinal");
  if (rc != FOT_OK) return rc;
  const int T_out = n_steps + (cur_pos ? 1 : 0);
  const size_t smem = (size_t)P * 2 * sizeof(double);
  if (smem > 48 * 1024) return fail(FOT_ERR_TOO_LARGE, "fot_predict_cv_device: too many pedestrians per query");
  fot_cv_kernel<<<n_q, 256, smem, (cudaStream_t)stream>>>(p_curr, p_prev, staleness, time_target, cur_pos, out, P, n_steps,
                                                          T_out, sgan_dt, obs_float32);
  return pred_end(stream);
}

extern "C" int fot_process_prediction_device(int device, void* stream, int n_q, int S, int P, int pred_len, const double* pred,
                                             const double* anchor, const double* staleness, double sgan_dt,
                                             const double* time_target, int n_steps, double* out) {
  if (n_q < 1 || S < 1 || P < 1 || pred_len < 1 || pred_len > kPredLenMax || n_steps < 1 || !pred || !time_target || !out ||
      !(sgan_dt > 0.0))
    return fail(FOT_ERR_ARG, "fot_process_prediction_device: bad argument (pred_len <= 64)");
  int rc = pred_begin(device, "fot_process_prediction_device: bad device ordinal");
  if (rc != FOT_OK) return rc;
  const int threads = std::min(256, std::max(32, (2 * P + 31) / 32 * 32));
  fot_resample_kernel<<<(unsigned)((long long)n_q * S), threads, 0, (cudaStream_t)stream>>>(pred, anchor, staleness, time_target, out, S, P,
                                                                                            pred_len, n_steps, sgan_dt);
  return pred_end(stream);
}

extern "C" int fot_select_best_sample_device(int device, void* stream, int n_q, int S, int P, int T, const double* samples,
                                             double* dist_scratch, int32_t* best_idx) {
  if (n_q < 1 || S < 1 || P < 1 || T < 1 || !samples || !dist_scratch || !best_idx)
    return fail(FOT_ERR_ARG, "fot_select_best_sample_device: bad argument");
  int rc = pred_begin(device, "fot_select_best_sample_device: bad device ordinal");
  if (rc != FOT_OK) return rc;
  const long long n = (long long)n_q * S;
  fot_best_dist_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(samples, dist_scratch, n_q, S, P * T);
  fot_best_pick_kernel<<<(n_q + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dist_scratch, best_idx, n_q, S);
  return pred_end(stream);
}

extern "C" int fot_prepend_current_device(int device, void* stream, int n_q, int S, int P, int T, const double* in,
                                          const int32_t* pick, const double* cur_pos, int conditional, double* out) {
  if (n_q < 1 || S < 1 || P < 1 || T < 1 || !in || !cur_pos || !out)
    return fail(FOT_ERR_ARG, "fot_prepend_current_device: bad argument");
  int rc = pred_begin(device, "fot_prepend_current_device: bad device ordinal");
  if (rc != FOT_OK) return rc;
  const long long blocks = (long long)n_q * (pick ? 1 : S);
  fot_prepend_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, pick, cur_pos, out, S, P, T, conditional);
  return pred_end(stream);
}

extern "C" int fot_safety_metrics_device(int device, void* stream, int n_q, int P, const double* ego, const double* ped_pos,
                                         const double* ped_vel, const int32_t* n_peds, double combined_radius,
                                         const double* offsets, int n_circ, double* out) {
  if (n_q < 1 || P < 0 || !ego || !out || (P > 0 && (!ped_pos || !ped_vel)) || n_circ < 0 || n_circ > FOT_MAX_CIRCLES ||
      (n_circ > 0 && !offsets))
    return fail(FOT_ERR_ARG, "fot_safety_metrics_device: bad argument");
  int rc = pred_begin(device, "fot_safety_metrics_device: bad device ordinal");
  if (rc != FOT_OK) return rc;
  CircleOffsets off{};
  for (int i = 0; i < n_circ; ++i) off.v[i] = offsets[i];
  const long long threads = (long long)n_q * 32;
  fot_safety_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ego, ped_pos, ped_vel, n_peds, out, n_q, P,
                                                                                         combined_radius, off, n_circ);
  return pred_end(stream);
}

#ifdef FOT_PHASE_CLOCKS
// tuning aid: read and reset the per-phase clock accumulators of fot_sweep_items
extern "C" int fot_debug_phase_clocks(unsigned long long out[16]) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out, fot::g_phase_clk, sizeof(unsigned long long) * 16));
  unsigned long long z[16] = {};
  CK(cudaMemcpyToSymbol(fot::g_phase_clk, z, sizeof z));
  return FOT_OK;
}
#endif
