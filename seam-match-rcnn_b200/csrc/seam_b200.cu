// C ABI of the B200-native SEAM Match-RCNN retrieval hot path (see include/seam_b200.h).
// Host-side argument checking, workspace carving, tensor-map encoding and kernel launches.
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/seam_b200.h"
#include "aggregate_fused.cuh"
#include "backward.cuh"
#include "fold.cuh"
#include "nlb_gemm.cuh"
#include "score_exact.cuh"
#include "score_tc.cuh"
#include "tower_tc.cuh"

using namespace seam;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct seam_handle {
  int device = 0;
  int num_sms = 0;
  float* fold = nullptr;            // folded weights (device)
  bool have_scorer = false;
  bool have_aggregator = false;
  PFN_encodeTiled encode = nullptr;
  uint64_t launches = 0;
  bool profiling = false;
  unsigned int* watchdog = nullptr;   // host-mapped record buffer of the device-side wait watchdogs
  // conv tower (f3): reorganised fp16 conv weights, biases, transposed linear weight, folded BatchNorm
  __half* tw_conv[4] = {nullptr, nullptr, nullptr, nullptr};
  float* tw_misc = nullptr;           // bias0..3 (256,256,256,1024) | lin_wt (1024*256) | lin_b | bn_scale | bn_shift
  bool have_tower = false;
  // developer overrides, read once at seam_create (never consulted on the hot path)
  int dbg_score_grid = 0, dbg_agg_grid = 0, dbg_cta_ns = 0, dbg_no_tm = 0, score_nseed = 4;
  struct Span { int kernel; cudaEvent_t a, b; };
  std::vector<Span> spans;            // recorded while profiling
  std::vector<cudaEvent_t> free_events;
  char err[512] = {0};
};

// brackets one kernel launch with CUDA events on the launching stream when profiling is on
struct ProfileScope {
  seam_handle* h;
  cudaStream_t stream;
  cudaEvent_t a = nullptr, b = nullptr;
  int kernel;
  static cudaEvent_t get(seam_handle* h) {
    cudaEvent_t e = nullptr;
    if (!h->free_events.empty()) { e = h->free_events.back(); h->free_events.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  }
  ProfileScope(seam_handle* h_, int kernel_, cudaStream_t s) : h(h_), stream(s), kernel(kernel_) {
    if (h->profiling && h->spans.size() < 65536) {
      a = get(h); b = get(h);
      cudaEventRecord(a, stream);
    }
  }
  ~ProfileScope() {
    if (a) {
      cudaEventRecord(b, stream);
      h->spans.push_back({kernel, a, b});
    }
  }
};

static_assert(exact::PLAN_MAX_GRID == score::MAX_GRID, "the re-score kernels carry a copy of the scorer's CTA ranges");
static thread_local char g_create_err[256] = "";

static unsigned int* g_watchdog_host[64] = {};

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static int fail(seam_handle* h, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  if (h) vsnprintf(h->err, sizeof(h->err), fmt, ap);
  else vsnprintf(g_create_err, sizeof(g_create_err), fmt, ap);
  va_end(ap);
  return code;
}

#define SEAM_CUDA(h, expr)                                                                             \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return fail(h, SEAM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                           \
  } while (0)

#define SEAM_LAUNCHED(h, name)                                                                      \
  do {                                                                                              \
    cudaError_t e__ = cudaGetLastError();                                                           \
    if (e__ != cudaSuccess)                                                                         \
      return fail(h, SEAM_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));      \
    ++(h)->launches;                                                                                \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Full tracks / full frame blocks are fetched with one 3-D tensor-map copy: box {256, 1 track, frames} over seq viewed as
// (1+Tmax, Q, 256) with the caller's strides.  Returns whether the map is usable.
static int encode_seq_map(seam_handle* h, const aggf::Params& p, int box_frames, CUtensorMap* tm) {
  memset(tm, 0, sizeof(*tm));
  if (h->dbg_no_tm) return 0;
  if (!p.seq || p.Q <= 0 || (p.track_stride * 4) % 16 != 0 || (p.frame_stride * 4) % 16 != 0 || p.track_stride < 256 ||
      p.frame_stride < 256 || box_frames > p.Tmax)
    return 0;
  const cuuint64_t dims[3] = {256, (cuuint64_t)p.Q, (cuuint64_t)(1 + p.Tmax)};
  const cuuint64_t strides[2] = {(cuuint64_t)p.track_stride * 4, (cuuint64_t)p.frame_stride * 4};
  const cuuint32_t box[3] = {256, 1, (cuuint32_t)box_frames};
  const cuuint32_t estr[3] = {1, 1, 1};
  return h->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.seq), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int TR>
static void launch_aggregate_warp(seam_handle* h, aggf::Params& p, int num_sms, cudaStream_t stream) {   // num_sms = CTA cap
  constexpr int NW = aggf::Cfg<TR>::NW;
  const int want = (p.Q + NW - 1) / NW;
  const int grid = want < num_sms ? want : num_sms;
  CUtensorMap tm;
  p.use_tm = p.Tmax == TR ? encode_seq_map(h, p, TR, &tm) : (memset(&tm, 0, sizeof(tm)), 0);
  aggf::aggregate_fused_warp_kernel<TR><<<grid, aggf::warp_threads<TR>(), aggf::warp_smem_bytes<TR>(), stream>>>(tm, p);
}
template <int GW>
static void launch_aggregate_group(seam_handle* h, aggf::Params& p, int num_sms, cudaStream_t stream) {
  constexpr int GPC = aggf::GWARPS / GW;
  const int want = (p.Q + GPC - 1) / GPC;
  const int grid = want < num_sms ? want : num_sms;
  CUtensorMap tm;
  p.use_tm = encode_seq_map(h, p, aggf::FB, &tm);
  aggf::aggregate_fused_group_kernel<GW><<<grid, aggf::GTHREADS, aggf::group_smem_bytes<GW>(), stream>>>(tm, p);
}

static int import_exchange(seam_handle* h, const seam_exchange* x, xchg::Exchange* e, const char* who) {
  if (!x) return fail(h, SEAM_ERR_BAD_ARG, "%s: exchange is null", who);
  if (x->world < 1 || x->world > SEAM_MAX_WORLD || x->rank < 0 || x->rank >= x->world)
    return fail(h, SEAM_ERR_BAD_ARG, "%s: world=%d rank=%d outside [1,%d]", who, x->world, x->rank, SEAM_MAX_WORLD);
  if (x->Q < 0 || x->k < 1 || x->k > SEAM_MAX_K || x->own_max < 0)
    return fail(h, SEAM_ERR_BAD_ARG, "%s: bad Q / k / own_max", who);
  if (x->q_lo[0] != 0 || x->q_lo[x->world] != x->Q) return fail(h, SEAM_ERR_BAD_ARG, "%s: q_lo must run from 0 to Q", who);
  if (!x->step || !x->done) return fail(h, SEAM_ERR_BAD_ARG, "%s: step / done are null", who);
  memset(e, 0, sizeof(*e));
  e->world = x->world;
  e->rank = x->rank;
  e->Q = x->Q;
  e->k = x->k;
  e->own_max = x->own_max;
  const bool have_final = x->final_score[0] != nullptr;
  for (int r = 0; r <= x->world; ++r) e->q_lo[r] = x->q_lo[r];
  for (int r = 0; r < x->world; ++r) {
    if (x->q_lo[r + 1] < x->q_lo[r] || x->q_lo[r + 1] - x->q_lo[r] > x->own_max)
      return fail(h, SEAM_ERR_BAD_ARG, "%s: q_lo not monotone or ownership above own_max", who);
    if (!x->q_all[r] || !x->list_margin[r] || !x->list_idx[r] || !x->flags[r])
      return fail(h, SEAM_ERR_BAD_ARG, "%s: null buffer for rank %d", who, r);
    if (have_final && (!x->final_score[r] || !x->final_margin[r] || !x->final_idx[r]))
      return fail(h, SEAM_ERR_BAD_ARG, "%s: final buffers must be given for all ranks or none", who);
    e->q_all[r] = x->q_all[r];
    e->list_margin[r] = x->list_margin[r];
    e->list_idx[r] = x->list_idx[r];
    e->final_score[r] = have_final ? x->final_score[r] : nullptr;
    e->final_margin[r] = have_final ? x->final_margin[r] : nullptr;
    e->final_idx[r] = have_final ? x->final_idx[r] : nullptr;
    e->flags[r] = x->flags[r];
  }
  e->step = x->step;
  e->done = x->done;
  e->q_all_mc = x->world > 1 ? x->q_all_mc : nullptr;
  const bool final_mc = x->world > 1 && have_final && x->final_score_mc && x->final_margin_mc && x->final_idx_mc;
  e->final_score_mc = final_mc ? x->final_score_mc : nullptr;
  e->final_margin_mc = final_mc ? x->final_margin_mc : nullptr;
  e->final_idx_mc = final_mc ? x->final_idx_mc : nullptr;
  return SEAM_OK;
}

__global__ void device_stamp_kernel(uint64_t* dst) { *dst = ptx::globaltimer_ns(); }

extern "C" {

int seam_abi_version(void) { return 3; }

size_t seam_exchange_sizeof(void) { return sizeof(seam_exchange); }

int seam_create(seam_handle** out, int device) {
  if (!out) return fail(nullptr, SEAM_ERR_BAD_ARG, "seam_create: out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, SEAM_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev) return fail(nullptr, SEAM_ERR_BAD_ARG, "device %d out of range", device);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, SEAM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, SEAM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  seam_handle* h = new (std::nothrow) seam_handle();
  if (!h) return fail(nullptr, SEAM_ERR_CUDA, "out of host memory");
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  DeviceGuard guard(device);
  if ((e = cudaMalloc(&h->fold, sizeof(float) * Fold::TOTAL)) != cudaSuccess) {
    delete h;
    return fail(nullptr, SEAM_ERR_CUDA, "cudaMalloc(fold): %s", cudaGetErrorString(e));
  }
  cudaMemset(h->fold, 0, sizeof(float) * Fold::TOTAL);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    cudaFree(h->fold);
    delete h;
    return fail(nullptr, SEAM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  }
  h->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  h->dbg_score_grid = env_int("SEAM_DEBUG_SCORE_GRID", 0);
  h->dbg_agg_grid = env_int("SEAM_DEBUG_AGG_GRID", 0);
  h->dbg_cta_ns = env_int("SEAM_DEBUG_CTA_NS", 0);
  h->dbg_no_tm = env_int("SEAM_DEBUG_AGG_NO_TM", 0);      // developer A/B: per-frame bulk copies instead of one tensor-map box
  h->score_nseed = env_int("SEAM_SCORE_NSEED", 4);
  // watchdog records live in host-mapped memory so that they survive a trapped launch: one buffer per
  // device for the life of the process (the device-side pointer must never dangle)
  if (device < 64) {
    if (!g_watchdog_host[device]) {
      unsigned int* hp = nullptr;
      if (cudaHostAlloc(reinterpret_cast<void**>(&hp), 1024, cudaHostAllocMapped) == cudaSuccess) {
        memset(hp, 0, 1024);
        unsigned int* dptr = nullptr;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), hp, 0) == cudaSuccess &&
            cudaMemcpyToSymbol(ptx::g_watchdog, &dptr, sizeof(dptr)) == cudaSuccess)
          g_watchdog_host[device] = hp;
      }
      cudaGetLastError();
    }
    h->watchdog = g_watchdog_host[device];
  }
  // opt in to large dynamic shared memory once
  cudaFuncSetAttribute(score::score_topk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)score::SMEM_BYTES);
#ifdef SEAM_DIAGNOSTIC_VARIANTS   // partial epilogues for measurements (wrong results): never in a product build
  cudaFuncSetAttribute(score::score_topk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)score::SMEM_BYTES);
  cudaFuncSetAttribute(score::score_topk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)score::SMEM_BYTES);
  cudaFuncSetAttribute(score::score_topk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)score::SMEM_BYTES);
  cudaFuncSetAttribute(score::score_topk_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)score::SMEM_BYTES);
#endif
  cudaFuncSetAttribute(score::score_topk_kernel<score::VAR_RANK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)score::SMEM_BYTES);
  cudaFuncSetAttribute(aggf::aggregate_fused_warp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)aggf::warp_smem_bytes<4>());
  cudaFuncSetAttribute(aggf::aggregate_fused_warp_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)aggf::warp_smem_bytes<10>());
  cudaFuncSetAttribute(aggf::aggregate_fused_warp_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)aggf::warp_smem_bytes<16>());
  cudaFuncSetAttribute(aggf::aggregate_fused_group_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)aggf::group_smem_bytes<2>());
  cudaFuncSetAttribute(aggf::aggregate_fused_group_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)aggf::group_smem_bytes<4>());
  cudaFuncSetAttribute(bwd::agg_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)bwd::agg_bwd_smem_bytes(bwd::BWD_MAX_T));
  cudaFuncSetAttribute(tower::conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tower::SMEM_BYTES);
  cudaFuncSetAttribute(nlbgemm::nlb_full_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((SEAM_MAX_T * 257 + 2 * SEAM_MAX_T) * sizeof(float)));
  if ((e = cudaGetLastError()) != cudaSuccess) {
    cudaFree(h->fold);
    delete h;
    return fail(nullptr, SEAM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  *out = h;
  return SEAM_OK;
}

void seam_destroy(seam_handle* h) {
  if (!h) return;
  {
    DeviceGuard guard(h->device);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->free_events) cudaEventDestroy(e);
    cudaFree(h->fold);
    for (int i = 0; i < 4; ++i) cudaFree(h->tw_conv[i]);
    cudaFree(h->tw_misc);
  }
  delete h;
}

int seam_watchdog_read(const seam_handle* h, uint32_t* out, int max_records) {
  if (!h || !h->watchdog || !out || max_records <= 0) return 0;
  const volatile unsigned int* w = h->watchdog;
  int n = (int)w[0];
  if (n > ptx::WATCHDOG_RECORDS) n = ptx::WATCHDOG_RECORDS;
  if (n > max_records) n = max_records;
  for (int i = 0; i < n * 8; ++i) out[i] = w[8 + i];
  return n;
}

int seam_device_stamp(seam_handle* h, uint64_t* dst, void* stream) {
  if (!h || !dst) return SEAM_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  device_stamp_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(dst);
  return SEAM_OK;
}

const char* seam_last_error(const seam_handle* h) { return h ? h->err : g_create_err; }

uint64_t seam_launch_count(const seam_handle* h) { return h ? h->launches : 0; }

int seam_profile_enable(seam_handle* h, int enable) {
  if (!h) return SEAM_ERR_BAD_ARG;
  h->profiling = enable != 0;
  return SEAM_OK;
}

int seam_profile_read(seam_handle* h, int kernel, double* total_ms, int* launches) {
  if (!h || !total_ms || !launches) return SEAM_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  double tot = 0.0;
  int n = 0;
  std::vector<seam_handle::Span> keep;
  for (auto& sp : h->spans) {
    if (sp.kernel != kernel) { keep.push_back(sp); continue; }
    SEAM_CUDA(h, cudaEventSynchronize(sp.b));
    float ms = 0.f;
    SEAM_CUDA(h, cudaEventElapsedTime(&ms, sp.a, sp.b));
    tot += ms;
    ++n;
    h->free_events.push_back(sp.a);
    h->free_events.push_back(sp.b);
  }
  h->spans.swap(keep);
  *total_ms = tot;
  *launches = n;
  return SEAM_OK;
}

int seam_load_weights(seam_handle* h, const seam_weights* w, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!w) return fail(h, SEAM_ERR_BAD_ARG, "seam_load_weights: weights struct is null");
  const float* const* ptrs = reinterpret_cast<const float* const*>(w);
  for (int i = 0; i < 13; ++i)
    if (!ptrs[i]) return fail(h, SEAM_ERR_BAD_ARG, "seam_load_weights: weight pointer %d is null", i);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  FoldIn in{w->theta_w, w->theta_b, w->phi_w, w->phi_b, w->g_w,    w->g_b,   w->W_w,
            w->W_b,     w->concat_w, w->att_w, w->att_b, w->last_w, w->last_b};
  fold_vectors_kernel<<<1, 256, 0, stream>>>(in, h->fold);
  SEAM_LAUNCHED(h, "fold_vectors_kernel");
  fold_matrix_kernel<<<256, 256, 0, stream>>>(w->W_w, w->g_w, h->fold);
  SEAM_LAUNCHED(h, "fold_matrix_kernel");
  fold_m16_kernel<<<256, 256, 0, stream>>>(h->fold);
  SEAM_LAUNCHED(h, "fold_m16_kernel");
  h->have_scorer = h->have_aggregator = true;
  return SEAM_OK;
}

int seam_load_scorer(seam_handle* h, const float* last_w, const float* last_b, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!last_w || !last_b) return fail(h, SEAM_ERR_BAD_ARG, "seam_load_scorer: null pointer");
  DeviceGuard guard(h->device);
  fold_scorer_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(last_w, last_b, h->fold);
  SEAM_LAUNCHED(h, "fold_scorer_kernel");
  h->have_scorer = true;
  return SEAM_OK;
}

// ------------------------------------------------------------------------------ aggregation
size_t seam_aggregate_workspace_bytes(int Q) {
  (void)Q;
  return 256;   // the fused kernel keeps every intermediate on the SM
}

static int aggregate_impl(seam_handle* h, const xchg::Exchange* x, int row0, int last, const float* seq,
                          const uint8_t* mask, const int32_t* lens, int Tmax, int Q, int64_t frame_stride,
                          int64_t track_stride, float* out, float* att, cudaStream_t stream, const char* who) {
  if (!h->have_aggregator) return fail(h, SEAM_ERR_STATE, "%s: weights not loaded", who);
  if (Q < 0 || Tmax < 0) return fail(h, SEAM_ERR_BAD_ARG, "%s: negative size", who);
  if (!x && Q == 0) return SEAM_OK;
  if (!x && !out) return fail(h, SEAM_ERR_BAD_ARG, "%s: out is null", who);
  if (Tmax > SEAM_MAX_T) return fail(h, SEAM_ERR_UNSUPPORTED, "%s: Tmax=%d exceeds %d", who, Tmax, SEAM_MAX_T);
  DeviceGuard guard(h->device);
  if (!x && Tmax == 0) {   // only the dummy row: every track is empty -> zero descriptors
    SEAM_CUDA(h, cudaMemsetAsync(out, 0, (size_t)Q * 256 * 4, stream));
    return SEAM_OK;
  }
  if (Q > 0 && Tmax > 0 && !seq) return fail(h, SEAM_ERR_BAD_ARG, "%s: null pointer", who);
  if (!aligned16(seq) || (out && !aligned16(out)) || (frame_stride & 3) || (track_stride & 3))
    return fail(h, SEAM_ERR_UNSUPPORTED, "%s: seq/out must be 16-byte aligned, strides multiples of 4", who);
  ProfileScope prof(h, SEAM_KERNEL_AGGREGATE, stream);
  aggf::Params p;
  memset(&p, 0, sizeof(p));
  p.seq = seq;
  p.mask = mask;
  p.lens = lens;
  p.Tmax = Tmax;
  p.Q = Q;
  p.frame_stride = frame_stride;
  p.track_stride = track_stride;
  p.fold = h->fold;
  p.out = out;
  p.att = att;
  if (x) {
    p.x_on = 1;
    p.x_last = last;
    p.x_row0 = row0;
    p.x = *x;
    // a rank without tracks in this call (or with empty tracks only: Tmax == 0) still has to take part in the
    // protocol: one CTA, zero tracks / zero-length tracks, descriptors = 0
    if (Tmax == 0) { p.Tmax = 1; p.lens = nullptr; p.mask = nullptr; p.seq = nullptr; p.Q = -Q; }
  }
  const int cap = h->dbg_agg_grid > 0 && h->dbg_agg_grid < h->num_sms ? h->dbg_agg_grid : h->num_sms;
  if (p.Q < 0) return fail(h, SEAM_ERR_UNSUPPORTED, "%s: Tmax == 0 in the sharded search", who);
  if (p.Q == 0) {
    // nothing to aggregate on this rank: only the signal is needed
    if (x && last) {
      aggf::signal_only_kernel<<<1, 32, 0, stream>>>(p.x);
      SEAM_LAUNCHED(h, "signal_only_kernel");
    }
    return SEAM_OK;
  }
  if (Tmax <= 4) launch_aggregate_warp<4>(h, p, cap, stream);
  else if (Tmax <= 10) launch_aggregate_warp<10>(h, p, cap, stream);
  else if (Tmax <= 16) launch_aggregate_warp<16>(h, p, cap, stream);
  else if (Tmax <= 32) launch_aggregate_group<2>(h, p, cap, stream);   // two warps per track
  else launch_aggregate_group<4>(h, p, cap, stream);                   // 33..64 frames: four warps per track
  SEAM_LAUNCHED(h, "aggregate kernel");
  return SEAM_OK;
}

int seam_aggregate(seam_handle* h, const float* seq, const uint8_t* mask, const int32_t* lens, int Tmax, int Q,
                   int64_t frame_stride, int64_t track_stride, float* out, float* att, void* workspace,
                   size_t workspace_bytes, void* stream_) {
  (void)workspace;
  (void)workspace_bytes;
  if (!h) return SEAM_ERR_BAD_ARG;
  return aggregate_impl(h, nullptr, 0, 0, seq, mask, lens, Tmax, Q, frame_stride, track_stride, out, att,
                        static_cast<cudaStream_t>(stream_), "seam_aggregate");
}

int seam_sharded_aggregate(seam_handle* h, const seam_exchange* x, const float* seq, const uint8_t* mask,
                           const int32_t* lens, int Tmax, int Qlocal, int64_t frame_stride, int64_t track_stride,
                           int row0, int last, float* att, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  xchg::Exchange e;
  int rc = import_exchange(h, x, &e, "seam_sharded_aggregate");
  if (rc != SEAM_OK) return rc;
  if (row0 < 0 || Qlocal < 0 || row0 + Qlocal > x->Q)
    return fail(h, SEAM_ERR_BAD_ARG, "seam_sharded_aggregate: rows [%d,%d) outside the %d queries", row0, row0 + Qlocal, x->Q);
  return aggregate_impl(h, &e, row0, last ? 1 : 0, seq, mask, lens, Tmax, Qlocal, frame_stride, track_stride, nullptr, att,
                        static_cast<cudaStream_t>(stream_), "seam_sharded_aggregate");
}

size_t seam_nlb_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 256;
  const size_t rows = (size_t)B * T;
  return 2 * align_up(rows * 256 * 4, 256) + align_up(rows * 2 * 4, 256);
}

int seam_nlb_forward(seam_handle* h, const float* x, int B, int T, float* z, void* workspace, size_t workspace_bytes,
                     void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_aggregator) return fail(h, SEAM_ERR_STATE, "seam_nlb_forward: weights not loaded");
  if (B < 0 || T < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_nlb_forward: negative size");
  if (B == 0 || T == 0) return SEAM_OK;
  if (T > SEAM_MAX_T) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_nlb_forward: T=%d exceeds %d", T, SEAM_MAX_T);
  if (!x || !z || !workspace) return fail(h, SEAM_ERR_BAD_ARG, "seam_nlb_forward: null pointer");
  if (workspace_bytes < seam_nlb_workspace_bytes(B, T))
    return fail(h, SEAM_ERR_STATE, "seam_nlb_forward: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  const size_t rows = (size_t)B * T;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* Xt = reinterpret_cast<float*>(ws);
  float* R = reinterpret_cast<float*>(ws + align_up(rows * 256 * 4, 256));
  float* sv = reinterpret_cast<float*>(ws + 2 * align_up(rows * 256 * 4, 256));
  const size_t smem = (size_t)(T * 257 + 2 * T) * sizeof(float);
  nlbgemm::nlb_full_front_kernel<<<B, 256, smem, stream>>>(x, T, h->fold, Xt, R, sv);
  SEAM_LAUNCHED(h, "nlb_full_front_kernel");
  nlbgemm::Params gp;
  gp.pooled = Xt;
  gp.R = R;
  gp.sv = sv;
  gp.fold = h->fold;
  gp.out = z;
  gp.rows = (int)rows;
  gp.T = T;
  dim3 ggrid((unsigned)((rows + nlbgemm::TM - 1) / nlbgemm::TM), 256 / nlbgemm::TN);
  {
    ProfileScope prof(h, SEAM_KERNEL_NLB_GEMM, stream);
    nlbgemm::nlb_gemm_simt_kernel<<<ggrid, 256, 0, stream>>>(gp);
    SEAM_LAUNCHED(h, "nlb_gemm_simt_kernel");
  }
  return SEAM_OK;
}

// ------------------------------------------------------------------------------ scorer
int seam_prepare_gallery(seam_handle* h, const float* g, int G, void* g16, float* cg, float* gstat, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_scorer) return fail(h, SEAM_ERR_STATE, "seam_prepare_gallery: scorer weights not loaded");
  if (G < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_prepare_gallery: negative size");
  if (!gstat) return fail(h, SEAM_ERR_BAD_ARG, "seam_prepare_gallery: gstat is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  SEAM_CUDA(h, cudaMemsetAsync(gstat, 0, 16, stream));
  if (G == 0) return SEAM_OK;
  if (!g || !g16 || !cg) return fail(h, SEAM_ERR_BAD_ARG, "seam_prepare_gallery: null pointer");
  if (!aligned16(g) || !aligned16(g16)) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_prepare_gallery: 16-byte alignment");
  {
    ProfileScope prof(h, SEAM_KERNEL_PREP_GALLERY, stream);
    exact::prep_gallery_kernel<<<(G + 7) / 8, 256, 0, stream>>>(g, G, h->fold, static_cast<__half*>(g16), cg, gstat);
    SEAM_LAUNCHED(h, "prep_gallery_kernel");
  }
  return SEAM_OK;
}

struct ScorePlan {
  int num_mtiles, ntiles_n, grid, P, CAP, nseed;
  int tb[score::MAX_GRID + 1];
  long long total_tiles;
  size_t off_a16, off_rq, off_anorm, off_thr, off_rowcnt, off_gmax, off_rowflag, off_rowbuf, off_cnt, off_rows, off_xpart, off_xdone, total;
};

// The (query tile, gallery tile) grid is linearised query-major and cut into one contiguous
// range per CTA (score_tc.cuh).  A query tile's gallery sweep is therefore shared by at most P
// CTAs ("pieces"); each piece owns 4 candidate sub-lists per row (one per epilogue thread of
// the row) of CAP entries, CAP >= 3 times the expected number of appends.
static ScorePlan plan_score(int num_sms, int Q, int G, int nseed_override = -1, int nseed_default = 4, int cap_grid = 0) {
  ScorePlan s;
  s.num_mtiles = (Q + score::BM - 1) / score::BM;
  s.ntiles_n = (G + score::BN - 1) / score::BN;
  if (s.num_mtiles < 1) s.num_mtiles = 1;
  if (s.ntiles_n < 1) s.ntiles_n = 1;
  s.total_tiles = (long long)s.num_mtiles * s.ntiles_n;
  s.grid = s.total_tiles < num_sms ? (int)s.total_tiles : num_sms;
  if (cap_grid > 0 && cap_grid < s.grid) s.grid = cap_grid;   // developer diagnostics only (seam_create reads the override)
  if (s.grid > score::MAX_GRID) s.grid = score::MAX_GRID;
  s.nseed = nseed_override >= 0 ? nseed_override : nseed_default;
  // Cost-balanced contiguous ranges.  A tile costs 1; a segment start costs SEG_COST (query tile load,
  // epilogue hand-over); a segment that holds its row's first gallery tile -- or is all a CTA has --
  // also sweeps min(nseed, length) threshold-only sample tiles (score_tc.cuh, segment_before).  walk()
  // cuts ranges of cost <= target; the smallest target whose last range also fits is found by bisection.
  const double SEG_COST = 0.35;
  const long long ntn = s.ntiles_n, total = s.total_tiles;
  auto walk = [&](double target, int* tb) -> double {   // returns the cost of the last CTA's range
    long long pos = 0;
    tb[0] = 0;
    for (int b = 0; b < s.grid; ++b) {
      const long long t0 = pos;
      double cost = 0.0;
      const bool last = b == s.grid - 1;
      while (pos < total) {
        const long long row_end = (pos / ntn + 1) * ntn;
        const long long seg_len = row_end - pos;
        const bool head = pos % ntn == 0;
        const double samples = (double)(seg_len < s.nseed ? seg_len : s.nseed);
        double over = SEG_COST + (head ? samples : 0.0);
        // the tail of a row that leaves no room for the next row's head is all this CTA sweeps: sampled
        if (!head && pos == t0 && (double)seg_len + over + SEG_COST + s.nseed + 1.0 > target) over += samples;
        long long n = seg_len;
        if (!last) {
          const double room = target - cost - over;
          if (room < 1.0 && pos > t0) break;             // not worth starting another segment
          const long long fit = room < 1.0 ? 1 : (long long)room;
          if (fit < n) n = fit;
        }
        cost += over + (double)n;
        pos += n;
        if (pos < row_end) break;                        // budget exhausted inside the row
      }
      tb[b + 1] = (int)pos;
      if (last) return cost;
    }
    return 0.0;
  };
  {
    // every query tile starts a segment, every CTA boundary at most one more
    const double all = (double)total + ((double)s.num_mtiles + s.grid) * (SEG_COST + s.nseed);
    double lo = (double)total / s.grid, hi = 2.0 * all / s.grid + s.nseed + 2.0;
    for (int it = 0; it < 40; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (walk(mid, s.tb) <= mid) hi = mid;
      else lo = mid;
    }
    walk(hi, s.tb);
    s.tb[s.grid] = (int)total;
  }
  // sub-list slots per row = most CTAs whose ranges touch one query tile's sweep (pieces are numbered
  // from the first CTA that touches the row)
  int pmax = 1;
  long long max_range = 1;
  {
    int b_first = 0;
    for (long long m = 0; m < s.num_mtiles; ++m) {
      const long long r0 = m * ntn, r1 = r0 + ntn;
      while (b_first + 1 < s.grid && s.tb[b_first + 1] <= r0) ++b_first;   // first CTA with tb[b+1] > r0
      int b_last = b_first;
      while (b_last + 1 < s.grid && s.tb[b_last + 1] < r1) ++b_last;
      if (b_last - b_first + 1 > pmax) pmax = b_last - b_first + 1;
    }
    for (int b = 0; b < s.grid; ++b)
      if (s.tb[b + 1] - s.tb[b] > max_range) max_range = s.tb[b + 1] - s.tb[b];
  }
  s.P = pmax;
  // Expected bytes one epilogue thread appends over a piece of T tiles.  A record is a quad of 4 adjacent
  // columns (16 bytes), appended when its maximum beats the row's bound: with a cold bound a whole tile
  // (16 quads) goes out, later about 32/t items per tile t (about 130 items sit above a row's 32-group
  // bound, a quarter of them in this thread's columns), nearly all in different quads; seeded pieces start
  // at the rate of tile nseed+1.  Segments shorter than 4*nseed tiles run unseeded and may append whole
  // tiles.  3x margin on the sparse part.
  const double T = (double)(max_range < s.ntiles_n ? max_range : s.ntiles_n);
  const double tile_bytes = score::QCOLS / 4 * 16.0;
  double need;
  if (s.nseed > 0) {
    const double short_tiles = T < 4.0 * s.nseed - 1.0 ? T : 4.0 * s.nseed - 1.0;
    need = short_tiles * tile_bytes;
    if (T >= 4.0 * s.nseed) {
      const double seeded = 16.0 * 3.0 * (16.0 + 32.0 * log((T + s.nseed) / s.nseed));
      if (seeded > need) need = seeded;
    }
  } else {
    need = 16.0 * 3.0 * (64.0 + 32.0 * log(T));
  }
  if (need > T * tile_bytes) need = T * tile_bytes;          // a thread cannot append more than it sees
  if (nseed_override == 0) {
    // rank-of-target variant: a thread appends the quads holding an element inside the band around the
    // target's value -- about 1-2 % of its columns for a target in the bulk of the ranking; budget 6 %
    const double band = 0.06 * T * score::QCOLS * 16.0;
    if (band > need) need = band;
  }
  need += tile_bytes;                                        // a list closes one tile before it is full
  int cap = 2 * score::QCOLS;                                // 8-byte units; power of two: a sub-list is aligned to its size
  while (cap * 8 < (int)need && cap < 8192) cap *= 2;
  s.CAP = cap;
  const size_t nlists = (size_t)s.P * score::NQ;
  size_t o = 0;
  s.off_a16 = o;     o += align_up((size_t)Q * 256 * 2, 256);
  s.off_rq = o;      o += align_up((size_t)Q * 4, 256);
  s.off_anorm = o;   o += align_up((size_t)Q * 4, 256);
  s.off_thr = o;     o += align_up((size_t)Q * 4, 256);
  s.off_rowcnt = o;  o += align_up((size_t)Q * nlists * 4, 256);
  s.off_gmax = o;    o += align_up((size_t)Q * nlists * score::GROUPS * 4, 256);
  s.off_rowflag = o; o += align_up((size_t)Q * 4, 256);
  s.off_rowbuf = o;  o += align_up((size_t)Q * nlists * s.CAP * 8 + (size_t)s.CAP * 8, 256);   // + slack to align the base
  s.off_cnt = o;     o += 256;
  s.off_rows = o;    o += align_up((size_t)Q * 4, 256);
  // exhaustive kernel: per-slice lists of rows split over several CTAs (units <= max(its grid, Q)), per-row counts
  const size_t xunits = (size_t)(Q > 2 * num_sms ? Q : 2 * num_sms);
  s.off_xpart = o;   o += align_up(xunits * 32 * 8, 256);
  s.off_xdone = o;   o += align_up((size_t)Q * 4, 256);
  s.total = o;
  return s;
}

size_t seam_score_workspace_bytes(const seam_handle* h, int Q, int G, int k) {
  (void)k;
  if (!h || Q <= 0 || G <= 0) return 256;
  return plan_score(h->num_sms, Q, G, -1, h->score_nseed, h->dbg_score_grid).total;
}

int seam_score_plan(const seam_handle* h, int Q, int G, int64_t* out) {
  if (!h || !out || Q <= 0 || G <= 0) return SEAM_ERR_BAD_ARG;
  const ScorePlan s = plan_score(h->num_sms, Q, G, -1, h->score_nseed, h->dbg_score_grid);
  const int64_t v[14] = {s.num_mtiles, s.ntiles_n, s.grid, s.P, s.CAP, (int64_t)s.off_a16,
                         (int64_t)s.off_rq, (int64_t)s.off_anorm, (int64_t)s.off_thr, (int64_t)s.off_rowcnt,
                         (int64_t)s.off_rowbuf, (int64_t)s.off_cnt, (int64_t)s.off_rows, (int64_t)s.total};
  for (int i = 0; i < 14; ++i) out[i] = v[i];
  return SEAM_OK;
}

// Pure host logic (no handle, no device): the tile ranges of the scorer's persistent CTAs.
int seam_score_partition(int num_sms, int Q, int G, int rank_variant, int32_t* bounds, int bounds_len, int32_t* out6) {
  if (num_sms < 1 || Q <= 0 || G <= 0 || !bounds || !out6) return SEAM_ERR_BAD_ARG;
  const ScorePlan s = plan_score(num_sms, Q, G, rank_variant ? 0 : -1);
  if (bounds_len < s.grid + 1) return SEAM_ERR_BAD_ARG;
  for (int b = 0; b <= s.grid; ++b) bounds[b] = s.tb[b];
  out6[0] = s.num_mtiles;
  out6[1] = s.ntiles_n;
  out6[2] = s.grid;
  out6[3] = s.P;
  out6[4] = s.CAP;
  out6[5] = s.nseed;
  return SEAM_OK;
}

__global__ void fill_empty_topk_kernel(float* s, float* d, int32_t* i, size_t n) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    s[t] = 0.f;
    d[t] = -INFINITY;
    i[t] = -1;
  }
}

static int encode_map_fp16_rows(seam_handle* h, CUtensorMap* map, const void* base, int rows, int box_rows) {
  const cuuint64_t dims[2] = {256, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {256 * 2};
  const cuuint32_t box[2] = {(cuuint32_t)score::BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, SEAM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SEAM_OK;
}

static int score_topk_impl(seam_handle* h, const xchg::Exchange* x, const float* q, int Q, const float* g,
                           const void* g16, const float* cg, const float* gstat, int G, int index_offset, int k,
                           float* out_score, float* out_margin, int32_t* out_idx, int32_t* stats, void* workspace,
                           size_t workspace_bytes, void* stream_) {
  if (!h->have_scorer) return fail(h, SEAM_ERR_STATE, "seam_score_topk: scorer weights not loaded");
  if (Q < 0 || G < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_topk: negative size");
  if (k < 1 || k > SEAM_MAX_K) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_score_topk: k=%d outside [1,%d]", k, SEAM_MAX_K);
  if (Q == 0 && !x) return SEAM_OK;
  if (!x && (!out_score || !out_margin || !out_idx)) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_topk: null output");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  if (stats) SEAM_CUDA(h, cudaMemsetAsync(stats, 0, 16, stream));
  if (x && (Q == 0 || G == 0))
    return fail(h, SEAM_ERR_UNSUPPORTED, "seam_sharded_score_topk: every rank needs queries and a non-empty gallery shard");
  const int x_on = x ? 1 : 0;
  xchg::Exchange xe;
  memset(&xe, 0, sizeof(xe));
  if (x) xe = *x;
  if (x) q = x->q_all[x->rank];      // alignment checks below; the kernels pick the step's parity half themselves
  if (G == 0) {
    const size_t n = (size_t)Q * k;
    fill_empty_topk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(out_score, out_margin, out_idx, n);
    SEAM_LAUNCHED(h, "fill_empty_topk_kernel");
    return SEAM_OK;
  }
  if (!q || !g || !g16 || !cg || !gstat || !workspace) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_topk: null pointer");
  if (!aligned16(q) || !aligned16(g) || !aligned16(g16) || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return fail(h, SEAM_ERR_UNSUPPORTED, "seam_score_topk: q/g/g16 need 16-byte, workspace 256-byte alignment");
  const ScorePlan s = plan_score(h->num_sms, Q, G, -1, h->score_nseed, h->dbg_score_grid);
  if (s.ntiles_n > (1 << 18))
    return fail(h, SEAM_ERR_UNSUPPORTED, "seam_score_topk: G=%d exceeds the 2^26 gallery rows one shard may hold", G);
  if (workspace_bytes < s.total)
    return fail(h, SEAM_ERR_STATE, "seam_score_topk: workspace too small (%zu < %zu)", workspace_bytes, s.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* a16 = reinterpret_cast<__half*>(ws + s.off_a16);
  float* rq = reinterpret_cast<float*>(ws + s.off_rq);
  float* anorm = reinterpret_cast<float*>(ws + s.off_anorm);
  uint32_t* thr = reinterpret_cast<uint32_t*>(ws + s.off_thr);
  uint32_t* rowcnt = reinterpret_cast<uint32_t*>(ws + s.off_rowcnt);
  float* gmax = reinterpret_cast<float*>(ws + s.off_gmax);
  uint32_t* rowflag = reinterpret_cast<uint32_t*>(ws + s.off_rowflag);
  // sub-lists are aligned to their (power-of-two) size so that none straddles a 4 GiB boundary
  uint2* rowbuf = reinterpret_cast<uint2*>(
      align_up(reinterpret_cast<uintptr_t>(ws + s.off_rowbuf), (size_t)s.CAP * 8));
  int32_t* counters = reinterpret_cast<int32_t*>(ws + s.off_cnt);
  int32_t* frows = reinterpret_cast<int32_t*>(ws + s.off_rows);
  const size_t xunits = (size_t)(Q > 2 * h->num_sms ? Q : 2 * h->num_sms);
  float* xpart_d = reinterpret_cast<float*>(ws + s.off_xpart);
  int32_t* xpart_i = reinterpret_cast<int32_t*>(ws + s.off_xpart + xunits * 32 * 4);
  int32_t* xdone = reinterpret_cast<int32_t*>(ws + s.off_xdone);

  {
    ProfileScope prof(h, SEAM_KERNEL_PREP_QUERIES, stream);
    exact::prep_queries_kernel<<<(Q + 7) / 8, 256, 0, stream>>>(q, Q, h->fold, a16, rq, anorm, thr, rowflag, counters, xdone, x_on, xe);
    SEAM_LAUNCHED(h, "prep_queries_kernel");
  }

  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = encode_map_fp16_rows(h, &tmA, a16, Q, score::BM)) != SEAM_OK) return rc;
  if ((rc = encode_map_fp16_rows(h, &tmB, g16, G, score::BN)) != SEAM_OK) return rc;
  score::Params sp;
  sp.Q = Q;
  sp.G = G;
  sp.num_mtiles = s.num_mtiles;
  sp.ntiles_n = s.ntiles_n;
  sp.total_tiles = s.total_tiles;
  sp.P = s.P;
  sp.CAP = s.CAP;
  sp.nseed = s.nseed;
  for (int b = 0; b <= s.grid; ++b) sp.tb[b] = s.tb[b];
#ifdef SEAM_DIAGNOSTIC_VARIANTS
  sp.mode = env_int("SEAM_DEBUG_SCORE_MODE", 0);   // developer diagnostics only (SEAM_BUILD_DIAGNOSTICS=1 builds)
#else
  sp.mode = 0;
#endif
  sp.cg = cg;
  sp.thr_global = thr;
  sp.rowcnt = rowcnt;
  sp.rowflag = rowflag;
  sp.rowbuf = rowbuf;
  sp.gmax = gmax;
  sp.rank_lo = nullptr;
  sp.rank_hi = nullptr;
  sp.rank_above = nullptr;
  // developer diagnostics: SEAM_DEBUG_CTA_NS=1 records per-CTA durations at the tail of the fallback-row list
  sp.cta_ns = (h->dbg_cta_ns && (size_t)Q * 4 >= 8192)
                  ? reinterpret_cast<unsigned long long*>(ws + s.off_rows + (((size_t)Q * 4 - 4096) & ~(size_t)7))
                  : nullptr;
  {
    ProfileScope prof(h, SEAM_KERNEL_SCORE, stream);
#ifdef SEAM_DIAGNOSTIC_VARIANTS
    if (sp.mode == 2) score::score_topk_kernel<2><<<s.grid, score::THREADS, score::SMEM_BYTES, stream>>>(tmA, tmB, sp);
    else if (sp.mode == 3) score::score_topk_kernel<3><<<s.grid, score::THREADS, score::SMEM_BYTES, stream>>>(tmA, tmB, sp);
    else if (sp.mode == 4) score::score_topk_kernel<4><<<s.grid, score::THREADS, score::SMEM_BYTES, stream>>>(tmA, tmB, sp);
    else if (sp.mode == 5) score::score_topk_kernel<5><<<s.grid, score::THREADS, score::SMEM_BYTES, stream>>>(tmA, tmB, sp);
    else
#endif
    score::score_topk_kernel<0><<<s.grid, score::THREADS, score::SMEM_BYTES, stream>>>(tmA, tmB, sp);
    SEAM_LAUNCHED(h, "score_topk_kernel");
  }

  exact::RescoreParams rp;
  rp.q = q;
  rp.g = g;
  rp.fold = h->fold;
  rp.rowbuf = rowbuf;
  rp.rowcnt = rowcnt;
  rp.rowflag = rowflag;
  rp.thr_global = thr;
  rp.gmax = gmax;
  rp.rq = rq;
  rp.anorm = anorm;
  rp.gstat = gstat;
  rp.Q = Q;
  rp.G = G;
  rp.nlists = s.P * score::NQ;
  rp.CAP = s.CAP;
  rp.k = k;
  rp.index_offset = index_offset;
  rp.out_score = out_score;
  rp.out_margin = out_margin;
  rp.out_idx = out_idx;
  rp.counters = counters;
  rp.fallback_rows = frows;
  rp.x_on = x_on;
  rp.x = xe;
  rp.plan.ntiles_n = s.ntiles_n;
  rp.plan.nb = s.grid;
  for (int b = 0; b <= s.grid; ++b) rp.plan.tb[b] = s.tb[b];
  {
    ProfileScope prof(h, SEAM_KERNEL_RESCORE, stream);
    exact::rescore_kernel<<<(Q + 7) / 8, 256, 0, stream>>>(rp);
    SEAM_LAUNCHED(h, "rescore_kernel");
  }

  exact::ExactParams ep;
  ep.q = q;
  ep.g = g;
  ep.fold = h->fold;
  ep.Q = Q;
  ep.G = G;
  ep.k = k;
  ep.index_offset = index_offset;
  ep.count = counters;
  ep.rows = frows;
  ep.out_score = out_score;
  ep.out_margin = out_margin;
  ep.out_idx = out_idx;
  ep.part_d = xpart_d;
  ep.part_i = xpart_i;
  ep.done = xdone;
  ep.x_on = x_on;
  ep.x = xe;
  {
    ProfileScope prof(h, SEAM_KERNEL_EXACT, stream);
    exact::exact_topk_kernel<<<2 * h->num_sms, 256, 0, stream>>>(ep);
    SEAM_LAUNCHED(h, "exact_topk_kernel");
  }
  if (stats) SEAM_CUDA(h, cudaMemcpyAsync(stats, counters, 4, cudaMemcpyDeviceToDevice, stream));
  return SEAM_OK;
}

int seam_score_topk(seam_handle* h, const float* q, int Q, const float* g, const void* g16, const float* cg,
                    const float* gstat, int G, int index_offset, int k, float* out_score, float* out_margin,
                    int32_t* out_idx, int32_t* stats, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  return score_topk_impl(h, nullptr, q, Q, g, g16, cg, gstat, G, index_offset, k, out_score, out_margin, out_idx, stats,
                         workspace, workspace_bytes, stream_);
}

int seam_search(seam_handle* h, const float* seq, const uint8_t* mask, const int32_t* lens, int Tmax, int Q,
                int64_t frame_stride, int64_t track_stride, float* q_out, const float* g, const void* g16, const float* cg,
                const float* gstat, int G, int index_offset, int k, float* out_score, float* out_margin, int32_t* out_idx,
                int32_t* stats, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!q_out && Q > 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_search: q_out is null");
  int rc = aggregate_impl(h, nullptr, 0, 0, seq, mask, lens, Tmax, Q, frame_stride, track_stride, q_out, nullptr,
                          static_cast<cudaStream_t>(stream_), "seam_search");
  if (rc != SEAM_OK) return rc;
  return score_topk_impl(h, nullptr, q_out, Q, g, g16, cg, gstat, G, index_offset, k, out_score, out_margin, out_idx, stats,
                         workspace, workspace_bytes, stream_);
}

int seam_sharded_score_topk(seam_handle* h, const seam_exchange* x, const float* g, const void* g16, const float* cg,
                            const float* gstat, int G, int index_offset, int32_t* stats, void* workspace,
                            size_t workspace_bytes, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  xchg::Exchange e;
  int rc = import_exchange(h, x, &e, "seam_sharded_score_topk");
  if (rc != SEAM_OK) return rc;
  return score_topk_impl(h, &e, nullptr, x->Q, g, g16, cg, gstat, G, index_offset, x->k, nullptr, nullptr, nullptr, stats,
                         workspace, workspace_bytes, stream_);
}

int seam_sharded_merge(seam_handle* h, const seam_exchange* x, float* out_score, float* out_margin, int32_t* out_idx,
                       void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  xchg::Exchange e;
  int rc = import_exchange(h, x, &e, "seam_sharded_merge");
  if (rc != SEAM_OK) return rc;
  if (!e.final_score[0] && (!out_score || !out_margin || !out_idx))
    return fail(h, SEAM_ERR_BAD_ARG, "seam_sharded_merge: no final buffers in the exchange and null outputs");
  DeviceGuard guard(h->device);
  const int own = e.q_lo[e.rank + 1] - e.q_lo[e.rank];
  const int grid = own > 0 ? (own + 7) / 8 : 1;
  ProfileScope prof(h, SEAM_KERNEL_MERGE, static_cast<cudaStream_t>(stream_));
  exact::merge_sharded_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(e, out_score, out_margin, out_idx);
  SEAM_LAUNCHED(h, "merge_sharded_kernel");
  return SEAM_OK;
}

int seam_score_dense(seam_handle* h, const float* q, int Q, const float* g, int G, float* x5, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_scorer) return fail(h, SEAM_ERR_STATE, "seam_score_dense: scorer weights not loaded");
  if (Q < 0 || G < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_dense: negative size");
  if (Q == 0 || G == 0) return SEAM_OK;
  if (!q || !g || !x5) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_dense: null pointer");
  if ((reinterpret_cast<uintptr_t>(x5) & 7u)) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_score_dense: x5 alignment");
  DeviceGuard guard(h->device);
  dim3 grid((G + 31) / 32, (Q + 31) / 32);
  if (grid.y > 65535) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_score_dense: Q too large for the dense path");
  exact::dense_logits_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(q, Q, g, G, h->fold, x5);
  SEAM_LAUNCHED(h, "dense_logits_kernel");
  return SEAM_OK;
}

int seam_score_prob(seam_handle* h, const float* q, int Q, const float* g, int G, float* prob, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_scorer) return fail(h, SEAM_ERR_STATE, "seam_score_prob: scorer weights not loaded");
  if (Q < 0 || G < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_prob: negative size");
  if (Q == 0 || G == 0) return SEAM_OK;
  if (!q || !g || !prob) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_prob: null pointer");
  DeviceGuard guard(h->device);
  dim3 grid((G + 31) / 32, (Q + 31) / 32);
  if (grid.y > 65535) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_score_prob: Q too large for the dense path");
  exact::dense_logits_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(q, Q, g, G, h->fold, prob);
  SEAM_LAUNCHED(h, "dense_logits_kernel<prob>");
  return SEAM_OK;
}

int seam_rank_fused_distances(seam_handle* h, const float* frames, const int32_t* start, int P, const float* g, int G,
                              const int32_t* target, int32_t* rank_avg, int32_t* rank_max, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_scorer) return fail(h, SEAM_ERR_STATE, "seam_rank_fused_distances: scorer weights not loaded");
  if (P < 0 || G <= 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_rank_fused_distances: bad size");
  if (P == 0) return SEAM_OK;
  if (!start || !g || !target || !rank_avg || !rank_max)
    return fail(h, SEAM_ERR_BAD_ARG, "seam_rank_fused_distances: null pointer");
  if ((frames && !aligned16(frames)) || !aligned16(g))
    return fail(h, SEAM_ERR_UNSUPPORTED, "seam_rank_fused_distances: 16-byte alignment");
  const int chunks = (G + exact::FD_ROWS - 1) / exact::FD_ROWS;
  if (chunks > 65535) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_rank_fused_distances: G too large");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  SEAM_CUDA(h, cudaMemsetAsync(rank_avg, 0, (size_t)P * 4, stream));
  SEAM_CUDA(h, cudaMemsetAsync(rank_max, 0, (size_t)P * 4, stream));
  exact::FusedDistParams fp;
  fp.frames = frames;
  fp.start = start;
  fp.g = g;
  fp.target = target;
  fp.fold = h->fold;
  fp.P = P;
  fp.G = G;
  fp.rank_avg = rank_avg;
  fp.rank_max = rank_max;
  exact::fused_dist_rank_kernel<<<dim3((unsigned)P, (unsigned)chunks), 256, 0, stream>>>(fp);
  SEAM_LAUNCHED(h, "fused_dist_rank_kernel");
  return SEAM_OK;
}

int seam_rank_of_target(seam_handle* h, const float* q, int Q, const float* g, int G, const int32_t* target,
                        int32_t* out_rank, float* out_margin, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_scorer) return fail(h, SEAM_ERR_STATE, "seam_rank_of_target: scorer weights not loaded");
  if (Q < 0 || G <= 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_rank_of_target: bad size");
  if (Q == 0) return SEAM_OK;
  if (!q || !g || !target || !out_rank) return fail(h, SEAM_ERR_BAD_ARG, "seam_rank_of_target: null pointer");
  if (!aligned16(q) || !aligned16(g)) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_rank_of_target: 16-byte alignment");
  DeviceGuard guard(h->device);
  exact::rank_of_target_kernel<<<Q, 256, 0, static_cast<cudaStream_t>(stream_)>>>(q, Q, g, G, target, h->fold, nullptr,
                                                                                   nullptr, out_rank, out_margin);
  SEAM_LAUNCHED(h, "rank_of_target_kernel");
  return SEAM_OK;
}

size_t seam_rank_workspace_bytes(const seam_handle* h, int Q, int G) {
  if (!h || Q <= 0 || G <= 0) return 256;
  return plan_score(h->num_sms, Q, G, 0, h->score_nseed, h->dbg_score_grid).total;
}

// Tensor-core path: prepare queries -> exact target margins + bands -> score_topk_kernel<VAR_RANK> (counts
// what is certainly above, appends what is inside the band) -> rank_resolve_kernel -> exhaustive kernel for
// the rows that could not be certified.  The workspace is laid out like seam_score_topk's (no sample
// tiles); the band, the target margins and the per-sub-list counts live in its group-maxima region.
int seam_rank_of_target_prepared(seam_handle* h, const float* q, int Q, const float* g, const void* g16,
                                 const float* cg, const float* gstat, int G, const int32_t* target, int32_t* out_rank,
                                 float* out_margin, int32_t* stats, void* workspace, size_t workspace_bytes,
                                 void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_scorer) return fail(h, SEAM_ERR_STATE, "seam_rank_of_target_prepared: scorer weights not loaded");
  if (Q < 0 || G <= 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_rank_of_target_prepared: bad size");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  if (stats) SEAM_CUDA(h, cudaMemsetAsync(stats, 0, 16, stream));
  if (Q == 0) return SEAM_OK;
  if (!q || !g || !g16 || !cg || !gstat || !target || !out_rank || !workspace)
    return fail(h, SEAM_ERR_BAD_ARG, "seam_rank_of_target_prepared: null pointer");
  if (!aligned16(q) || !aligned16(g) || !aligned16(g16) || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return fail(h, SEAM_ERR_UNSUPPORTED, "seam_rank_of_target_prepared: q/g/g16 need 16-byte, workspace 256-byte alignment");
  const ScorePlan s = plan_score(h->num_sms, Q, G, 0, h->score_nseed, h->dbg_score_grid);
  if (s.ntiles_n > (1 << 18))
    return fail(h, SEAM_ERR_UNSUPPORTED, "seam_rank_of_target_prepared: G=%d exceeds the 2^26 gallery rows one shard may hold", G);
  if (workspace_bytes < s.total)
    return fail(h, SEAM_ERR_STATE, "seam_rank_of_target_prepared: workspace too small (%zu < %zu)", workspace_bytes, s.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* a16 = reinterpret_cast<__half*>(ws + s.off_a16);
  float* rq = reinterpret_cast<float*>(ws + s.off_rq);
  float* anorm = reinterpret_cast<float*>(ws + s.off_anorm);
  uint32_t* thr = reinterpret_cast<uint32_t*>(ws + s.off_thr);
  uint32_t* rowcnt = reinterpret_cast<uint32_t*>(ws + s.off_rowcnt);
  float* gmax = reinterpret_cast<float*>(ws + s.off_gmax);
  uint32_t* rowflag = reinterpret_cast<uint32_t*>(ws + s.off_rowflag);
  uint2* rowbuf = reinterpret_cast<uint2*>(align_up(reinterpret_cast<uintptr_t>(ws + s.off_rowbuf), (size_t)s.CAP * 8));
  int32_t* counters = reinterpret_cast<int32_t*>(ws + s.off_cnt);
  int32_t* frows = reinterpret_cast<int32_t*>(ws + s.off_rows);
  const int nlists = s.P * score::NQ;
  // the group-maxima region (Q x nlists x 16 words) is free in this variant
  int32_t* above = reinterpret_cast<int32_t*>(gmax);
  float* lo = gmax + (size_t)Q * nlists;
  float* hi = lo + Q;
  float* dtarget = hi + Q;

  xchg::Exchange no_x;
  memset(&no_x, 0, sizeof(no_x));
  exact::prep_queries_kernel<<<(Q + 7) / 8, 256, 0, stream>>>(q, Q, h->fold, a16, rq, anorm, thr, rowflag, counters, nullptr, 0, no_x);
  SEAM_LAUNCHED(h, "prep_queries_kernel");
  exact::rank_prep_kernel<<<(Q + 7) / 8, 256, 0, stream>>>(q, Q, g, target, h->fold, rq, anorm, gstat, lo, hi, dtarget,
                                                          above, nlists);
  SEAM_LAUNCHED(h, "rank_prep_kernel");

  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = encode_map_fp16_rows(h, &tmA, a16, Q, score::BM)) != SEAM_OK) return rc;
  if ((rc = encode_map_fp16_rows(h, &tmB, g16, G, score::BN)) != SEAM_OK) return rc;
  score::Params sp;
  sp.Q = Q;
  sp.G = G;
  sp.num_mtiles = s.num_mtiles;
  sp.ntiles_n = s.ntiles_n;
  sp.total_tiles = s.total_tiles;
  sp.P = s.P;
  sp.CAP = s.CAP;
  sp.nseed = 0;
  sp.mode = score::VAR_RANK;
  for (int b = 0; b <= s.grid; ++b) sp.tb[b] = s.tb[b];
  sp.cg = cg;
  sp.thr_global = thr;
  sp.rowcnt = rowcnt;
  sp.rowflag = rowflag;
  sp.rowbuf = rowbuf;
  sp.gmax = gmax;
  sp.cta_ns = nullptr;
  sp.rank_lo = lo;
  sp.rank_hi = hi;
  sp.rank_above = above;
  score::score_topk_kernel<score::VAR_RANK><<<s.grid, score::THREADS, score::SMEM_BYTES, stream>>>(tmA, tmB, sp);
  SEAM_LAUNCHED(h, "score_topk_kernel<rank>");

  exact::RankResolveParams rp;
  rp.q = q;
  rp.g = g;
  rp.fold = h->fold;
  rp.rowbuf = rowbuf;
  rp.rowcnt = rowcnt;
  rp.rowflag = rowflag;
  rp.above = above;
  rp.lo = lo;
  rp.hi = hi;
  rp.dtarget = dtarget;
  rp.rq = rq;
  rp.gstat = gstat;
  rp.target = target;
  rp.Q = Q;
  rp.G = G;
  rp.nlists = nlists;
  rp.CAP = s.CAP;
  rp.out_rank = out_rank;
  rp.out_margin = out_margin;
  rp.counters = counters;
  rp.fallback_rows = frows;
  rp.plan.ntiles_n = s.ntiles_n;
  rp.plan.nb = s.grid;
  for (int b = 0; b <= s.grid; ++b) rp.plan.tb[b] = s.tb[b];
  exact::rank_resolve_kernel<<<(Q + 7) / 8, 256, 0, stream>>>(rp);
  SEAM_LAUNCHED(h, "rank_resolve_kernel");
  exact::rank_of_target_kernel<<<2 * h->num_sms, 256, 0, stream>>>(q, Q, g, G, target, h->fold, counters, frows, out_rank,
                                                                  out_margin);
  SEAM_LAUNCHED(h, "rank_of_target_kernel");
  if (stats) SEAM_CUDA(h, cudaMemcpyAsync(stats, counters, 4, cudaMemcpyDeviceToDevice, stream));
  return SEAM_OK;
}

// ------------------------------------------------------------------------------ backward (f4)
int seam_aggregate_backward(seam_handle* h, const seam_weights* w, const float* seq, const uint8_t* mask,
                            const int32_t* lens, int Tmax, int Q, int64_t frame_stride, int64_t track_stride,
                            const float* dout, float* dseq, const seam_weight_grads* grads, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!w || !grads) return fail(h, SEAM_ERR_BAD_ARG, "seam_aggregate_backward: weights / grads struct is null");
  if (Q < 0 || Tmax < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_aggregate_backward: negative size");
  if (Q == 0 || Tmax == 0) return SEAM_OK;
  if (Tmax > bwd::BWD_MAX_T)
    return fail(h, SEAM_ERR_UNSUPPORTED, "seam_aggregate_backward: Tmax=%d exceeds the %d frames per track the training path supports",
                Tmax, bwd::BWD_MAX_T);
  if (!seq || !dout || !dseq) return fail(h, SEAM_ERR_BAD_ARG, "seam_aggregate_backward: null pointer");
  const float* const* wp = reinterpret_cast<const float* const*>(w);
  float* const* gp = reinterpret_cast<float* const*>(grads);
  for (int i = 0; i < 11; ++i)
    if (!wp[i] || !gp[i]) return fail(h, SEAM_ERR_BAD_ARG, "seam_aggregate_backward: null weight / gradient pointer %d", i);
  DeviceGuard guard(h->device);
  bwd::AggBwdParams p;
  p.seq = seq;
  p.mask = mask;
  p.lens = lens;
  p.Tmax = Tmax;
  p.Q = Q;
  p.frame_stride = frame_stride;
  p.track_stride = track_stride;
  p.dout = dout;
  p.dseq = dseq;
  p.w = {w->theta_w, w->theta_b, w->phi_w, w->phi_b, w->g_w, w->g_b, w->W_w, w->W_b, w->concat_w, w->att_w, w->att_b};
  p.g = {grads->theta_w, grads->theta_b, grads->phi_w, grads->phi_b, grads->g_w, grads->g_b,
         grads->W_w,     grads->W_b,     grads->concat_w, grads->att_w, grads->att_b};
  bwd::agg_backward_kernel<<<Q, 256, bwd::agg_bwd_smem_bytes(Tmax), static_cast<cudaStream_t>(stream_)>>>(p);
  SEAM_LAUNCHED(h, "agg_backward_kernel");
  return SEAM_OK;
}

int seam_score_dense_backward(seam_handle* h, const float* last_w, const float* q, int Q, const float* g, int G,
                              const float* dx5, float* dq, float* dg, float* dlast_w, float* dlast_b, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (Q < 0 || G < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_score_dense_backward: negative size");
  if (Q == 0 || G == 0) return SEAM_OK;
  if (!last_w || !q || !g || !dx5 || !dq || !dg || !dlast_w || !dlast_b)
    return fail(h, SEAM_ERR_BAD_ARG, "seam_score_dense_backward: null pointer");
  if ((reinterpret_cast<uintptr_t>(dx5) & 7u)) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_score_dense_backward: dx5 alignment");
  DeviceGuard guard(h->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  bwd::scorer_backward_q_kernel<<<Q, 256, 0, stream>>>(q, Q, g, G, last_w, dx5, dq, dlast_w, dlast_b);
  SEAM_LAUNCHED(h, "scorer_backward_q_kernel");
  bwd::scorer_backward_g_kernel<<<G, 256, 0, stream>>>(q, Q, g, G, last_w, dx5, dg);
  SEAM_LAUNCHED(h, "scorer_backward_g_kernel");
  return SEAM_OK;
}

// ------------------------------------------------------------------------------ conv tower (f3)
namespace {
constexpr int TW_GUARD = 192;                                    // rows a shifted A box may reach past a layer's last position
constexpr size_t TW_OFF_B[4] = {0, 256, 512, 768};               // biases inside tw_misc (floats)
constexpr size_t TW_OFF_WT = 1792, TW_OFF_LB = TW_OFF_WT + 1024 * 256, TW_OFF_SC = TW_OFF_LB + 256,
                 TW_OFF_SH = TW_OFF_SC + 256, TW_MISC_TOTAL = TW_OFF_SH + 256;
const int TW_HW[5] = {14, 12, 10, 8, 6};
struct TowerPlan {
  size_t off[5], total;
};
TowerPlan plan_tower(int K) {
  TowerPlan t;
  size_t o = 0;
  for (int l = 0; l < 5; ++l) {
    t.off[l] = o;
    const size_t rows = (size_t)K * TW_HW[l] * TW_HW[l] + (l < 4 ? TW_GUARD : 0);
    o += align_up(rows * (l < 4 ? 256 : 1024) * 2, 1024);
  }
  t.total = o;
  return t;
}
int encode_map_fp16_2d(seam_handle* h, CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {inner * 2};
  const cuuint32_t box[2] = {(cuuint32_t)tower::BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, SEAM_ERR_CUDA, "cuTensorMapEncodeTiled (tower) failed with CUresult %d", (int)r);
  return SEAM_OK;
}
}  // namespace

int seam_tower_load_weights(seam_handle* h, const float* const* conv_w, const float* const* conv_b, const float* lin_w,
                            const float* lin_b, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                            const float* bn_var, float bn_eps, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!conv_w || !conv_b || !lin_w || !lin_b || !bn_gamma || !bn_beta || !bn_mean || !bn_var)
    return fail(h, SEAM_ERR_BAD_ARG, "seam_tower_load_weights: null pointer");
  for (int l = 0; l < 4; ++l)
    if (!conv_w[l] || !conv_b[l]) return fail(h, SEAM_ERR_BAD_ARG, "seam_tower_load_weights: null conv parameter %d", l);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  for (int l = 0; l < 4; ++l) {
    const int cout = l < 3 ? 256 : 1024;
    if (!h->tw_conv[l]) SEAM_CUDA(h, cudaMalloc(&h->tw_conv[l], (size_t)cout * 2304 * 2));
  }
  if (!h->tw_misc) SEAM_CUDA(h, cudaMalloc(&h->tw_misc, TW_MISC_TOTAL * 4));
  for (int l = 0; l < 4; ++l) {
    const int cout = l < 3 ? 256 : 1024;
    tower::conv_weight_rows_kernel<<<cout, 256, 0, stream>>>(conv_w[l], h->tw_conv[l], cout);
    SEAM_LAUNCHED(h, "conv_weight_rows_kernel");
    SEAM_CUDA(h, cudaMemcpyAsync(h->tw_misc + TW_OFF_B[l], conv_b[l], (size_t)cout * 4, cudaMemcpyDeviceToDevice, stream));
  }
  SEAM_CUDA(h, cudaMemcpyAsync(h->tw_misc + TW_OFF_LB, lin_b, 256 * 4, cudaMemcpyDeviceToDevice, stream));
  tower::tower_fold_kernel<<<256, 256, 0, stream>>>(lin_w, bn_gamma, bn_beta, bn_mean, bn_var, bn_eps, h->tw_misc + TW_OFF_WT,
                                                    h->tw_misc + TW_OFF_SC, h->tw_misc + TW_OFF_SH);
  SEAM_LAUNCHED(h, "tower_fold_kernel");
  h->have_tower = true;
  return SEAM_OK;
}

size_t seam_tower_workspace_bytes(int K) { return K > 0 ? plan_tower(K).total : 1024; }

int seam_tower_forward(seam_handle* h, const float* x, int K, float* out, const int64_t* dst_row, void* workspace,
                       size_t workspace_bytes, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (!h->have_tower) return fail(h, SEAM_ERR_STATE, "seam_tower_forward: tower weights not loaded");
  if (K < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_tower_forward: negative size");
  if (K == 0) return SEAM_OK;
  if ((long long)K * 196 > 0x7fffff00ll) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_tower_forward: K too large");
  if (!x || !out || !workspace) return fail(h, SEAM_ERR_BAD_ARG, "seam_tower_forward: null pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023u)) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_tower_forward: workspace must be 1 KB aligned");
  const TowerPlan tp = plan_tower(K);
  if (workspace_bytes < tp.total) return fail(h, SEAM_ERR_STATE, "seam_tower_forward: workspace too small (%zu < %zu)", workspace_bytes, tp.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(h->device);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  ProfileScope prof(h, SEAM_KERNEL_TOWER, stream);
  tower::nchw_to_rows_kernel<<<dim3((unsigned)K, 8), 256, 0, stream>>>(x, reinterpret_cast<__half*>(ws + tp.off[0]), K);
  SEAM_LAUNCHED(h, "nchw_to_rows_kernel");
  for (int l = 0; l < 4; ++l) {
    const int H = TW_HW[l], cout = l < 3 ? 256 : 1024;
    const long long rows_in = (long long)K * H * H;
    CUtensorMap tmA, tmB;
    int rc;
    if ((rc = encode_map_fp16_2d(h, &tmA, ws + tp.off[l], 256, (uint64_t)rows_in + TW_GUARD, tower::BM)) != SEAM_OK) return rc;
    if ((rc = encode_map_fp16_2d(h, &tmB, h->tw_conv[l], 2304, (uint64_t)cout, tower::BN)) != SEAM_OK) return rc;
    tower::ConvParams cp;
    cp.H = H;
    cp.W = H;
    cp.K = K;
    cp.Cout = cout;
    cp.rows_in = rows_in;
    cp.m_tiles = (int)((rows_in + tower::BM - 1) / tower::BM);
    cp.n_tiles = cout / tower::BN;
    cp.bias = h->tw_misc + TW_OFF_B[l];
    cp.out = reinterpret_cast<__half*>(ws + tp.off[l + 1]);
    const long long total = (long long)cp.m_tiles * cp.n_tiles;
    const int grid = total < h->num_sms ? (int)total : h->num_sms;
    tower::conv3x3_kernel<<<grid, tower::THREADS, tower::SMEM_BYTES, stream>>>(tmA, tmB, cp);
    SEAM_LAUNCHED(h, "conv3x3_kernel");
  }
  tower::PoolLinearParams pp;
  pp.a4 = reinterpret_cast<const __half*>(ws + tp.off[4]);
  pp.wt = h->tw_misc + TW_OFF_WT;
  pp.lin_b = h->tw_misc + TW_OFF_LB;
  pp.bn_scale = h->tw_misc + TW_OFF_SC;
  pp.bn_shift = h->tw_misc + TW_OFF_SH;
  pp.dst_row = reinterpret_cast<const long long*>(dst_row);
  pp.out = out;
  pp.K = K;
  tower::pool_linear_bn_kernel<<<(K + tower::PL_ROIS - 1) / tower::PL_ROIS, 256, 0, stream>>>(pp);
  SEAM_LAUNCHED(h, "pool_linear_bn_kernel");
  return SEAM_OK;
}

int seam_upload_tracks(seam_handle* h, const float* seq_host, int Tmax, int Q, int lo, int hi, float* seq_dev,
                       void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (Tmax < 0 || Q < 0 || lo < 0 || hi < lo || hi > Q) return fail(h, SEAM_ERR_BAD_ARG, "seam_upload_tracks: bad range");
  if (Tmax == 0 || hi == lo) return SEAM_OK;
  if (!seq_host || !seq_dev) return fail(h, SEAM_ERR_BAD_ARG, "seam_upload_tracks: null pointer");
  DeviceGuard guard(h->device);
  const size_t n = (size_t)(hi - lo);
  SEAM_CUDA(h, cudaMemcpy2DAsync(seq_dev + n * 256, n * 1024, seq_host + ((size_t)Q + (size_t)lo) * 256,
                                 (size_t)Q * 1024, n * 1024, (size_t)Tmax, cudaMemcpyHostToDevice,
                                 static_cast<cudaStream_t>(stream_)));
  return SEAM_OK;
}

int seam_merge_topk(seam_handle* h, const float* scores, const float* margins, const int32_t* idx, int N, int Q,
                    int k, float* out_score, float* out_margin, int32_t* out_idx, void* stream_) {
  if (!h) return SEAM_ERR_BAD_ARG;
  if (N < 1 || Q < 0) return fail(h, SEAM_ERR_BAD_ARG, "seam_merge_topk: bad size");
  if (k < 1 || k > SEAM_MAX_K) return fail(h, SEAM_ERR_UNSUPPORTED, "seam_merge_topk: k=%d outside [1,%d]", k, SEAM_MAX_K);
  if (Q == 0) return SEAM_OK;
  if (!margins || !idx || !out_score || !out_margin || !out_idx)    // scores may be null: recomputed from the margins
    return fail(h, SEAM_ERR_BAD_ARG, "seam_merge_topk: null pointer");
  DeviceGuard guard(h->device);
  ProfileScope prof(h, SEAM_KERNEL_MERGE, static_cast<cudaStream_t>(stream_));
  exact::merge_topk_kernel<<<(Q + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      scores, margins, idx, N, Q, k, out_score, out_margin, out_idx);
  SEAM_LAUNCHED(h, "merge_topk_kernel");
  return SEAM_OK;
}

}  // extern "C"
