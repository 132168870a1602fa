// extern "C" surface of libkgnet_b200.so (declared in include/kgnet_b200.h).
#include "common.cuh"
#include "decode.cuh"

#include <map>
#include <mutex>
#include <vector>

namespace kg {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}
const char* last_error() { return g_error.c_str(); }

// ---- stage timing -----------------------------------------------------------------------------
struct StageTiming {
  std::mutex mu;
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  struct Rec { int id; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  cudaEvent_t open_ev[KG_MAX_STAGES] = {};
  size_t next = 0;
  cudaEvent_t get() {
    if (next == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[next++];
  }
};
static StageTiming g_timing;
bool timing_enabled() { return g_timing.enabled; }
void stage_begin(int id, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_timing.mu);
  cudaEvent_t e = g_timing.get();
  cudaEventRecord(e, s);
  g_timing.open_ev[id] = e;
}
void StageScope::restage(int new_id) {
  if (!timing_enabled() || new_id == id) return;
  std::lock_guard<std::mutex> lock(g_timing.mu);
  g_timing.open_ev[new_id] = g_timing.open_ev[id];
  id = new_id;
}
void stage_end(int id, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_timing.mu);
  cudaEvent_t e = g_timing.get();
  cudaEventRecord(e, s);
  g_timing.recs.push_back({id, g_timing.open_ev[id], e});
}

// Cached device buffers of kg_decode_host (grown on demand, never shrunk), one set per CUDA device.
struct HostDecodeCtx {
  void* d_in = nullptr; size_t in_bytes = 0;
  void* d_ws = nullptr; size_t ws_bytes = 0;
  void* d_out = nullptr; size_t out_bytes = 0;
  int* h_status = nullptr;     // pinned
};
static std::mutex g_hd_mu;
static std::map<int, HostDecodeCtx> g_hd_by_device;

static int grow(void** p, size_t* have, size_t need) {
  if (*have >= need) return KG_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  KG_CUDA_CHECK(cudaMalloc(p, need));
  *have = need;
  return KG_OK;
}

}  // namespace kg

using namespace kg;

extern "C" {

const char* kg_last_error(void) { return kg::last_error(); }
int kg_abi_version(void) { return KG_ABI_VERSION; }

int kg_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  KG_CUDA_CHECK(cudaGetDevice(&dev));
  KG_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  KG_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

int kg_timing_enable(int on) {
  std::lock_guard<std::mutex> lock(g_timing.mu);
  g_timing.enabled = on != 0;
  g_timing.recs.clear();
  g_timing.next = 0;
  return KG_OK;
}

int kg_timing_collect(float* ms_per_stage, int* launches_per_stage, int n_stages) {
  KG_REQUIRE(ms_per_stage != nullptr && n_stages > 0, "kg_timing_collect: bad arguments");
  KG_CUDA_CHECK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_timing.mu);
  for (int i = 0; i < n_stages; ++i) { ms_per_stage[i] = 0.f; if (launches_per_stage) launches_per_stage[i] = 0; }
  for (auto& r : g_timing.recs) {
    float ms = 0.f;
    KG_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.id < n_stages) { ms_per_stage[r.id] += ms; if (launches_per_stage) launches_per_stage[r.id] += 1; }
  }
  g_timing.recs.clear();
  g_timing.next = 0;
  return KG_OK;
}

size_t kg_decode_workspace_bytes(const kg_decode_config* cfg, const kg_decode_scale* scales) {
  return kg::decode_workspace_bytes(cfg, scales);
}

int kg_decode(const kg_decode_config* cfg, const kg_decode_scale* scales, const kg_decode_outputs* out, void* d_workspace,
              size_t workspace_bytes, void* stream, int* n_launches) {
  return kg::decode_launch(cfg, scales, out, d_workspace, workspace_bytes, (cudaStream_t)stream, n_launches);
}

int kg_decode_host(const kg_decode_config* cfg, const float* const* h_kp, const float* const* h_short,
                   const float* const* h_mid, const int* H, const int* W, const int* box_scale, double* h_dets,
                   int* h_det_count, void* stream_) {
  KG_REQUIRE(cfg && h_kp && h_short && h_mid && H && W && box_scale && h_dets && h_det_count,
             "kg_decode_host: null argument");
  KG_REQUIRE(cfg->n_scales >= 1 && cfg->n_scales <= KG_MAX_SCALES, "kg_decode_host: n_scales=%d", cfg->n_scales);
  cudaStream_t stream = (cudaStream_t)stream_;
  std::lock_guard<std::mutex> lock(g_hd_mu);
  int dev = 0;
  KG_CUDA_CHECK(cudaGetDevice(&dev));
  HostDecodeCtx& g_hd = g_hd_by_device[dev];
  if (g_hd.h_status == nullptr) KG_CUDA_CHECK(cudaMallocHost(&g_hd.h_status, sizeof(int)));
  kg_decode_scale sc[KG_MAX_SCALES] = {};
  size_t in_bytes = 0;
  for (int s = 0; s < cfg->n_scales; ++s) {
    sc[s].H = H[s]; sc[s].W = W[s]; sc[s].box_scale = box_scale[s];
    KG_REQUIRE(H[s] > 0 && W[s] > 0 && cfg->N > 0, "kg_decode_host: bad shape");
    in_bytes += align_up((size_t)cfg->N * 55 * H[s] * W[s] * sizeof(float), 256) + 512;
  }
  const size_t ws = kg::decode_workspace_bytes(cfg, sc);
  if (ws == 0) return KG_ERR_INVALID;
  const size_t det_bytes = sizeof(double) * 5 * cfg->N * cfg->max_boxes;
  const size_t out_bytes = align_up(det_bytes, 256) + align_up(sizeof(int) * cfg->N, 256) + 256;
  KG_TRY(grow(&g_hd.d_in, &g_hd.in_bytes, in_bytes));
  KG_TRY(grow(&g_hd.d_ws, &g_hd.ws_bytes, ws));
  KG_TRY(grow(&g_hd.d_out, &g_hd.out_bytes, out_bytes));
  Arena a(g_hd.d_in, g_hd.in_bytes);
  for (int s = 0; s < cfg->n_scales; ++s) {
    const size_t hw = (size_t)H[s] * W[s], n = cfg->N;
    float* kp = a.take<float>(n * 5 * hw); float* sh = a.take<float>(n * 10 * hw); float* mid = a.take<float>(n * 40 * hw);
    KG_CUDA_CHECK(cudaMemcpyAsync(kp, h_kp[s], n * 5 * hw * sizeof(float), cudaMemcpyHostToDevice, stream));
    KG_CUDA_CHECK(cudaMemcpyAsync(sh, h_short[s], n * 10 * hw * sizeof(float), cudaMemcpyHostToDevice, stream));
    KG_CUDA_CHECK(cudaMemcpyAsync(mid, h_mid[s], n * 40 * hw * sizeof(float), cudaMemcpyHostToDevice, stream));
    sc[s].d_kp = kp; sc[s].d_short = sh; sc[s].d_mid = mid;
  }
  Arena o(g_hd.d_out, g_hd.out_bytes);
  kg_decode_outputs out = {};
  out.d_dets = o.take<double>((size_t)5 * cfg->N * cfg->max_boxes);
  out.d_det_count = o.take<int>(cfg->N);
  out.d_status = o.take<int>(1);
  KG_TRY(kg::decode_launch(cfg, sc, &out, g_hd.d_ws, g_hd.ws_bytes, stream, nullptr));
  KG_CUDA_CHECK(cudaMemcpyAsync(h_dets, out.d_dets, det_bytes, cudaMemcpyDeviceToHost, stream));
  KG_CUDA_CHECK(cudaMemcpyAsync(h_det_count, out.d_det_count, sizeof(int) * cfg->N, cudaMemcpyDeviceToHost, stream));
  KG_CUDA_CHECK(cudaMemcpyAsync(g_hd.h_status, out.d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
  KG_CUDA_CHECK(cudaStreamSynchronize(stream));
  const int status = *g_hd.h_status;
  if (status != 0) {
    kg::set_error("kg_decode_host: device list overflow (status=%d): raise max_peaks/max_boxes", status);
    return KG_ERR_CAPACITY;
  }
  return KG_OK;
}

int kg_skeletons_to_boxes_host(const double* h_skeletons, int n, int box_scale, int apply_refine, uint8_t* h_keep,
                               double* h_boxes, int* n_boxes) {
  return kg::skeletons_to_boxes_host(h_skeletons, n, box_scale, apply_refine, h_keep, h_boxes, n_boxes);
}

int kg_nms_host(const double* h_boxes, int n, double nms_thresh, double* h_out, int* n_out) {
  return kg::nms_host(h_boxes, n, nms_thresh, h_out, n_out);
}

}  // extern "C"
