// Shared helpers for the kgnet_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include "../../include/kgnet_b200.h"

namespace kg {

// ---- error reporting (C-ABI: int return codes + kg_last_error()) ------------------------------
// status codes: kg_status in include/kgnet_b200.h (KG_OK, KG_ERR_*), shared with the C-ABI

void set_error(const char* fmt, ...);
const char* last_error();

#define KG_CUDA_CHECK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      kg::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return KG_ERR_CUDA;                                                                    \
    }                                                                                            \
  } while (0)

#define KG_REQUIRE(cond, ...)                  \
  do {                                         \
    if (!(cond)) {                             \
      kg::set_error(__VA_ARGS__);              \
      return KG_ERR_INVALID;               \
    }                                          \
  } while (0)

#define KG_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != KG_OK) return _r; \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base; size_t size; size_t off;
  Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <class T> T* take(size_t n) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return off <= size; }
};

// ---- optional per-stage device timing (kg_timing_enable / kg_timing_collect) ---------------------
// Stage ids: 0 vote, 1 blur+peak, 2 sort+group+boxes, 3 nms, 8.. network stages (see net.cu).
constexpr int KG_MAX_STAGES = 256;
bool timing_enabled();
void stage_begin(int id, cudaStream_t s);
void stage_end(int id, cudaStream_t s);
struct StageScope {
  int id; cudaStream_t s;
  StageScope(int id_, cudaStream_t s_) : id(id_), s(s_) { if (timing_enabled()) stage_begin(id, s); }
  void restage(int new_id);   // re-attribute the open interval to another stage
  ~StageScope() { if (timing_enabled()) stage_end(id, s); }
};

// ---- split-bf16 activations: v ~= float(hi) + float(lo) ---------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}

}  // namespace kg
