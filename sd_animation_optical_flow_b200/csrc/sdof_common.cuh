// Shared host/device helpers for libsdof_b200.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sdof_b200.h"

namespace sdof {

// ---- error plumbing (thread-local text behind sdof_last_error) -------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define SDOF_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::sdof::fail(SDOF_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define SDOF_CUDA(expr)                                                                            \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::sdof::fail(SDOF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                                     \
  } while (0)

// after a kernel launch: surface launch-configuration errors without syncing
#define SDOF_LAUNCH_CHECK(name)                                                                    \
  do {                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess)                                                                         \
      return ::sdof::fail(SDOF_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    ::sdof::count_launch();                                                                        \
  } while (0)

inline cudaStream_t as_stream(sdof_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached multiProcessorCount of the current device

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// grid sized for a grid-stride loop: enough CTAs for `work_items` but capped at a
// whole number of waves (multiple of the SM count).
inline int grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  int64_t need = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl may be scheduled while its predecessor in the
// stream (or CUDA-graph chain) is still draining; it must call pdl_wait() before touching anything the predecessor wrote
// (a no-op when the launch carries no PDL attribute).  Work that does not depend on the predecessor -- staging filters into
// shared memory, index arithmetic -- goes BEFORE pdl_wait() and overlaps the predecessor's tail.  pdl_trigger() lets the NEXT
// kernel's launch begin early in the same way.  SDOF_PDL=0 in the environment switches the attribute off (A/B).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
// L2 residency of a buffer read many times (the correlation pyramid: 20 lookups per pair): bytes of L2 this device sets aside
// for persisting lines (0 = unsupported or switched off with SDOF_L2_PERSIST=0; the limit is raised once per device).
size_t l2_persist_bytes();
struct L2Window {        // accesses to [ptr, ptr + bytes) keep their lines in the persisting part of L2
  const void* ptr = nullptr;
  size_t bytes = 0;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_win(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, L2Window win, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  const size_t persist = win.ptr ? l2_persist_bytes() : 0;
  if (persist > 0 && win.bytes > 0) {
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = const_cast<void*>(win.ptr);
    attr[n].val.accessPolicyWindow.num_bytes = win.bytes;
    attr[n].val.accessPolicyWindow.hitRatio = win.bytes <= persist ? 1.0f : (float)((double)persist / (double)win.bytes);
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl_win(kernel, grid, block, smem, st, L2Window{}, args...);
}

// FlowHead.conv2 as "taps first" (raft_glue.cu): y[p][t] (t = 3*(dy+1) + (dx+1), two floats each) = <x[p,:], w[t][:, :]>; the 3x3
// convolution at pixel p is bias + the sum over the in-image neighbours q = p + (dy, dx) of y[q][t].  ONE definition of that sum
// (order: bias, then t = 0..8) for the coords-update kernel and for the kernels that apply the update on the fly (lookup, convf1),
// so that all of them produce bit-identical coordinates.
#ifdef __CUDACC__
__device__ __forceinline__ float2 flowhead2_gather(const float* __restrict__ y, float2 bias, int64_t p, int yy, int xx, int h, int w) {
  float s0 = bias.x, s1 = bias.y;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ny = yy + t / 3 - 1, nx = xx + t % 3 - 1;
    if ((unsigned)ny < (unsigned)h && (unsigned)nx < (unsigned)w) {
      const float2 q = *reinterpret_cast<const float2*>(y + (p + (int64_t)(t / 3 - 1) * w + (t % 3 - 1)) * 18 + t * 2);
      s0 += q.x;
      s1 += q.y;
    }
  }
  return make_float2(s0, s1);
}
#endif

// host-side tables (cubic_table.cpp)
const int16_t* cubic_table_i16_host();  // [1024][16]
const float* cubic_table_f32_host();    // [1024][16]
void ellipse_half_widths_host(int ksize, int32_t* out);

}  // namespace sdof
