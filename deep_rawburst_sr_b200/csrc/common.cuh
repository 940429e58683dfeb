// Shared helpers for libdbsr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/dbsr_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdbsr_b200 is written for sm_100a (B200) only"
#endif

namespace dbsr {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define DBSR_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      dbsr::set_error(__VA_ARGS__);        \
      return 1;                            \
    }                                      \
  } while (0)

// ---- element access -----------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DBSR_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == DBSR_ACT_LRELU) return v > 0.0f ? v : 0.1f * v;
  return v;
}

// device-side copy of a view (passed by value in kernel params)
struct View {
  void* data;
  int n, h, w, c, c_off, c_pitch, dtype;
};
inline View make_view(const dbsr_nhwc_t* v) {
  View r;
  if (v == nullptr) {
    r.data = nullptr; r.n = r.h = r.w = r.c = r.c_off = r.c_pitch = r.dtype = 0;
    return r;
  }
  r.data = v->data; r.n = v->n; r.h = v->h; r.w = v->w; r.c = v->c; r.c_off = v->c_off;
  r.c_pitch = v->c_pitch; r.dtype = v->dtype;
  return r;
}
inline bool view_ok(const dbsr_nhwc_t* v) {
  return v && v->data && v->n > 0 && v->h > 0 && v->w > 0 && v->c > 0 && v->c_off >= 0 &&
         v->c_off + v->c <= v->c_pitch && (v->dtype == DBSR_F32 || v->dtype == DBSR_BF16);
}
inline size_t elem_size(int dtype) { return dtype == DBSR_F32 ? 4 : 2; }

__device__ __forceinline__ float view_ld(const View& v, long long pix, int ch) {
  long long off = pix * v.c_pitch + v.c_off + ch;
  if (v.dtype == DBSR_F32) return __ldg(reinterpret_cast<const float*>(v.data) + off);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(v.data)[off]);
}
__device__ __forceinline__ void view_st(const View& v, long long pix, int ch, float val) {
  long long off = pix * v.c_pitch + v.c_off + ch;
  if (v.dtype == DBSR_F32) reinterpret_cast<float*>(v.data)[off] = val;
  else reinterpret_cast<__nv_bfloat16*>(v.data)[off] = __float2bfloat16_rn(val);
}

// ---- programmatic dependent launch ---------------------------------------------------------------------
// Kernels launched through launch_pdl may be scheduled while the preceding kernel of the stream is still draining
// (hides launch latency between the ~140 dependent launches of one forward); they call griddep_wait() before their
// first global-memory access.  griddep_wait() is a no-op for a normally launched kernel.
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Early trigger: once every CTA of this grid has executed it, the CTAs of the NEXT kernel of the stream may be scheduled on
// whatever SMs (and shared memory) are free; they run their prologue -- and, for static weights, their weight fetches -- and
// then block in their own griddep_wait() until this grid has completed and flushed.  Ordering of data is unchanged.
// Only worth it -- and only harmless -- when this grid leaves SMs idle: early dependents hold their shared memory while they
// wait, which on a full GPU takes SMs from the kernels of a concurrent stream (measured: -5 % at 32 bursts when every kernel
// triggered early).  The tensor-core kernels get the decision from the host (ConvTcParams::early_trigger); the small kernels
// use griddep_launch_dependents_if_small().
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents_if_small() {
  if (gridDim.x * gridDim.y * gridDim.z <= 64u) griddep_launch_dependents();
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(args)...);     // errors surface through check_launch()
}
#endif

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Host-side caches (function attributes already set, SM counts, identity tiles) are kept PER DEVICE: a process may drive
// several GPUs (nn.DataParallel style), and cudaFuncSetAttribute / cudaMalloc act on the current device only.
constexpr int MAX_DEVICES = 64;
inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
  return d;
}

}  // namespace dbsr
