// Shared declarations for libnfb200.so (sm_100a only).
// The library is a C-ABI replacement for the array kernels behind the reference's
// neural_arch.backends.Backend plugin (backends/backend.py:9-322); see include/nfb200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#define NFB_API extern "C" __attribute__((visibility("default")))

// dtype codes shared with the Python host (nfb200/_lib.py)
enum NfbDType { NFB_F32 = 0, NFB_F64 = 1, NFB_I32 = 2, NFB_I64 = 3, NFB_BOOL = 4, NFB_BF16 = 5 };

#define NFB_MAX_DIMS 8

// ---- error plumbing: every entry point returns 0 on success, <0 on failure ----
int nfb_fail(const char* file, int line, const char* fmt, ...);
#define NFB_FAIL(...) nfb_fail(__FILE__, __LINE__, __VA_ARGS__)
#define NFB_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      (void)cudaGetLastError(); /* reported here: do not let it surface at the next launch check */ \
      return NFB_FAIL("%s -> %s", #expr, cudaGetErrorString(_e));                   \
    }                                                                               \
  } while (0)
// after a kernel launch
#define NFB_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    nfb_count_launch();                                                             \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) return NFB_FAIL("kernel launch -> %s", cudaGetErrorString(_e)); \
  } while (0)

cudaStream_t nfb_cur_stream();
cudaStream_t nfb_aux_stream();          // side stream for optimizer-only gradients (NULL = none)
unsigned long long nfb_capture_epoch();  // bumps at graph-capture begin / end
void nfb_count_launch();
int nfb_sm_count();
int* nfb_index_error_flag();   // device int set to 1 by a kernel that met an out-of-range index / class id (NULL on allocation failure)
extern "C" int nfb_malloc(size_t nbytes, void** out);
extern "C" int nfb_free(void* p);

static inline int64_t nfb_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL).  Kernels of the training step are launched with the programmatic stream
// serialization attribute; each starts with nfb_pdl_wait() (= griddepcontrol.wait: blocks until the preceding kernel in
// the stream has completed and its writes are visible) followed by nfb_pdl_trigger() (lets the NEXT kernel's blocks be
// scheduled early, where they sit in their own nfb_pdl_wait()).  The launch latency and block ramp-up of kernel N+1 then
// overlap the last wave of kernel N; with ~350 short launches per step that is a measurable share of the step.
bool nfb_pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t nfb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = nfb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// grid sizing for bandwidth-bound grid-stride kernels: a multiple of the SM count
static inline int nfb_grid_for(int64_t work_items, int per_block, int waves_cap = 16) {
  int64_t blocks = nfb_cdiv(work_items, per_block);
  int64_t cap = (int64_t)nfb_sm_count() * waves_cap;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- device helpers ----
#ifdef __CUDACC__
__device__ __forceinline__ void nfb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void nfb_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// both, at the top of a kernel that has no prologue worth overlapping
__device__ __forceinline__ void nfb_pdl_enter() { nfb_pdl_wait(); nfb_pdl_trigger(); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; `red` must hold >= 32 floats; every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// generic load/store between storage type and float (activations may be fp32 or bf16)
template <typename T> __device__ __forceinline__ float ld_f(const T* p, int64_t i);
template <> __device__ __forceinline__ float ld_f<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}
template <typename T> __device__ __forceinline__ void st_f(T* p, int64_t i, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, int64_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

// 4-wide vector access for fp32 / bf16 rows (16 B / 8 B per access)
struct f4 { float x, y, z, w; };
template <typename T> __device__ __forceinline__ f4 ld4(const T* p);
template <> __device__ __forceinline__ f4 ld4<float>(const float* p) {
  float4 v = *reinterpret_cast<const float4*>(p);
  return {v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ f4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return {fa.x, fa.y, fb.x, fb.y};
}
template <typename T> __device__ __forceinline__ void st4(T* p, f4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, f4 v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, f4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
#endif
