// Runtime: device context, stream-ordered caching allocator, copies, streams/events, CUDA graphs.
// Replaces what the reference gets from CuPy in backends/cuda_backend.py:24-58 (device + memory pool)
// and backends/cuda_kernels.py:490-526 (event timing).
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

static thread_local char g_err[1024] = "";
static int g_device = -1;
static int g_sm_count = 148;
static cudaStream_t g_stream = nullptr;  // current compute stream
static unsigned long long g_launches = 0;

int nfb_fail(const char* file, int line, const char* fmt, ...) {
  char msg[768];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", base ? base + 1 : file, line, msg);
  return -1;
}
cudaStream_t nfb_cur_stream() { return g_stream; }
void nfb_count_launch() { ++g_launches; }
int nfb_sm_count() { return g_sm_count; }

NFB_API const char* nfb_last_error() { return g_err; }
NFB_API const char* nfb_version() { return "nfb200 0.1 (sm_100a)"; }
NFB_API unsigned long long nfb_launch_count() { return g_launches; }

NFB_API int nfb_device_count(int* out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out = 0;
    cudaGetLastError();
    return NFB_FAIL("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
  }
  *out = n;
  return 0;
}

NFB_API int nfb_init(int device) {
  if (g_device == device && g_stream) return 0;
  int n = 0;
  NFB_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return NFB_FAIL("device %d out of range (have %d)", device, n);
  NFB_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  NFB_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return NFB_FAIL("libnfb200 is built for sm_100a only; device %d is sm_%d%d", device, p.major, p.minor);
  g_sm_count = p.multiProcessorCount;
  NFB_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  g_device = device;
  return 0;
}
NFB_API int nfb_current_device() { return g_device; }
NFB_API int nfb_get_sm_count() { return g_sm_count; }

NFB_API int nfb_device_name(char* buf, int len) {
  cudaDeviceProp p;
  NFB_CUDA(cudaGetDeviceProperties(&p, g_device < 0 ? 0 : g_device));
  snprintf(buf, len, "%s", p.name);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Caching allocator. All work is issued on one compute stream (plus a comm stream fenced by
// events), so a freed block can be handed out again immediately: the next writer is ordered
// after the last reader by the stream itself.
// ---------------------------------------------------------------------------------------------
static std::mutex g_mu;
static std::multimap<size_t, void*> g_free;             // size -> block
static std::unordered_map<void*, size_t> g_live;        // block -> size
static size_t g_bytes_live = 0, g_bytes_cached = 0, g_bytes_peak = 0;
static unsigned long long g_n_cuda_malloc = 0;

static size_t round_size(size_t n) {
  if (n == 0) n = 1;
  if (n < (1u << 20)) return (n + 511) & ~(size_t)511;
  return (n + ((2u << 20) - 1)) & ~(size_t)((2u << 20) - 1);
}

static void release_cached_locked() {
  for (auto& kv : g_free) cudaFree(kv.second);
  g_free.clear();
  g_bytes_cached = 0;
}

NFB_API int nfb_malloc(size_t nbytes, void** out) {
  if (g_device < 0) return NFB_FAIL("nfb_init was not called");
  size_t sz = round_size(nbytes);
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_free.lower_bound(sz);
  // accept a cached block when it wastes < 1/8 of its size
  if (it != g_free.end() && it->first - sz <= (it->first >> 3)) {
    *out = it->second;
    g_bytes_cached -= it->first;
    g_live[*out] = it->first;
    g_bytes_live += it->first;
    g_free.erase(it);
  } else {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cudaDeviceSynchronize();
      release_cached_locked();
      e = cudaMalloc(&p, sz);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return NFB_FAIL("cudaMalloc(%zu) -> %s (live %zu B)", sz, cudaGetErrorString(e), g_bytes_live);
      }
    }
    ++g_n_cuda_malloc;
    *out = p;
    g_live[p] = sz;
    g_bytes_live += sz;
  }
  if (g_bytes_live > g_bytes_peak) g_bytes_peak = g_bytes_live;
  return 0;
}

NFB_API int nfb_free(void* p) {
  if (!p) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_live.find(p);
  if (it == g_live.end()) return NFB_FAIL("nfb_free: unknown pointer %p", p);
  size_t sz = it->second;
  g_live.erase(it);
  g_bytes_live -= sz;
  g_free.emplace(sz, p);
  g_bytes_cached += sz;
  return 0;
}

NFB_API int nfb_empty_cache() {
  std::lock_guard<std::mutex> lk(g_mu);
  cudaDeviceSynchronize();
  release_cached_locked();
  return 0;
}

// stats[0]=live bytes, [1]=cached bytes, [2]=peak live, [3]=#cudaMalloc calls
NFB_API int nfb_mem_stats(unsigned long long* stats) {
  std::lock_guard<std::mutex> lk(g_mu);
  stats[0] = g_bytes_live;
  stats[1] = g_bytes_cached;
  stats[2] = g_bytes_peak;
  stats[3] = g_n_cuda_malloc;
  return 0;
}

NFB_API int nfb_host_alloc(size_t nbytes, void** out) {
  NFB_CUDA(cudaMallocHost(out, nbytes ? nbytes : 1));
  return 0;
}
NFB_API int nfb_host_free(void* p) {
  NFB_CUDA(cudaFreeHost(p));
  return 0;
}

// ---- copies (async on the current stream; *_sync variants block the host) ----
NFB_API int nfb_memcpy_h2d(void* dst, const void* src, size_t n) {
  if (!n) return 0;
  NFB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, g_stream));
  NFB_CUDA(cudaStreamSynchronize(g_stream));  // src may be pageable / freed by the caller
  return 0;
}
NFB_API int nfb_memcpy_h2d_async(void* dst, const void* src, size_t n) {
  if (!n) return 0;
  NFB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, g_stream));
  return 0;
}
NFB_API int nfb_memcpy_d2h(void* dst, const void* src, size_t n) {
  if (!n) return 0;
  NFB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, g_stream));
  NFB_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}
NFB_API int nfb_memcpy_d2h_async(void* dst, const void* src, size_t n) {
  if (!n) return 0;
  NFB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, g_stream));
  return 0;
}
NFB_API int nfb_memcpy_d2d(void* dst, const void* src, size_t n) {
  if (!n) return 0;
  NFB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, g_stream));
  return 0;
}
NFB_API int nfb_memset(void* dst, int value, size_t n) {
  if (!n) return 0;
  NFB_CUDA(cudaMemsetAsync(dst, value, n, g_stream));
  return 0;
}

// ---- streams and events ----
NFB_API int nfb_stream_create(void** out) {
  cudaStream_t s;
  NFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *out = (void*)s;
  return 0;
}
NFB_API int nfb_stream_destroy(void* s) {
  NFB_CUDA(cudaStreamDestroy((cudaStream_t)s));
  return 0;
}
NFB_API void* nfb_get_stream() { return (void*)g_stream; }
// auxiliary stream for work that only the optimizer consumes (weight / bias / norm-parameter gradients); NULL = none.
// g_capture_epoch changes at every graph-capture boundary: cross-stream events recorded on one side of a boundary must
// not be waited on from the other side.
static cudaStream_t g_aux_stream = nullptr;
static unsigned long long g_capture_epoch = 0;
cudaStream_t nfb_aux_stream() { return g_aux_stream; }
unsigned long long nfb_capture_epoch() { return g_capture_epoch; }
NFB_API int nfb_set_aux_stream(void* s) { g_aux_stream = (cudaStream_t)s; return 0; }
NFB_API int nfb_set_stream(void* s) {
  g_stream = (cudaStream_t)s;
  return 0;
}
NFB_API int nfb_stream_sync(void* s) {
  NFB_CUDA(cudaStreamSynchronize(s ? (cudaStream_t)s : g_stream));
  return 0;
}
NFB_API int nfb_device_sync() {
  NFB_CUDA(cudaDeviceSynchronize());
  return 0;
}
NFB_API int nfb_event_create(void** out) {
  cudaEvent_t e;
  NFB_CUDA(cudaEventCreate(&e));
  *out = (void*)e;
  return 0;
}
NFB_API int nfb_event_create_notiming(void** out) {
  cudaEvent_t e;
  NFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out = (void*)e;
  return 0;
}
NFB_API int nfb_event_destroy(void* e) {
  NFB_CUDA(cudaEventDestroy((cudaEvent_t)e));
  return 0;
}
NFB_API int nfb_event_record(void* e, void* stream) {
  NFB_CUDA(cudaEventRecord((cudaEvent_t)e, stream ? (cudaStream_t)stream : g_stream));
  return 0;
}
// record as an EXTERNAL event node when the stream is being captured: such an event keeps its timestamp of the last
// graph replay and can be read with nfb_event_elapsed_ms (a plainly captured event cannot)
NFB_API int nfb_event_record_external(void* e, void* stream) {
  cudaStream_t s = stream ? (cudaStream_t)stream : g_stream;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  NFB_CUDA(cudaStreamIsCapturing(s, &st));
  if (st == cudaStreamCaptureStatusActive) NFB_CUDA(cudaEventRecordWithFlags((cudaEvent_t)e, s, cudaEventRecordExternal));
  else NFB_CUDA(cudaEventRecord((cudaEvent_t)e, s));
  return 0;
}
NFB_API int nfb_event_sync(void* e) {
  NFB_CUDA(cudaEventSynchronize((cudaEvent_t)e));
  return 0;
}
NFB_API int nfb_event_elapsed_ms(void* a, void* b, float* ms) {
  NFB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return 0;
}
NFB_API int nfb_stream_wait_event(void* stream, void* e) {
  NFB_CUDA(cudaStreamWaitEvent(stream ? (cudaStream_t)stream : g_stream, (cudaEvent_t)e, 0));
  return 0;
}

// ---- CUDA graphs: capture the launch-bound steady-state step once, replay it ----
NFB_API int nfb_graph_begin_capture() {
  ++g_capture_epoch;
  NFB_CUDA(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeRelaxed));
  return 0;
}
NFB_API int nfb_graph_end_capture(void** out_exec) {
  ++g_capture_epoch;
  cudaGraph_t g;
  NFB_CUDA(cudaStreamEndCapture(g_stream, &g));
  cudaGraphExec_t ex;
  cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return NFB_FAIL("cudaGraphInstantiate -> %s", cudaGetErrorString(e));
  *out_exec = (void*)ex;
  return 0;
}
NFB_API int nfb_graph_launch(void* exec) {
  NFB_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, g_stream));
  return 0;
}
NFB_API int nfb_graph_destroy(void* exec) {
  NFB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec));
  return 0;
}

// programmatic dependent launch switch (default on; env NFB_PDL=0 or nfb_set_pdl(0) turns it off)
static int g_pdl = -1;
bool nfb_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("NFB_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl == 1;
}
NFB_API int nfb_set_pdl(int on) { g_pdl = on ? 1 : 0; return 0; }

// L2 flush helper for benchmarks: writes a buffer larger than the 126 MB L2, then streams a second one through it with
// loads, so that the cache is left full of CLEAN foreign lines (a write-only flush leaves 126 MB of dirty lines whose
// write-back is then billed to the kernel under test)
static void* g_flush = nullptr;
static const size_t kFlushBytes = 256u << 20;
__global__ void k_flush_read(const uint4* __restrict__ p, size_t n, unsigned* sink) {
  unsigned acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x9e3779b9u) *sink = acc;   // never true for the zero-filled buffer; keeps the loads alive
}
NFB_API int nfb_flush_l2() {
  if (!g_flush) {
    NFB_CUDA(cudaMalloc(&g_flush, 2 * kFlushBytes + 256));
    NFB_CUDA(cudaMemsetAsync(g_flush, 0, 2 * kFlushBytes + 256, g_stream));
  }
  NFB_CUDA(cudaMemsetAsync(g_flush, 0, kFlushBytes, g_stream));
  k_flush_read<<<nfb_sm_count() * 8, 256, 0, g_stream>>>((const uint4*)((char*)g_flush + kFlushBytes), kFlushBytes / 16,
                                                        (unsigned*)((char*)g_flush + 2 * kFlushBytes));
  NFB_CUDA(cudaGetLastError());
  return 0;
}
