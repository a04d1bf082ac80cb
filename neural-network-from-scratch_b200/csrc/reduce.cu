// Reductions over a [outer, R, inner] view (the host permutes/reshapes so the reduced axes are the
// middle block): sum / mean / max / min / prod, argmax / argmin, and the squared-L2 norm used by
// gradient clipping. Semantics follow numpy_backend.py:138-160 of the reference (np.sum/mean/max/
// min/argmax/argmin); NaNs propagate through max/min as in NumPy.
#include "common.cuh"
#include <map>
#include <utility>
#include <float.h>

enum RedOp { R_SUM = 0, R_MEAN = 1, R_MAX = 2, R_MIN = 3, R_PROD = 4 };

template <typename A> __device__ __forceinline__ A red_identity(int op);
template <> __device__ __forceinline__ float red_identity<float>(int op) {
  return op == R_MAX ? -INFINITY : (op == R_MIN ? INFINITY : (op == R_PROD ? 1.f : 0.f));
}
template <> __device__ __forceinline__ double red_identity<double>(int op) {
  return op == R_MAX ? -INFINITY : (op == R_MIN ? INFINITY : (op == R_PROD ? 1.0 : 0.0));
}
template <> __device__ __forceinline__ long long red_identity<long long>(int op) {
  return op == R_MAX ? LLONG_MIN : (op == R_MIN ? LLONG_MAX : (op == R_PROD ? 1 : 0));
}
template <typename A> __device__ __forceinline__ A red_combine(int op, A a, A b) {
  switch (op) {
    case R_MAX: return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
    case R_MIN: return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
    case R_PROD: return a * b;
    default: return a + b;
  }
}
template <typename A> __device__ __forceinline__ A shfl_xor_t(A v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

template <typename A> __device__ __forceinline__ A warp_red(int op, A v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_combine<A>(op, v, shfl_xor_t<A>(v, o));
  return v;
}

// inner == 1: one block per (outer, chunk) row segment
template <typename T, typename A, typename O>
__global__ void k_reduce_rows(int op, const T* __restrict__ x, O* __restrict__ out, int64_t R, int64_t chunk_len, int chunks, double post_scale) {
  __shared__ A red[32];
  int64_t row = blockIdx.x / chunks, ch = blockIdx.x % chunks;
  int64_t lo = ch * chunk_len, hi = lo + chunk_len < R ? lo + chunk_len : R;
  const T* p = x + row * R;
  A acc = red_identity<A>(op);
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = red_combine<A>(op, acc, (A)p[i]);
  acc = warp_red<A>(op, acc);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[w] = acc;
  __syncthreads();
  if (w == 0) {
    A r = lane < nw ? red[lane] : red_identity<A>(op);
    r = warp_red<A>(op, r);
    if (lane == 0) out[blockIdx.x] = (O)(post_scale == 1.0 ? r : (A)(r * (A)post_scale));
  }
}

// inner > 1: block = 32 columns x 8 row-lanes; grid = (ceil(inner/32), chunks, outer)
template <typename T, typename A, typename O>
__global__ void k_reduce_cols(int op, const T* __restrict__ x, O* __restrict__ out, int64_t R, int64_t inner, int64_t chunk_len, int chunks, double post_scale) {
  __shared__ A red[8][33];
  int64_t o = blockIdx.z, ch = blockIdx.y;
  int64_t col = (int64_t)blockIdx.x * 32 + threadIdx.x;
  int64_t lo = ch * chunk_len, hi = lo + chunk_len < R ? lo + chunk_len : R;
  A acc = red_identity<A>(op);
  if (col < inner) {
    const T* p = x + o * R * inner + col;
    for (int64_t r = lo + threadIdx.y; r < hi; r += 8) acc = red_combine<A>(op, acc, (A)p[r * inner]);
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < inner) {
    A r = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) r = red_combine<A>(op, r, red[k][threadIdx.x]);
    out[(o * chunks + ch) * inner + col] = (O)(post_scale == 1.0 ? r : (A)(r * (A)post_scale));
  }
}

template <typename T, typename A>
static int reduce_typed(int op, const T* x, T* out, int64_t outer, int64_t R, int64_t inner) {
  cudaStream_t s = nfb_cur_stream();
  int sms = nfb_sm_count();
  double scale = (op == R_MEAN) ? 1.0 / (double)R : 1.0;
  int kop = (op == R_MEAN) ? R_SUM : op;
  if (inner == 1) {
    int chunks = 1;
    if (outer < 2 * sms && R >= 1 << 16) {
      chunks = (int)((4 * (int64_t)sms + outer - 1) / outer);
      int64_t max_chunks = R / 4096;
      if (chunks > max_chunks) chunks = (int)max_chunks;
      if (chunks < 1) chunks = 1;
    }
    int threads = R / chunks >= 2048 ? 512 : (R / chunks >= 256 ? 256 : 64);
    if (chunks == 1) {
      k_reduce_rows<T, A, T><<<(unsigned)outer, threads, 0, s>>>(kop, x, out, R, R, 1, scale);
      NFB_LAUNCH_CHECK();
      return 0;
    }
    int64_t chunk_len = nfb_cdiv(R, chunks);
    A* tmp = nullptr;
    if (nfb_malloc(sizeof(A) * outer * chunks, (void**)&tmp)) return -1;
    k_reduce_rows<T, A, A><<<(unsigned)(outer * chunks), threads, 0, s>>>(kop, x, tmp, R, chunk_len, chunks, 1.0);
    nfb_count_launch();
    k_reduce_rows<A, A, T><<<(unsigned)outer, 64, 0, s>>>(kop, tmp, out, chunks, chunks, 1, scale);
    nfb_free(tmp);
    NFB_LAUNCH_CHECK();
    return 0;
  }
  if (outer > 65535) return NFB_FAIL("nfb_reduce: outer extent %lld too large for the column kernel", (long long)outer);
  int64_t col_blocks = nfb_cdiv(inner, 32);
  int chunks = 1;
  if (col_blocks * outer < 2 * sms && R >= 1024) {
    chunks = (int)((4 * (int64_t)sms) / (col_blocks * outer));
    int64_t max_chunks = R / 64;
    if (chunks > max_chunks) chunks = (int)max_chunks;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
  }
  dim3 block(32, 8);
  if (chunks == 1) {
    k_reduce_cols<T, A, T><<<dim3((unsigned)col_blocks, 1, (unsigned)outer), block, 0, s>>>(kop, x, out, R, inner, R, 1, scale);
    NFB_LAUNCH_CHECK();
    return 0;
  }
  int64_t chunk_len = nfb_cdiv(R, chunks);
  A* tmp = nullptr;
  if (nfb_malloc(sizeof(A) * outer * chunks * inner, (void**)&tmp)) return -1;
  k_reduce_cols<T, A, A><<<dim3((unsigned)col_blocks, (unsigned)chunks, (unsigned)outer), block, 0, s>>>(kop, x, tmp, R, inner, chunk_len, chunks, 1.0);
  nfb_count_launch();
  k_reduce_cols<A, A, T><<<dim3((unsigned)col_blocks, 1, (unsigned)outer), block, 0, s>>>(kop, tmp, out, chunks, inner, chunks, 1, scale);
  nfb_free(tmp);
  NFB_LAUNCH_CHECK();
  return 0;
}

NFB_API int nfb_reduce(int op, int dtype, const void* x, void* out, int64_t outer, int64_t R, int64_t inner) {
  if (outer * inner == 0) return 0;
  if (R == 0) return NFB_FAIL("nfb_reduce: zero-size reduction axis");
  switch (dtype) {
    case NFB_F32: return reduce_typed<float, float>(op, (const float*)x, (float*)out, outer, R, inner);
    case NFB_F64: return reduce_typed<double, double>(op, (const double*)x, (double*)out, outer, R, inner);
    case NFB_I64: return reduce_typed<long long, long long>(op, (const long long*)x, (long long*)out, outer, R, inner);
  }
  return NFB_FAIL("nfb_reduce: unsupported dtype %d (cast to f32/f64/i64 first)", dtype);
}

// ---- argmax / argmin: first occurrence wins, NaN wins over everything (NumPy semantics) ----
template <typename T>
__device__ __forceinline__ bool arg_better(bool is_max, T v, int64_t i, T bv, int64_t bi) {
  bool vnan = (v != v), bnan = (bv != bv);
  if (bnan) return vnan && i < bi;  // keep the first NaN
  if (vnan) return true;
  if (v == bv) return i < bi;
  return is_max ? (v > bv) : (v < bv);
}
template <typename T>
__global__ void k_argreduce(int is_max, const T* __restrict__ x, int64_t* __restrict__ out, int64_t R, int64_t inner) {
  __shared__ T sv[256];
  __shared__ int64_t si[256];
  int64_t o = blockIdx.x / inner, c = blockIdx.x % inner;
  const T* p = x + o * R * inner + c;
  T bv = p[0];
  int64_t bi = 0;
  for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {
    T v = p[r * inner];
    if (arg_better<T>(is_max, v, r, bv, bi)) { bv = v; bi = r; }
  }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      if (arg_better<T>(is_max, sv[threadIdx.x + s], si[threadIdx.x + s], sv[threadIdx.x], si[threadIdx.x])) {
        sv[threadIdx.x] = sv[threadIdx.x + s];
        si[threadIdx.x] = si[threadIdx.x + s];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = si[0];
}
NFB_API int nfb_argreduce(int is_max, int dtype, const void* x, int64_t* out, int64_t outer, int64_t R, int64_t inner) {
  if (outer * inner == 0) return 0;
  if (R == 0) return NFB_FAIL("nfb_argreduce: attempt to get argmax of an empty sequence");
  int64_t blocks = outer * inner;
  if (blocks > 2147483647LL) return NFB_FAIL("nfb_argreduce: too many output elements");
  int threads = R >= 256 ? 256 : (R >= 64 ? 64 : 32);
  cudaStream_t s = nfb_cur_stream();
  switch (dtype) {
    case NFB_F32: k_argreduce<float><<<(unsigned)blocks, threads, 0, s>>>(is_max, (const float*)x, out, R, inner); break;
    case NFB_F64: k_argreduce<double><<<(unsigned)blocks, threads, 0, s>>>(is_max, (const double*)x, out, R, inner); break;
    case NFB_I32: k_argreduce<int32_t><<<(unsigned)blocks, threads, 0, s>>>(is_max, (const int32_t*)x, out, R, inner); break;
    case NFB_I64: k_argreduce<int64_t><<<(unsigned)blocks, threads, 0, s>>>(is_max, (const int64_t*)x, out, R, inner); break;
    case NFB_BOOL: k_argreduce<uint8_t><<<(unsigned)blocks, threads, 0, s>>>(is_max, (const uint8_t*)x, out, R, inner); break;
    default: return NFB_FAIL("nfb_argreduce: unsupported dtype %d", dtype);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}

// ---- out[c] (+)= sum_r x[r, c] for a row-major (R, C) matrix with pitch ld: bias gradients (functional/utils.py:33-96
// reduce_gradient for a (T, n) -> (n) broadcast).  fp32 or bf16 input, fp32 output; two deterministic stages.
template <typename T>
__global__ void k_colsum_stage1(const T* __restrict__ x, int64_t R, int C, int64_t ld, int64_t chunk_len, float* __restrict__ tmp) {
  __shared__ float red[8][65];
  int c0 = blockIdx.x * 64 + threadIdx.x * 2;
  int64_t lo = (int64_t)blockIdx.y * chunk_len, hi = lo + chunk_len < R ? lo + chunk_len : R;
  float a0 = 0.f, a1 = 0.f;
  if (c0 + 1 < C) {
    for (int64_t r = lo + threadIdx.y; r < hi; r += 8) {
      a0 += ld_f<T>(x, r * ld + c0);
      a1 += ld_f<T>(x, r * ld + c0 + 1);
    }
  } else if (c0 < C) {
    for (int64_t r = lo + threadIdx.y; r < hi; r += 8) a0 += ld_f<T>(x, r * ld + c0);
  }
  red[threadIdx.y][threadIdx.x * 2] = a0;
  red[threadIdx.y][threadIdx.x * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.y < 2) {
    int cc = threadIdx.x * 2 + threadIdx.y;
    int col = blockIdx.x * 64 + cc;
    if (col < C) {
      float r = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) r += red[k][cc];
      tmp[(int64_t)blockIdx.y * C + col] = r;
    }
  }
}
__global__ void k_colsum_stage2(const float* __restrict__ tmp, int chunks, int C, float* __restrict__ out, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int b = 0; b < chunks; ++b) a += tmp[(int64_t)b * C + c];
  out[c] = accumulate ? out[c] + a : a;
}
// single-launch variant: 16-byte loads (8 bf16 / 4 fp32 columns per thread), block (32, 8) sums a row chunk of its
// 32 x VEC columns, and the LAST block to finish a column strip (device counter, __threadfence) adds the chunk partials
// in a fixed order -- deterministic, one kernel instead of two.
template <typename T>
__global__ void __launch_bounds__(256)
k_colsum_fused(const T* __restrict__ x, int64_t R, int C, int64_t ld, int64_t chunk_len, float* __restrict__ tmp, unsigned* __restrict__ counters,
               float* __restrict__ out, int accumulate) {
  nfb_pdl_enter();
  constexpr int VEC = 16 / (int)sizeof(T);
  __shared__ float red[8][32 * VEC + 1];
  __shared__ int s_last;
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int64_t lo = (int64_t)blockIdx.y * chunk_len, hi = lo + chunk_len < R ? lo + chunk_len : R;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  if (c0 < C) {
    int64_t r = lo + threadIdx.y;
    for (; r + 56 < hi; r += 64) {   // eight independent 16-byte loads in flight per thread
      uint4 u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = *reinterpret_cast<const uint4*>(x + (r + 8 * j) * ld + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (sizeof(T) == 4) {
          const float* f = reinterpret_cast<const float*>(&u[j]);
#pragma unroll
          for (int k = 0; k < VEC; ++k) acc[k] += f[k];
        } else {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[j]);
#pragma unroll
          for (int k = 0; k < VEC / 2; ++k) { float2 f = __bfloat1622float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
        }
      }
    }
    for (; r < hi; r += 8) {
      uint4 u0 = *reinterpret_cast<const uint4*>(x + r * ld + c0);
      if (sizeof(T) == 4) {
        const float* f = reinterpret_cast<const float*>(&u0);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] += f[k];
      } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u0);
#pragma unroll
        for (int k = 0; k < VEC / 2; ++k) { float2 f = __bfloat1622float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) red[threadIdx.y][threadIdx.x * VEC + k] = acc[k];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int cc = tid; cc < 32 * VEC; cc += 256) {
    const int col = blockIdx.x * 32 * VEC + cc;
    if (col < C) {
      float r = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) r += red[k][cc];
      tmp[(int64_t)blockIdx.y * C + col] = r;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&counters[blockIdx.x], 1u) == gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // fixed-order sum of the chunk partials: thread (tx, ty) adds the partials ty, ty + 8, ... of its VEC columns (independent 16-byte
  // loads, all in flight together), then the eight ty-sums meet in shared memory in ty order.  (One column per thread with 64
  // partials, eight in flight, was eight dependent L2 round trips: ~5 us of an 11 us kernel, profiles/r02_ncu_attn_norm_colsum.txt.)
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  if (c0 < C) {
#pragma unroll 8
    for (unsigned b = threadIdx.y; b < gridDim.y; b += 8) {
      const float4* tp = reinterpret_cast<const float4*>(tmp + (int64_t)b * C + c0);
#pragma unroll
      for (int k4 = 0; k4 < VEC / 4; ++k4) {
        const float4 t = __ldcg(tp + k4);
        acc[4 * k4] += t.x; acc[4 * k4 + 1] += t.y; acc[4 * k4 + 2] += t.z; acc[4 * k4 + 3] += t.w;
      }
    }
  }
  __syncthreads();   // (every thread of the last block is here: s_last is block-uniform)
#pragma unroll
  for (int k = 0; k < VEC; ++k) red[threadIdx.y][threadIdx.x * VEC + k] = acc[k];
  __syncthreads();
  for (int cc = tid; cc < 32 * VEC; cc += 256) {
    const int col = blockIdx.x * 32 * VEC + cc;
    if (col < C) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[k][cc];
      out[col] = accumulate ? out[col] + a : a;
    }
  }
  if (tid == 0) counters[blockIdx.x] = 0;   // ready for the next launch on this stream
}

NFB_API int nfb_colsum(int dtype, const void* x, int64_t R, int C, int64_t ld, float* out, int accumulate) {
  if (C == 0) return 0;
  cudaStream_t s = nfb_cur_stream();
  const int esz = dtype == NFB_F32 ? 4 : 2, vec = 16 / esz;
  if ((dtype == NFB_F32 || dtype == NFB_BF16) && C % vec == 0 && ld % vec == 0 && ((uintptr_t)x % 16 == 0) && R >= 64 && nfb_cdiv(C, 32 * vec) <= 4096) {
    int strips = (int)nfb_cdiv(C, 32 * vec);
    // ~4 blocks per SM, and a chunk length that is a MULTIPLE OF 64 rows: the kernel's fast loop takes 64 rows per trip (8 loads
    // in flight per thread); a 125-row chunk (296 blocks over 4096 rows) ran its second half one load at a time -- 10.0 us for
    // (4096, 2304) bf16, profiles/r02_bw_kernels.txt
    int chunks = (int)((4LL * nfb_sm_count() + strips - 1) / strips);
    int64_t max_chunks = R / 64 > 0 ? R / 64 : 1;
    if (chunks > max_chunks) chunks = (int)max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t chunk_len = (nfb_cdiv(R, chunks) + 63) / 64 * 64;
    chunks = (int)nfb_cdiv(R, chunk_len);
    // chunk partials live in a grow-only workspace PER STREAM (not in the stream-ordered caching allocator): bias gradients
    // may be computed on the weight-gradient side stream, where an allocator block freed "in stream order" of the compute
    // stream could be recycled while this kernel still runs
    static std::map<cudaStream_t, std::pair<float*, size_t>> ws;
    static std::map<cudaStream_t, unsigned*> ws_counters;
    std::pair<float*, size_t>& w = ws[s];
    const size_t need = (size_t)chunks * C;
    if (w.second < need) {
      if (w.first) NFB_CUDA(cudaFree(w.first));   // implicit device sync: earlier users have finished
      NFB_CUDA(cudaMalloc((void**)&w.first, sizeof(float) * need));
      w.second = need;
    }
    unsigned*& cnt = ws_counters[s];
    if (!cnt) {
      NFB_CUDA(cudaMalloc((void**)&cnt, 4096 * sizeof(unsigned)));
      NFB_CUDA(cudaMemset(cnt, 0, 4096 * sizeof(unsigned)));
    }
    float* tmp = w.first;
    dim3 grid(strips, chunks), block(32, 8);
    if (dtype == NFB_F32) nfb_launch_pdl(k_colsum_fused<float>, grid, block, 0, s, (const float*)x, R, C, ld, chunk_len, tmp, cnt, out, accumulate);
    else nfb_launch_pdl(k_colsum_fused<__nv_bfloat16>, grid, block, 0, s, (const __nv_bfloat16*)x, R, C, ld, chunk_len, tmp, cnt, out, accumulate);
    NFB_LAUNCH_CHECK();
    return 0;
  }
  int col_blocks = (C + 63) / 64;
  int chunks = (int)((4LL * nfb_sm_count() + col_blocks - 1) / col_blocks);
  int64_t max_chunks = R / 32 > 0 ? R / 32 : 1;
  if (chunks > max_chunks) chunks = (int)max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t chunk_len = nfb_cdiv(R > 0 ? R : 1, chunks);
  chunks = (int)nfb_cdiv(R > 0 ? R : 1, chunk_len);
  float* tmp = nullptr;
  if (nfb_malloc(sizeof(float) * (size_t)chunks * C, (void**)&tmp)) return -1;
  dim3 grid(col_blocks, chunks), block(32, 8);
  if (dtype == NFB_F32) k_colsum_stage1<float><<<grid, block, 0, s>>>((const float*)x, R, C, ld, chunk_len, tmp);
  else if (dtype == NFB_BF16) k_colsum_stage1<__nv_bfloat16><<<grid, block, 0, s>>>((const __nv_bfloat16*)x, R, C, ld, chunk_len, tmp);
  else { nfb_free(tmp); return NFB_FAIL("nfb_colsum: unsupported dtype %d", dtype); }
  nfb_count_launch();
  k_colsum_stage2<<<(C + 255) / 256, 256, 0, s>>>(tmp, chunks, C, out, accumulate);
  nfb_free(tmp);
  NFB_LAUNCH_CHECK();
  return 0;
}
