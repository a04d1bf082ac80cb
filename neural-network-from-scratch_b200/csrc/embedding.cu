// Embedding row gather and the deterministic (bit-exact) scatter-add backward.
// Reference: nn/embedding.py:27-63 -- forward W[idx.astype(int32)], backward a sequential loop
// `weight_grad[idx[i]] += flat_grad[i]` in ascending i. To reproduce that bit for bit the backward
// stable-sorts (index, position) pairs and then sums every segment in ascending position order,
// one fp32 add at a time, with no atomics. Segments are parallel over columns and over rows.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

static int* g_index_error = nullptr;  // device flag: set when an index is out of range

static int ensure_flag() {
  if (!g_index_error) {
    NFB_CUDA(cudaMalloc(&g_index_error, sizeof(int)));
    NFB_CUDA(cudaMemsetAsync(g_index_error, 0, sizeof(int), nfb_cur_stream()));
  }
  return 0;
}

// the flag is shared with the loss kernels (class ids outside the vocabulary), csrc/softmax_ce.cu
int* nfb_index_error_flag() {
  if (ensure_flag()) return nullptr;
  return g_index_error;
}

template <typename TI>
__device__ __forceinline__ int64_t load_index(const TI* idx, int64_t i, int64_t V, int* err) {
  int64_t v = (int64_t)idx[i];
  if (v < 0) v += V;
  if (v < 0 || v >= V) {
    *err = 1;
    v = 0;
  }
  return v;
}

// out[i, :] = W[idx[i], :]   one warp per output row, 128-bit copies when d % 4 == 0
template <typename TI, typename TW, typename TO>
__global__ void k_gather_rows(const TW* __restrict__ W, const TI* __restrict__ idx, TO* __restrict__ out, int64_t n, int d, int64_t V, int* err) {
  int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = warp; i < n; i += nwarps) {
    int64_t r = load_index<TI>(idx, i, V, err);
    const TW* src = W + r * d;
    TO* dst = out + i * d;
    if (d % 4 == 0) {
      for (int c = lane * 4; c < d; c += 128) st4<TO>(dst + c, ld4<TW>(src + c));
    } else {
      for (int c = lane; c < d; c += 32) st_f<TO>(dst, c, ld_f<TW>(src, c));
    }
  }
}

// idx_dtype: NFB_I32 / NFB_I64 ; wdt: dtype of W (F32/BF16) ; odt: dtype of out (F32/BF16)
NFB_API int nfb_gather_rows(int idx_dtype, int wdt, int odt, const void* W, const void* idx, void* out, int64_t n, int d, int64_t V) {
  if (n == 0 || d == 0) return 0;
  if (ensure_flag()) return -1;
  cudaStream_t s = nfb_cur_stream();
  int grid = nfb_grid_for(n, 8);
#define GO(TI, TW, TO) k_gather_rows<TI, TW, TO><<<grid, 256, 0, s>>>((const TW*)W, (const TI*)idx, (TO*)out, n, d, V, g_index_error)
  if (wdt == NFB_F32 && odt == NFB_F32) { if (idx_dtype == NFB_I32) GO(int32_t, float, float); else GO(int64_t, float, float); }
  else if (wdt == NFB_F32 && odt == NFB_BF16) { if (idx_dtype == NFB_I32) GO(int32_t, float, __nv_bfloat16); else GO(int64_t, float, __nv_bfloat16); }
  else if (wdt == NFB_BF16 && odt == NFB_BF16) { if (idx_dtype == NFB_I32) GO(int32_t, __nv_bfloat16, __nv_bfloat16); else GO(int64_t, __nv_bfloat16, __nv_bfloat16); }
  else return NFB_FAIL("nfb_gather_rows: unsupported dtype pair (%d,%d)", wdt, odt);
#undef GO
  NFB_LAUNCH_CHECK();
  return 0;
}

// returns 1 in *flag when any gather/scatter since the last call saw an out-of-range index
NFB_API int nfb_index_error_check(int* flag) {
  *flag = 0;
  if (!g_index_error) return 0;
  NFB_CUDA(cudaMemcpyAsync(flag, g_index_error, sizeof(int), cudaMemcpyDeviceToHost, nfb_cur_stream()));
  NFB_CUDA(cudaStreamSynchronize(nfb_cur_stream()));
  if (*flag) NFB_CUDA(cudaMemsetAsync(g_index_error, 0, sizeof(int), nfb_cur_stream()));
  return 0;
}

template <typename TI>
__global__ void k_prep_keys(const TI* __restrict__ idx, int32_t* __restrict__ keys, int32_t* __restrict__ pos, int64_t n, int64_t V, int* err) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (int32_t)load_index<TI>(idx, i, V, err);
  pos[i] = (int32_t)i;
}

// grid = (n candidate segment starts, ceil(d/128) column groups); 32 threads, one float4 column each.
#define LONG_SEG 16
template <typename TG>
__global__ void k_scatter_add_sorted(const int32_t* __restrict__ keys, const int32_t* __restrict__ pos, const TG* __restrict__ grad,
                                     float* __restrict__ dW, int64_t n, int d, int accumulate, int32_t* __restrict__ long_list) {
  int64_t i = blockIdx.x;
  int32_t key = keys[i];
  if (i > 0 && keys[i - 1] == key) return;  // not the first element of its segment
  // long segments (an index repeated more than LONG_SEG times: BERT's position / segment-id tables) go to the
  // cooperative kernel below -- this one-warp loop chases pos[j] -> grad row serially, ~1 us per row
  if (long_list && i + LONG_SEG < n && keys[i + LONG_SEG] == key) {
    if (blockIdx.y == 0 && threadIdx.x == 0) long_list[1 + atomicAdd(long_list, 1)] = (int32_t)i;
    return;
  }
  int c = (blockIdx.y * 32 + threadIdx.x) * 4;
  if (c >= d) return;
  bool vec = (d % 4 == 0);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int64_t j = i; j < n && keys[j] == key; ++j) {
    const TG* g = grad + (int64_t)pos[j] * d + c;
    if (vec) {
      f4 v = ld4<TG>(g);
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    } else {
      a0 += ld_f<TG>(g, 0);
      if (c + 1 < d) a1 += ld_f<TG>(g, 1);
      if (c + 2 < d) a2 += ld_f<TG>(g, 2);
      if (c + 3 < d) a3 += ld_f<TG>(g, 3);
    }
  }
  float* o = dW + (int64_t)key * d + c;
  if (vec) {
    float4 prev = accumulate ? *reinterpret_cast<float4*>(o) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(o) = accumulate ? make_float4(prev.x + a0, prev.y + a1, prev.z + a2, prev.w + a3) : make_float4(a0, a1, a2, a3);
  } else {
    o[0] = accumulate ? o[0] + a0 : a0;
    if (c + 1 < d) o[1] = accumulate ? o[1] + a1 : a1;
    if (c + 2 < d) o[2] = accumulate ? o[2] + a2 : a2;
    if (c + 3 < d) o[3] = accumulate ? o[3] + a3 : a3;
  }
}

// Long segments, one block (128 threads) per (segment, 128-column group).  Still the reference's order -- row j of the
// segment is added after row j-1, one fp32 add at a time (nn/embedding.py:54-55) -- but the memory side is parallel:
// 64 gradient rows are fetched at once (16 independent 16-byte loads per thread) into shared memory, then every
// thread walks ITS scalar column down the 64 staged rows.  long_list[0] = number of segments, long_list[1..] = starts.
template <typename TG>
__global__ void __launch_bounds__(128)
k_scatter_add_long(const int32_t* __restrict__ keys, const int32_t* __restrict__ pos, const TG* __restrict__ grad, float* __restrict__ dW,
                   int64_t n, int d, int accumulate, const int32_t* __restrict__ long_list) {
  __shared__ float sbuf[64][128 + 4];
  __shared__ int32_t spos[64];
  __shared__ int64_t s_len;
  if ((int)blockIdx.x >= long_list[0]) return;
  const int64_t i = long_list[1 + blockIdx.x];
  const int32_t key = keys[i];
  const int t = threadIdx.x, c4 = t & 31, rg = t >> 5;
  const int colbase = blockIdx.y * 128;
  if (t == 0) {   // segment end: gallop, then bisect (keys are sorted)
    int64_t lo = i, step = 1;
    while (lo + step < n && keys[lo + step] == key) { lo += step; step <<= 1; }
    int64_t hi = lo + step < n ? lo + step : n;   // keys[lo] == key, keys[hi] != key (or hi == n)
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (keys[mid] == key) lo = mid; else hi = mid;
    }
    s_len = hi - i;
  }
  __syncthreads();
  const int64_t L = s_len;
  const bool vec = (d % 4 == 0);
  const int ncols = min(128, d - colbase);   // columns this block owns
  float acc = 0.f;
  for (int64_t base = 0; base < L; base += 64) {
    const int rows = (int)min((int64_t)64, L - base);
    if (t < rows) spos[t] = pos[i + base + t];
    __syncthreads();
    f4 v[16];
    const int c = colbase + c4 * 4;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int row = rg * 16 + r;
      v[r] = {0.f, 0.f, 0.f, 0.f};
      if (row < rows && c < d) {
        const TG* g = grad + (int64_t)spos[row] * d + c;
        if (vec) v[r] = ld4<TG>(g);
        else {
          v[r].x = ld_f<TG>(g, 0);
          if (c + 1 < d) v[r].y = ld_f<TG>(g, 1);
          if (c + 2 < d) v[r].z = ld_f<TG>(g, 2);
          if (c + 3 < d) v[r].w = ld_f<TG>(g, 3);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int row = rg * 16 + r;
      sbuf[row][c4 * 4 + 0] = v[r].x; sbuf[row][c4 * 4 + 1] = v[r].y; sbuf[row][c4 * 4 + 2] = v[r].z; sbuf[row][c4 * 4 + 3] = v[r].w;
    }
    __syncthreads();
    if (t < ncols) {
#pragma unroll 8
      for (int row = 0; row < rows; ++row) acc += sbuf[row][t];   // strictly in segment order
    }
    __syncthreads();
  }
  if (t < ncols) {
    float* o = dW + (int64_t)key * d + colbase + t;
    *o = accumulate ? *o + acc : acc;
  }
}

// dW[idx[i], :] (+)= sum over i in ascending order of grad[i, :].  Rows of dW that no index touches
// are left alone: the caller zero-fills dW first when accumulate == 0.
NFB_API int nfb_scatter_add_rows(int idx_dtype, int gdt, const void* grad, const void* idx, float* dW, int64_t n, int d, int64_t V, int accumulate) {
  if (n == 0 || d == 0) return 0;
  if (n > 2147483647LL || V > 2147483647LL) return NFB_FAIL("nfb_scatter_add_rows: sizes exceed int32");
  if (ensure_flag()) return -1;
  cudaStream_t s = nfb_cur_stream();
  int32_t* buf = nullptr;
  if (nfb_malloc(sizeof(int32_t) * 4 * (size_t)n, (void**)&buf)) return -1;
  int32_t *keys_in = buf, *pos_in = buf + n, *keys_out = buf + 2 * n, *pos_out = buf + 3 * n;
  unsigned g1 = (unsigned)nfb_cdiv(n, 256);
  if (idx_dtype == NFB_I32) k_prep_keys<int32_t><<<g1, 256, 0, s>>>((const int32_t*)idx, keys_in, pos_in, n, V, g_index_error);
  else k_prep_keys<int64_t><<<g1, 256, 0, s>>>((const int64_t*)idx, keys_in, pos_in, n, V, g_index_error);
  nfb_count_launch();
  int end_bit = 1;
  while (end_bit < 32 && (1LL << end_bit) < V) ++end_bit;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, pos_in, pos_out, (int)n, 0, end_bit, s);
  void* tmp = nullptr;
  if (nfb_malloc(tmp_bytes, &tmp)) { nfb_free(buf); return -1; }
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, pos_in, pos_out, (int)n, 0, end_bit, s);
  nfb_count_launch();
  if (e != cudaSuccess) { nfb_free(tmp); nfb_free(buf); return NFB_FAIL("radix sort -> %s", cudaGetErrorString(e)); }
  dim3 grid((unsigned)n, (unsigned)nfb_cdiv(d, 128));
  // segments longer than LONG_SEG are collected into a list and summed by the cooperative kernel
  const int64_t max_long = n / (LONG_SEG + 1) + 1;
  int32_t* long_list = nullptr;
  if (nfb_malloc(sizeof(int32_t) * (size_t)(max_long + 1), (void**)&long_list)) { nfb_free(tmp); nfb_free(buf); return -1; }
  NFB_CUDA(cudaMemsetAsync(long_list, 0, sizeof(int32_t), s));
  dim3 grid_long((unsigned)max_long, (unsigned)nfb_cdiv(d, 128));
  if (gdt == NFB_F32) {
    k_scatter_add_sorted<float><<<grid, 32, 0, s>>>(keys_out, pos_out, (const float*)grad, dW, n, d, accumulate, long_list);
    nfb_count_launch();
    k_scatter_add_long<float><<<grid_long, 128, 0, s>>>(keys_out, pos_out, (const float*)grad, dW, n, d, accumulate, long_list);
  } else if (gdt == NFB_BF16) {
    k_scatter_add_sorted<__nv_bfloat16><<<grid, 32, 0, s>>>(keys_out, pos_out, (const __nv_bfloat16*)grad, dW, n, d, accumulate, long_list);
    nfb_count_launch();
    k_scatter_add_long<__nv_bfloat16><<<grid_long, 128, 0, s>>>(keys_out, pos_out, (const __nv_bfloat16*)grad, dW, n, d, accumulate, long_list);
  } else { nfb_free(long_list); nfb_free(tmp); nfb_free(buf); return NFB_FAIL("nfb_scatter_add_rows: unsupported grad dtype %d", gdt); }
  nfb_free(long_list);
  nfb_free(tmp);
  nfb_free(buf);
  NFB_LAUNCH_CHECK();
  return 0;
}
