// Row softmax forward/backward and the fused cross-entropy forward+backward.
// Reference math: functional/activation.py:65-147 (softmax: e^{x-max} / max(sum, 1e-8);
// backward p*(g - sum(g*p))) and functional/loss.py:15-112 (loss = -log(p[t] + 1e-8), gradient
// (p - onehot) * scale, epsilon ignored in the gradient).
// The CE kernel keeps a whole logits row in shared memory (50257 fp32 = 196 KB < 227 KB) so the
// logits are read from HBM once and dlogits written once: 8 B/element algorithmic.
#include "common.cuh"

// ---------------- softmax: small rows (cols <= 1024), one warp per row, row in registers -------
template <int PER_LANE>
__global__ void k_softmax_fwd_warp(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int cols) {
  int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * cols;
  float v[PER_LANE];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < PER_LANE; ++k) {
    int c = k * 32 + lane;
    v[k] = c < cols ? xr[c] : -INFINITY;
    m = fmaxf(m, v[k]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < PER_LANE; ++k) {
    int c = k * 32 + lane;
    v[k] = c < cols ? expf(v[k] - m) : 0.f;
    s += v[k];
  }
  s = fmaxf(warp_sum(s), 1e-8f);
  float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < PER_LANE; ++k) {
    int c = k * 32 + lane;
    if (c < cols) y[row * cols + c] = v[k] * inv;
  }
}

// ---------------- softmax: long rows, one block per row (online max/sum, then write) -----------
__global__ void k_softmax_fwd_block(const float* __restrict__ x, float* __restrict__ y, int cols) {
  __shared__ float red[32];
  int64_t row = blockIdx.x;
  const float* xr = x + row * cols;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, xr[c]);
  m = block_max(m, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s += expf(xr[c] - m);
  s = fmaxf(block_sum(s, red), 1e-8f);
  float inv = 1.f / s;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) y[row * cols + c] = expf(xr[c] - m) * inv;
}

NFB_API int nfb_softmax_fwd(const float* x, float* y, int64_t rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  cudaStream_t s = nfb_cur_stream();
  if (cols <= 1024) {
    int wpb = 4;
    unsigned grid = (unsigned)nfb_cdiv(rows, wpb);
    int per = (cols + 31) / 32;
    if (per <= 1) k_softmax_fwd_warp<1><<<grid, wpb * 32, 0, s>>>(x, y, rows, cols);
    else if (per <= 2) k_softmax_fwd_warp<2><<<grid, wpb * 32, 0, s>>>(x, y, rows, cols);
    else if (per <= 4) k_softmax_fwd_warp<4><<<grid, wpb * 32, 0, s>>>(x, y, rows, cols);
    else if (per <= 8) k_softmax_fwd_warp<8><<<grid, wpb * 32, 0, s>>>(x, y, rows, cols);
    else if (per <= 16) k_softmax_fwd_warp<16><<<grid, wpb * 32, 0, s>>>(x, y, rows, cols);
    else k_softmax_fwd_warp<32><<<grid, wpb * 32, 0, s>>>(x, y, rows, cols);
  } else {
    k_softmax_fwd_block<<<(unsigned)rows, 512, 0, s>>>(x, y, cols);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}

// dx = p * (g - sum(g*p)) per row
__global__ void k_softmax_bwd(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ dx, int cols) {
  __shared__ float red[32];
  int64_t row = blockIdx.x;
  const float* pr = p + row * cols;
  const float* gr = g + row * cols;
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s += pr[c] * gr[c];
  s = block_sum(s, red);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dx[row * cols + c] = pr[c] * (gr[c] - s);
}
NFB_API int nfb_softmax_bwd(const float* p, const float* g, float* dx, int64_t rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  int threads = cols >= 2048 ? 512 : (cols >= 256 ? 128 : 32);
  k_softmax_bwd<<<(unsigned)rows, threads, 0, nfb_cur_stream()>>>(p, g, dx, cols);
  NFB_LAUNCH_CHECK();
  return 0;
}

// ---------------- fused cross entropy ---------------------------------------------------------
// One block per row (grid-stride over rows). Row cached in dynamic shared memory when it fits.
// loss[row] = -log(p[t] + 1e-8); dlogits[row, :] = (p - onehot(t)) * grad_scale   (optional)
// logits row stride = ld elements (ld >= V; rows may be padded for 16-B alignment).
template <typename TD, bool CACHE, bool VEC>
__global__ void __launch_bounds__(1024)
k_cross_entropy(const float* __restrict__ logits, const void* __restrict__ targets, int tgt_is_i64, float* __restrict__ loss,
                TD* __restrict__ dlogits, int64_t rows, int V, int64_t ld, int64_t dld, float grad_scale) {
  extern __shared__ float srow[];
  __shared__ float red[32];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* xr = logits + row * ld;
    float m = -INFINITY;
    if (VEC) {
      int v4 = V / 4;
      for (int i = threadIdx.x; i < v4; i += blockDim.x) {
        float4 a = reinterpret_cast<const float4*>(xr)[i];
        if (CACHE) reinterpret_cast<float4*>(srow)[i] = a;
        m = fmaxf(m, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
      }
      for (int c = v4 * 4 + threadIdx.x; c < V; c += blockDim.x) {
        float a = xr[c];
        if (CACHE) srow[c] = a;
        m = fmaxf(m, a);
      }
    } else {
      for (int c = threadIdx.x; c < V; c += blockDim.x) {
        float a = xr[c];
        if (CACHE) srow[c] = a;
        m = fmaxf(m, a);
      }
    }
    m = block_max(m, red);  // contains the __syncthreads that publishes srow
    float s = 0.f;
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
      float e = expf((CACHE ? srow[c] : xr[c]) - m);
      if (CACHE) srow[c] = e;
      s += e;
    }
    s = fmaxf(block_sum(s, red), 1e-8f);
    float inv = 1.f / s;
    int64_t t = tgt_is_i64 ? ((const int64_t*)targets)[row] : (int64_t)((const int32_t*)targets)[row];
    if (t < 0) t += V;  // numpy negative indexing
    if (threadIdx.x == 0) {
      float pt = (CACHE ? srow[t] : expf(xr[t] - m)) * inv;
      loss[row] = -logf(pt + 1e-8f);
    }
    if (!CACHE) __syncthreads();  // dlogits may alias logits: finish reading xr[t] first
    if (dlogits) {
      TD* dr = dlogits + row * dld;
      if (VEC && CACHE) {
        int v4 = V / 4;
        for (int i = threadIdx.x; i < v4; i += blockDim.x) {
          float4 e = reinterpret_cast<const float4*>(srow)[i];
          int c = i * 4;
          f4 o = {e.x * inv, e.y * inv, e.z * inv, e.w * inv};
          if (t >= c && t < c + 4) {
            if (t == c) o.x -= 1.f; else if (t == c + 1) o.y -= 1.f; else if (t == c + 2) o.z -= 1.f; else o.w -= 1.f;
          }
          o.x *= grad_scale; o.y *= grad_scale; o.z *= grad_scale; o.w *= grad_scale;
          st4<TD>(dr + c, o);
        }
        for (int c = v4 * 4 + threadIdx.x; c < V; c += blockDim.x) {
          float pv = srow[c] * inv;
          if (c == t) pv -= 1.f;
          st_f<TD>(dr, c, pv * grad_scale);
        }
      } else {
        for (int c = threadIdx.x; c < V; c += blockDim.x) {
          float pv = (CACHE ? srow[c] : expf(xr[c] - m)) * inv;
          if (c == t) pv -= 1.f;
          st_f<TD>(dr, c, pv * grad_scale);
        }
      }
    }
    __syncthreads();  // srow reused by the next row
  }
}

template <typename TD>
static int ce_launch(const float* logits, const void* targets, int tgt_is_i64, float* loss, TD* dlogits, int64_t rows, int V,
                     int64_t ld, int64_t dld, float grad_scale) {
  cudaStream_t s = nfb_cur_stream();
  size_t smem = (size_t)V * sizeof(float);
  bool cache = smem <= 200 * 1024;
  bool vec = (ld % 4 == 0) && ((uintptr_t)logits % 16 == 0) && (!dlogits || ((dld % 4 == 0) && ((uintptr_t)dlogits % 16 == 0)));
  int threads = V >= 8192 ? 1024 : (V >= 1024 ? 256 : 64);
  int64_t grid = rows;
  int64_t cap = (int64_t)nfb_sm_count() * (cache && smem > 100 * 1024 ? 1 : 4) * 4;
  if (grid > cap) grid = cap;
#define LAUNCH(C, VV)                                                                                                        \
  do {                                                                                                                       \
    if (C) cudaFuncSetAttribute(k_cross_entropy<TD, C, VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    k_cross_entropy<TD, C, VV><<<(unsigned)grid, threads, C ? smem : 0, s>>>(logits, targets, tgt_is_i64, loss, dlogits, rows, V, ld, dld, grad_scale); \
  } while (0)
  if (cache && vec) LAUNCH(true, true);
  else if (cache) LAUNCH(true, false);
  else LAUNCH(false, false);
#undef LAUNCH
  NFB_LAUNCH_CHECK();
  return 0;
}

// ddt: dtype of dlogits (NFB_F32 or NFB_BF16); dlogits may be NULL (forward only) and may alias logits (fp32, same ld).
NFB_API int nfb_cross_entropy(const float* logits, const void* targets, int tgt_is_i64, float* loss, void* dlogits, int ddt,
                              int64_t rows, int V, int64_t ld, int64_t dld, float grad_scale) {
  if (rows == 0) return 0;
  if (ld < V) return NFB_FAIL("nfb_cross_entropy: ld %lld < V %d", (long long)ld, V);
  if (!dlogits || ddt == NFB_F32) return ce_launch<float>(logits, targets, tgt_is_i64, loss, (float*)dlogits, rows, V, ld, dld, grad_scale);
  if (ddt == NFB_BF16) return ce_launch<__nv_bfloat16>(logits, targets, tgt_is_i64, loss, (__nv_bfloat16*)dlogits, rows, V, ld, dld, grad_scale);
  return NFB_FAIL("nfb_cross_entropy: unsupported dlogits dtype %d", ddt);
}
