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
                TD* __restrict__ dlogits, int64_t rows, int V, int64_t ld, int64_t dld, float grad_scale, int* __restrict__ err) {
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
    // a class id outside [-V, V) (NumPy raises IndexError; typical sources: an ignore_index such as -100, corrupt labels) must
    // not index the row: NaN loss, zero gradient row, and the device-side index-error flag the host can poll
    const bool bad = t < 0 || t >= V;
    if (bad) t = -1;
    if (threadIdx.x == 0) {
      if (bad) {
        loss[row] = __int_as_float(0x7fc00000);
        if (err) *err = 1;
      } else {
        float pt = (CACHE ? srow[t] : expf(xr[t] - m)) * inv;
        loss[row] = -logf(pt + 1e-8f);
      }
    }
    if (bad) inv = 0.f;
    const float gs = bad ? 0.f : grad_scale;
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
          o.x *= gs; o.y *= gs; o.z *= gs; o.w *= gs;
          st4<TD>(dr + c, o);
        }
        for (int c = v4 * 4 + threadIdx.x; c < V; c += blockDim.x) {
          float pv = srow[c] * inv;
          if (c == t) pv -= 1.f;
          st_f<TD>(dr, c, pv * gs);
        }
      } else {
        for (int c = threadIdx.x; c < V; c += blockDim.x) {
          float pv = (CACHE ? srow[c] : expf(xr[c] - m)) * inv;
          if (c == t) pv -= 1.f;
          st_f<TD>(dr, c, pv * gs);
        }
      }
    }
    __syncthreads();  // srow reused by the next row
  }
}

// exp(x - m) for x <= m as ONE FFMA + MUFU.EX2 (x * log2e - m * log2e); ~2 ulp, no overflow possible since the argument
// is <= 0.  expf() costs ~10 instructions and made the wide-row kernels issue-bound (206 M exponentials per LM-head step).
// ex2.approx.ftz: exp2f() without it wraps the MUFU in a subnormal-range rescue (FSETP + 2 predicated FMULs per value); a
// probability below 2^-126 flushed to zero changes nothing at any tolerance this library states.
__device__ __forceinline__ float exp_sub_fast(float x, float m_log2e) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(x, 1.4426950408889634f, -m_log2e)));
  return r;
}

// ---- register-resident variant for wide rows (the LM-head shapes: V = 50257, 30522).  One 1024-thread block per SM
// keeps a whole logits row in registers (NV4 x 4 values per thread, element 4*(k*1024 + t) .. +3), so the row is read
// from HBM exactly once, never touches shared memory, and the stores of row r overlap the loads of row r+1:
// max -> block_max -> exp/sum -> block_sum -> scale/store.  Logits may be fp32 or bf16 (TL); dlogits fp32 or bf16 (TD)
// and may alias the logits buffer.
template <typename TL, typename TD, int NV4>
__global__ void __launch_bounds__(1024, 1)
k_cross_entropy_reg(const TL* __restrict__ logits, const void* __restrict__ targets, int tgt_is_i64, float* __restrict__ loss,
                    TD* __restrict__ dlogits, int64_t rows, int V, int64_t ld, int64_t dld, float grad_scale, int* __restrict__ err) {
  __shared__ float red[32];
  const int t = threadIdx.x;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const TL* xr = logits + row * ld;
    f4 x[NV4];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int c = (k * 1024 + t) * 4;
      if (c < V) {   // the row pitch is padded to a multiple of 4 elements, the pad holds garbage: mask by column
        x[k] = ld4<TL>(xr + c);
        if (c + 1 >= V) x[k].y = -INFINITY;
        if (c + 2 >= V) x[k].z = -INFINITY;
        if (c + 3 >= V) x[k].w = -INFINITY;
      } else {
        x[k] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      }
      m = fmaxf(m, fmaxf(fmaxf(x[k].x, x[k].y), fmaxf(x[k].z, x[k].w)));
    }
    int64_t tg = tgt_is_i64 ? ((const int64_t*)targets)[row] : (int64_t)((const int32_t*)targets)[row];
    if (tg < 0) tg += V;  // numpy negative indexing
    const bool bad = tg < 0 || tg >= V;   // see k_cross_entropy: NaN loss, zero gradient row, error flag
    if (bad) {
      tg = -1;
      if (t == 0) { loss[row] = __int_as_float(0x7fc00000); if (err) *err = 1; }
    }
    m = block_max(m, red);
    const float ml2 = m * 1.4426950408889634f;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      x[k].x = exp_sub_fast(x[k].x, ml2); x[k].y = exp_sub_fast(x[k].y, ml2); x[k].z = exp_sub_fast(x[k].z, ml2); x[k].w = exp_sub_fast(x[k].w, ml2);
      s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
    }
    s = fmaxf(block_sum(s, red), 1e-8f);
    const float inv = 1.f / s;
    // the thread that owns the target column publishes its probability
    {
      const int tv = (int)(tg >> 2), tk = tv >> 10, tt = tv & 1023, te = (int)(tg & 3);
      if (tt == t) {
#pragma unroll
        for (int k = 0; k < NV4; ++k)
          if (k == tk) {
            float e = te == 0 ? x[k].x : te == 1 ? x[k].y : te == 2 ? x[k].z : x[k].w;
            loss[row] = -logf(e * inv + 1e-8f);
            if (dlogits) {  // (p - onehot): subtract after scaling, below
              if (te == 0) x[k].x -= s; else if (te == 1) x[k].y -= s; else if (te == 2) x[k].z -= s; else x[k].w -= s;
            }
          }
      }
    }
    if (dlogits) {
      TD* dr = dlogits + row * dld;
      const float sc = bad ? 0.f : inv * grad_scale;
#pragma unroll
      for (int k = 0; k < NV4; ++k) {
        const int c = (k * 1024 + t) * 4;
        if (c + 3 < V) {
          st4<TD>(dr + c, {x[k].x * sc, x[k].y * sc, x[k].z * sc, x[k].w * sc});
        } else if (c < V) {
          st_f<TD>(dr, c, x[k].x * sc);
          if (c + 1 < V) st_f<TD>(dr, c + 1, x[k].y * sc);
          if (c + 2 < V) st_f<TD>(dr, c + 2, x[k].z * sc);
        }
      }
    }
  }
}

// ---- bf16 logits, wide rows: the row lives in SHARED memory as the raw bf16 it was loaded as (100 KB at V = 50257),
// fetched by ONE bulk-async copy (cp.async.bulk global -> shared, mbarrier completion: no registers, no LSU issue), and
// double-buffered: one 1024-thread block per SM runs the three shared-memory passes (max, sum of exp, scale + store) of
// row r while the copy of its next row lands in the other buffer.  exp is evaluated twice per element; that is cheaper
// than a third trip through memory.  dlogits may alias the logits buffer (a block only ever prefetches its OWN next row).
__device__ __forceinline__ void ce_unpack8(const uint4& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}
template <typename TD>
__global__ void __launch_bounds__(1024, 1)
k_cross_entropy_bulk(const __nv_bfloat16* __restrict__ logits, const void* __restrict__ targets, int tgt_is_i64, float* __restrict__ loss,
                     TD* __restrict__ dlogits, int64_t rows, int V, int64_t ld, int64_t dld, float grad_scale, int* __restrict__ err) {
  // One 1024-thread CTA per SM, TWO row buffers: the bulk copy of row r+1 is in flight while the three shared-memory passes
  // (max, exp-sum, gradient) of row r run, and the gradient stores of row r drain under the passes of row r+1.  (The
  // single-buffer 2-CTA version serialised copy -> passes per CTA and ran at memory time + MUFU time: 202 us for 4096 x 50257.)
  extern __shared__ __align__(128) unsigned char ce_smem[];
  __shared__ float red[32];
  __shared__ __align__(8) unsigned long long bar[2];
  const int t = threadIdx.x;
  const int nv = (V + 7) >> 3;                 // 16-byte vectors per row (the pitch is padded to a multiple of 8 elements)
  const uint32_t row_bytes = (uint32_t)nv * 16u;
  const uint32_t buf_pitch = (row_bytes + 127u) & ~127u;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[0]), smem_a = (uint32_t)__cvta_generic_to_shared(ce_smem);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a + 8u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto fetch = [&](int64_t row, int b) {       // thread 0 only; every generic read of buffer b is behind a __syncthreads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a + 8u * b), "r"(row_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_a + buf_pitch * b), "l"(logits + row * ld), "r"(row_bytes), "r"(bar_a + 8u * b) : "memory");
  };
  if (t == 0 && (int64_t)blockIdx.x < rows) fetch(blockIdx.x, 0);
  int it = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
    const int b = it & 1;
    if (t == 0 && row + gridDim.x < rows) fetch(row + gridDim.x, b ^ 1);
    const uint4* srow = reinterpret_cast<const uint4*>(ce_smem + (size_t)buf_pitch * b);
    int64_t tg = tgt_is_i64 ? ((const int64_t*)targets)[row] : (int64_t)((const int32_t*)targets)[row];
    if (tg < 0) tg += V;  // numpy negative indexing
    const bool bad = tg < 0 || tg >= V;   // see k_cross_entropy: NaN loss, zero gradient row, error flag
    if (bad) tg = -1;
    {
      const uint32_t phase = (uint32_t)(it >> 1) & 1u;
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar_a + 8u * b), "r"(phase) : "memory");
      }
    }
    float m = -INFINITY;
    for (int i = t; i < nv; i += 1024) {
      float v[8];
      ce_unpack8(srow[i], v);
      const int c = i * 8;
      if (c + 8 > V) {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (c + k >= V) v[k] = -INFINITY;   // the pad columns hold garbage
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) m = fmaxf(m, v[k]);
    }
    m = block_max(m, red);
    const float ml2 = m * 1.4426950408889634f;
    float s = 0.f;
    for (int i = t; i < nv; i += 1024) {
      float v[8];
      ce_unpack8(srow[i], v);
      const int c = i * 8;
      if (c + 8 <= V) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s += exp_sub_fast(v[k], ml2);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) s += (c + k < V) ? exp_sub_fast(v[k], ml2) : 0.f;
      }
    }
    s = fmaxf(block_sum(s, red), 1e-8f);
    const float inv = 1.f / s;
    if (t == 0) {
      if (bad) {
        loss[row] = __int_as_float(0x7fc00000);
        if (err) *err = 1;
      } else {
        float xt = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(srow)[tg]);
        loss[row] = -logf(exp_sub_fast(xt, ml2) * inv + 1e-8f);
      }
    }
    if (dlogits) {
      TD* dr = dlogits + row * dld;
      const float sc = bad ? 0.f : inv * grad_scale;
      const int tvec = (int)(tg >> 3), te = (int)(tg & 7);
      for (int i = t; i < nv; i += 1024) {
        float v[8];
        ce_unpack8(srow[i], v);
        const int c = i * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = exp_sub_fast(v[k], ml2) * sc;
        if (i == tvec) {
#pragma unroll
          for (int k = 0; k < 8; ++k) if (k == te) v[k] -= grad_scale;
        }
        if (c + 8 <= V) {
          st4<TD>(dr + c, {v[0], v[1], v[2], v[3]});
          st4<TD>(dr + c + 4, {v[4], v[5], v[6], v[7]});
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) if (c + k < V) st_f<TD>(dr, c + k, v[k]);
        }
      }
    }
    __syncthreads();   // everyone is done with buffer b before the copy of row r+2 overwrites it
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
    k_cross_entropy<TD, C, VV><<<(unsigned)grid, threads, C ? smem : 0, s>>>(logits, targets, tgt_is_i64, loss, dlogits, rows, V, ld, dld, grad_scale, nfb_index_error_flag()); \
  } while (0)
  if (cache && vec) LAUNCH(true, true);
  else if (cache) LAUNCH(true, false);
  else LAUNCH(false, false);
#undef LAUNCH
  NFB_LAUNCH_CHECK();
  return 0;
}

// wide rows: the register-resident kernel (one block per SM)
template <typename TL, typename TD>
static int ce_launch_reg(const TL* logits, const void* targets, int tgt_is_i64, float* loss, TD* dlogits, int64_t rows, int V,
                         int64_t ld, int64_t dld, float grad_scale) {
  cudaStream_t s = nfb_cur_stream();
  int nv4 = (int)nfb_cdiv(nfb_cdiv(V, 4), 1024);
  int64_t grid = rows < nfb_sm_count() ? rows : nfb_sm_count();
#define LR(NV) k_cross_entropy_reg<TL, TD, NV><<<(unsigned)grid, 1024, 0, s>>>(logits, targets, tgt_is_i64, loss, dlogits, rows, V, ld, dld, grad_scale, nfb_index_error_flag())
  if (nv4 <= 4) LR(4);
  else if (nv4 <= 8) LR(8);
  else if (nv4 <= 13) LR(13);
  else return NFB_FAIL("cross entropy: V=%d exceeds the register-resident kernel", V);
#undef LR
  NFB_LAUNCH_CHECK();
  return 0;
}

static int g_ce_bulk = 1;
NFB_API int nfb_cross_entropy_set_bulk(int on) { g_ce_bulk = on ? 1 : 0; return 0; }  // test hook: 0 = register-resident kernel for bf16 logits too

// ldt / ddt: dtypes of logits / dlogits (NFB_F32 or NFB_BF16); dlogits may be NULL (forward only) and may alias logits
// (same dtype and pitch).
NFB_API int nfb_cross_entropy_ex(int ldt, const void* logits, const void* targets, int tgt_is_i64, float* loss, void* dlogits, int ddt,
                                 int64_t rows, int V, int64_t ld, int64_t dld, float grad_scale) {
  if (rows == 0) return 0;
  if (ld < V) return NFB_FAIL("nfb_cross_entropy: ld %lld < V %d", (long long)ld, V);
  if (ldt != NFB_F32 && ldt != NFB_BF16) return NFB_FAIL("nfb_cross_entropy: unsupported logits dtype %d", ldt);
  if (dlogits && ddt != NFB_F32 && ddt != NFB_BF16) return NFB_FAIL("nfb_cross_entropy: unsupported dlogits dtype %d", ddt);
  const int lsz = ldt == NFB_F32 ? 4 : 2, dsz = ddt == NFB_F32 ? 4 : 2;
  // 4-element vectors: 16-byte (fp32) / 8-byte (bf16) aligned rows
  bool vec = (ld % 4 == 0) && ((uintptr_t)logits % (4 * lsz) == 0) && (!dlogits || ((dld % 4 == 0) && ((uintptr_t)dlogits % (4 * dsz) == 0)));
  if (vec && V >= 8192 && V <= 13 * 4096) {
    if (ldt == NFB_F32) {
      if (!dlogits || ddt == NFB_F32) return ce_launch_reg<float, float>((const float*)logits, targets, tgt_is_i64, loss, (float*)dlogits, rows, V, ld, dld, grad_scale);
      return ce_launch_reg<float, __nv_bfloat16>((const float*)logits, targets, tgt_is_i64, loss, (__nv_bfloat16*)dlogits, rows, V, ld, dld, grad_scale);
    }
    // bf16 logits: shared-memory rows + bulk-async copies, one block per SM (needs 16-byte aligned rows)
    if (ld % 8 == 0 && (uintptr_t)logits % 16 == 0 && g_ce_bulk) {
      const size_t smem = 2 * (((size_t)((V + 7) / 8) * 16 + 127) & ~(size_t)127);   // two row buffers
      cudaStream_t s = nfb_cur_stream();
      int64_t grid = nfb_sm_count();
      if (grid > rows) grid = rows;
      if (!dlogits || ddt == NFB_BF16) {
        NFB_CUDA(cudaFuncSetAttribute(k_cross_entropy_bulk<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_cross_entropy_bulk<__nv_bfloat16><<<(unsigned)grid, 1024, smem, s>>>((const __nv_bfloat16*)logits, targets, tgt_is_i64, loss, (__nv_bfloat16*)dlogits, rows, V, ld, dld, grad_scale, nfb_index_error_flag());
      } else {
        NFB_CUDA(cudaFuncSetAttribute(k_cross_entropy_bulk<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_cross_entropy_bulk<float><<<(unsigned)grid, 1024, smem, s>>>((const __nv_bfloat16*)logits, targets, tgt_is_i64, loss, (float*)dlogits, rows, V, ld, dld, grad_scale, nfb_index_error_flag());
      }
      NFB_LAUNCH_CHECK();
      return 0;
    }
    if (!dlogits || ddt == NFB_BF16) return ce_launch_reg<__nv_bfloat16, __nv_bfloat16>((const __nv_bfloat16*)logits, targets, tgt_is_i64, loss, (__nv_bfloat16*)dlogits, rows, V, ld, dld, grad_scale);
    return ce_launch_reg<__nv_bfloat16, float>((const __nv_bfloat16*)logits, targets, tgt_is_i64, loss, (float*)dlogits, rows, V, ld, dld, grad_scale);
  }
  if (ldt != NFB_F32) return NFB_FAIL("nfb_cross_entropy: bf16 logits need 8-byte aligned rows and 8192 <= V <= 53248 (V=%d, ld=%lld)", V, (long long)ld);
  if (!dlogits || ddt == NFB_F32) return ce_launch<float>((const float*)logits, targets, tgt_is_i64, loss, (float*)dlogits, rows, V, ld, dld, grad_scale);
  return ce_launch<__nv_bfloat16>((const float*)logits, targets, tgt_is_i64, loss, (__nv_bfloat16*)dlogits, rows, V, ld, dld, grad_scale);
}

// fp32 logits (the original entry point)
NFB_API int nfb_cross_entropy(const float* logits, const void* targets, int tgt_is_i64, float* loss, void* dlogits, int ddt,
                              int64_t rows, int V, int64_t ld, int64_t dld, float grad_scale) {
  return nfb_cross_entropy_ex(NFB_F32, logits, targets, tgt_is_i64, loss, dlogits, ddt, rows, V, ld, dld, grad_scale);
}
