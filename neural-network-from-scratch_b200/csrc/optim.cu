// Fused optimizer updates over flat parameter arenas (one launch per arena or per tensor),
// plus the global-norm gradient clip. HBM-bound: 28 B/param with fp32 moments, 44 B/param with the
// reference's fp64 moments.
// Reference semantics (optim/adam.py:165-252, optim/adamw.py:278-372), including its quirks:
//   * moments are accumulated in float64 from grad.astype(float64)        (MT = double variant)
//   * epsilon sits INSIDE the square root: denom = sqrt(v_hat + eps)
//   * Adam multiplies small learning rates by min(5, 0.1/lr) when 0 < lr <= 0.02
//   * AdamW uses step_size = lr / bias_correction1 on an already bias-corrected m_hat
//   * the update is cast to the parameter dtype and clipped to [-10, 10]
//   * non-finite parameters are scrubbed with nan_to_num(nan=0, posinf=1, neginf=-1)
#include "common.cuh"

// per-launch scalars, computed once per block by thread 0 (two double pow() per THREAD cost more than the update itself)
struct AdamScalars { double lr_eff, inv_bc1, inv_bc2; float gs, wd_factor; };
__device__ __forceinline__ void adam_scalars(AdamScalars* sc, int variant, double lr, double beta1, double beta2, int step_host,
                                             const int* step_dev, float gsign, const float* grad_scale, float wd_factor,
                                             const double* lr_dev, double weight_decay) {
  if (threadIdx.x == 0) {
    // learning rate: a host value, or a device scalar a scheduler rewrites between replays of a captured step
    if (lr_dev) {
      lr = *lr_dev;
      wd_factor = (variant == 1 && weight_decay != 0.0) ? (float)(1.0 - lr * weight_decay) : 1.f;
    }
    sc->wd_factor = wd_factor;
    // bias corrections from the step count: a host value, or a device counter so that a captured CUDA graph of the
    // training step stays valid from one replay to the next
    const int step = step_dev ? *step_dev : step_host;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    double lr_d;
    if (variant == 0) lr_d = (lr <= 0.02 && lr > 0.0) ? lr * fmin(5.0, 0.1 / lr) : lr;
    else lr_d = lr / bc1;
    sc->lr_eff = lr_d; sc->inv_bc1 = 1.0 / bc1; sc->inv_bc2 = 1.0 / bc2;
    sc->gs = (grad_scale ? *grad_scale : 1.f) * gsign;
  }
  __syncthreads();
}

template <typename MT>
__device__ __forceinline__ float adam_one(int variant, float pi, float gi, MT& mi, MT& vi, MT beta1, MT beta2, MT eps, MT lr_eff,
                                          MT inv_bc1, MT inv_bc2, float wd_factor, float wd_grad) {
  if (variant == 0 && wd_grad != 0.f) gi = __fadd_rn(gi, __fmul_rn(wd_grad, pi));  // Adam: L2 term in the gradient (fp32)
  MT gd = (MT)gi;
  mi = beta1 * mi + (MT(1) - beta1) * gd;
  vi = beta2 * vi + (MT(1) - beta2) * (gd * gd);
  MT mhat = mi * inv_bc1;
  MT vhat = vi * inv_bc2;
  MT denom = sqrt(vhat + eps);
  float upd = (float)(lr_eff * mhat / denom);
  if (upd == upd) upd = fminf(fmaxf(upd, -10.f), 10.f);  // np.clip keeps NaN (scrubbed below)
  float np_;
  if (variant == 1 && wd_factor != 1.f) np_ = __fsub_rn(__fmul_rn(pi, wd_factor), upd);  // AdamW decoupled decay
  else np_ = __fsub_rn(pi, upd);
  if (!isfinite(np_)) np_ = (np_ != np_) ? 0.f : (np_ > 0.f ? 1.f : -1.f);
  return np_;
}

// scalar variant (fp64 moments = the reference's exact state dtype; also the tail of the vector kernel)
template <typename MT>
__global__ void k_adam(int variant, float* __restrict__ p, const float* __restrict__ g, MT* __restrict__ m, MT* __restrict__ v,
                       __nv_bfloat16* __restrict__ p_bf16, int64_t n0, int64_t n, double lr, MT beta1, MT beta2, MT eps, float wd_factor, float wd_grad,
                       int step_host, const int* step_dev, float gsign, const float* grad_scale, const double* lr_dev, double weight_decay) {
  __shared__ AdamScalars sc;
  adam_scalars(&sc, variant, lr, (double)beta1, (double)beta2, step_host, step_dev, gsign, grad_scale, wd_factor, lr_dev, weight_decay);
  wd_factor = sc.wd_factor;
  const MT lr_eff = (MT)sc.lr_eff, inv_bc1 = (MT)sc.inv_bc1, inv_bc2 = (MT)sc.inv_bc2;
  const float gs = sc.gs;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    MT mi = m[i], vi = v[i];
    float np_ = adam_one<MT>(variant, p[i], g[i] * gs, mi, vi, beta1, beta2, eps, lr_eff, inv_bc1, inv_bc2, wd_factor, wd_grad);
    m[i] = mi;
    v[i] = vi;
    p[i] = np_;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(np_);
  }
}

// fp32-moment variant, 128-bit accesses: 4 x LDG.128 + 3 x STG.128 + 1 x STG.64 per 4 parameters (30 B/param), two
// independent vectors per thread and iteration so that 8 loads are in flight per thread
__global__ void __launch_bounds__(256)
k_adam_vec4(int variant, float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
            uint2* __restrict__ p_bf16, int64_t n4, double lr, float beta1, float beta2, float eps, float wd_factor, float wd_grad,
            int step_host, const int* step_dev, float gsign, const float* grad_scale, const double* lr_dev, double weight_decay) {
  __shared__ AdamScalars sc;
  adam_scalars(&sc, variant, lr, (double)beta1, (double)beta2, step_host, step_dev, gsign, grad_scale, wd_factor, lr_dev, weight_decay);
  wd_factor = sc.wd_factor;
  const float lr_eff = (float)sc.lr_eff, inv_bc1 = (float)sc.inv_bc1, inv_bc2 = (float)sc.inv_bc2;
  const float gs = sc.gs;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const int64_t i1 = i0 + stride;
    const bool two = i1 < n4;
    float4 pa = p[i0], ga = g[i0], ma = m[i0], va = v[i0];
    float4 pb, gb, mb, vb;
    if (two) { pb = p[i1]; gb = g[i1]; mb = m[i1]; vb = v[i1]; }
#define ADAM_LANE(P, G, M, V) P = adam_one<float>(variant, P, G * gs, M, V, beta1, beta2, eps, lr_eff, inv_bc1, inv_bc2, wd_factor, wd_grad)
    ADAM_LANE(pa.x, ga.x, ma.x, va.x); ADAM_LANE(pa.y, ga.y, ma.y, va.y); ADAM_LANE(pa.z, ga.z, ma.z, va.z); ADAM_LANE(pa.w, ga.w, ma.w, va.w);
    m[i0] = ma; v[i0] = va; p[i0] = pa;
    if (p_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pa.x, pa.y), hi = __floats2bfloat162_rn(pa.z, pa.w);
      p_bf16[i0] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
    if (two) {
      ADAM_LANE(pb.x, gb.x, mb.x, vb.x); ADAM_LANE(pb.y, gb.y, mb.y, vb.y); ADAM_LANE(pb.z, gb.z, mb.z, vb.z); ADAM_LANE(pb.w, gb.w, mb.w, vb.w);
      m[i1] = mb; v[i1] = vb; p[i1] = pb;
      if (p_bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(pb.x, pb.y), hi = __floats2bfloat162_rn(pb.z, pb.w);
        p_bf16[i1] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
#undef ADAM_LANE
  }
}

// variant: 0 = Adam, 1 = AdamW. moment_dtype: NFB_F32 or NFB_F64. step is 1-based (after increment); when step_dev
// (a device int) is given it overrides `step`.
// lr_dev (a device double, optional) overrides `lr` the same way: learning-rate schedulers keep working across replays
// of a captured training step.
NFB_API int nfb_adam_step_ex(int variant, int moment_dtype, float* p, const float* g, void* m, void* v, void* p_bf16, int64_t n,
                             double lr, double beta1, double beta2, double eps, double weight_decay, int step, int maximize,
                             const float* grad_scale, const int* step_dev, const double* lr_dev) {
  if (n == 0) return 0;
  float wd_factor = (variant == 1 && weight_decay != 0.0) ? (float)(1.0 - lr * weight_decay) : 1.f;
  float wd_grad = (variant == 0) ? (float)weight_decay : 0.f;
  float gsign = maximize ? -1.f : 1.f;
  cudaStream_t s = nfb_cur_stream();
  if (moment_dtype == NFB_F64) {
    k_adam<double><<<nfb_grid_for(n, 256, 8), 256, 0, s>>>(variant, p, g, (double*)m, (double*)v, (__nv_bfloat16*)p_bf16, 0, n, lr, beta1, beta2, eps, wd_factor, wd_grad, step, step_dev, gsign, grad_scale, lr_dev, weight_decay);
  } else if (moment_dtype == NFB_F32) {
    bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0) && ((uintptr_t)p_bf16 % 8 == 0);
    int64_t n4 = vec ? n / 4 : 0;
    if (n4 > 0) {
      k_adam_vec4<<<nfb_grid_for(nfb_cdiv(n4, 2), 256, 8), 256, 0, s>>>(variant, (float4*)p, (const float4*)g, (float4*)m, (float4*)v, (uint2*)p_bf16, n4, lr,
                                                                       (float)beta1, (float)beta2, (float)eps, wd_factor, wd_grad, step, step_dev, gsign, grad_scale, lr_dev, weight_decay);
      if (n4 * 4 < n) nfb_count_launch();
    }
    if (n4 * 4 < n)
      k_adam<float><<<1, 256, 0, s>>>(variant, p, g, (float*)m, (float*)v, (__nv_bfloat16*)p_bf16, n4 * 4, n, lr, (float)beta1, (float)beta2, (float)eps, wd_factor, wd_grad, step, step_dev, gsign, grad_scale, lr_dev, weight_decay);
  } else return NFB_FAIL("nfb_adam_step: moment dtype %d", moment_dtype);
  NFB_LAUNCH_CHECK();
  return 0;
}

NFB_API int nfb_adam_step(int variant, int moment_dtype, float* p, const float* g, void* m, void* v, void* p_bf16, int64_t n,
                          double lr, double beta1, double beta2, double eps, double weight_decay, int step, int maximize,
                          const float* grad_scale, const int* step_dev) {
  return nfb_adam_step_ex(variant, moment_dtype, p, g, m, v, p_bf16, n, lr, beta1, beta2, eps, weight_decay, step, maximize, grad_scale, step_dev, nullptr);
}

__global__ void k_add_i32(int* p, int delta) { *p += delta; }
// *counter += delta on the device (optimizer step counters that live inside a captured graph)
NFB_API int nfb_add_i32(int* counter, int delta) {
  k_add_i32<<<1, 1, 0, nfb_cur_stream()>>>(counter, delta);
  NFB_LAUNCH_CHECK();
  return 0;
}

// SGD with momentum / dampening / nesterov / L2 (optim/sgd.py:87-115)
__global__ void k_sgd(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, __nv_bfloat16* __restrict__ p_bf16, int64_t n,
                      float lr, float momentum, float dampening, float wd, int nesterov, int first_step, float gsign, const double* lr_dev) {
  if (lr_dev) lr = (float)*lr_dev;     // device-side learning rate: schedulers act on replays of a captured step
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i], gi = g[i] * gsign;
    if (wd != 0.f) gi = __fadd_rn(gi, __fmul_rn(wd, pi));
    if (momentum != 0.f) {
      float b = first_step ? gi : __fadd_rn(__fmul_rn(momentum, buf[i]), __fmul_rn(1.f - dampening, gi));
      buf[i] = b;
      gi = nesterov ? __fadd_rn(gi, __fmul_rn(momentum, b)) : b;
    }
    float np_ = __fsub_rn(pi, __fmul_rn(lr, gi));
    p[i] = np_;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(np_);
  }
}
NFB_API int nfb_sgd_step_ex(float* p, const float* g, float* buf, void* p_bf16, int64_t n, double lr, double momentum, double dampening,
                            double weight_decay, int nesterov, int first_step, int maximize, const double* lr_dev) {
  if (n == 0) return 0;
  k_sgd<<<nfb_grid_for(n, 256, 8), 256, 0, nfb_cur_stream()>>>(p, g, buf, (__nv_bfloat16*)p_bf16, n, (float)lr, (float)momentum, (float)dampening,
                                                                 (float)weight_decay, nesterov, first_step, maximize ? -1.f : 1.f, lr_dev);
  NFB_LAUNCH_CHECK();
  return 0;
}
NFB_API int nfb_sgd_step(float* p, const float* g, float* buf, void* p_bf16, int64_t n, double lr, double momentum, double dampening,
                         double weight_decay, int nesterov, int first_step, int maximize) {
  return nfb_sgd_step_ex(p, g, buf, p_bf16, n, lr, momentum, dampening, weight_decay, nesterov, first_step, maximize, nullptr);
}

// ---- global-norm clipping: sumsq accumulates sum(g^2) over one or more buffers in double ----
__global__ void k_sumsq_partial(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = (double)g[i];
    acc += v * v;
  }
  acc = warp_sum_d(acc);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = acc;
  __syncthreads();
  if (w == 0) {
    double r = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    r = warp_sum_d(r);
    if (lane == 0) partial[blockIdx.x] = r;
  }
}
__global__ void k_sumsq_final(const double* __restrict__ partial, int nblocks, double* __restrict__ total, int accumulate) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partial[i];
  acc = warp_sum_d(acc);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) r += red[k];
    *total = accumulate ? *total + r : r;
  }
}
// total (device double) = (accumulate ? total : 0) + sum(g^2)
NFB_API int nfb_sumsq(const float* g, int64_t n, double* total, int accumulate) {
  int grid = nfb_grid_for(n > 0 ? n : 1, 1024, 2);
  double* partial = nullptr;
  if (nfb_malloc(sizeof(double) * grid, (void**)&partial)) return -1;
  cudaStream_t s = nfb_cur_stream();
  k_sumsq_partial<<<grid, 256, 0, s>>>(g, n, partial);
  nfb_count_launch();
  k_sumsq_final<<<1, 256, 0, s>>>(partial, grid, total, accumulate);
  nfb_free(partial);
  NFB_LAUNCH_CHECK();
  return 0;
}
// coef = min(1, max_norm / (sqrt(total) + eps)); norm_out = sqrt(total)   (both device floats)
__global__ void k_clip_coef(const double* __restrict__ total, float max_norm, float eps, float* __restrict__ coef, float* __restrict__ norm_out) {
  double nrm = sqrt(*total);
  if (norm_out) *norm_out = (float)nrm;
  double c = (double)max_norm / (nrm + (double)eps);
  *coef = c < 1.0 ? (float)c : 1.f;
}
NFB_API int nfb_clip_coef(const double* total, float max_norm, float eps, float* coef, float* norm_out) {
  k_clip_coef<<<1, 1, 0, nfb_cur_stream()>>>(total, max_norm, eps, coef, norm_out);
  NFB_LAUNCH_CHECK();
  return 0;
}
// g *= *coef  (device scalar), vectorised
__global__ void k_scale_by(float* __restrict__ g, int64_t n4, int64_t n, const float* __restrict__ coef) {
  float c = *coef;
  if (c == 1.f) return;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride) {
    float4 a = reinterpret_cast<float4*>(g)[v];
    a.x *= c; a.y *= c; a.z *= c; a.w *= c;
    reinterpret_cast<float4*>(g)[v] = a;
  }
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  g[t] *= c;
}
NFB_API int nfb_scale_by_device_scalar(float* g, int64_t n, const float* coef) {
  if (n == 0) return 0;
  bool vec = ((uintptr_t)g % 16 == 0);
  int64_t n4 = vec ? n / 4 : 0;
  k_scale_by<<<nfb_grid_for(vec && n4 ? n4 : n, 256), 256, 0, nfb_cur_stream()>>>(g, n4, n, coef);
  NFB_LAUNCH_CHECK();
  return 0;
}

// ---- loss-scale unscaling with per-parameter norm clipping (optimization/grad_scaler.py:131-196) ----
// The gradient buffer is cut into chunks of at most UNSCALE_CHUNK elements, none crossing a parameter boundary; the host
// builds the chunk table (seg, start, len) once per parameter set.  Pass 1: per parameter, sum((g/scale)^2) in double and
// the smallest parameter index holding a non-finite unscaled value.  Pass 2: g = g/scale * clip(param) for the
// parameters in front of that index (the reference stops at the first bad parameter).
struct UnscaleChunk { int seg; int len; int64_t start; };
__global__ void __launch_bounds__(256)
k_unscale_stats(const float* __restrict__ g, const UnscaleChunk* __restrict__ chunks, float scale, double* __restrict__ norms2,
                int* __restrict__ first_bad) {
  const UnscaleChunk c = chunks[blockIdx.x];
  const float* p = g + c.start;
  double acc = 0.0;
  bool bad = false;
  const int head = (int)(((16 - ((uintptr_t)p & 15)) & 15) >> 2);   // scalars in front of the first 16-byte boundary
  const int h = head < c.len ? head : c.len;
  const int n4 = (c.len - h) >> 2;
  for (int i = threadIdx.x; i < h; i += blockDim.x) {
    float u = __fdiv_rn(p[i], scale);
    bad |= !isfinite(u);
    acc += (double)u * (double)u;
  }
  const float4* p4 = reinterpret_cast<const float4*>(p + h);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 a = p4[i];
    float u0 = __fdiv_rn(a.x, scale), u1 = __fdiv_rn(a.y, scale), u2 = __fdiv_rn(a.z, scale), u3 = __fdiv_rn(a.w, scale);
    bad |= !(isfinite(u0) && isfinite(u1) && isfinite(u2) && isfinite(u3));
    acc += (double)u0 * u0 + (double)u1 * u1 + (double)u2 * u2 + (double)u3 * u3;
  }
  for (int i = h + n4 * 4 + threadIdx.x; i < c.len; i += blockDim.x) {
    float u = __fdiv_rn(p[i], scale);
    bad |= !isfinite(u);
    acc += (double)u * (double)u;
  }
  __shared__ double red[8];
  __shared__ int any_bad;
  if (threadIdx.x == 0) any_bad = 0;
  __syncthreads();
  if (bad) any_bad = 1;
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int k = 0; k < 8; ++k) r += red[k];
    if (any_bad) atomicMin(first_bad, c.seg);
    else atomicAdd(&norms2[c.seg], r);
  }
}
__global__ void __launch_bounds__(256)
k_unscale_apply(float* __restrict__ g, const UnscaleChunk* __restrict__ chunks, float scale, float clip_threshold,
                const double* __restrict__ norms2, const int* __restrict__ first_bad) {
  const UnscaleChunk c = chunks[blockIdx.x];
  if (c.seg >= *first_bad) return;
  float coef = 1.f;
  if (clip_threshold > 0.f) {
    const float nrm = (float)sqrt(norms2[c.seg]);                 // np.linalg.norm of an fp32 array is an fp32 scalar
    if (nrm > clip_threshold) coef = __fdiv_rn(clip_threshold, __fadd_rn(nrm, 1e-6f));
  }
  float* p = g + c.start;
  const int head = (int)(((16 - ((uintptr_t)p & 15)) & 15) >> 2);
  const int h = head < c.len ? head : c.len;
  const int n4 = (c.len - h) >> 2;
  for (int i = threadIdx.x; i < h; i += blockDim.x) p[i] = __fmul_rn(__fdiv_rn(p[i], scale), coef);
  float4* p4 = reinterpret_cast<float4*>(p + h);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 a = p4[i];
    a.x = __fmul_rn(__fdiv_rn(a.x, scale), coef); a.y = __fmul_rn(__fdiv_rn(a.y, scale), coef);
    a.z = __fmul_rn(__fdiv_rn(a.z, scale), coef); a.w = __fmul_rn(__fdiv_rn(a.w, scale), coef);
    p4[i] = a;
  }
  for (int i = h + n4 * 4 + threadIdx.x; i < c.len; i += blockDim.x) p[i] = __fmul_rn(__fdiv_rn(p[i], scale), coef);
}
__global__ void k_unscale_init(double* norms2, int nseg, int* first_bad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += gridDim.x * blockDim.x) norms2[i] = 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *first_bad = nseg;
}
// chunks: device array of nchunks {int seg; int len; int64 start} records (16 bytes each).  norms2[nseg] receives
// sum((g/scale)^2) per parameter, *first_bad the first parameter index with a non-finite value (nseg if none).
NFB_API int nfb_grad_unscale(float* g, const void* chunks, int nchunks, int nseg, float scale, float clip_threshold, double* norms2,
                             int* first_bad) {
  if (nseg <= 0) return 0;
  cudaStream_t s = nfb_cur_stream();
  k_unscale_init<<<nfb_cdiv(nseg, 256), 256, 0, s>>>(norms2, nseg, first_bad);
  if (nchunks > 0) {
    nfb_count_launch();
    k_unscale_stats<<<nchunks, 256, 0, s>>>(g, (const UnscaleChunk*)chunks, scale, norms2, first_bad);
    nfb_count_launch();
    k_unscale_apply<<<nchunks, 256, 0, s>>>(g, (const UnscaleChunk*)chunks, scale, clip_threshold, norms2, first_bad);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}

// ---- Lion (optim/lion.py:129-175): c = b1*m + (1-b1)*g ; p -= clip(fp32(lr*sign(c) + lr*wd*p), +-1) ; m = b2*m + (1-b2)*g ----
// The interpolation / update are formed in the moment dtype (fp64 = the reference's arithmetic, fp32 = throughput mode), the
// parameter subtraction in fp32 as NumPy does it; non-finite parameters are scrubbed (nan -> 0, +-inf -> +-1).
template <typename MT>
__device__ __forceinline__ float lion_one(float pi, float gi, MT& mi, MT beta1, MT beta2, double lr, double lr_wd) {
  const MT gd = (MT)gi;
  const MT c = beta1 * mi + (MT(1) - beta1) * gd;
  const double sgn = (c > MT(0)) ? 1.0 : ((c < MT(0)) ? -1.0 : ((c == c) ? 0.0 : (double)c));   // np.sign keeps NaN
  float upd = (float)(lr * sgn + lr_wd * (double)pi);
  if (upd == upd) upd = fminf(fmaxf(upd, -1.f), 1.f);
  mi = beta2 * mi + (MT(1) - beta2) * gd;
  float np_ = __fsub_rn(pi, upd);
  if (!isfinite(np_)) np_ = (np_ != np_) ? 0.f : (np_ > 0.f ? 1.f : -1.f);
  return np_;
}
template <typename MT>
__global__ void k_lion(float* __restrict__ p, const float* __restrict__ g, MT* __restrict__ m, __nv_bfloat16* __restrict__ p_bf16, int64_t n0,
                       int64_t n, double lr, MT beta1, MT beta2, double wd, float gsign, const float* grad_scale, const double* lr_dev) {
  if (lr_dev) lr = *lr_dev;
  const double lr_wd = lr * wd;
  const float gs = (grad_scale ? *grad_scale : 1.f) * gsign;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    MT mi = m[i];
    const float np_ = lion_one<MT>(p[i], g[i] * gs, mi, beta1, beta2, lr, lr_wd);
    m[i] = mi;
    p[i] = np_;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(np_);
  }
}
// fp32 momentum, 128-bit accesses: 3 x LDG.128 + 2 x STG.128 + 1 x STG.64 per 4 parameters (22 B/param)
__global__ void __launch_bounds__(256)
k_lion_vec4(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, uint2* __restrict__ p_bf16, int64_t n4, double lr,
            float beta1, float beta2, double wd, float gsign, const float* grad_scale, const double* lr_dev) {
  if (lr_dev) lr = *lr_dev;
  const double lr_wd = lr * wd;
  const float gs = (grad_scale ? *grad_scale : 1.f) * gsign;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pa = p[i], ga = g[i], ma = m[i];
    pa.x = lion_one<float>(pa.x, ga.x * gs, ma.x, beta1, beta2, lr, lr_wd);
    pa.y = lion_one<float>(pa.y, ga.y * gs, ma.y, beta1, beta2, lr, lr_wd);
    pa.z = lion_one<float>(pa.z, ga.z * gs, ma.z, beta1, beta2, lr, lr_wd);
    pa.w = lion_one<float>(pa.w, ga.w * gs, ma.w, beta1, beta2, lr, lr_wd);
    m[i] = ma;
    p[i] = pa;
    if (p_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pa.x, pa.y), hi = __floats2bfloat162_rn(pa.z, pa.w);
      p_bf16[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}
// moment_dtype: NFB_F32 or NFB_F64.  grad_scale (device float) and lr_dev (device double) are optional, as in nfb_adam_step_ex.
NFB_API int nfb_lion_step(int moment_dtype, float* p, const float* g, void* m, void* p_bf16, int64_t n, double lr, double beta1, double beta2,
                          double weight_decay, int maximize, const float* grad_scale, const double* lr_dev) {
  if (n == 0) return 0;
  const float gsign = maximize ? -1.f : 1.f;
  cudaStream_t s = nfb_cur_stream();
  if (moment_dtype == NFB_F64) {
    k_lion<double><<<nfb_grid_for(n, 256, 8), 256, 0, s>>>(p, g, (double*)m, (__nv_bfloat16*)p_bf16, 0, n, lr, beta1, beta2, weight_decay, gsign, grad_scale, lr_dev);
  } else if (moment_dtype == NFB_F32) {
    const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m) % 16 == 0) && ((uintptr_t)p_bf16 % 8 == 0);
    const int64_t n4 = vec ? n / 4 : 0;
    if (n4 > 0) {
      k_lion_vec4<<<nfb_grid_for(n4, 256, 8), 256, 0, s>>>((float4*)p, (const float4*)g, (float4*)m, (uint2*)p_bf16, n4, lr, (float)beta1, (float)beta2,
                                                          weight_decay, gsign, grad_scale, lr_dev);
      if (n4 * 4 < n) nfb_count_launch();
    }
    if (n4 * 4 < n)
      k_lion<float><<<1, 256, 0, s>>>(p, g, (float*)m, (__nv_bfloat16*)p_bf16, n4 * 4, n, lr, (float)beta1, (float)beta2, weight_decay, gsign, grad_scale, lr_dev);
  } else return NFB_FAIL("nfb_lion_step: moment dtype %d", moment_dtype);
  NFB_LAUNCH_CHECK();
  return 0;
}
