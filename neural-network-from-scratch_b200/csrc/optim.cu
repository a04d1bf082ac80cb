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

template <typename MT>
__global__ void k_adam(int variant, float* __restrict__ p, const float* __restrict__ g, MT* __restrict__ m, MT* __restrict__ v,
                       __nv_bfloat16* __restrict__ p_bf16, int64_t n, double lr, MT beta1, MT beta2, MT eps, float wd_factor, float wd_grad,
                       int step_host, const int* __restrict__ step_dev, float gsign, const float* __restrict__ grad_scale) {
  // bias corrections from the step count: a host value, or a device counter so that a captured CUDA graph of the
  // training step stays valid from one replay to the next
  const int step = step_dev ? *step_dev : step_host;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  double lr_d;
  if (variant == 0) lr_d = (lr <= 0.02 && lr > 0.0) ? lr * fmin(5.0, 0.1 / lr) : lr;
  else lr_d = lr / bc1;
  const MT lr_eff = (MT)lr_d, inv_bc1 = (MT)(1.0 / bc1), inv_bc2 = (MT)(1.0 / bc2);
  float gs = grad_scale ? *grad_scale : 1.f;
  gs *= gsign;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * gs;
    if (variant == 0 && wd_grad != 0.f) gi = __fadd_rn(gi, __fmul_rn(wd_grad, pi));  // Adam: L2 term in the gradient (fp32)
    MT gd = (MT)gi;
    MT mi = beta1 * m[i] + (MT(1) - beta1) * gd;
    MT vi = beta2 * v[i] + (MT(1) - beta2) * (gd * gd);
    m[i] = mi;
    v[i] = vi;
    MT mhat = mi * inv_bc1;
    MT vhat = vi * inv_bc2;
    MT denom = sqrt(vhat + eps);
    float upd = (float)(lr_eff * mhat / denom);
    if (upd == upd) upd = fminf(fmaxf(upd, -10.f), 10.f);  // np.clip keeps NaN (scrubbed below)
    float np_;
    if (variant == 1 && wd_factor != 1.f) np_ = __fsub_rn(__fmul_rn(pi, wd_factor), upd);  // AdamW decoupled decay
    else np_ = __fsub_rn(pi, upd);
    if (!isfinite(np_)) np_ = (np_ != np_) ? 0.f : (np_ > 0.f ? 1.f : -1.f);
    p[i] = np_;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(np_);
  }
}

// variant: 0 = Adam, 1 = AdamW. moment_dtype: NFB_F32 or NFB_F64. step is 1-based (after increment); when step_dev
// (a device int) is given it overrides `step`.
NFB_API int nfb_adam_step(int variant, int moment_dtype, float* p, const float* g, void* m, void* v, void* p_bf16, int64_t n,
                          double lr, double beta1, double beta2, double eps, double weight_decay, int step, int maximize,
                          const float* grad_scale, const int* step_dev) {
  if (n == 0) return 0;
  float wd_factor = (variant == 1 && weight_decay != 0.0) ? (float)(1.0 - lr * weight_decay) : 1.f;
  float wd_grad = (variant == 0) ? (float)weight_decay : 0.f;
  float gsign = maximize ? -1.f : 1.f;
  int grid = nfb_grid_for(n, 256, 8);
  cudaStream_t s = nfb_cur_stream();
  if (moment_dtype == NFB_F64)
    k_adam<double><<<grid, 256, 0, s>>>(variant, p, g, (double*)m, (double*)v, (__nv_bfloat16*)p_bf16, n, lr, beta1, beta2, eps, wd_factor, wd_grad, step, step_dev, gsign, grad_scale);
  else if (moment_dtype == NFB_F32)
    k_adam<float><<<grid, 256, 0, s>>>(variant, p, g, (float*)m, (float*)v, (__nv_bfloat16*)p_bf16, n, lr, (float)beta1, (float)beta2, (float)eps, wd_factor, wd_grad, step, step_dev, gsign, grad_scale);
  else return NFB_FAIL("nfb_adam_step: moment dtype %d", moment_dtype);
  NFB_LAUNCH_CHECK();
  return 0;
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
                      float lr, float momentum, float dampening, float wd, int nesterov, int first_step, float gsign) {
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
NFB_API int nfb_sgd_step(float* p, const float* g, float* buf, void* p_bf16, int64_t n, double lr, double momentum, double dampening,
                         double weight_decay, int nesterov, int first_step, int maximize) {
  if (n == 0) return 0;
  k_sgd<<<nfb_grid_for(n, 256, 8), 256, 0, nfb_cur_stream()>>>(p, g, buf, (__nv_bfloat16*)p_bf16, n, (float)lr, (float)momentum, (float)dampening,
                                                                 (float)weight_decay, nesterov, first_step, maximize ? -1.f : 1.f);
  NFB_LAUNCH_CHECK();
  return 0;
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
