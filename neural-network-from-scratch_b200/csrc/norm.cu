// LayerNorm / RMSNorm forward + backward, one warp per row with the row held in registers
// (two-pass statistics like np.mean / np.var, warp-shuffle reductions, 128-bit accesses).
// Reference math: nn/normalization.py:101-200 (LayerNorm) and :269-341 (RMSNorm);
// models/language/gpt2.py:98-116 (the GPT-2 local RMSNorm). HBM-bound: fwd 8 B/elem, bwd 12 B/elem
// in fp32.
//   x / residual stream : fp32 or bf16 (TX)       y and dy : fp32 or bf16 (TY)      dx : TX
// Optional fusions: `res` (forward: s = x + res is normalised and written back to `sum_out`) and
// `dres` (backward: dx += dres, the gradient arriving over the skip connection).
#include "common.cuh"
#include <stdlib.h>

#define WARPS_PER_BLOCK 4

template <int VPL, typename TX, typename TY, bool RMS>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_norm_fwd(const TX* __restrict__ x, const TX* __restrict__ res, TX* __restrict__ sum_out, const float* __restrict__ gamma,
           const float* __restrict__ beta, TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
           int64_t rows, int cols, float eps) {
  nfb_pdl_enter();
  int lane = threadIdx.x & 31;
  int64_t warp_global = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  float inv_n = 1.f / (float)cols;
  for (int64_t row = warp_global; row < rows; row += nwarps) {
    const TX* xr = x + row * cols;
    f4 v[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int c = k * 128 + lane * 4;
      if (c < cols) {
        v[k] = ld4<TX>(xr + c);
        if (res) {
          f4 r = ld4<TX>(res + row * cols + c);
          v[k].x += r.x; v[k].y += r.y; v[k].z += r.z; v[k].w += r.w;
          st4<TX>(sum_out + row * cols + c, v[k]);
        }
      } else v[k] = {0.f, 0.f, 0.f, 0.f};
    }
    float mean = 0.f;
    if (!RMS) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      mean = warp_sum(s) * inv_n;
    }
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int c = k * 128 + lane * 4;
      if (c < cols) {
        float a = v[k].x - mean, b = v[k].y - mean, cc = v[k].z - mean, d = v[k].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    float var = warp_sum(q) * inv_n;
    float rstd = 1.f / sqrtf(var + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int c = k * 128 + lane * 4;
      if (c < cols) {
        float4 g = gamma ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 b = beta ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        f4 o;
        o.x = g.x * ((v[k].x - mean) * rstd) + b.x;
        o.y = g.y * ((v[k].y - mean) * rstd) + b.y;
        o.z = g.z * ((v[k].z - mean) * rstd) + b.z;
        o.w = g.w * ((v[k].w - mean) * rstd) + b.w;
        st4<TY>(y + row * cols + c, o);
      }
    }
  }
}

// backward: 8 warps per block, (at most) two blocks per SM, each warp walks rows with a grid stride.  Its lanes own
// fixed columns, so dgamma / dbeta partials accumulate in registers over all of the warp's rows and are combined
// through shared memory once per block into `partial` ([gridDim.x][NP][cols], NP = 1 for RMSNorm (no beta), 2 for
// LayerNorm); k_norm_param_reduce then sums the block partials in a fixed order (deterministic, one launch).
// The row is kept RAW in registers (x as loaded, dy packed as loaded) and x_hat / g*gamma are recomputed in the
// second phase: that keeps the kernel at <= 128 registers so that 16 warps per SM have their rows in flight.
#define BWD_WARPS 8
template <int VPL, typename TX, typename TY, bool RMS>
__global__ void __launch_bounds__(BWD_WARPS * 32, 2)
k_norm_bwd(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean_in,
           const float* __restrict__ rstd_in, const TX* __restrict__ dres, TX* __restrict__ dx, float* __restrict__ partial,
           int64_t rows, int cols, float clip, __nv_bfloat16* __restrict__ dx_twin) {
  nfb_pdl_enter();
  constexpr int NP = RMS ? 1 : 2;
  __shared__ float sh[BWD_WARPS][VPL * 128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * BWD_WARPS + w;
  const int64_t nwarps = (int64_t)gridDim.x * BWD_WARPS;
  const float inv_n = 1.f / (float)cols;
  f4 dg[VPL], db[RMS ? 1 : VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) dg[k] = {0.f, 0.f, 0.f, 0.f};
  if (!RMS) {
#pragma unroll
    for (int k = 0; k < (RMS ? 1 : VPL); ++k) db[k] = {0.f, 0.f, 0.f, 0.f};
  }
  for (int64_t row = warp_global; row < rows; row += nwarps) {
    const float mean = (RMS || !mean_in) ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    f4 xv[VPL], gv[VPL];
    // (prefetching this warp's NEXT row into L2 here was tried in round 2: 13.7 -> 15.2 us under ncu, reverted)
    if (dres) {   // the skip-connection gradient is only needed at the very end: start pulling its lines into L2 / L1 now
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c = k * 128 + lane * 4;
        if (c < cols) asm volatile("prefetch.global.L2 [%0];" ::"l"(dres + row * cols + c));
      }
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = k * 128 + lane * 4;
      if (c < cols) {
        xv[k] = ld4<TX>(x + row * cols + c);
        gv[k] = ld4<TY>(dy + row * cols + c);
      } else {
        xv[k] = {mean, mean, mean, mean};
        gv[k] = {0.f, 0.f, 0.f, 0.f};
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = k * 128 + lane * 4;
      const float4 gm = (gamma && c < cols) ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
      // x_hat in place of x; g stays raw (dgamma / dbeta need it), g*gamma is formed on the fly
      xv[k] = {(xv[k].x - mean) * rstd, (xv[k].y - mean) * rstd, (xv[k].z - mean) * rstd, (xv[k].w - mean) * rstd};
      dg[k].x += gv[k].x * xv[k].x; dg[k].y += gv[k].y * xv[k].y; dg[k].z += gv[k].z * xv[k].z; dg[k].w += gv[k].w * xv[k].w;
      if (!RMS) { db[k].x += gv[k].x; db[k].y += gv[k].y; db[k].z += gv[k].z; db[k].w += gv[k].w; }
      gv[k] = {gv[k].x * gm.x, gv[k].y * gm.y, gv[k].z * gm.z, gv[k].w * gm.w};
      s1 += (gv[k].x + gv[k].y) + (gv[k].z + gv[k].w);
      s2 += (gv[k].x * xv[k].x + gv[k].y * xv[k].y) + (gv[k].z * xv[k].z + gv[k].w * xv[k].w);
    }
    const float m1 = RMS ? 0.f : warp_sum(s1) * inv_n;
    const float m2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = k * 128 + lane * 4;
      if (c < cols) {
        f4 o;
        o.x = rstd * (gv[k].x - m1 - xv[k].x * m2);
        o.y = rstd * (gv[k].y - m1 - xv[k].y * m2);
        o.z = rstd * (gv[k].z - m1 - xv[k].z * m2);
        o.w = rstd * (gv[k].w - m1 - xv[k].w * m2);
        if (dres) {
          f4 r = ld4<TX>(dres + row * cols + c);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (clip > 0.f) {  // the reference's +-clip on every upstream gradient, fused into the producer
          o.x = fminf(fmaxf(o.x, -clip), clip); o.y = fminf(fmaxf(o.y, -clip), clip);
          o.z = fminf(fmaxf(o.z, -clip), clip); o.w = fminf(fmaxf(o.w, -clip), clip);
        }
        st4<TX>(dx + row * cols + c, o);
        if (dx_twin) st4<__nv_bfloat16>(dx_twin + row * cols + c, o);  // bf16 copy for the next dgrad/wgrad GEMMs
      }
    }
  }
  if (!partial) return;
#pragma unroll
  for (int which = 0; which < NP; ++which) {
    if (which) __syncthreads();
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const f4 a = which == 0 ? dg[k] : db[RMS ? 0 : k];
      *reinterpret_cast<float4*>(&sh[w][k * 128 + lane * 4]) = make_float4(a.x, a.y, a.z, a.w);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int ww = 0; ww < BWD_WARPS; ++ww) a += sh[ww][c];
      partial[((int64_t)blockIdx.x * NP + which) * cols + c] = a;
    }
  }
}

// partial is [nblocks][NP][cols]: block (8, 32) owns 8 columns; threadIdx.y strides the row-blocks (<= 10 independent
// loads per thread for 296 row-blocks, all in flight at once), then a fixed-order shared-memory sum over threadIdx.y and
// a write / accumulate into dgamma (which = 0) or dbeta (which = 1).  Deterministic.
__device__ __forceinline__ void norm_param_reduce_body(const float* __restrict__ partial, int nblocks, int np, int cols, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta, int accumulate, int block) {
  __shared__ float red[32][9];
  const int i = block * 8 + threadIdx.x;   // index into [np][cols]
  const int total = np * cols;
  float a = 0.f;
  if (i < total) {
#pragma unroll 4
    for (int b = threadIdx.y; b < nblocks; b += 32) a += __ldcg(partial + (int64_t)b * total + i);
  }
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && i < total) {
    float r = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 32; ++k) r += red[k][threadIdx.x];
    const int which = i / cols, c = i % cols;
    float* dst = which == 0 ? dgamma : dbeta;
    if (dst) dst[c] = accumulate ? dst[c] + r : r;
  }
}
__global__ void __launch_bounds__(256)
k_norm_param_reduce(const float* __restrict__ partial, int nblocks, int np, int cols, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, int accumulate) {
  nfb_pdl_enter();
  norm_param_reduce_body(partial, nblocks, np, cols, dgamma, dbeta, accumulate, blockIdx.x);
}
// Up to NR_BATCH parameter reductions in ONE launch (the auxiliary-stream path queues them: nothing but the optimizer / the
// gradient all-reduce reads dgamma / dbeta, so the reductions of several norm layers wait for a flush point and go together).
constexpr int NR_BATCH = 8;
struct NormReduceJob { const float* partial; float* dgamma; float* dbeta; int nblocks, np, cols, accumulate; };
struct NormReduceBatch { int n; int blk0[NR_BATCH + 1]; NormReduceJob job[NR_BATCH]; };
__global__ void __launch_bounds__(256)
k_norm_param_reduce_batched(const __grid_constant__ NormReduceBatch b) {
  int j = 0;
  while (j + 1 < b.n && (int)blockIdx.x >= b.blk0[j + 1]) ++j;
  const NormReduceJob& q = b.job[j];
  norm_param_reduce_body(q.partial, q.nblocks, q.np, q.cols, q.dgamma, q.dbeta, q.accumulate, (int)blockIdx.x - b.blk0[j]);
}

// generic fallback (cols not a multiple of 4 or > 1024): one block per row, three passes.
template <bool RMS>
__global__ void k_norm_fwd_generic(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int cols, float eps) {
  __shared__ float red[32];
  int64_t row = blockIdx.x;
  const float* xr = x + row * cols;
  float mean = 0.f;
  if (!RMS) {
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) s += xr[c];
    mean = block_sum(s, red) / (float)cols;
  }
  float q = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) { float d = xr[c] - mean; q += d * d; }
  float var = block_sum(q, red) / (float)cols;
  float rstd = 1.f / sqrtf(var + eps);
  if (threadIdx.x == 0) { if (mean_out) mean_out[row] = mean; rstd_out[row] = rstd; }
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float h = (xr[c] - mean) * rstd;
    y[row * cols + c] = (gamma ? gamma[c] : 1.f) * h + (beta ? beta[c] : 0.f);
  }
}
template <bool RMS>
__global__ void k_norm_bwd_generic(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dx, int cols) {
  __shared__ float red[32];
  int64_t row = blockIdx.x;
  float mean = (RMS || !mean_in) ? 0.f : mean_in[row], rstd = rstd_in[row];
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float gh = dy[row * cols + c] * (gamma ? gamma[c] : 1.f);
    float xh = (x[row * cols + c] - mean) * rstd;
    s1 += gh;
    s2 += gh * xh;
  }
  float m1 = RMS ? 0.f : block_sum(s1, red) / (float)cols;
  float m2 = block_sum(s2, red) / (float)cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float gh = dy[row * cols + c] * (gamma ? gamma[c] : 1.f);
    float xh = (x[row * cols + c] - mean) * rstd;
    dx[row * cols + c] = rstd * (gh - m1 - xh * m2);
  }
}
// dgamma[c] = sum_rows dy*xhat ; dbeta[c] = sum_rows dy   (generic path)
template <bool RMS>
__global__ void k_norm_param_generic(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_in,
                                     const float* __restrict__ rstd_in, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int64_t rows, int cols, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float a = 0.f, b = 0.f;
  for (int64_t r = 0; r < rows; ++r) {
    float mean = (RMS || !mean_in) ? 0.f : mean_in[r];
    float g = dy[r * cols + c];
    a += g * ((x[r * cols + c] - mean) * rstd_in[r]);
    b += g;
  }
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + a : a;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + b : b;
}

static int pick_vpl(int cols) {
  if (cols % 4 != 0 || cols > 1024) return 0;
  int v = (cols + 127) / 128;
  if (v <= 1) return 1;
  if (v <= 2) return 2;
  if (v <= 4) return 4;
  if (v <= 6) return 6;
  return 8;
}

template <typename TX, typename TY, bool RMS>
static int norm_fwd_t(const void* x, const void* res, void* sum_out, const float* gamma, const float* beta, void* y,
                      float* mean, float* rstd, int64_t rows, int cols, float eps) {
  int vpl = pick_vpl(cols);
  cudaStream_t s = nfb_cur_stream();
  int grid = nfb_grid_for(rows, WARPS_PER_BLOCK, 8);
  if (const char* e = getenv("NFB_NORM_FWD_GRID")) {   // measurement hook (profiles/tools/sweep_norm.py)
    int g = atoi(e);
    if (g > 0 && g < grid) grid = g;
  }
#define LF(V) nfb_launch_pdl(k_norm_fwd<V, TX, TY, RMS>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, s, (const TX*)x, (const TX*)res, (TX*)sum_out, gamma, beta, (TY*)y, mean, rstd, rows, cols, eps)
  switch (vpl) {
    case 1: LF(1); break;
    case 2: LF(2); break;
    case 4: LF(4); break;
    case 6: LF(6); break;
    case 8: LF(8); break;
    default: return NFB_FAIL("norm fwd: cols=%d needs the generic path", cols);
  }
#undef LF
  NFB_LAUNCH_CHECK();
  return 0;
}

// ---- auxiliary-stream parameter reductions: the block partials of a backward go to one of NR_RING grow-only workspaces (not the
// stream-ordered allocator), the reduction is QUEUED, and norm_flush_pending launches up to NR_BATCH of them as one kernel on
// the auxiliary stream behind an event of the compute stream.  A workspace is fenced by the event of the launch that read it.
struct AuxWs { float* buf; size_t n; cudaEvent_t done; unsigned long long epoch; bool used; bool pending; };
constexpr int NR_RING = 2 * NR_BATCH;
static AuxWs g_nr_ws[NR_RING] = {};
static int g_nr_next = 0;
static NormReduceBatch g_nr_pending = {};
static AuxWs* g_nr_owner[NR_BATCH] = {};
static cudaStream_t g_nr_stream = nullptr;     // the compute stream the queued backward kernels ran on
static cudaEvent_t g_nr_ready = nullptr;

static int norm_flush_pending() {
  if (g_nr_pending.n == 0) return 0;
  NormReduceBatch& b = g_nr_pending;
  int blocks = 0;
  for (int j = 0; j < b.n; ++j) { b.blk0[j] = blocks; blocks += (b.job[j].np * b.job[j].cols + 7) / 8; }
  b.blk0[b.n] = blocks;
  cudaStream_t aux = nfb_aux_stream();
  cudaStream_t where = (aux && aux != g_nr_stream) ? aux : g_nr_stream;
  if (where != g_nr_stream) {
    if (!g_nr_ready) NFB_CUDA(cudaEventCreateWithFlags(&g_nr_ready, cudaEventDisableTiming));
    NFB_CUDA(cudaEventRecord(g_nr_ready, g_nr_stream));
    NFB_CUDA(cudaStreamWaitEvent(where, g_nr_ready, 0));
  }
  nfb_count_launch();
  k_norm_param_reduce_batched<<<blocks, dim3(8, 32), 0, where>>>(b);
  for (int j = 0; j < b.n; ++j) {
    AuxWs* w = g_nr_owner[j];
    NFB_CUDA(cudaEventRecord(w->done, where));
    w->used = true;
    w->pending = false;
    w->epoch = nfb_capture_epoch();
  }
  b.n = 0;
  NFB_LAUNCH_CHECK();
  return 0;
}
// Launch the queued norm-parameter reductions now (the host calls this before it joins / fences the auxiliary stream).
NFB_API int nfb_norm_flush() { return norm_flush_pending(); }

template <typename TX, typename TY, bool RMS>
static int norm_bwd_t(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dres,
                      void* dx, float* dgamma, float* dbeta, int accumulate, int64_t rows, int cols, float clip, void* dx_twin) {
  int vpl = pick_vpl(cols);
  cudaStream_t s = nfb_cur_stream();
  // two 8-warp blocks per SM, every warp with a whole row (x, dy, dres) in flight; a warp handles several rows so the
  // dgamma / dbeta partials amortise
  int64_t want = nfb_cdiv(rows, BWD_WARPS);
  // equal work per warp: with r = ceil(want / resident blocks) rows per warp, cdiv(want, r) blocks leave no warp a row short
  // (T = 4096: 256 blocks x 2 rows instead of 296 blocks x 1.73 -> 15.5 -> 13.5 us, profiles/r01_norm_grid_sweep.txt)
  const int64_t cap = 2LL * nfb_sm_count();
  int grid = (int)(want <= cap ? want : nfb_cdiv(want, nfb_cdiv(want, cap)));
  if (const char* e = getenv("NFB_NORM_BWD_GRID")) {   // measurement hook (profiles/tools/sweep_norm.py)
    int g = atoi(e);
    if (g > 0 && g <= want) grid = g;
  }
  if (grid < 1) grid = 1;
  constexpr int NP = RMS ? 1 : 2;
  float* partial = nullptr;
  bool need_params = dgamma || dbeta;
  // With an auxiliary stream the (tiny, optimizer-only) parameter reduction leaves the critical path: the block partials go
  // to one of two grow-only workspaces instead of the stream-ordered allocator, the reduce kernel runs on the aux stream
  // behind an event, and the workspace is fenced by the event of the reduce that last read it.
  cudaStream_t aux = need_params ? nfb_aux_stream() : nullptr;
  if (aux == s) aux = nullptr;
  AuxWs* w = nullptr;
  const size_t part_elems = (size_t)NP * cols * grid;
  if (aux) {
    w = &g_nr_ws[g_nr_next];
    g_nr_next = (g_nr_next + 1) % NR_RING;
    if (w->pending && norm_flush_pending()) return -1;   // the ring wrapped onto a reduction that has not been launched yet
    if (!w->done) NFB_CUDA(cudaEventCreateWithFlags(&w->done, cudaEventDisableTiming));
    if (w->n < part_elems) {
      if (w->buf) NFB_CUDA(cudaFree(w->buf));   // implicit device sync
      NFB_CUDA(cudaMalloc((void**)&w->buf, sizeof(float) * part_elems));
      w->n = part_elems;
      w->used = false;
    }
    if (w->used && w->epoch == nfb_capture_epoch()) NFB_CUDA(cudaStreamWaitEvent(s, w->done, 0));
    partial = w->buf;
  } else if (need_params && nfb_malloc(sizeof(float) * part_elems, (void**)&partial)) return -1;
#define LB(V) nfb_launch_pdl(k_norm_bwd<V, TX, TY, RMS>, dim3(grid), dim3(BWD_WARPS * 32), 0, s, (const TY*)dy, (const TX*)x, gamma, mean, rstd, (const TX*)dres, (TX*)dx, partial, rows, cols, clip, (__nv_bfloat16*)dx_twin)
  switch (vpl) {
    case 1: LB(1); break;
    case 2: LB(2); break;
    case 4: LB(4); break;
    case 6: LB(6); break;
    case 8: LB(8); break;
    default: if (partial && !aux) nfb_free(partial); return NFB_FAIL("norm bwd: cols=%d needs the generic path", cols);
  }
#undef LB
  if (need_params) {
    if (aux) {
      // queued: launched with its neighbours by norm_flush_pending (NR_BATCH queued, nfb_norm_flush, or the ring wrapping)
      bool clash = g_nr_pending.n > 0 && g_nr_stream != s;
      for (int j = 0; j < g_nr_pending.n && !clash; ++j)   // the same parameter twice in one batch would race: keep them in order
        clash = (dgamma && g_nr_pending.job[j].dgamma == dgamma) || (dbeta && g_nr_pending.job[j].dbeta == dbeta);
      if (clash && norm_flush_pending()) return -1;
      g_nr_stream = s;
      NormReduceJob& q = g_nr_pending.job[g_nr_pending.n];
      q.partial = partial; q.dgamma = dgamma; q.dbeta = dbeta; q.nblocks = grid; q.np = NP; q.cols = cols; q.accumulate = accumulate;
      g_nr_owner[g_nr_pending.n] = w;
      ++g_nr_pending.n;
      w->pending = true;
      if (g_nr_pending.n == NR_BATCH && norm_flush_pending()) return -1;
    } else {
      nfb_count_launch();
      nfb_launch_pdl(k_norm_param_reduce, dim3((NP * cols + 7) / 8), dim3(8, 32), 0, s, partial, grid, NP, cols, dgamma, dbeta, accumulate);
      nfb_free(partial);
    }
  }
  NFB_LAUNCH_CHECK();
  return 0;
}

// kind: 0 = LayerNorm, 1 = RMSNorm. xdt / ydt: NFB_F32 or NFB_BF16.
NFB_API int nfb_norm_fwd(int kind, int xdt, int ydt, const void* x, const void* res, void* sum_out, const float* gamma,
                         const float* beta, void* y, float* mean, float* rstd, int64_t rows, int cols, float eps) {
  if (rows == 0) return 0;
  if (pick_vpl(cols) == 0) {
    if (xdt != NFB_F32 || ydt != NFB_F32 || res) return NFB_FAIL("nfb_norm_fwd: generic path (cols=%d) is fp32 only, no residual", cols);
    cudaStream_t s = nfb_cur_stream();
    if (kind == 0) k_norm_fwd_generic<false><<<(unsigned)rows, 256, 0, s>>>((const float*)x, gamma, beta, (float*)y, mean, rstd, cols, eps);
    else k_norm_fwd_generic<true><<<(unsigned)rows, 256, 0, s>>>((const float*)x, gamma, beta, (float*)y, mean, rstd, cols, eps);
    NFB_LAUNCH_CHECK();
    return 0;
  }
#define DISPATCH(TX, TY)                                                                                            \
  return kind == 0 ? norm_fwd_t<TX, TY, false>(x, res, sum_out, gamma, beta, y, mean, rstd, rows, cols, eps)        \
                   : norm_fwd_t<TX, TY, true>(x, res, sum_out, gamma, beta, y, mean, rstd, rows, cols, eps)
  if (xdt == NFB_F32 && ydt == NFB_F32) { DISPATCH(float, float); }
  if (xdt == NFB_F32 && ydt == NFB_BF16) { DISPATCH(float, __nv_bfloat16); }
  if (xdt == NFB_BF16 && ydt == NFB_BF16) { DISPATCH(__nv_bfloat16, __nv_bfloat16); }
#undef DISPATCH
  return NFB_FAIL("nfb_norm_fwd: unsupported dtype pair (%d,%d)", xdt, ydt);
}

NFB_API int nfb_norm_bwd(int kind, int xdt, int ydt, const void* dy, const void* x, const float* gamma, const float* mean,
                         const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, int accumulate,
                         int64_t rows, int cols, float clip, void* dx_twin) {
  if (rows == 0) return 0;
  if (pick_vpl(cols) == 0) {
    if (xdt != NFB_F32 || ydt != NFB_F32 || dres || dx_twin || clip > 0.f) return NFB_FAIL("nfb_norm_bwd: generic path (cols=%d) is fp32 only, no fusions", cols);
    cudaStream_t s = nfb_cur_stream();
    if (kind == 0) {
      k_norm_bwd_generic<false><<<(unsigned)rows, 256, 0, s>>>((const float*)dy, (const float*)x, gamma, mean, rstd, (float*)dx, cols);
      nfb_count_launch();
      if (dgamma || dbeta) k_norm_param_generic<false><<<(cols + 127) / 128, 128, 0, s>>>((const float*)dy, (const float*)x, mean, rstd, dgamma, dbeta, rows, cols, accumulate);
    } else {
      k_norm_bwd_generic<true><<<(unsigned)rows, 256, 0, s>>>((const float*)dy, (const float*)x, gamma, mean, rstd, (float*)dx, cols);
      nfb_count_launch();
      if (dgamma || dbeta) k_norm_param_generic<true><<<(cols + 127) / 128, 128, 0, s>>>((const float*)dy, (const float*)x, mean, rstd, dgamma, dbeta, rows, cols, accumulate);
    }
    NFB_LAUNCH_CHECK();
    return 0;
  }
#define DISPATCH(TX, TY)                                                                                                   \
  return kind == 0 ? norm_bwd_t<TX, TY, false>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, accumulate, rows, cols, clip, dx_twin)  \
                   : norm_bwd_t<TX, TY, true>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, accumulate, rows, cols, clip, dx_twin)
  if (xdt == NFB_F32 && ydt == NFB_F32) { DISPATCH(float, float); }
  if (xdt == NFB_F32 && ydt == NFB_BF16) { DISPATCH(float, __nv_bfloat16); }
  if (xdt == NFB_BF16 && ydt == NFB_BF16) { DISPATCH(__nv_bfloat16, __nv_bfloat16); }
#undef DISPATCH
  return NFB_FAIL("nfb_norm_bwd: unsupported dtype pair (%d,%d)", xdt, ydt);
}
