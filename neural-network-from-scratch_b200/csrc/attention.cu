// Fused flash-style attention forward + backward (bf16 operands, fp32 softmax statistics and
// accumulators, warp-shuffle row reductions) and RoPE, for head_dim 64.
//
// Semantics follow the model-zoo attention cores of the reference:
//   GPT-2 (models/language/gpt2.py:191-215): s = q.k^T / sqrt(hd); s = where(tril == 0, -1e4, s); softmax; @ v
//   BERT  (models/language/bert.py:186-213): s = q.k^T / sqrt(hd) + (1 - m)[:, None, None, :] * -1e4
//   ViT   (models/vision/vision_transformer.py:117-126): s = (q.k^T) * hd^-0.5
// mask_mode bit 0 = causal (masked scores are REPLACED by -1e4, as the reference does; since
// exp(-1e4 - rowmax) == 0 in fp32, key blocks that are entirely above the diagonal are skipped),
// bit 1 = additive per-key mask, fp32 (B, S_kv).
//
// Layout: q/k/v/o element (b, h, s, d) lives at  base + b*sb + s*ss + h*sh + d  (d contiguous), so
// the packed (B, S, 3, H, D) output of the fused qkv projection is consumed in place and the output
// is written directly in (B, S, H*D) order: no transposes are materialised.
//
// Kernels: k_flash_fwd (O, LSE), k_attn_delta (rowsum(dO*O)), k_flash_bwd_dkv (one CTA per 64-key
// block, loops over query blocks), k_flash_bwd_dq (one CTA per 64-query block, loops over key
// blocks).  Two backward kernels instead of one with atomics keeps dQ deterministic.
// Tensor-core work uses mma.sync.m16n8k16 bf16 (HMMA); QK^T/PV are ~8% of the step FLOPs,
// moving them to tcgen05 is listed as future work in DESIGN.md.
#include "common.cuh"

#define HD 64          // head dim
#define BR 64          // query rows per CTA
#define BC 64          // keys per inner block
#define NTHREADS 128   // 4 warps x 16 rows

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* o;
  float* lse;                 // (B, H, S)
  const __nv_bfloat16 *dout;  // same layout as o
  const float* delta;         // (B, H, S)
  __nv_bfloat16 *dq, *dk, *dv;  // same layout as q/k/v
  int B, H, S, Skv;
  long long q_sb, q_ss, q_sh;      // q / dq strides (elements)
  long long kv_sb, kv_ss, kv_sh;   // k, v, dk, dv strides
  int kv_group;                    // query heads per key/value head (grouped-query attention); k/v/dk/dv have H / kv_group heads
  long long o_sb, o_ss, o_sh;      // o / dout strides
  float scale;
  int mask_mode;
  const float* key_mask;     // (B, Skv) additive, fp32
};

// ------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
// D(16x8, fp32) += A(16x16, bf16 row) * B(16x8, bf16 col)
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// A tile is [64 rows][64 bf16] = 128 B per row = 8 chunks of 16 B; chunk c of row r is stored at
// chunk (c ^ (r & 7)) so that ldmatrix (8 rows x 16 B) is bank-conflict free.
__device__ __forceinline__ __nv_bfloat16* tile_addr(__nv_bfloat16* base, int row, int chunk) {
  return base + row * HD + ((chunk ^ (row & 7)) << 3);
}
// cooperative async load of 64 rows x 64 cols (rows >= nrows_valid are zero-filled)
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* smem, const __nv_bfloat16* g, long long row_stride, int row0, int nrows_total) {
#pragma unroll
  for (int i = 0; i < (64 * 8) / NTHREADS; ++i) {
    int e = threadIdx.x + i * NTHREADS;
    int r = e >> 3, c = e & 7;
    bool ok = (row0 + r) < nrows_total;
    const __nv_bfloat16* src = g + (long long)(ok ? (row0 + r) : 0) * row_stride + c * 8;
    cp_async16(tile_addr(smem, r, c), src, ok);
  }
}
// A-operand fragments (16 rows x 64 cols of a row-major tile) for this warp's rows r0..r0+15
__device__ __forceinline__ void load_a_frags(uint32_t (*f)[4], __nv_bfloat16* tile, int r0, int lane) {
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    int chunk = ks * 2 + (lane >> 4);
    ldsm_x4(f[ks][0], f[ks][1], f[ks][2], f[ks][3], tile_addr(tile, row, chunk));
  }
}
// acc[nt] (16 x 8 per n-tile, 8 n-tiles = 64 columns) += A(16 x 64) . T^T where T is a row-major [64][64]
// tile: contraction over T's columns (B operand read NON-transposed: k = col of T, n = row of T).
__device__ __forceinline__ void mma_a_bt(float (*acc)[4], const uint32_t (*a)[4], __nv_bfloat16* T, int lane) {
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // pairs of n-tiles
      int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
      int chunk = ks * 2 + ((lane >> 3) & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b0, b1, b2, b3, tile_addr(T, row, chunk));
      mma16816(acc[np * 2], a[ks], b0, b1);
      mma16816(acc[np * 2 + 1], a[ks], b2, b3);
    }
  }
}
// acc[nt] += P(16 x 64, A-fragments over the tile's ROWS) . T where T is row-major [64 rows][64 cols]:
// contraction over T's rows (B operand read TRANSPOSED: k = row of T, n = col of T).
__device__ __forceinline__ void mma_a_b(float (*acc)[4], const uint32_t (*a)[4], __nv_bfloat16* T, int lane) {
#pragma unroll
  for (int ks = 0; ks < BC / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      int chunk = np * 2 + (lane >> 4);
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(b0, b1, b2, b3, tile_addr(T, row, chunk));
      mma16816(acc[np * 2], a[ks], b0, b1);
      mma16816(acc[np * 2 + 1], a[ks], b2, b3);
    }
  }
}
// C-fragments (16 x 64 fp32, 8 n-tiles) -> bf16 A-fragments (4 k-steps of 16)
__device__ __forceinline__ void c_to_a(uint32_t (*a)[4], const float (*c)[4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf16(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

// score for (query i, key j) given raw dot s: scale, causal replace, key mask, out-of-range keys
__device__ __forceinline__ float apply_mask(float s, int i, int j, const AttnParams& p, const float* kmask_row) {
  if (j >= p.Skv) return -INFINITY;
  float v = s * p.scale;
  if ((p.mask_mode & 1) && j > i) v = -1e4f;
  if (p.mask_mode & 2) v += kmask_row[j];
  return v;
}

// ================================================================== forward
__global__ void __launch_bounds__(NTHREADS) k_flash_fwd(const AttnParams p) {
  nfb_pdl_enter();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __nv_bfloat16* sQ = (__nv_bfloat16*)smem_raw;
  __nv_bfloat16* sK = sQ + BR * HD;          // 2 stages
  __nv_bfloat16* sV = sK + 2 * BC * HD;      // 2 stages
  // causal: the last query blocks see the most keys -- schedule them first so the grid does not end on its longest CTAs
  const int qb = (p.mask_mode & 1) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* Q = p.q + b * p.q_sb + h * p.q_sh;
  const __nv_bfloat16* K = p.k + b * p.kv_sb + (h / p.kv_group) * p.kv_sh;
  const __nv_bfloat16* V = p.v + b * p.kv_sb + (h / p.kv_group) * p.kv_sh;
  const float* kmask = (p.mask_mode & 2) ? p.key_mask + (long long)b * p.Skv : nullptr;
  const int q0 = qb * BR;
  int nkv = (p.Skv + BC - 1) / BC;
  if (p.mask_mode & 1) { int lim = (q0 + BR + BC - 1) / BC; nkv = min(nkv, lim); }  // skip fully masked blocks

  load_tile_async(sQ, Q, p.q_ss, q0, p.S);
  load_tile_async(sK, K, p.kv_ss, 0, p.Skv);
  load_tile_async(sV, V, p.kv_ss, 0, p.Skv);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4];
  load_a_frags(qf, sQ, warp * 16, lane);

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows g and g+8
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;

  for (int kb = 0; kb < nkv; ++kb) {
    int st = kb & 1;
    if (kb + 1 < nkv) {
      load_tile_async(sK + (st ^ 1) * BC * HD, K, p.kv_ss, (kb + 1) * BC, p.Skv);
      load_tile_async(sV + (st ^ 1) * BC * HD, V, p.kv_ss, (kb + 1) * BC, p.Skv);
      cp_async_commit();
    }
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_a_bt(s, qf, sK + st * BC * HD, lane);
    // mask + online softmax (rows g / g+8 of this warp; 4 lanes share a row)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      int j = kb * BC + nt * 8 + 2 * t;
      s[nt][0] = apply_mask(s[nt][0], row_a, j, p, kmask);
      s[nt][1] = apply_mask(s[nt][1], row_a, j + 1, p, kmask);
      s[nt][2] = apply_mask(s[nt][2], row_b, j, p, kmask);
      s[nt][3] = apply_mask(s[nt][3], row_b, j + 1, p, kmask);
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float nm0 = fmaxf(m0, mx0), nm1 = fmaxf(m1, mx1);
    float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - nm0), c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - nm1);
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = __expf(s[nt][0] - nm0); s[nt][1] = __expf(s[nt][1] - nm0);
      s[nt][2] = __expf(s[nt][2] - nm1); s[nt][3] = __expf(s[nt][3] - nm1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
    m0 = nm0; m1 = nm1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
    uint32_t pf[4][4];
    c_to_a(pf, s);
    mma_a_b(o, pf, sV + st * BC * HD, lane);
    if (kb + 1 < nkv) cp_async_wait<0>();
    __syncthreads();
  }
  // finalize: O = o / l ; LSE = m + log(l)
  float inv0 = 1.f / fmaxf(l0, 1e-8f), inv1 = 1.f / fmaxf(l1, 1e-8f);  // the reference's 1e-8 floor (activation.py:95)
  __nv_bfloat16* O = p.o + b * p.o_sb + h * p.o_sh;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    int d = nt * 8 + 2 * t;
    if (row_a < p.S) *reinterpret_cast<uint32_t*>(O + (long long)row_a * p.o_ss + d) = pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
    if (row_b < p.S) *reinterpret_cast<uint32_t*>(O + (long long)row_b * p.o_ss + d) = pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
  }
  if (t == 0 && p.lse) {
    float* L = p.lse + (long long)bh * p.S;
    if (row_a < p.S) L[row_a] = m0 + logf(fmaxf(l0, 1e-8f));
    if (row_b < p.S) L[row_b] = m1 + logf(fmaxf(l1, 1e-8f));
  }
}

// ================================================================== delta = rowsum(dO * O)
__global__ void k_attn_delta(const AttnParams p, float* __restrict__ delta) {
  nfb_pdl_enter();
  int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // over B*H*S
  long long total = (long long)p.B * p.H * p.S;
  if (row >= total) return;
  int s = (int)(row % p.S);
  int bh = (int)(row / p.S), b = bh / p.H, h = bh % p.H;
  const __nv_bfloat16* o = p.o + b * p.o_sb + h * p.o_sh + (long long)s * p.o_ss;
  const __nv_bfloat16* d = p.dout + b * p.o_sb + h * p.o_sh + (long long)s * p.o_ss;
  __nv_bfloat162 ov = *reinterpret_cast<const __nv_bfloat162*>(o + lane * 2);
  __nv_bfloat162 dv = *reinterpret_cast<const __nv_bfloat162*>(d + lane * 2);
  float2 of = __bfloat1622float2(ov), df = __bfloat1622float2(dv);
  float acc = warp_sum(of.x * df.x + of.y * df.y);
  if (lane == 0) delta[row] = acc;
}

// ================================================================== backward: dK, dV
// One CTA per 64-key block; warp w owns keys [16w, 16w+16).  Works on S^T so P^T / dS^T come out of
// the tensor cores already in A-operand layout for the dV / dK products.
__global__ void __launch_bounds__(NTHREADS) k_flash_bwd_dkv(const AttnParams p) {
  nfb_pdl_enter();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __nv_bfloat16* sK = (__nv_bfloat16*)smem_raw;
  __nv_bfloat16* sV = sK + BC * HD;
  __nv_bfloat16* sQ = sV + BC * HD;          // 2 stages
  __nv_bfloat16* sdO = sQ + 2 * BR * HD;     // 2 stages
  float* sLse = (float*)(sdO + 2 * BR * HD); // 2 x 64
  float* sDelta = sLse + 2 * BR;             // 2 x 64
  // blockIdx.y = (batch, KEY/VALUE head); with grouped-query attention the iterations run over the kv_group query heads
  // that share this head, so dK / dV accumulate over the whole group in registers
  const int Hkv = p.H / p.kv_group;
  const int kb = blockIdx.x, b = blockIdx.y / Hkv, h = blockIdx.y % Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* K = p.k + b * p.kv_sb + h * p.kv_sh;
  const __nv_bfloat16* V = p.v + b * p.kv_sb + h * p.kv_sh;
  const float* kmask = (p.mask_mode & 2) ? p.key_mask + (long long)b * p.Skv : nullptr;
  const int k0 = kb * BC;
  const int nq = (p.S + BR - 1) / BR;
  const int qb_start = (p.mask_mode & 1) ? (k0 / BR) : 0;  // query blocks entirely above the diagonal see nothing
  const int nqb = max(nq - qb_start, 0), niter = nqb * p.kv_group;

  auto load_q_stage = [&](int stage, int it) {
    const int hq = h * p.kv_group + it / nqb, qb = qb_start + it % nqb;
    load_tile_async(sQ + stage * BR * HD, p.q + b * p.q_sb + hq * p.q_sh, p.q_ss, qb * BR, p.S);
    load_tile_async(sdO + stage * BR * HD, p.dout + b * p.o_sb + hq * p.o_sh, p.o_ss, qb * BR, p.S);
    if (threadIdx.x < BR) {
      int r = qb * BR + threadIdx.x;
      const long long row0 = ((long long)b * p.H + hq) * p.S;
      sLse[stage * BR + threadIdx.x] = r < p.S ? p.lse[row0 + r] : 0.f;
      sDelta[stage * BR + threadIdx.x] = r < p.S ? p.delta[row0 + r] : 0.f;
    }
  };
  load_tile_async(sK, K, p.kv_ss, k0, p.Skv);
  load_tile_async(sV, V, p.kv_ss, k0, p.Skv);
  if (niter > 0) load_q_stage(0, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(kf, sK, warp * 16, lane);
  load_a_frags(vf, sV, warp * 16, lane);
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const int key_a = k0 + warp * 16 + g, key_b = key_a + 8;
  const float km_a = (kmask && key_a < p.Skv) ? kmask[key_a] : 0.f, km_b = (kmask && key_b < p.Skv) ? kmask[key_b] : 0.f;

  for (int it = 0; it < niter; ++it) {
    const int st = it & 1, qb = qb_start + it % nqb;
    if (it + 1 < niter) { load_q_stage(st ^ 1, it + 1); cp_async_commit(); }
    __nv_bfloat16* tQ = sQ + st * BR * HD;
    __nv_bfloat16* tdO = sdO + st * BR * HD;
    // S^T (16 keys x 64 queries) = K . Q^T
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_a_bt(s, kf, tQ, lane);
    // P^T = exp(score - LSE[q])
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      int qi = nt * 8 + 2 * t;          // column (query) within the block
      int i0 = qb * BR + qi, i1 = i0 + 1;
      float L0 = sLse[st * BR + qi], L1 = sLse[st * BR + qi + 1];
      float v00 = s[nt][0] * p.scale, v01 = s[nt][1] * p.scale, v10 = s[nt][2] * p.scale, v11 = s[nt][3] * p.scale;
      if (p.mask_mode & 1) {
        if (key_a > i0) v00 = -1e4f; if (key_a > i1) v01 = -1e4f;
        if (key_b > i0) v10 = -1e4f; if (key_b > i1) v11 = -1e4f;
      }
      v00 += km_a; v01 += km_a; v10 += km_b; v11 += km_b;
      bool a_ok = key_a < p.Skv, b_ok = key_b < p.Skv;
      s[nt][0] = (a_ok && i0 < p.S) ? __expf(v00 - L0) : 0.f;
      s[nt][1] = (a_ok && i1 < p.S) ? __expf(v01 - L1) : 0.f;
      s[nt][2] = (b_ok && i0 < p.S) ? __expf(v10 - L0) : 0.f;
      s[nt][3] = (b_ok && i1 < p.S) ? __expf(v11 - L1) : 0.f;
    }
    uint32_t pf[4][4];
    c_to_a(pf, s);
    // dV += P^T . dO      (contraction over queries = rows of the dO tile)
    mma_a_b(dv, pf, tdO, lane);
    // dP^T = V . dO^T
    float dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    mma_a_bt(dp, vf, tdO, lane);
    // dS^T = P^T * (dP^T - delta[q]) * scale ; causal-replaced positions carry no gradient
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      int qi = nt * 8 + 2 * t;
      int i0 = qb * BR + qi, i1 = i0 + 1;
      float D0 = sDelta[st * BR + qi], D1 = sDelta[st * BR + qi + 1];
      float d00 = s[nt][0] * (dp[nt][0] - D0) * p.scale, d01 = s[nt][1] * (dp[nt][1] - D1) * p.scale;
      float d10 = s[nt][2] * (dp[nt][2] - D0) * p.scale, d11 = s[nt][3] * (dp[nt][3] - D1) * p.scale;
      if (p.mask_mode & 1) {
        if (key_a > i0) d00 = 0.f; if (key_a > i1) d01 = 0.f;
        if (key_b > i0) d10 = 0.f; if (key_b > i1) d11 = 0.f;
      }
      s[nt][0] = d00; s[nt][1] = d01; s[nt][2] = d10; s[nt][3] = d11;
    }
    c_to_a(pf, s);
    // dK += dS^T . Q
    mma_a_b(dk, pf, tQ, lane);
    if (it + 1 < niter) cp_async_wait<0>();
    __syncthreads();
  }
  __nv_bfloat16* dK = p.dk + b * p.kv_sb + h * p.kv_sh;
  __nv_bfloat16* dV = p.dv + b * p.kv_sb + h * p.kv_sh;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    int d = nt * 8 + 2 * t;
    if (key_a < p.Skv) {
      *reinterpret_cast<uint32_t*>(dK + (long long)key_a * p.kv_ss + d) = pack_bf16(dk[nt][0], dk[nt][1]);
      *reinterpret_cast<uint32_t*>(dV + (long long)key_a * p.kv_ss + d) = pack_bf16(dv[nt][0], dv[nt][1]);
    }
    if (key_b < p.Skv) {
      *reinterpret_cast<uint32_t*>(dK + (long long)key_b * p.kv_ss + d) = pack_bf16(dk[nt][2], dk[nt][3]);
      *reinterpret_cast<uint32_t*>(dV + (long long)key_b * p.kv_ss + d) = pack_bf16(dv[nt][2], dv[nt][3]);
    }
  }
}

// ================================================================== backward: dQ
__global__ void __launch_bounds__(NTHREADS) k_flash_bwd_dq(const AttnParams p) {
  nfb_pdl_enter();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __nv_bfloat16* sQ = (__nv_bfloat16*)smem_raw;
  __nv_bfloat16* sdO = sQ + BR * HD;
  __nv_bfloat16* sK = sdO + BR * HD;        // 2 stages
  __nv_bfloat16* sV = sK + 2 * BC * HD;     // 2 stages
  // causal: the last query blocks see the most keys -- schedule them first so the grid does not end on its longest CTAs
  const int qb = (p.mask_mode & 1) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* Q = p.q + b * p.q_sb + h * p.q_sh;
  const __nv_bfloat16* K = p.k + b * p.kv_sb + (h / p.kv_group) * p.kv_sh;
  const __nv_bfloat16* V = p.v + b * p.kv_sb + (h / p.kv_group) * p.kv_sh;
  const __nv_bfloat16* dO = p.dout + b * p.o_sb + h * p.o_sh;
  const float* kmask = (p.mask_mode & 2) ? p.key_mask + (long long)b * p.Skv : nullptr;
  const int q0 = qb * BR;
  int nkv = (p.Skv + BC - 1) / BC;
  if (p.mask_mode & 1) { int lim = (q0 + BR + BC - 1) / BC; nkv = min(nkv, lim); }
  load_tile_async(sQ, Q, p.q_ss, q0, p.S);
  load_tile_async(sdO, dO, p.o_ss, q0, p.S);
  load_tile_async(sK, K, p.kv_ss, 0, p.Skv);
  load_tile_async(sV, V, p.kv_ss, 0, p.Skv);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4], dof[4][4];
  load_a_frags(qf, sQ, warp * 16, lane);
  load_a_frags(dof, sdO, warp * 16, lane);
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  const float* LSE = p.lse + (long long)bh * p.S;
  const float* DEL = p.delta + (long long)bh * p.S;
  const float La = row_a < p.S ? LSE[row_a] : 0.f, Lb = row_b < p.S ? LSE[row_b] : 0.f;
  const float Da = row_a < p.S ? DEL[row_a] : 0.f, Db = row_b < p.S ? DEL[row_b] : 0.f;
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

  for (int kb = 0; kb < nkv; ++kb) {
    int st = kb & 1;
    if (kb + 1 < nkv) {
      load_tile_async(sK + (st ^ 1) * BC * HD, K, p.kv_ss, (kb + 1) * BC, p.Skv);
      load_tile_async(sV + (st ^ 1) * BC * HD, V, p.kv_ss, (kb + 1) * BC, p.Skv);
      cp_async_commit();
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    mma_a_bt(s, qf, sK + st * BC * HD, lane);     // S  = Q . K^T
    mma_a_bt(dp, dof, sV + st * BC * HD, lane);   // dP = dO . V^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      int j = kb * BC + nt * 8 + 2 * t;
      float p00 = __expf(apply_mask(s[nt][0], row_a, j, p, kmask) - La), p01 = __expf(apply_mask(s[nt][1], row_a, j + 1, p, kmask) - La);
      float p10 = __expf(apply_mask(s[nt][2], row_b, j, p, kmask) - Lb), p11 = __expf(apply_mask(s[nt][3], row_b, j + 1, p, kmask) - Lb);
      float d00 = p00 * (dp[nt][0] - Da) * p.scale, d01 = p01 * (dp[nt][1] - Da) * p.scale;
      float d10 = p10 * (dp[nt][2] - Db) * p.scale, d11 = p11 * (dp[nt][3] - Db) * p.scale;
      if (p.mask_mode & 1) {
        if (j > row_a) d00 = 0.f; if (j + 1 > row_a) d01 = 0.f;
        if (j > row_b) d10 = 0.f; if (j + 1 > row_b) d11 = 0.f;
      }
      s[nt][0] = d00; s[nt][1] = d01; s[nt][2] = d10; s[nt][3] = d11;
    }
    uint32_t dsf[4][4];
    c_to_a(dsf, s);
    mma_a_b(dq, dsf, sK + st * BC * HD, lane);    // dQ += dS . K
    if (kb + 1 < nkv) cp_async_wait<0>();
    __syncthreads();
  }
  __nv_bfloat16* dQ = p.dq + b * p.q_sb + h * p.q_sh;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    int d = nt * 8 + 2 * t;
    if (row_a < p.S) *reinterpret_cast<uint32_t*>(dQ + (long long)row_a * p.q_ss + d) = pack_bf16(dq[nt][0], dq[nt][1]);
    if (row_b < p.S) *reinterpret_cast<uint32_t*>(dQ + (long long)row_b * p.q_ss + d) = pack_bf16(dq[nt][2], dq[nt][3]);
  }
}

// ================================================================== host entry points
static int check_attn(int head_dim, long long q_ss, long long kv_ss, long long o_ss) {
  if (head_dim != HD) return NFB_FAIL("flash attention: head_dim %d unsupported (this build handles %d)", head_dim, HD);
  if (q_ss % 8 || kv_ss % 8 || o_ss % 8) return NFB_FAIL("flash attention: row strides must be multiples of 8 elements (16 B)");
  return 0;
}

// tcgen05 / TMEM / TMA forward (csrc/attention_tc.cu); the mma.sync kernel below stays as the path for layouts the
// tensor maps cannot address (unaligned bases / strides)
bool nfb_flash_fwd_tc_eligible(const void* q, const void* k, const void* v, const void* o, int head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh,
                               int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_ss, int64_t o_sh, int64_t o_sb);
int nfb_flash_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int S, int Skv, int64_t q_sb, int64_t q_ss,
                     int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh, float scale, int mask_mode,
                     const float* key_mask, int kv_group);
int nfb_flash_bwd_tc(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, const float* delta, void* dq, void* dk, void* dv,
                     int B, int H, int S, int Skv, int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb,
                     int64_t o_ss, int64_t o_sh, float scale, int mask_mode, const float* key_mask, int kv_group);
static int g_attn_tc = 3;   // bit 0: tcgen05 forward, bit 1: tcgen05 backward
NFB_API int nfb_flash_attn_set_tc(int on) { g_attn_tc = on; return 0; }   // test hook: 0 = mma.sync kernels only, 1 = tcgen05 forward only, 3 = both (default)

static int check_group(int H, int Hkv) {
  if (Hkv <= 0 || H % Hkv) return NFB_FAIL("flash attention: %d query heads are not a multiple of %d key/value heads", H, Hkv);
  return 0;
}
// dtype must be NFB_BF16.  lse: (B,H,S) fp32 (may be NULL for inference).  key_mask: (B,Skv) fp32 or NULL.
// Grouped-query form: q / o / lse have H heads, k / v have Hkv heads (H % Hkv == 0); query head h reads key/value head
// h / (H / Hkv) in place -- the repeated K/V of nn/modern_attention.py:137-161 never exists in memory.
NFB_API int nfb_flash_attn_fwd_gqa(int dtype, const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv, int S,
                                   int head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh,
                                   int64_t o_sb, int64_t o_ss, int64_t o_sh, int Skv, float scale, int mask_mode, const float* key_mask) {
  if (dtype != NFB_BF16) return NFB_FAIL("flash attention operands must be bf16");
  if (B == 0 || H == 0 || S == 0) return 0;
  if (check_group(H, Hkv)) return -1;
  const int kv_group = H / Hkv;
  if (check_attn(head_dim, q_ss, kv_ss, o_ss)) return -1;
  if ((mask_mode & 2) && !key_mask) return NFB_FAIL("flash attention: mask_mode bit 1 set but key_mask is NULL");
  if ((g_attn_tc & 1) && nfb_flash_fwd_tc_eligible(q, k, v, o, head_dim, q_sb, q_ss, q_sh, kv_sb, kv_ss, kv_sh, o_ss, o_sh, o_sb))
    return nfb_flash_fwd_tc(q, k, v, o, lse, B, H, S, Skv, q_sb, q_ss, q_sh, kv_sb, kv_ss, kv_sh, o_sb, o_ss, o_sh, scale, mask_mode, key_mask, kv_group);
  AttnParams p = {};
  p.kv_group = kv_group;
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.o = (__nv_bfloat16*)o; p.lse = lse;
  p.B = B; p.H = H; p.S = S; p.Skv = Skv;
  p.q_sb = q_sb; p.q_ss = q_ss; p.q_sh = q_sh; p.kv_sb = kv_sb; p.kv_ss = kv_ss; p.kv_sh = kv_sh; p.o_sb = o_sb; p.o_ss = o_ss; p.o_sh = o_sh;
  p.scale = scale; p.mask_mode = mask_mode; p.key_mask = key_mask;
  size_t smem = (size_t)(BR * HD + 4 * BC * HD) * 2;
  dim3 grid((unsigned)nfb_cdiv(S, BR), (unsigned)(B * H));
  nfb_launch_pdl(k_flash_fwd, grid, dim3(NTHREADS), smem, nfb_cur_stream(), p);
  NFB_LAUNCH_CHECK();
  return 0;
}

NFB_API int nfb_flash_attn_fwd(int dtype, const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int S,
                               int head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh,
                               int64_t o_sb, int64_t o_ss, int64_t o_sh, int Skv, float scale, int mask_mode, const float* key_mask) {
  return nfb_flash_attn_fwd_gqa(dtype, q, k, v, o, lse, B, H, H, S, head_dim, q_sb, q_ss, q_sh, kv_sb, kv_ss, kv_sh, o_sb, o_ss, o_sh, Skv, scale,
                                mask_mode, key_mask);
}

// dq has H heads; dk / dv have Hkv heads (k / v strides) and receive the SUM over the H / Hkv query heads of each group,
// accumulated inside the kernels (TMEM / registers), not by a separate reduction.
NFB_API int nfb_flash_attn_bwd_gqa(int dtype, const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                                   void* dq, void* dk, void* dv, int B, int H, int Hkv, int S, int head_dim, int64_t q_sb, int64_t q_ss,
                                   int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh, int Skv,
                                   float scale, int mask_mode, const float* key_mask) {
  if (dtype != NFB_BF16) return NFB_FAIL("flash attention operands must be bf16");
  if (B == 0 || H == 0 || S == 0) return 0;
  if (check_attn(head_dim, q_ss, kv_ss, o_ss)) return -1;
  if (check_group(H, Hkv)) return -1;
  const int kv_group = H / Hkv;
  AttnParams p = {};
  p.kv_group = kv_group;
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.o = (__nv_bfloat16*)const_cast<void*>(o);
  p.dout = (const __nv_bfloat16*)dout; p.lse = const_cast<float*>(lse);
  p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  p.B = B; p.H = H; p.S = S; p.Skv = Skv;
  p.q_sb = q_sb; p.q_ss = q_ss; p.q_sh = q_sh; p.kv_sb = kv_sb; p.kv_ss = kv_ss; p.kv_sh = kv_sh; p.o_sb = o_sb; p.o_ss = o_ss; p.o_sh = o_sh;
  p.scale = scale; p.mask_mode = mask_mode; p.key_mask = key_mask;
  if ((g_attn_tc & 2) && nfb_flash_fwd_tc_eligible(q, k, v, dout, head_dim, q_sb, q_ss, q_sh, kv_sb, kv_ss, kv_sh, o_ss, o_sh, o_sb) &&
      ((((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv | (uintptr_t)o) & 15) == 0))
    return nfb_flash_bwd_tc(q, k, v, o, dout, lse, nullptr, dq, dk, dv, B, H, S, Skv, q_sb, q_ss, q_sh, kv_sb, kv_ss, kv_sh, o_sb, o_ss, o_sh, scale,
                            mask_mode, key_mask, kv_group);
  float* delta = nullptr;
  if (nfb_malloc(sizeof(float) * (size_t)B * H * S, (void**)&delta)) return -1;
  p.delta = delta;
  cudaStream_t s = nfb_cur_stream();
  long long rows = (long long)B * H * S;
  nfb_launch_pdl(k_attn_delta, dim3((unsigned)nfb_cdiv(rows, 8)), dim3(256), 0, s, p, delta);
  nfb_count_launch();
  size_t smem_dkv = (size_t)(2 * BC * HD + 4 * BR * HD) * 2 + 4 * BR * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(k_flash_bwd_dkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dkv);
    cudaFuncSetAttribute(k_flash_bwd_dq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * BR * HD + 4 * BC * HD) * 2));
    cfg = true;
  }
  nfb_launch_pdl(k_flash_bwd_dkv, dim3((unsigned)nfb_cdiv(Skv, BC), (unsigned)(B * Hkv)), dim3(NTHREADS), smem_dkv, s, p);
  nfb_count_launch();
  size_t smem_dq = (size_t)(2 * BR * HD + 4 * BC * HD) * 2;
  nfb_launch_pdl(k_flash_bwd_dq, dim3((unsigned)nfb_cdiv(S, BR), (unsigned)(B * H)), dim3(NTHREADS), smem_dq, s, p);
  nfb_free(delta);
  NFB_LAUNCH_CHECK();
  return 0;
}

NFB_API int nfb_flash_attn_bwd(int dtype, const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                               void* dq, void* dk, void* dv, int B, int H, int S, int head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh,
                               int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh, int Skv, float scale,
                               int mask_mode, const float* key_mask) {
  return nfb_flash_attn_bwd_gqa(dtype, q, k, v, o, dout, lse, dq, dk, dv, B, H, H, S, head_dim, q_sb, q_ss, q_sh, kv_sb, kv_ss, kv_sh, o_sb, o_ss,
                                o_sh, Skv, scale, mask_mode, key_mask);
}

// ================================================================== RoPE (models/language/gpt2.py:25-95)
// x' = x*cos + rotate_half(x)*sin.  The cos/sin tables (S x D/2, fp32) are built on the host with the
// reference's exact NumPy expressions (RotaryEmbedding._build_cache) so the angles are bit-identical;
// inverse=1 applies the transpose of the rotation (the backward).  Operates in place on n_parts consecutive
// (B,S,H,D)-strided tensors (q and k of a packed qkv buffer): element (part, b, s, h, d) at
// x + part*part_stride + b*sb + s*ss + h*sh + d.
template <typename T> struct RopeVec;
template <> struct RopeVec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* v) { float4 a = *reinterpret_cast<const float4*>(p); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
  __device__ static void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct RopeVec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
  }
  __device__ static void store(__nv_bfloat16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
// one thread rotates N (d, d + D/2) pairs with 16-byte data accesses (N = 8 for bf16, 4 for fp32) and 16-byte table
// loads; 32-bit index arithmetic (the launcher checks the element count)
template <typename T>
__global__ void k_rope(T* __restrict__ x, const float* __restrict__ cos_t, const float* __restrict__ sin_t, int n_parts, long long part_stride,
                       int B, int S, int H, int D, long long sb, long long ss, long long sh, int inverse) {
  nfb_pdl_enter();
  constexpr int N = RopeVec<T>::N;
  const unsigned half = D / 2, groups = half / N;
  const unsigned total = (unsigned)n_parts * B * S * H * groups, stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned d = (i % groups) * N;
    unsigned r = i / groups;
    const unsigned h = r % H; r /= H;
    const unsigned sidx = r % S; r /= S;
    const unsigned b = r % B;
    const unsigned part = r / B;
    float c[N], sn[N], x1[N], x2[N], y1[N], y2[N];
#pragma unroll
    for (int k = 0; k < N; k += 4) {
      float4 cv = __ldg(reinterpret_cast<const float4*>(cos_t + sidx * half + d + k)), sv = __ldg(reinterpret_cast<const float4*>(sin_t + sidx * half + d + k));
      c[k] = cv.x; c[k + 1] = cv.y; c[k + 2] = cv.z; c[k + 3] = cv.w;
      sn[k] = sv.x; sn[k + 1] = sv.y; sn[k + 2] = sv.z; sn[k + 3] = sv.w;
    }
    T* p = x + part * part_stride + b * sb + (long long)sidx * ss + (long long)h * sh;
    RopeVec<T>::load(p + d, x1);
    RopeVec<T>::load(p + d + half, x2);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float sk = inverse ? -sn[k] : sn[k];
      y1[k] = x1[k] * c[k] - x2[k] * sk;
      y2[k] = x2[k] * c[k] + x1[k] * sk;
    }
    RopeVec<T>::store(p + d, y1);
    RopeVec<T>::store(p + d + half, y2);
  }
}
// any even head_dim / alignment: one pair per thread
template <typename T>
__global__ void k_rope_scalar(T* __restrict__ x, const float* __restrict__ cos_t, const float* __restrict__ sin_t, int n_parts, long long part_stride,
                              int B, int S, int H, int D, long long sb, long long ss, long long sh, int inverse) {
  nfb_pdl_enter();
  const int half = D / 2;
  long long total = (long long)n_parts * B * S * H * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int d = (int)(i % half);
    long long r = i / half;
    int h = (int)(r % H); r /= H;
    int s = (int)(r % S); r /= S;
    int b = (int)(r % B);
    int part = (int)(r / B);
    float c = cos_t[s * half + d], sn = sin_t[s * half + d];
    if (inverse) sn = -sn;
    T* p = x + part * part_stride + b * sb + (long long)s * ss + (long long)h * sh;
    float x1 = ld_f<T>(p, d), x2 = ld_f<T>(p, d + half);
    st_f<T>(p, d, x1 * c - x2 * sn);
    st_f<T>(p, d + half, x2 * c + x1 * sn);
  }
}
// Interleaved-pair rotary embedding of nn/positional.py:140-366 (RotaryPositionalEmbedding): pairs are (2j, 2j+1),
//   y[2j] = x[2j] cos_j - x[2j+1] sin_j ;  y[2j+1] = x[2j+1] cos_j + x[2j] sin_j ;  tables (S, D/2).
// One thread rotates the N/2 pairs of one 16-byte vector in place; inverse = transpose rotation (the gradient).
template <typename T, bool VEC>
__global__ void k_rope_pairs(T* __restrict__ x, const float* __restrict__ cos_t, const float* __restrict__ sin_t, int B, int S, int H, int D,
                             long long sb, long long ss, long long sh, int inverse) {
  nfb_pdl_enter();
  constexpr int N = VEC ? RopeVec<T>::N : 2;
  const int half = D / 2, groups = D / N;
  const long long total = (long long)B * S * H * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % groups) * N;
    long long r = i / groups;
    const int h = (int)(r % H); r /= H;
    const int sidx = (int)(r % S);
    const long long b = r / S;
    T* p = x + b * sb + (long long)sidx * ss + (long long)h * sh + d;
    float v[N], y[N];
    if constexpr (VEC) RopeVec<T>::load(p, v);
    else { v[0] = ld_f<T>(p, 0); v[1] = ld_f<T>(p, 1); }
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      const float c = __ldg(cos_t + (long long)sidx * half + d / 2 + j);
      float sn = __ldg(sin_t + (long long)sidx * half + d / 2 + j);
      if (inverse) sn = -sn;
      y[2 * j] = v[2 * j] * c - v[2 * j + 1] * sn;
      y[2 * j + 1] = v[2 * j + 1] * c + v[2 * j] * sn;
    }
    if constexpr (VEC) RopeVec<T>::store(p, y);
    else { st_f<T>(p, 0, y[0]); st_f<T>(p, 1, y[1]); }
  }
}
NFB_API int nfb_rope_interleaved(int dtype, void* x, const float* cos_t, const float* sin_t, int B, int S, int H, int head_dim, int64_t sb,
                                 int64_t ss, int64_t sh, int inverse) {
  if (head_dim <= 0 || head_dim % 2) return NFB_FAIL("nfb_rope_interleaved: head_dim %d must be positive and even", head_dim);
  const int D = head_dim, vn = dtype == NFB_BF16 ? 8 : 4;
  const bool vec = (D % vn == 0) && !(sb % vn || ss % vn || sh % vn) && ((uintptr_t)x % 16 == 0);
  const long long total = (long long)B * S * H * (D / (vec ? vn : 2));
  if (total == 0) return 0;
  const int grid = nfb_grid_for(total, 256);
  cudaStream_t st = nfb_cur_stream();
#define RP(T, V) nfb_launch_pdl(k_rope_pairs<T, V>, dim3(grid), dim3(256), 0, st, (T*)x, cos_t, sin_t, B, S, H, D, sb, ss, sh, inverse)
  if (dtype == NFB_BF16) { if (vec) RP(__nv_bfloat16, true); else RP(__nv_bfloat16, false); }
  else if (dtype == NFB_F32) { if (vec) RP(float, true); else RP(float, false); }
  else return NFB_FAIL("nfb_rope_interleaved: unsupported dtype %d", dtype);
#undef RP
  NFB_LAUNCH_CHECK();
  return 0;
}
NFB_API int nfb_rope(int dtype, void* x, const float* cos_t, const float* sin_t, int n_parts, int64_t part_stride, int B, int S, int H,
                     int head_dim, int64_t sb, int64_t ss, int64_t sh, int inverse) {
  if (head_dim <= 0 || head_dim % 2) return NFB_FAIL("nfb_rope: head_dim %d must be positive and even", head_dim);
  int D = head_dim;
  const int vn = dtype == NFB_BF16 ? 8 : 4;   // elements per 16-byte access
  bool vec = ((D / 2) % vn == 0) && !(sb % vn || ss % vn || sh % vn || part_stride % vn) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)cos_t % 16 == 0) &&
             ((uintptr_t)sin_t % 16 == 0) && ((long long)n_parts * B * S * H * (D / 2 / vn) < (1LL << 31));
  long long total = (long long)n_parts * B * S * H * (vec ? D / 2 / vn : D / 2);
  if (total == 0) return 0;
  int grid = nfb_grid_for(total, 256);
  cudaStream_t st = nfb_cur_stream();
#define RL(K, T) nfb_launch_pdl(K<T>, dim3(grid), dim3(256), 0, st, (T*)x, cos_t, sin_t, n_parts, part_stride, B, S, H, D, sb, ss, sh, inverse)
  if (dtype == NFB_BF16) { if (vec) RL(k_rope, __nv_bfloat16); else RL(k_rope_scalar, __nv_bfloat16); }
  else if (dtype == NFB_F32) { if (vec) RL(k_rope, float); else RL(k_rope_scalar, float); }
  else return NFB_FAIL("nfb_rope: unsupported dtype %d", dtype);
#undef RL
  NFB_LAUNCH_CHECK();
  return 0;
}
