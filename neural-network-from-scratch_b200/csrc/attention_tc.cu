// Flash attention FORWARD on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by
// TMA), head_dim 64.  Same semantics and (b, h, s, d)-strided layouts as csrc/attention.cu (SURVEY 8a row a6:
// models/language/gpt2.py:183-215, models/language/bert.py:155-231, models/vision/vision_transformer.py:100-134):
//   scores = q.k^T * scale ; causal: scores[j > i] REPLACED by -1e4 ; key mask: scores += mask[b, j] ; softmax ; . v
//
// One CTA per (batch, head, PAIR of 128-row query tiles) -- 320 threads:
//   warps 0-3  softmax group 0 (query tile A: thread = one query row, TMEM lanes 32*warp ..)
//   warps 4-7  softmax group 1 (query tile B)
//   warp  8    TMA producer: Q tiles once, then K / V blocks of 128 keys through a 2-stage ring (4-D tensor maps over
//              the strided (b, s, h, d) tensors: the packed qkv projection output is read in place)
//   warp  9    MMA issuer (one elected lane): S_g = Q_g K^T (128x128x64) and O_g += P_g V (128x64x128) for g = 0, 1
// The two query tiles ping-pong: while group g exponentiates S_g the tensor core works on the other group's
// products, and both groups share every K / V block (half the K/V traffic of one tile per CTA).
// TMEM (512 columns): S_0 [0,128) S_1 [128,256) O_0 [256,320) O_1 [320,384), fp32.
// Softmax is online in the log2 domain with LAZY rescaling of the O accumulator (it lives in TMEM, so a rescale is a
// tcgen05.ld / tcgen05.st round trip): the running maximum is only advanced when some row of the warp grew by more
// than 2^8, otherwise the stale maximum is kept (exact: the final division by the row sum uses the same maximum).
// P is rounded to bf16 and handed to the P.V product through shared memory in the K-major SWIZZLE_128B layout.
#include "common.cuh"
#ifndef NFB_ATTN_POLY_FWD_DEFAULT
#define NFB_ATTN_POLY_FWD_DEFAULT 0   // set from the measurements in profiles/r02_attn_poly.txt
#define NFB_ATTN_POLY_BWD_DEFAULT 0
#endif
#include <cuda.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <tuple>

namespace {

constexpr int ATC_THREADS = 320;
constexpr int TILE_Q = 128, TILE_K = 128, HDIM = 64;
constexpr float LOG2E = 1.4426950408889634f;

struct AttnTcParams {
  __nv_bfloat16* o;
  float* lse;
  int B, H, S, Skv;
  long long o_sb, o_ss, o_sh;
  float scale;
  int mask_mode;
  const float* key_mask;
  unsigned long long* timeline;   // debug: clock64 stamps of CTA (0,0) (NULL in production)
  int kv_group;                   // query heads per key/value head (grouped-query attention): K/V maps have H / kv_group heads
};

// ---------------------------------------------------------------- PTX helpers (same conventions as gemm_tc.cu)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 8000000u) {
      printf("nfb attention_tc: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {   // 2^x, one MUFU op (x <= ~8 here; -inf -> 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA pipe (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-4 polynomial (relative error
// < 5e-5, far inside the bf16 rounding P receives), 2^n by an integer add into the exponent field.  Both attention kernels are
// bound by the 16 ex2/clk/SM MUFU rate at head_dim 64; a share of the exponentials of every row goes through here instead
// (POLY of every 4 element pairs / elements: 0 = none).  x <= ~8; anything below -127 (masked keys, -inf) comes out as 0.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -127.f);
  const float magic = 12582912.f;                 // 1.5 * 2^23: adding it leaves round(x) in the low mantissa bits
  const float xf = x + magic;
  const float f = x - (xf - magic);
  float pl = fmaf(f, 0.009618129f, 0.05550411f);   // 2^f = sum (f ln 2)^k / k!
  pl = fmaf(pl, f, 0.2402265f);
  pl = fmaf(pl, f, 0.6931472f);
  pl = fmaf(pl, f, 1.0f);
  return __uint_as_float(__float_as_uint(pl) + (__float_as_uint(xf) << 23));
}
template <int POLY>
__device__ __forceinline__ float ex2_mix(float x, int slot) {   // slot: position of the element within its group of 4
  if (slot < POLY) return exp2_poly(x);
  return ex2(x);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_128B (see gemm_tc.cu)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn, int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 1u << 7;                       // a_format = BF16
  d |= 1u << 10;                      // b_format = BF16
  d |= (uint32_t)(a_mn & 1) << 15;
  d |= (uint32_t)(b_mn & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// per-CTA (start, end) in globaltimer ns + SM id, slots 256.. (debug)
__device__ __forceinline__ void tl_cta(const AttnTcParams& p, int which) {
  if (!p.timeline) return;
  const int cta = blockIdx.x;
  if (cta >= 256) return;
  if (which == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.timeline[256 + 3 * cta + 2] = sm; }
  p.timeline[256 + 3 * cta + which] = gtimer();
}
__device__ __forceinline__ void tl(const AttnTcParams& p, int slot) {
  if (p.timeline && blockIdx.x == 0 && slot < 256) p.timeline[slot] = (unsigned long long)clock64();
}

// shared-memory carve-up (offsets from the 1024-byte aligned base)
constexpr int KV_STAGES = 3;                     // K/V ring depth: block i+2 is requested while block i is being consumed
constexpr int SM_Q = 0;                          // 2 x 16 KB   (query tiles A, B: [128 rows][64] K-major)
constexpr int SM_K = 32768;                      // KV_STAGES x 16 KB ([128 keys][64] K-major)
constexpr int SM_V = SM_K + KV_STAGES * 16384;   // KV_STAGES x 16 KB ([128 keys][64]: MN-major B operand, two 64-key chunks)
constexpr int SM_P = SM_V + KV_STAGES * 16384;   // 2 groups x 32 KB ([128 rows][128 keys] bf16 as two 64-key K-major panels)
constexpr int SM_BAR = SM_P + 65536;             // barriers + TMEM slot
constexpr int SM_MASK = SM_BAR + 256;            // KV_STAGES x 128 fp32: additive key mask (log2 domain) of the staged block; -inf beyond Skv
constexpr int SM_TOTAL = SM_MASK + KV_STAGES * 512 + 1024;    // + alignment slack

template <int POLY>
__global__ void __launch_bounds__(ATC_THREADS, 1)
k_flash_fwd_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + SM_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // KV_STAGES
  uint64_t* kv_empty = bars + 5;      // KV_STAGES
  uint64_t* s_full = bars + 9;        // 2 (per group)
  uint64_t* p_full = bars + 11;       // 2 (per group, 128 arrivals)
  uint64_t* pv_done = bars + 13;      // 2 (per group): P_g(i).V(i) has completed -> P_g buffer and O_g are safe to touch
  uint32_t* tmem_slot = (uint32_t*)(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool causal = (p.mask_mode & 1) != 0;
  // 1-D grid, tile-pair-major; causal: the last tile pairs see the most keys -- ALL of those CTAs are scheduled first
  // (longest-processing-time order over the 148 single-CTA SMs)
  const int nbh = p.B * p.H, npairs = (int)gridDim.x / nbh;
  const int pidx = (int)blockIdx.x / nbh, bh = (int)blockIdx.x % nbh, b = bh / p.H, h = bh % p.H;
  const int pair = causal ? (npairs - 1 - pidx) : pidx;
  const int q0A = pair * 2 * TILE_Q, q0B = q0A + TILE_Q;
  const int nkv_all = (p.Skv + TILE_K - 1) / TILE_K;
  const int nA = causal ? min(nkv_all, q0A / TILE_K + 1) : nkv_all;
  const int nB = (q0B < p.S) ? (causal ? min(nkv_all, q0B / TILE_K + 1) : nkv_all) : 0;
  const int nMax = max(nA, nB);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  nfb_pdl_wait();      // everything above overlapped the previous kernel's tail
  nfb_pdl_trigger();
  if (threadIdx.x == 0) { tl(p, 0); tl_cta(p, 0); }   // (debug stamps: the CTA's span starts behind the dependency wait)

  if (warp == 8) {
    // ===================== TMA producer (lane 0) + key-mask staging (all lanes) =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, nB > 0 ? 32768u : 16384u);
      tma_load_4d(smem + SM_Q, &tmQ, q_full, 0, h, q0A, b);
      if (nB > 0) tma_load_4d(smem + SM_Q + 16384, &tmQ, q_full, 0, h, q0B, b);
    }
    const float* kmask_p = (p.mask_mode & 2) ? p.key_mask + (long long)b * p.Skv : nullptr;
    float* smask = reinterpret_cast<float*>(smem + SM_MASK);
    for (int i = 0; i < nMax; ++i) {
      const int st = i % KV_STAGES;
      mbar_wait(&kv_empty[st], ((i / KV_STAGES) & 1) ^ 1);
      const int j0 = i * TILE_K;
      if (kmask_p || j0 + TILE_K > p.Skv) {
        // additive mask of this key block in the log2 domain; keys beyond Skv get -inf (their exponential is exactly 0)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kj = j0 + lane * 4 + e;
          smask[st * 128 + lane * 4 + e] = kj < p.Skv ? (kmask_p ? __ldg(kmask_p + kj) * LOG2E : 0.f) : -INFINITY;
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(&kv_full[st], 32768u);   // release: the mask words above are visible to whoever acquires kv_full
        tma_load_4d(smem + SM_K + st * 16384, &tmK, &kv_full[st], 0, h / p.kv_group, i * TILE_K, b);
        tma_load_4d(smem + SM_V + st * 16384, &tmV, &kv_full[st], 0, h / p.kv_group, i * TILE_K, b);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(0, 0, TILE_Q, TILE_K);   // S = Q K^T : A, B K-major, N = 128
      const uint32_t idesc_o = make_idesc(0, 1, TILE_Q, HDIM);     // O += P V  : A K-major, B MN-major, N = 64
      const uint32_t sQ = smem_u32(smem + SM_Q), sK = smem_u32(smem + SM_K), sV = smem_u32(smem + SM_V), sP = smem_u32(smem + SM_P);
      auto issue_s = [&](int g, int st) {
#pragma unroll
        for (int k = 0; k < HDIM / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)(g * 128), make_smem_desc(sQ + g * 16384 + k * 32, 16, 1024),
                    make_smem_desc(sK + st * 16384 + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
      };
      auto issue_pv = [&](int g, int st, bool acc) {
#pragma unroll
        for (int k = 0; k < TILE_K / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)(256 + g * 64), make_smem_desc(sP + g * 32768 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc(sV + st * 16384 + k * 2048, 8192, 1024), idesc_o, (acc || k > 0) ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (nA > 0) { issue_s(0, 0); umma_commit(&s_full[0]); }
      if (nB > 0) { issue_s(1, 0); umma_commit(&s_full[1]); }
      for (int i = 0; i < nMax; ++i) {
        const int st = i % KV_STAGES, stn = (i + 1) % KV_STAGES;
        const bool more = (i + 1 < nMax);
        if (more) { mbar_wait(&kv_full[stn], ((i + 1) / KV_STAGES) & 1); tc_fence_after(); }
        // per group: once P_g(i) is in shared memory, S_g(i) has been consumed -> issue S_g(i+1) FIRST (the group's next
        // softmax can start as soon as it lands), then the P_g(i).V(i) product
        if (i < nA) {
          mbar_wait(&p_full[0], i & 1);
          tl(p, 16 + 4 * i);
          tc_fence_after();
          if (i + 1 < nA) { issue_s(0, stn); umma_commit(&s_full[0]); }
          issue_pv(0, st, i > 0);
          umma_commit(&pv_done[0]);
        }
        tl(p, 17 + 4 * i);
        if (i < nB) {
          mbar_wait(&p_full[1], i & 1);
          tl(p, 18 + 4 * i);
          tc_fence_after();
          if (i + 1 < nB) { issue_s(1, stn); umma_commit(&s_full[1]); }
          issue_pv(1, st, i > 0);
          umma_commit(&pv_done[1]);
        }
        umma_commit(&kv_empty[st]);   // every product that reads K / V block i has been issued
        tl(p, 19 + 4 * i);
      }
    }
  } else {
    // ===================== softmax groups (thread = query row) =====================
    const int g = warp >> 2, quarter = warp & 3;
    const int n = g == 0 ? nA : nB;
    const int r = quarter * 32 + lane;
    const int qi = (g == 0 ? q0A : q0B) + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t sP_row = smem_u32(smem + SM_P) + (uint32_t)g * 32768u + (uint32_t)r * 128u;
    const int sw = r & 7;
    const float sl2 = p.scale * LOG2E;
    const float* kmask = (p.mask_mode & 2) ? p.key_mask + (long long)b * p.Skv : nullptr;
    float m2 = -INFINITY, l = 0.f;    // running maximum (log2 domain) and row sum
    for (int i = 0; i < n; ++i) {
      mbar_wait(&s_full[g], i & 1);
      if (quarter == 0 && lane == 0) tl(p, 64 + g * 64 + 8 * i);
      tc_fence_after();
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(lane_addr + (uint32_t)(g * 128 + c * 32), sr + c * 32);
      tmem_ld_wait();
      if (quarter == 0 && lane == 0) tl(p, 64 + g * 64 + 8 * i + 1);
      float t[128];
      float tmul = 1.f;   // exponent of element j = t[j] * tmul - m2
      const int j0 = i * TILE_K;
      const bool diag = causal && (j0 + TILE_K - 1 > (g == 0 ? q0A : q0B));   // some (row, key) of this block has key > row
      const bool tail = j0 + TILE_K > p.Skv;
      float bm;
      if (!diag && !tail && !kmask) {
        // eight independent running maxima: a single chain of 127 dependent FMNMX would cost ~500 clk of pure latency
        // (there are only two softmax warps per scheduler to hide it)
        float mx[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) mx[a] = __uint_as_float(sr[a]);
#pragma unroll
        for (int j = 8; j < 128; ++j) mx[j & 7] = fmaxf(mx[j & 7], __uint_as_float(sr[j]));
        bm = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) * sl2;
        // the scores stay RAW: the scale rides in the FFMA that forms the exponent below (one instruction per element
        // instead of a multiply here and a subtract there)
#pragma unroll
        for (int j = 0; j < 128; ++j) t[j] = __uint_as_float(sr[j]);
        tmul = sl2;
      } else if (!diag) {
        // key mask and / or ragged tail, no causal diagonal: one FFMA per element with the staged log2-domain mask
        // (-inf beyond Skv); 16-byte broadcast reads of the mask row
        mbar_wait(&kv_full[i % KV_STAGES], (i / KV_STAGES) & 1);   // acquire the producer's mask words (long complete: S(i) needed this stage)
        const float4* sm4 = reinterpret_cast<const float4*>(smem + SM_MASK + (i % KV_STAGES) * 512);
        float mx[8];
#pragma unroll
        for (int j4 = 0; j4 < 32; ++j4) {
          const float4 mk = sm4[j4];
          t[4 * j4 + 0] = fmaf(__uint_as_float(sr[4 * j4 + 0]), sl2, mk.x);
          t[4 * j4 + 1] = fmaf(__uint_as_float(sr[4 * j4 + 1]), sl2, mk.y);
          t[4 * j4 + 2] = fmaf(__uint_as_float(sr[4 * j4 + 2]), sl2, mk.z);
          t[4 * j4 + 3] = fmaf(__uint_as_float(sr[4 * j4 + 3]), sl2, mk.w);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) mx[a] = t[a];
#pragma unroll
        for (int j = 8; j < 128; ++j) mx[j & 7] = fmaxf(mx[j & 7], t[j]);
        bm = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      } else if (!tail && !kmask) {
        // causal diagonal block: keys j0 + j > qi are REPLACED by -1e4 -- as the RAW score whose scaled value that is, so this block
        // too leaves the scale to the exponent's FFMA (compare + select per element, no multiply)
        const int lim = qi - j0;
        const float negraw = -1e4f / p.scale;
        float mx[8];
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          const float v = (j > lim) ? negraw : __uint_as_float(sr[j]);
          t[j] = v;
          if (j < 8) mx[j] = v; else mx[j & 7] = fmaxf(mx[j & 7], v);
        }
        bm = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) * sl2;
        tmul = sl2;
      } else {
        bm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          const int kj = j0 + j;
          float v = __uint_as_float(sr[j]) * sl2;
          if (causal && kj > qi) v = -1e4f * LOG2E;          // the reference REPLACES masked scores by -1e4
          if (kmask) v += (kj < p.Skv ? __ldg(kmask + kj) : 0.f) * LOG2E;
          if (kj >= p.Skv) v = -INFINITY;
          t[j] = v;
          bm = fmaxf(bm, v);
        }
      }
      // lazy rescale: advance the maximum only when some row of this warp grew by more than 2^8
      const bool need = (bm - m2) > 8.f;   // also true on the first block (m2 = -inf)
      const bool rescale = __any_sync(0xffffffffu, need);
      if (quarter == 0 && lane == 0) tl(p, 64 + g * 64 + 8 * i + 2);
      float c = 1.f;
      if (rescale) {
        const float m_new = fmaxf(m2, bm);
        c = (m2 == -INFINITY) ? 0.f : ex2(m2 - m_new);
        l *= c;
        m2 = m_new;
      }
      // S_g(i) was issued BEFORE P_g(i-1).V(i-1): wait for that product before overwriting its P operand (and before
      // touching O_g); it has normally retired long before the row maximum above is known
      if (i > 0) { mbar_wait(&pv_done[g], (i - 1) & 1); tc_fence_after(); }
      float rsa[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      // P (bf16) -> shared memory, K-major SWIZZLE_128B: 16-byte slot c of row r at c ^ (r & 7), two 64-key panels
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2_mix<POLY>(fmaf(t[cc * 8 + 2 * e], tmul, -m2), e), p1 = ex2_mix<POLY>(fmaf(t[cc * 8 + 2 * e + 1], tmul, -m2), e);
          rsa[(2 * e) & 7] += p0;
          rsa[(2 * e + 1) & 7] += p1;
          __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
          w[e] = *reinterpret_cast<uint32_t*>(&b2);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP_row + (uint32_t)(cc >> 3) * 16384u + (uint32_t)(((cc & 7) ^ sw) * 16)),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      }
      l += ((rsa[0] + rsa[1]) + (rsa[2] + rsa[3])) + ((rsa[4] + rsa[5]) + (rsa[6] + rsa[7]));
      if (quarter == 0 && lane == 0) tl(p, 64 + g * 64 + 8 * i + 3);
      // the (rare) rescale of O_g comes last, when the score registers are dead.  O_g holds the products of blocks < i,
      // all complete (pv_done above); P.V of this block is only issued after the arrive below.
      if (rescale && i > 0) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t orr[32];
          tmem_ld32(lane_addr + (uint32_t)(256 + g * 64 + cc * 32), orr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * c);
          tmem_st32(lane_addr + (uint32_t)(256 + g * 64 + cc * 32), orr);
        }
        tmem_st_wait();
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      if (quarter == 0 && lane == 0) tl(p, 64 + g * 64 + 8 * i + 4);
    }
    if (n > 0) {
      mbar_wait(&pv_done[g], (n - 1) & 1);
      tc_fence_after();
      const float lc = fmaxf(l, 1e-8f);   // the reference's 1e-8 floor (functional/activation.py:95)
      const float inv = 1.f / lc;
      __nv_bfloat16* orow = p.o + b * p.o_sb + h * p.o_sh + (long long)qi * p.o_ss;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t orr[32];
        tmem_ld32(lane_addr + (uint32_t)(256 + g * 64 + cc * 32), orr);
        tmem_ld_wait();
        if (qi < p.S) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(orr[c * 8 + 2 * e]) * inv, __uint_as_float(orr[c * 8 + 2 * e + 1]) * inv);
              w[e] = *reinterpret_cast<uint32_t*>(&b2);
            }
            *reinterpret_cast<uint4*>(orow + cc * 32 + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      if (p.lse && qi < p.S) p.lse[(long long)bh * p.S + qi] = (m2 + log2f(lc)) * 0.6931471805599453f;
    }
  }

  if (threadIdx.x == 0) tl(p, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (threadIdx.x == 0) tl_cta(p, 1);
}


// =====================================================================================================================
// BACKWARD.  One CTA per (batch, head, 128-key block j) -- 192 threads: warps 0-3 = elementwise group (thread = one
// query row of the current 128-query block i), warp 4 = TMA producer, warp 5 = MMA issuer.  For every query block i
// (i >= j when causal) five tensor-core products run on tiles that TMA delivers as plain [128 rows][64] SWIZZLE_128B
// boxes -- no transposed copy of anything is ever made, the transposes are expressed as MN-major UMMA operands:
//     S   = Q_i K_j^T          (A = Q_i K-major,       B = K_j K-major,   128x128x64)   -> TMEM cols [0,128)
//     dP  = dO_i V_j^T         (A = dO_i K-major,      B = V_j K-major,   128x128x64)   -> TMEM cols [128,256)
//     dV += P^T dO_i           (A = P  MN-major (M=k), B = dO_i MN-major, 128x64x128)   -> TMEM cols [256,320)
//     dK += dS^T Q_i           (A = dS MN-major (M=k), B = Q_i MN-major,  128x64x128)   -> TMEM cols [320,384)
//     dQ_i = dS K_j            (A = dS K-major  (M=q), B = K_j MN-major,  128x64x128)   -> TMEM cols [384,448)
// P = exp2(S*scale*log2e - LSE_i*log2e) and dS = P o (dP - delta_i) * scale are formed by the elementwise group
// straight from TMEM (thread = row, so LSE_i / delta_i are per-thread scalars) and written as bf16 into ONE shared
// memory layout ([two 64-key panels][128 query rows][128 B], SWIZZLE_128B) that serves both as the K-major operand
// (M = queries) and, with LBO = 16 KB, as the MN-major operand (M = keys).  dK_j / dV_j accumulate in TMEM over the
// whole query loop; the dQ_i partial of each (i, j) pair is drained TMEM -> registers -> shared memory (bf16) and added
// straight into the (zero-initialised, strided) bf16 dq tensor by a TMA reduce-add (cp.reduce.async.bulk.tensor .add):
// the L2 reduction units turned out to be the limiter of this kernel (~0.75 TB/s of fp32 adds), bf16 partials halve
// that traffic and remove the fp32 accumulation buffer and its conversion pass.  Masked scores carry P = 0, hence dS = 0 (the reference REPLACES them
// by -1e4, models/language/gpt2.py:194-200, so no gradient flows through them).
constexpr int BWD_THREADS = 320;        // warps 0-7 elementwise (two column halves x four lane quarters), warp 8 TMA, warp 9 MMA
constexpr int SB_K = 0;                 // K_j   16 KB
constexpr int SB_V = 16384;             // V_j   16 KB
constexpr int SB_Q = 32768;             // Q_i   2 stages x 16 KB
constexpr int SB_DO = 65536;            // dO_i  2 stages x 16 KB
constexpr int SB_P = 98304;             // P     32 KB
constexpr int SB_DS = 131072;           // dS    32 KB
constexpr int SB_DQ = 163840;           // dQ staging, bf16 [128 rows][128 B] = 16 KB (32 KB reserved)
constexpr int SB_BAR = 196608;
constexpr int SB_MASK = SB_BAR + 256;   // 128 fp32: additive key mask of this CTA's key block (log2 domain), -inf beyond Skv
constexpr int SB_TOTAL = SB_MASK + 512 + 1024 + 1024;

struct AttnBwdTcParams {
  const float* lse;     // (B, H, S) natural log
  const float* delta;   // (B, H, S) rowsum(dO o O): from k_attn_bwd_prep (or the caller)
  const __nv_bfloat16* o;   // forward output, same (b, s, h, d) strides as dO
  long long o_sb, o_ss, o_sh;
  __nv_bfloat16 *dk, *dv;
  long long kv_sb, kv_ss, kv_sh;
  int B, H, S, Skv;
  float scale;
  int mask_mode;
  const float* key_mask;
  unsigned long long* timeline;
  int kv_group;   // query heads per key/value head: k / v / dk / dv have H / kv_group heads, the CTA loops over the group
  int kv_split;   // CTAs per key/value head (few-kv-head grids): CTA head h reads K/V head h / kv_split and writes dk/dv slice h
  int dq_f32;     // 1: tmdQ is the 3-D fp32 map of a (B*H, S, 64) accumulation buffer (long sequences), 0: the 4-D bf16 map of dq
};
__device__ __forceinline__ void tlb(const AttnBwdTcParams& p, int slot) {
  if (p.timeline && blockIdx.x == 0 && slot < 256) p.timeline[slot] = (unsigned long long)clock64();
}

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// Issue order of the MMA thread (tensor-core products retire in issue order) and the matching order of the elementwise
// warps; iteration `it` handles query block i0 + it:
//   MMA :  S(0) dP(0) | loop it: [p_full(it)] dV(it) ; S(it+1) ; [ds_full(it)] dK(it) ; [dq_free(it-1)] dQ(it) ; dP(it+1)
//   ELEM:  loop it: [s_full(it)] P(it) ; (it > 0: [dq_full(it-1)] drain dQ(it-1)) ; [dp_full(it)] dS(it)   ... final drain
// S(it+1) is issued as soon as P(it) has been written (the score columns are free then), so the next score block is in TMEM
// long before the elementwise warps finish the dQ drain and dS of block it; the dS buffer of block it is only rewritten
// after dq_full(it).
template <int POLY>
__global__ void __launch_bounds__(BWD_THREADS, 1)
k_flash_bwd_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmdQ, const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + SB_BAR);
  uint64_t* kv_full = bars;          // 1
  uint64_t* q_full = bars + 1;       // 2 stages (Q_i + dO_i)
  uint64_t* q_empty = bars + 3;      // 2
  uint64_t* s_full = bars + 5;       // S ready (and: every product issued before it has retired)
  uint64_t* dp_full = bars + 6;      // dP ready
  uint64_t* p_full = bars + 7;       // P in smem (256 arrivals)
  uint64_t* ds_full = bars + 8;      // dS in smem (256 arrivals)
  uint64_t* dq_full = bars + 9;      // dQ_i partial in TMEM; dK(it) / dQ(it) have retired -> dS buffer free
  uint64_t* dq_free = bars + 10;     // dQ TMEM columns drained (256 arrivals)
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool causal = (p.mask_mode & 1) != 0;
  // 1-D grid, key-block-major: with a causal mask key block 0 meets every query block -- all of those CTAs go first
  // CTA = (key block, batch, KEY/VALUE head h).  Its iterations run over the kv_group query heads sharing that head times
  // the query blocks that see the key block, so dK / dV accumulate over the whole group in TMEM (grouped-query attention
  // needs no separate reduction); iteration `it` -> query head hq_of(it), first query row q0_of(it).
  const int Hkv = p.H / p.kv_group;
  const int nbh = p.B * Hkv;
  const int jb = (int)blockIdx.x / nbh, bh = (int)blockIdx.x % nbh, b = bh / Hkv, h = bh % Hkv;
  const int k0 = jb * TILE_K;
  const int nq = (p.S + TILE_Q - 1) / TILE_Q;
  const int i0 = causal ? min(jb, nq) : 0;      // query blocks entirely above the diagonal see none of these keys
  const int nqb = nq - i0;
  const int niter = nqb * p.kv_group;
  auto hq_of = [&](int it) { return p.kv_group == 1 ? h : h * p.kv_group + it / nqb; };
  auto q0_of = [&](int it) { return (i0 + (p.kv_group == 1 ? it : it % nqb)) * TILE_Q; };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmdQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(p_full, 256); mbar_init(ds_full, 256); mbar_init(dq_full, 1); mbar_init(dq_free, 256);
    fence_barrier_init();
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  nfb_pdl_wait();
  nfb_pdl_trigger();
  if (threadIdx.x == 0) tlb(p, 0);

  if (warp == 8) {
    // ===================== TMA producer (lane 0) =====================
    if (niter > 0) {
      if (lane == 0) {
        mbar_expect_tx(kv_full, 32768u);
        tma_load_4d(smem + SB_K, &tmK, kv_full, 0, h / p.kv_split, k0, b);
        tma_load_4d(smem + SB_V, &tmV, kv_full, 0, h / p.kv_split, k0, b);
      }
      for (int it = 0; it < niter; ++it) {
        const int st = it & 1;
        const int q0 = q0_of(it), hq = hq_of(it);
        mbar_wait(&q_empty[st], ((it >> 1) & 1) ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&q_full[st], 32768u);
          tma_load_4d(smem + SB_Q + st * 16384, &tmQ, &q_full[st], 0, hq, q0, b);
          tma_load_4d(smem + SB_DO + st * 16384, &tmdO, &q_full[st], 0, hq, q0, b);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0 && niter > 0) {
      const uint32_t idesc_s = make_idesc(0, 0, 128, 128);    // S, dP : A K-major, B K-major, N = 128
      const uint32_t idesc_kv = make_idesc(1, 1, 128, 64);    // dV, dK: A MN-major, B MN-major, N = 64
      const uint32_t idesc_q = make_idesc(0, 1, 128, 64);     // dQ    : A K-major, B MN-major, N = 64
      const uint32_t sK = smem_u32(smem + SB_K), sV = smem_u32(smem + SB_V), sQ = smem_u32(smem + SB_Q), sdO = smem_u32(smem + SB_DO);
      const uint32_t sP = smem_u32(smem + SB_P), sdS = smem_u32(smem + SB_DS);
      auto issue_s = [&](int st) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + 0u, make_smem_desc(sQ + st * 16384 + k * 32, 16, 1024), make_smem_desc(sK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      auto issue_dp = [&](int st) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + 128u, make_smem_desc(sdO + st * 16384 + k * 32, 16, 1024), make_smem_desc(sV + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(dp_full);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int it = 0; it < niter; ++it) {
        const int st = it & 1;
        const bool more = it + 1 < niter;
        // dV += P^T dO_i : contraction over the 128 query rows (8 steps of 16 rows = 2048 B in both operands)
        mbar_wait(p_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + 256u, make_smem_desc(sP + k * 2048, 16384, 1024), make_smem_desc(sdO + st * 16384 + k * 2048, 8192, 1024), idesc_kv,
                    (it > 0 || k > 0) ? 1u : 0u);
        // S(it+1) goes out right behind dV(it): the S columns were consumed when P(it) was written (p_full), so the next score block
        // -- and the ~1.6 k clk from its issue to the elementwise warps seeing s_full -- no longer waits for dS(it); its s_full
        // also certifies dV(it), i.e. that the P buffer may be rewritten
        if (more) {
          mbar_wait(&q_full[st ^ 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(st ^ 1);
        }
        mbar_wait(ds_full, it & 1);
        tc_fence_after();
        // dK += dS^T Q_i ; dQ_i = dS K_j
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + 320u, make_smem_desc(sdS + k * 2048, 16384, 1024), make_smem_desc(sQ + st * 16384 + k * 2048, 8192, 1024), idesc_kv,
                    (it > 0 || k > 0) ? 1u : 0u);
        if (it > 0) { mbar_wait(dq_free, (it - 1) & 1); tc_fence_after(); }   // the previous dQ partial has been drained
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + 384u, make_smem_desc(sdS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024), make_smem_desc(sK + k * 2048, 8192, 1024), idesc_q,
                    k > 0 ? 1u : 0u);
        umma_commit(dq_full);          // dQ_i ready; dS buffer and the Q_i / dO_i stage are no longer read
        umma_commit(&q_empty[st]);
        if (more) issue_dp(st ^ 1);    // dP columns were consumed when dS(it) was written
      }
    }
  } else {
    // ===================== elementwise warps: thread = (query row, 64-key half) =====================
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // this thread's 64 keys are one 64-key panel of the [panel][row][128 B] operand layout
    const uint32_t sP_row = smem_u32(smem + SB_P) + (uint32_t)half * 16384u + (uint32_t)r * 128u;
    const uint32_t sdS_row = smem_u32(smem + SB_DS) + (uint32_t)half * 16384u + (uint32_t)r * 128u;
    const uint32_t sdQ_row = smem_u32(smem + SB_DQ) + (uint32_t)r * 128u;   // dQ staging row (bf16, 64 columns)
    const int sw = r & 7;
    const float sl2 = p.scale * LOG2E;
    const float* kmask = (p.mask_mode & 2) ? p.key_mask + (long long)b * p.Skv : nullptr;
    const int kb0 = k0 + half * 64;                       // first key of this thread's half
    const bool tail = kb0 + 64 > p.Skv;
    // key mask / ragged tail: stage this CTA's 128 mask values once (log2 domain, -inf beyond Skv)
    const bool use_smask = (kmask != nullptr) || (k0 + TILE_K > p.Skv);
    const float* smask_h = reinterpret_cast<const float*>(smem + SB_MASK) + half * 64;
    if (use_smask) {
      if (threadIdx.x < 128) {
        const int kj = k0 + (int)threadIdx.x;
        reinterpret_cast<float*>(smem + SB_MASK)[threadIdx.x] = kj < p.Skv ? (kmask ? __ldg(kmask + kj) * LOG2E : 0.f) : -INFINITY;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }

    auto drain_dq = [&](int it_d) {
      // dQ partial of iteration it_d: TMEM -> registers -> shared memory (fp32, one 32-column SWIZZLE_128B box per half) -> TMA reduce-add
      mbar_wait(dq_full, it_d & 1);
      tc_fence_after();
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous reduce has read the staging buffer
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t dq[32];
      tmem_ld32(lane_addr + (uint32_t)(384 + half * 32), dq);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);
      if (p.dq_f32) {
        // fp32 staging: one [128 rows][32 columns = 128 B] SWIZZLE_128B box per column half, added into the fp32 buffer
        const uint32_t row32 = smem_u32(smem + SB_DQ) + (uint32_t)half * 16384u + (uint32_t)r * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row32 + (uint32_t)((c ^ sw) * 16)),
                       "r"(dq[4 * c]), "r"(dq[4 * c + 1]), "r"(dq[4 * c + 2]), "r"(dq[4 * c + 3]) : "memory");
      } else {
      // bf16 staging box [128 rows][64 columns = 128 B], SWIZZLE_128B: this thread's 32 columns are 16-byte slots 4*half .. +3
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(dq[c * 8 + 2 * e]), __uint_as_float(dq[c * 8 + 2 * e + 1]));
          w[e] = *reinterpret_cast<uint32_t*>(&b2);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdQ_row + (uint32_t)(((half * 4 + c) ^ sw) * 16)),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        if (p.dq_f32) {
          const int bhq = b * p.H + hq_of(it_d);
          tma_reduce_add_3d(&tmdQ, smem_u32(smem + SB_DQ), 0, q0_of(it_d), bhq);
          tma_reduce_add_3d(&tmdQ, smem_u32(smem + SB_DQ) + 16384u, 32, q0_of(it_d), bhq);
        } else {
          tma_reduce_add_4d(&tmdQ, smem_u32(smem + SB_DQ), 0, hq_of(it_d), q0_of(it_d), b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };

    // the row's LSE is loaded one block AHEAD: as the first instruction of an iteration its ~1.2 k clk global-load latency sat
    // between dS(it-1) and P(it) on the critical path of these warps (profiles/r02_attn_timeline.txt)
    auto lse_of = [&](int it_l) -> float {
      const int qi_l = q0_of(it_l) + r;
      const long long row_l = ((long long)b * p.H + hq_of(it_l)) * p.S;
      // invalid (padding) rows: LSE = +inf makes P = 0, so nothing of them reaches dV / dK
      return qi_l < p.S ? p.lse[row_l + qi_l] : INFINITY;
    };
    auto delta_of = [&](int it_l) -> float {
      const int qi_l = q0_of(it_l) + r;
      const long long row_l = ((long long)b * p.H + hq_of(it_l)) * p.S;
      return (p.delta && qi_l < p.S) ? p.delta[row_l + qi_l] : 0.f;
    };
    float lse_next = niter > 0 ? lse_of(0) : 0.f;
    float dlt_next = niter > 0 ? delta_of(0) : 0.f;
    for (int it = 0; it < niter; ++it) {
      const int q0 = q0_of(it), qi = q0 + r;
      const long long row0 = ((long long)b * p.H + hq_of(it)) * p.S;   // (b, query head) row of lse / delta
      const float lse2 = lse_next * LOG2E;
      const float dlt_ext = dlt_next;
      if (it + 1 < niter) { lse_next = lse_of(it + 1); dlt_next = delta_of(it + 1); }
      // ---- P = exp2(score - LSE)
      mbar_wait(s_full, it & 1);
      if (threadIdx.x == 0) tlb(p, 16 + 8 * it);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld32(lane_addr + (uint32_t)(half * 64), sr);
      tmem_ld32(lane_addr + (uint32_t)(half * 64 + 32), sr + 32);
      tmem_ld_wait();
      uint32_t pk[32];   // P as packed bf16 pairs (kept for dS)
      const bool diag = causal && (kb0 + 63 > q0);
      if (!diag && !tail && !kmask) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float p0 = ex2_mix<POLY>(fmaf(__uint_as_float(sr[2 * j]), sl2, -lse2), j & 3), p1 = ex2_mix<POLY>(fmaf(__uint_as_float(sr[2 * j + 1]), sl2, -lse2), j & 3);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
          pk[j] = *reinterpret_cast<uint32_t*>(&b2);
        }
      } else if (!diag) {
        // key mask and / or ragged tail, no causal diagonal: staged log2-domain mask (-inf beyond Skv), broadcast 16-byte reads
        const float4* sm4 = reinterpret_cast<const float4*>(smask_h);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 mk = sm4[j4];
          const float p0 = ex2(fmaf(__uint_as_float(sr[4 * j4 + 0]), sl2, mk.x) - lse2), p1 = ex2(fmaf(__uint_as_float(sr[4 * j4 + 1]), sl2, mk.y) - lse2);
          const float p2 = ex2(fmaf(__uint_as_float(sr[4 * j4 + 2]), sl2, mk.z) - lse2), p3 = ex2(fmaf(__uint_as_float(sr[4 * j4 + 3]), sl2, mk.w) - lse2);
          __nv_bfloat162 b01 = __floats2bfloat162_rn(p0, p1), b23 = __floats2bfloat162_rn(p2, p3);
          pk[2 * j4] = *reinterpret_cast<uint32_t*>(&b01);
          pk[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&b23);
        }
      } else if (!tail && !kmask) {
        // causal diagonal: keys kb0 + j > qi are REPLACED by -1e4, whose exponential underflows to exactly 0
        const int lim = qi - kb0;
        const float pneg = ex2(-1e4f * LOG2E - lse2);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float p0 = ex2_mix<POLY>(fmaf(__uint_as_float(sr[2 * j]), sl2, -lse2), j & 3), p1 = ex2_mix<POLY>(fmaf(__uint_as_float(sr[2 * j + 1]), sl2, -lse2), j & 3);
          if (2 * j > lim) p0 = pneg;
          if (2 * j + 1 > lim) p1 = pneg;
          __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
          pk[j] = *reinterpret_cast<uint32_t*>(&b2);
        }
      } else {
        const float neg = -1e4f * LOG2E;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int kj = kb0 + 2 * j + e;
            float t = __uint_as_float(sr[2 * j + e]) * sl2;
            if (causal && kj > qi) t = neg;                                   // REPLACED by -1e4
            if (kmask) t += (kj < p.Skv ? __ldg(kmask + kj) : 0.f) * LOG2E;
            v[e] = (kj < p.Skv) ? ex2(t - lse2) : 0.f;
          }
          __nv_bfloat162 b2 = __floats2bfloat162_rn(v[0], v[1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&b2);
        }
      }
      // (P buffer: dV(it-1) retired before s_full(it), which was committed after it)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP_row + (uint32_t)((c ^ sw) * 16)),
                     "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
      if (threadIdx.x == 0) tlb(p, 16 + 8 * it + 1);
      // ---- the dQ partial of the previous block (its products ran while the exponentials above were computed)
      if (it > 0) drain_dq(it - 1);
      if (threadIdx.x == 0) tlb(p, 16 + 8 * it + 2);
      // ---- dS = P o (dP - delta) * scale   (dS buffer: dK(it-1), dQ(it-1) retired -- dq_full(it-1) awaited in the drain)
      const float dlt = dlt_ext;   // delta_i = rowsum(dO_i o O_i): formed by k_attn_bwd_prep (or passed in), loaded one block ahead
      if (threadIdx.x == 0) tlb(p, 16 + 8 * it + 5);
      mbar_wait(dp_full, it & 1);
      if (threadIdx.x == 0) tlb(p, 16 + 8 * it + 3);
      tc_fence_after();
      tmem_ld32(lane_addr + (uint32_t)(128 + half * 64), sr);
      tmem_ld32(lane_addr + (uint32_t)(128 + half * 64 + 32), sr + 32);
      tmem_ld_wait();
      const float dlt_s = dlt * p.scale;   // dS = P (dP - delta) scale = P * fma(dP, scale, -delta scale): one instruction less per element
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[j]));
        const float d0 = pf.x * fmaf(__uint_as_float(sr[2 * j]), p.scale, -dlt_s), d1 = pf.y * fmaf(__uint_as_float(sr[2 * j + 1]), p.scale, -dlt_s);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(d0, d1);
        pk[j] = *reinterpret_cast<uint32_t*>(&b2);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS_row + (uint32_t)((c ^ sw) * 16)),
                     "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(ds_full);
      if (threadIdx.x == 0) tlb(p, 16 + 8 * it + 4);
    }
    if (niter > 0) drain_dq(niter - 1);
    if (threadIdx.x == 0) tlb(p, 1);
    // ---- epilogue: dK_j, dV_j (thread = key row, 32 of the 64 columns); dq_full(niter-1) certifies every product
    const int kj = k0 + r;
    {
      __nv_bfloat16* dkrow = p.dk + b * p.kv_sb + h * p.kv_sh + (long long)kj * p.kv_ss + half * 32;
      __nv_bfloat16* dvrow = p.dv + b * p.kv_sb + h * p.kv_sh + (long long)kj * p.kv_ss + half * 32;
      if (niter > 0) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          uint32_t acc[32];
          tmem_ld32(lane_addr + (uint32_t)((which == 0 ? 256 : 320) + half * 32), acc);   // warp-collective: every lane executes it
          tmem_ld_wait();
          if (kj < p.Skv) {
            __nv_bfloat16* dst = which == 0 ? dvrow : dkrow;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(acc[c * 8 + 2 * e]), __uint_as_float(acc[c * 8 + 2 * e + 1]));
                w[e] = *reinterpret_cast<uint32_t*>(&b2);
              }
              *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      } else if (kj < p.Skv) {
        // no query block sees these keys (cannot happen for causal self-attention with S >= Skv, kept for safety)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(dkrow + c * 8) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(dvrow + c * 8) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    // the staging buffer must have been read before the CTA (and its shared memory) goes away; the reductions themselves
    // complete at the latest with the grid
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (threadIdx.x == 0) tlb(p, 2);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// In front of the backward: dq := 0 (skipped when dq is NULL: the fp32 accumulation path clears its own buffer) and
// delta[b, h, s] = sum_d dO[b, s, h, d] * O[b, s, h, d] (skipped when delta is NULL: the caller supplied it).  Eight threads per
// (b, s, h) row, one 16-byte access each; consecutive rows walk the heads first, so a warp touches 4 x 128 contiguous bytes
// of the (b, s, h, d) tensors.
__global__ void __launch_bounds__(256)
k_attn_bwd_prep(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta, __nv_bfloat16* __restrict__ dq,
                int B, int H, int S, long long o_sb, long long o_ss, long long o_sh, long long q_sb, long long q_ss, long long q_sh) {
  nfb_pdl_enter();
  // 32-bit index arithmetic (the launcher checks B*S*H*8 < 2^31): four 64-bit divisions per row were most of this kernel's
  // instructions (ncu: 62 % issue-active at 18 % of DRAM bandwidth)
  const unsigned total = (unsigned)B * (unsigned)S * (unsigned)H * 8u;
  const unsigned start = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;   // stride is a multiple of 32
  for (unsigned i0 = start & ~7u; i0 < total; i0 += stride) {   // the 8 threads of a row stay together (shuffles below)
    const unsigned i = i0 + (start & 7u);
    const int c = (int)(i & 7u);
    unsigned row = i >> 3;
    const unsigned t1 = row / (unsigned)H;
    const int hh = (int)(row - t1 * (unsigned)H);
    const unsigned bq = t1 / (unsigned)S;
    const int sidx = (int)(t1 - bq * (unsigned)S);
    const long long bb = (long long)bq;
    if (dq) *reinterpret_cast<uint4*>(dq + bb * q_sb + (long long)sidx * q_ss + (long long)hh * q_sh + c * 8) = make_uint4(0, 0, 0, 0);
    if (delta) {
      const long long off = bb * o_sb + (long long)sidx * o_ss + (long long)hh * o_sh + c * 8;
      const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(o + off)), b4 = __ldg(reinterpret_cast<const uint4*>(dout + off));
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a4);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b4);
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = __bfloat1622float2(a2[e]), fb = __bfloat1622float2(b2[e]);
        acc0 = fmaf(fa.x, fb.x, acc0);
        acc1 = fmaf(fa.y, fb.y, acc1);
      }
      float v = acc0 + acc1;
      const unsigned gmask = 0xFFu << (threadIdx.x & 24);   // the row's 8 lanes (other rows of the warp may be past the end)
      v += __shfl_xor_sync(gmask, v, 1);
      v += __shfl_xor_sync(gmask, v, 2);
      v += __shfl_xor_sync(gmask, v, 4);
      if (c == 0) delta[(bb * H + hh) * S + sidx] = v;
    }
  }
}

// dq (bf16, strided (b, s, h, d)) = fp32 accumulation buffer (B, H, S, 64)   [long-sequence path]
__global__ void k_dq_convert(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int B, int H, int S, long long sb, long long ss, long long sh) {
  nfb_pdl_enter();
  const long long total = (long long)B * H * S * 8;   // 8 threads per row, 8 elements each
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7) * 8;
    const long long row = i >> 3;
    const int sidx = (int)(row % S);
    const long long bhh = row / S;
    const int hh = (int)(bhh % H);
    const long long bb = bhh / H;
    const float4 a = *reinterpret_cast<const float4*>(acc + row * 64 + c), bq = *reinterpret_cast<const float4*>(acc + row * 64 + c + 4);
    __nv_bfloat162 w0 = __floats2bfloat162_rn(a.x, a.y), w1 = __floats2bfloat162_rn(a.z, a.w), w2 = __floats2bfloat162_rn(bq.x, bq.y), w3 = __floats2bfloat162_rn(bq.z, bq.w);
    *reinterpret_cast<uint4*>(dq + bb * sb + (long long)sidx * ss + (long long)hh * sh + c) =
        make_uint4(*reinterpret_cast<uint32_t*>(&w0), *reinterpret_cast<uint32_t*>(&w1), *reinterpret_cast<uint32_t*>(&w2), *reinterpret_cast<uint32_t*>(&w3));
  }
}

// dk / dv (bf16, k / v strides, Hkv heads) = sum over the `split` partial heads of a contiguous (B, Skv, Hkv*split, 64) bf16 buffer
__global__ void k_kv_split_sum(const __nv_bfloat16* __restrict__ part, __nv_bfloat16* __restrict__ out, int B, int Skv, int Hkv, int split,
                               long long sb, long long ss, long long sh) {
  nfb_pdl_enter();
  const long long total = (long long)B * Skv * Hkv * 8;   // 8 threads per (b, s, head) row, 8 elements each
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7) * 8;
    long long r = i >> 3;
    const int hk = (int)(r % Hkv); r /= Hkv;
    const int sidx = (int)(r % Skv);
    const long long bb = r / Skv;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* src = part + ((bb * Skv + sidx) * (long long)(Hkv * split) + (long long)hk * split) * 64 + c;
    for (int j = 0; j < split; ++j) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + (long long)j * 64);
      const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(v2[e]); acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
    }
    __nv_bfloat162 w0 = __floats2bfloat162_rn(acc[0], acc[1]), w1 = __floats2bfloat162_rn(acc[2], acc[3]), w2 = __floats2bfloat162_rn(acc[4], acc[5]),
                   w3 = __floats2bfloat162_rn(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + bb * sb + (long long)sidx * ss + (long long)hk * sh + c) =
        make_uint4(*reinterpret_cast<uint32_t*>(&w0), *reinterpret_cast<uint32_t*>(&w1), *reinterpret_cast<uint32_t*>(&w2), *reinterpret_cast<uint32_t*>(&w3));
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int get_encode_fn() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return NFB_FAIL("cuTensorMapEncodeTiled entry point unavailable");
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

typedef std::tuple<const void*, int, int, int, long long, long long, long long> MapKey;
std::map<MapKey, CUtensorMap> g_map_cache;
std::mutex g_map_mu;

// 4-D bf16 map over element (b, s, h, d) at base + b*sb + s*ss + h*sh + d: dims (d = 64, h, s, b), box (64, 1, 128, 1)
int make_map_bshd(CUtensorMap* out, const void* base, int B, int S, int H, long long sb, long long ss, long long sh) {
  MapKey key(base, B, S, H, sb, ss, sh);
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *out = it->second; return 0; }
  }
  if (get_encode_fn()) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)HDIM, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
  // a stride is irrelevant for an extent of 1 but must still be a legal (16-byte multiple, non-zero) value
  cuuint64_t strides[3] = {(cuuint64_t)(H > 1 ? sh : HDIM) * 2, (cuuint64_t)(S > 1 ? ss : HDIM) * 2, (cuuint64_t)(B > 1 ? sb : HDIM) * 2};
  cuuint32_t box[4] = {(cuuint32_t)HDIM, 1, (cuuint32_t)TILE_K, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return NFB_FAIL("attention_tc: cuTensorMapEncodeTiled failed (%d) B=%d S=%d H=%d strides=(%lld,%lld,%lld)", (int)r, B, S, H, sb, ss, sh);
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}


unsigned long long* g_attn_timeline = nullptr;
}  // namespace

// 3-D fp32 map over the dQ accumulation buffer (B*H, S, 64): dims (64, S, B*H), box (32, 128, 1)
int make_map_dq_f32(CUtensorMap* out, const float* base, int BH, int S) {
  if (get_encode_fn()) return -1;
  cuuint64_t dims[3] = {64, (cuuint64_t)S, (cuuint64_t)BH};
  cuuint64_t strides[2] = {64 * 4, (cuuint64_t)S * 64 * 4};
  cuuint32_t box[3] = {32, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return NFB_FAIL("attention_tc: cuTensorMapEncodeTiled (dQ fp32) failed (%d) BH=%d S=%d", (int)r, BH, S);
  return 0;
}
static int g_poly_fwd = -1, g_poly_bwd = -1;   // -1: not set yet -> NFB_ATTN_POLY="f,b" from the environment, else the defaults below
static void poly_defaults() {
  if (g_poly_fwd >= 0) return;
  int f = NFB_ATTN_POLY_FWD_DEFAULT, b = NFB_ATTN_POLY_BWD_DEFAULT;
  if (const char* e = getenv("NFB_ATTN_POLY")) sscanf(e, "%d,%d", &f, &b);
  g_poly_fwd = f < 0 ? 0 : (f > 2 ? 2 : f);
  g_poly_bwd = b < 0 ? 0 : (b > 2 ? 2 : b);
}
// share of the softmax exponentials evaluated by a polynomial on the FMA pipe instead of MUFU.EX2: n of every 4 (0, 1 or 2)
NFB_API int nfb_flash_attn_set_poly(int fwd, int bwd) {
  g_poly_fwd = fwd < 0 ? 0 : (fwd > 2 ? 2 : fwd);
  g_poly_bwd = bwd < 0 ? 0 : (bwd > 2 ? 2 : bwd);
  return 0;
}
static int g_dq_f32_blocks = 4;   // more key blocks than this per query row -> fp32 dQ accumulation (bf16 partial sums lose ~2^-9 per add)
NFB_API int nfb_flash_attn_set_dq_f32_blocks(int n) { g_dq_f32_blocks = n; return 0; }   // test hook: 0 = always fp32, large = always bf16
NFB_API int nfb_flash_attn_set_timeline(void* dev_buf) { g_attn_timeline = (unsigned long long*)dev_buf; return 0; }   // debug: 256 x u64

// true when the tcgen05 forward can take this problem (16-byte aligned bases and strides; head_dim 64)
bool nfb_flash_fwd_tc_eligible(const void* q, const void* k, const void* v, const void* o, int head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh,
                               int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_ss, int64_t o_sh, int64_t o_sb) {
  if (head_dim != HDIM) return false;
  if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15) != 0) return false;
  if ((q_sb | q_ss | q_sh | kv_sb | kv_ss | kv_sh | o_ss | o_sh | o_sb) % 8 != 0) return false;
  if (q_sb < 0 || q_ss <= 0 || q_sh < 0 || kv_sb < 0 || kv_ss <= 0 || kv_sh < 0) return false;
  return true;
}

int nfb_flash_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int S, int Skv, int64_t q_sb, int64_t q_ss,
                     int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh, float scale, int mask_mode,
                     const float* key_mask, int kv_group) {
  CUtensorMap tmQ, tmK, tmV;
  if (make_map_bshd(&tmQ, q, B, S, H, q_sb, q_ss, q_sh)) return -1;
  if (make_map_bshd(&tmK, k, B, Skv, H / kv_group, kv_sb, kv_ss, kv_sh)) return -1;
  if (make_map_bshd(&tmV, v, B, Skv, H / kv_group, kv_sb, kv_ss, kv_sh)) return -1;
  AttnTcParams p = {};
  p.kv_group = kv_group;
  p.o = (__nv_bfloat16*)o; p.lse = lse; p.B = B; p.H = H; p.S = S; p.Skv = Skv;
  p.o_sb = o_sb; p.o_ss = o_ss; p.o_sh = o_sh; p.scale = scale; p.mask_mode = mask_mode; p.key_mask = key_mask;
  p.timeline = g_attn_timeline;
  static bool configured = false;
  if (!configured) {
    NFB_CUDA(cudaFuncSetAttribute(k_flash_fwd_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    NFB_CUDA(cudaFuncSetAttribute(k_flash_fwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    NFB_CUDA(cudaFuncSetAttribute(k_flash_fwd_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    configured = true;
  }
  dim3 grid((unsigned)(nfb_cdiv(S, 2 * TILE_Q) * B * H));
  poly_defaults();
  if (g_poly_fwd == 1) NFB_CUDA(nfb_launch_pdl(k_flash_fwd_tc<1>, grid, dim3(ATC_THREADS), (size_t)SM_TOTAL, nfb_cur_stream(), tmQ, tmK, tmV, p));
  else if (g_poly_fwd == 2) NFB_CUDA(nfb_launch_pdl(k_flash_fwd_tc<2>, grid, dim3(ATC_THREADS), (size_t)SM_TOTAL, nfb_cur_stream(), tmQ, tmK, tmV, p));
  else NFB_CUDA(nfb_launch_pdl(k_flash_fwd_tc<0>, grid, dim3(ATC_THREADS), (size_t)SM_TOTAL, nfb_cur_stream(), tmQ, tmK, tmV, p));
  NFB_LAUNCH_CHECK();
  return 0;
}

// backward on the tensor cores; `delta` = rowsum(dO o O) (B, H, S) may be NULL (k_attn_bwd_prep forms it from o and dO).  dq/dk/dv use the q / kv strides.
int nfb_flash_bwd_tc(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, const float* delta, void* dq, void* dk, void* dv,
                     int B, int H, int S, int Skv, int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb,
                     int64_t o_ss, int64_t o_sh, float scale, int mask_mode, const float* key_mask, int kv_group) {
  CUtensorMap tmQ, tmK, tmV, tmdO, tmdQ;
  if (make_map_bshd(&tmQ, q, B, S, H, q_sb, q_ss, q_sh)) return -1;
  if (make_map_bshd(&tmK, k, B, Skv, H / kv_group, kv_sb, kv_ss, kv_sh)) return -1;
  if (make_map_bshd(&tmV, v, B, Skv, H / kv_group, kv_sb, kv_ss, kv_sh)) return -1;
  if (make_map_bshd(&tmdO, dout, B, S, H, o_sb, o_ss, o_sh)) return -1;
  cudaStream_t st = nfb_cur_stream();
  // dQ partials (one per key block) are summed by L2 reduce-adds: straight into the bf16 dq for short key ranges, into an
  // fp32 (B*H, S, 64) buffer converted afterwards when more than g_dq_f32_blocks partials meet in one element
  const bool dq_f32 = nfb_cdiv(Skv, TILE_K) > g_dq_f32_blocks;
  float* acc = nullptr;
  if (dq_f32) {
    const size_t acc_bytes = sizeof(float) * (size_t)B * H * S * 64;
    if (nfb_malloc(acc_bytes, (void**)&acc)) return -1;
    if (make_map_dq_f32(&tmdQ, acc, B * H, S)) { nfb_free(acc); return -1; }
    NFB_CUDA(cudaMemsetAsync(acc, 0, acc_bytes, st));
  } else {
  if (make_map_bshd(&tmdQ, dq, B, S, H, q_sb, q_ss, q_sh)) return -1;
  }
  // ONE preparation kernel in front of the backward: it clears dq (which accumulates through reduce-adds; any stride pattern) and
  // forms delta = rowsum(dO o O) per (b, h, query row) unless the caller passed it.  (Until round 2 the delta of each staged
  // query block was computed inside the backward kernel by its TMA-producer warp -- that single warp needed ~6.9 k clk per
  // block and PACED the whole loop: 49.4 -> 38.8 us at B8 H12 S512 causal with delta supplied, profiles/r02_attn_timeline.txt --
  // and dq was cleared by a 2-D memset node.)
  float* delta_buf = nullptr;
  if (!delta && nfb_malloc(sizeof(float) * (size_t)B * H * S, (void**)&delta_buf)) { if (acc) nfb_free(acc); return -1; }
  {
    const long long items = (long long)B * S * H * 8;
    if (items >= (1LL << 31)) { if (acc) nfb_free(acc); if (delta_buf) nfb_free(delta_buf); return NFB_FAIL("attention_tc backward: B*S*H = %lld rows exceed the preparation kernel's 32-bit index", items / 8); }
    nfb_launch_pdl(k_attn_bwd_prep, dim3((unsigned)nfb_grid_for(items, 256)), dim3(256), 0, st, (const __nv_bfloat16*)o, (const __nv_bfloat16*)dout,
                   delta_buf, dq_f32 ? (__nv_bfloat16*)nullptr : (__nv_bfloat16*)dq, B, H, S, (long long)o_sb, (long long)o_ss, (long long)o_sh,
                   (long long)q_sb, (long long)q_ss, (long long)q_sh);
    nfb_count_launch();
  }
  // grouped-query grids with few key/value heads (multi-query attention: one) would leave most SMs idle: split each
  // group over kv_split CTAs, each writing its partial dK / dV into a slice of a scratch buffer summed afterwards
  const int Hkv = H / kv_group;
  const long long kv_sb_out = kv_sb, kv_ss_out = kv_ss, kv_sh_out = kv_sh;
  int kv_split = 1;
  {
    const long long base = (long long)nfb_cdiv(Skv, TILE_K) * B * Hkv;
    for (int sdiv = 1; sdiv <= kv_group; ++sdiv) {
      if (kv_group % sdiv) continue;
      kv_split = sdiv;
      if (base * sdiv >= 148) break;
    }
  }
  __nv_bfloat16* part = nullptr;
  if (kv_split > 1) {
    if (nfb_malloc(sizeof(__nv_bfloat16) * 2 * (size_t)B * Skv * Hkv * kv_split * HDIM, (void**)&part)) { if (acc) nfb_free(acc); if (delta_buf) nfb_free(delta_buf); return -1; }
  }
  AttnBwdTcParams p = {};
  p.lse = lse; p.delta = delta ? delta : delta_buf; p.o = (const __nv_bfloat16*)o; p.o_sb = o_sb; p.o_ss = o_ss; p.o_sh = o_sh;
  p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  if (part) {   // partial dK | dV: two contiguous (B, Skv, Hkv * kv_split, 64) halves
    const long long hp = (long long)Hkv * kv_split;
    p.dk = part; p.dv = part + (size_t)B * Skv * hp * HDIM;
    kv_sb = (long long)Skv * hp * HDIM; kv_ss = hp * HDIM; kv_sh = HDIM;
  }
  p.kv_split = kv_split;
  p.kv_sb = kv_sb; p.kv_ss = kv_ss; p.kv_sh = kv_sh; p.B = B; p.H = H; p.S = S; p.Skv = Skv; p.scale = scale; p.mask_mode = mask_mode; p.key_mask = key_mask;
  p.timeline = g_attn_timeline;
  p.kv_group = kv_group / kv_split;
  p.dq_f32 = dq_f32 ? 1 : 0;
  static bool configured = false;
  if (!configured) {
    NFB_CUDA(cudaFuncSetAttribute(k_flash_bwd_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_TOTAL));
    NFB_CUDA(cudaFuncSetAttribute(k_flash_bwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_TOTAL));
    NFB_CUDA(cudaFuncSetAttribute(k_flash_bwd_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_TOTAL));
    configured = true;
  }
  dim3 grid((unsigned)(nfb_cdiv(Skv, TILE_K) * B * Hkv * kv_split));
  poly_defaults();
  // programmatic dependent launch behind k_attn_bwd_prep (small blocks, no shared memory: the backward CTAs co-reside and run their
  // prologue -- barrier init, TMEM allocation, descriptor prefetch -- under its tail; every global access sits behind griddepcontrol.wait)
  if (g_poly_bwd == 1) NFB_CUDA(nfb_launch_pdl(k_flash_bwd_tc<1>, grid, dim3(BWD_THREADS), (size_t)SB_TOTAL, st, tmQ, tmK, tmV, tmdO, tmdQ, p));
  else if (g_poly_bwd == 2) NFB_CUDA(nfb_launch_pdl(k_flash_bwd_tc<2>, grid, dim3(BWD_THREADS), (size_t)SB_TOTAL, st, tmQ, tmK, tmV, tmdO, tmdQ, p));
  else NFB_CUDA(nfb_launch_pdl(k_flash_bwd_tc<0>, grid, dim3(BWD_THREADS), (size_t)SB_TOTAL, st, tmQ, tmK, tmV, tmdO, tmdQ, p));
  if (part) {
    const long long vecs = (long long)B * Skv * Hkv * 8;
    const long long half_elems = (long long)B * Skv * Hkv * kv_split * HDIM;
    nfb_launch_pdl(k_kv_split_sum, dim3((unsigned)nfb_grid_for(vecs, 256)), dim3(256), 0, st, (const __nv_bfloat16*)part, (__nv_bfloat16*)dk, B, Skv, Hkv,
                   kv_split, kv_sb_out, kv_ss_out, kv_sh_out);
    nfb_count_launch();
    nfb_launch_pdl(k_kv_split_sum, dim3((unsigned)nfb_grid_for(vecs, 256)), dim3(256), 0, st, (const __nv_bfloat16*)(part + half_elems), (__nv_bfloat16*)dv, B,
                   Skv, Hkv, kv_split, kv_sb_out, kv_ss_out, kv_sh_out);
    nfb_count_launch();
    nfb_free(part);
  }
  if (delta_buf) nfb_free(delta_buf);
  if (dq_f32) {
    const long long vecs = (long long)B * H * S * 8;
    nfb_launch_pdl(k_dq_convert, dim3((unsigned)nfb_grid_for(vecs, 256)), dim3(256), 0, st, (const float*)acc, (__nv_bfloat16*)dq, B, H, S,
                   (long long)q_sb, (long long)q_ss, (long long)q_sh);
    nfb_count_launch();
    nfb_free(acc);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}
