// bf16 tensor-core GEMM for sm_100a: tcgen05.mma (UMMA) with the accumulator in TMEM, operands
// staged in shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B), a persistent warp-specialised
// CTA per SM, and a fused epilogue (bias, tanh-GELU/ReLU, residual add, clip, bf16/fp32 store,
// split-K accumulation).  This is the throughput path for every Linear / logits contraction of
// the Transformer step -- forward x.W^T, dgrad dY.W and wgrad dY^T.X (the
// reference does these with np.matmul, functional/arithmetic.py:441-575, nn/optimized.py:231-317).
//
//   C[M,N] (+)= epi( sum_k opA[m,k] * opB[k,n] + bias[n] ) + residual[m,n]
//   transA = 0: A stored (M,K) row-major  -> "K-major"  UMMA operand
//   transA = 1: A stored (K,M) row-major  -> "MN-major" UMMA operand (no transpose materialised)
//   transB = 1: B stored (N,K) row-major  -> K-major    (y = x W^T with W (out,in))
//   transB = 0: B stored (K,N) row-major  -> MN-major   (dX = dY W ; dW = dY^T X)
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected lane), warps 2..9 = epilogue (two warps per 32-lane TMEM quarter 32*(warp%4) .. +31).
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), double-buffered TMEM accumulator
// full/empty mbarriers (MMA <-> epilogue), static persistent tile schedule.
#include "common.cuh"
#include <cuda.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <tuple>

#define BLOCK_M 128
#define BLOCK_K 64
#define UMMA_K 16
#define GEMM_THREADS 320

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;       // operand major-ness (1 = MN-major)
  int num_m, num_n, split_k, k_blocks_per_split, k_blocks_total;
  void* C;
  int c_bf16;
  long long ldc;
  const float* bias;
  const float* residual;
  long long ldr;
  void* aux;            // optional bf16 pre-activation output (same shape as C)
  long long ldaux;
  int epilogue;         // 0 none, 1 gelu_tanh, 2 relu
  int accumulate;       // C += (fp32 only, red.global.add)
  float clip;           // > 0: clamp the result to [-clip, clip]
  int vec_ok;           // C / residual / aux pitches and bases allow 128-bit (fp32) / 64-bit (bf16) accesses
};

// ---------------------------------------------------------------- PTX helpers
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
    if (++spins > 4000000u) {
      printf("nfb gemm_tc: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]   (bf16 x bf16 -> fp32)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor), SWIZZLE_128B:
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor (InstrDescriptor): fp32 accumulate, bf16 x bf16, dense
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn, int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 1u << 7;                       // a_format = BF16
  d |= 1u << 10;                      // b_format = BF16
  d |= (uint32_t)(a_mn & 1) << 15;    // a_major: 0 = K, 1 = MN
  d |= (uint32_t)(b_mn & 1) << 16;    // b_major
  d |= (uint32_t)(n >> 3) << 17;      // n_dim
  d |= (uint32_t)(m >> 4) << 24;      // m_dim
  return d;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  float inner = 0.79788456080286535588f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(inner));
}

template <int BLOCK_N, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 8 * 32 * 32 * 4;   // per-epilogue-warp 32x32 fp32 transpose buffers (XOR swizzled)
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B needs 1024-B alignment
  uint8_t* stage_base = smem;
  float* staging = (float*)(smem + STAGES * L::STAGE_BYTES);
  uint64_t* full_bar = (uint64_t*)((uint8_t*)staging + L::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mn = p.num_m * p.num_n;
  const int total_tiles = tiles_mn * p.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int ks = tile / tiles_mn, mn = tile % tiles_mn;
        int m_blk = mn % p.num_m, n_blk = mn / p.num_m;
        int kb0 = ks * p.k_blocks_per_split;
        int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m_blk * BLOCK_M);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m_blk * BLOCK_M + j * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmB, &full_bar[stage], k0, n_blk * BLOCK_N);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n_blk * BLOCK_N + j * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.a_mn, p.b_mn, BLOCK_M, BLOCK_N);
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;  // bytes per UMMA_K
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int ks = tile / tiles_mn;
        int kb0 = ks * p.k_blocks_per_split;
        int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint32_t sa = smem_u32(stage_base + stage * L::STAGE_BYTES);
          uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31); the pair splits the
    // 32-column chunks even/odd.  Per chunk: tcgen05.ld (thread = row, 32 columns) -> XOR-swizzled smem transpose
    // (conflict-free 128-bit STS/LDS) -> each thread owns 4 consecutive columns of 8 rows -> fused bias /
    // activation / residual / clip -> 128-bit (fp32) or 64-bit (bf16) coalesced global stores.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float4* st4 = reinterpret_cast<float4*>(staging) + (warp - 2) * 256;
    const int g = lane & 7, rsub = lane >> 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int ks = tile / tiles_mn, mn = tile % tiles_mn;
      int m_blk = mn % p.num_m, n_blk = mn / p.num_m;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_blk * BLOCK_M + q * 32;
      const bool add_bias = p.bias != nullptr && ks == 0;
      const bool add_res = p.residual != nullptr && ks == 0;
#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
        const int col0 = n_blk * BLOCK_N + c0;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
        tmem_ld_wait();
#pragma unroll
        for (int gg = 0; gg < 8; ++gg)
          st4[lane * 8 + (gg ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * gg]), __uint_as_float(r[4 * gg + 1]),
                                                          __uint_as_float(r[4 * gg + 2]), __uint_as_float(r[4 * gg + 3]));
        __syncwarp();
        const int col = col0 + g * 4;
        const bool full4 = (col + 3 < p.N) && p.vec_ok;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (add_bias) {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (col + e < p.N) bv[e] = p.bias[col + e];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rsub;
          const int row = row0 + rr;
          float4 t4 = st4[rr * 8 + (g ^ (rr & 7))];
          if (row < p.M && col < p.N) {
            float v[4] = {t4.x + bv[0], t4.y + bv[1], t4.z + bv[2], t4.w + bv[3]};
            if (full4) {
              if (p.aux) {
                __nv_bfloat162 a0 = __floats2bfloat162_rn(v[0], v[1]), a1 = __floats2bfloat162_rn(v[2], v[3]);
                uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
                *reinterpret_cast<uint2*>((__nv_bfloat16*)p.aux + (long long)row * p.ldaux + col) = u;
              }
              if (p.epilogue == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = gelu_tanh_f(v[e]);
              } else if (p.epilogue == 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
              }
              if (add_res) {
                float4 rv = *reinterpret_cast<const float4*>(p.residual + (long long)row * p.ldr + col);
                v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
              }
              if (p.clip > 0.f) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fminf(fmaxf(v[e], -p.clip), p.clip);
              }
              if (p.c_bf16) {
                __nv_bfloat162 a0 = __floats2bfloat162_rn(v[0], v[1]), a1 = __floats2bfloat162_rn(v[2], v[3]);
                uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
                *reinterpret_cast<uint2*>((__nv_bfloat16*)p.C + (long long)row * p.ldc + col) = u;
              } else {
                float* dst = (float*)p.C + (long long)row * p.ldc + col;
                if (p.accumulate) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                } else {
                  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (col + e >= p.N) break;
                float x = v[e];
                if (p.aux) ((__nv_bfloat16*)p.aux)[(long long)row * p.ldaux + col + e] = __float2bfloat16_rn(x);
                if (p.epilogue == 1) x = gelu_tanh_f(x);
                else if (p.epilogue == 2) x = fmaxf(x, 0.f);
                if (add_res) x += p.residual[(long long)row * p.ldr + col + e];
                if (p.clip > 0.f) x = fminf(fmaxf(x, -p.clip), p.clip);
                if (p.c_bf16) ((__nv_bfloat16*)p.C)[(long long)row * p.ldc + col + e] = __float2bfloat16_rn(x);
                else {
                  float* dst = (float*)p.C + (long long)row * p.ldc + col + e;
                  if (p.accumulate) atomicAdd(dst, x);
                  else *dst = x;
                }
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encode_fn() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return NFB_FAIL("cuTensorMapEncodeTiled entry point unavailable");
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

typedef std::tuple<const void*, long long, long long, long long, int, int> MapKey;
static std::map<MapKey, CUtensorMap> g_map_cache;
static std::mutex g_map_mu;

// 2-D bf16 tensor map: dim0 (contiguous) x dim1, row pitch ld elements, box {box0, box1}, 128-B swizzle
static int make_map(CUtensorMap* out, const void* ptr, long long dim0, long long dim1, long long ld, int box0, int box1) {
  MapKey key(ptr, dim0, dim1, ld, box0, box1);
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *out = it->second; return 0; }
  }
  if (get_encode_fn()) return -1;
  if (((uintptr_t)ptr & 15) != 0) return NFB_FAIL("gemm_bf16: operand base %p is not 16-byte aligned", ptr);
  if ((ld * 2) % 16 != 0) return NFB_FAIL("gemm_bf16: operand leading dimension %lld must be a multiple of 8 elements", ld);
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return NFB_FAIL("cuTensorMapEncodeTiled failed (%d) dims=(%lld,%lld) ld=%lld box=(%d,%d)", (int)r, dim0, dim1, ld, box0, box1);
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

template <int BLOCK_N, int STAGES>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  static bool configured = false;
  if (!configured) {
    NFB_CUDA(cudaFuncSetAttribute(k_gemm_tc<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  k_gemm_tc<BLOCK_N, STAGES><<<grid, GEMM_THREADS, L::TOTAL, nfb_cur_stream()>>>(tmA, tmB, p);
  NFB_LAUNCH_CHECK();
  return 0;
}

static int g_force_block_n = 0;
NFB_API int nfb_gemm_bf16_force_tile(int block_n) { g_force_block_n = block_n; return 0; }

// C (+)= epi(opA . opB + bias) + residual.  A, B bf16; C fp32 (c_dtype = NFB_F32) or bf16.
// split_k: 0 = choose automatically (only used when accumulate != 0), >= 1 = explicit.
NFB_API int nfb_gemm_bf16(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
                          void* C, int c_dtype, int64_t ldc, const float* bias, const float* residual, int64_t ldr, void* aux,
                          int64_t ldaux, int epilogue, int accumulate, int split_k, float clip) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0) return NFB_FAIL("gemm_bf16: K must be positive");
  if (c_dtype != NFB_F32 && c_dtype != NFB_BF16) return NFB_FAIL("gemm_bf16: C dtype %d", c_dtype);
  if (accumulate && c_dtype != NFB_F32) return NFB_FAIL("gemm_bf16: accumulate needs an fp32 C");
  const int sms = nfb_sm_count();
  const int num_m = (int)nfb_cdiv(M, BLOCK_M);
  // tile-N choice: best wave efficiency over {256, 192, 128, 64}, larger tile wins ties
  int cand[4] = {256, 192, 128, 64};
  int best_bn = 128;
  double best_eff = -1.0;
  for (int i = 0; i < 4; ++i) {
    int bn = cand[i];
    long long tiles = (long long)num_m * nfb_cdiv(N, bn);
    long long waves = nfb_cdiv(tiles, sms);
    double useful = (double)M * N / ((double)num_m * BLOCK_M * nfb_cdiv(N, bn) * bn);  // padding waste
    double eff = (double)tiles / (double)(waves * sms) * useful * (bn >= 128 ? 1.0 : 0.8);
    if (eff > best_eff + 0.03) { best_eff = eff; best_bn = bn; }
  }
  if (g_force_block_n) best_bn = g_force_block_n;
  const int num_n = (int)nfb_cdiv(N, best_bn);
  const int k_blocks = (int)nfb_cdiv(K, BLOCK_K);
  int sk = 1;
  if (accumulate) {
    if (split_k >= 1) sk = split_k;
    else {
      long long tiles = (long long)num_m * num_n;
      if (tiles < sms && k_blocks >= 8) {
        sk = (int)(sms / tiles);
        if (sk > k_blocks / 4) sk = k_blocks / 4;
        if (sk < 1) sk = 1;
      }
    }
  } else if (split_k > 1) {
    return NFB_FAIL("gemm_bf16: split_k > 1 requires accumulate (C pre-initialised)");
  }
  if (sk > 1 && epilogue != 0) return NFB_FAIL("gemm_bf16: split-K cannot be combined with an activation epilogue");
  int kbps = (int)nfb_cdiv(k_blocks, sk);
  sk = (int)nfb_cdiv(k_blocks, kbps);

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = transA ? 1 : 0;
  p.b_mn = transB ? 0 : 1;
  p.num_m = num_m; p.num_n = num_n; p.split_k = sk; p.k_blocks_per_split = kbps; p.k_blocks_total = k_blocks;
  p.C = C; p.c_bf16 = (c_dtype == NFB_BF16); p.ldc = ldc;
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.aux = aux; p.ldaux = ldaux;
  p.epilogue = epilogue; p.accumulate = accumulate; p.clip = clip;
  {
    bool ok = (ldc % 4 == 0) && ((uintptr_t)C % 16 == 0);
    if (residual) ok = ok && (ldr % 4 == 0) && ((uintptr_t)residual % 16 == 0);
    if (aux) ok = ok && (ldaux % 4 == 0) && ((uintptr_t)aux % 8 == 0);
    p.vec_ok = ok ? 1 : 0;
  }

  CUtensorMap tmA, tmB;
  if (!p.a_mn) { if (make_map(&tmA, A, K, M, lda, BLOCK_K, BLOCK_M)) return -1; }
  else { if (make_map(&tmA, A, M, K, lda, 64, BLOCK_K)) return -1; }
  if (!p.b_mn) { if (make_map(&tmB, B, K, N, ldb, BLOCK_K, best_bn)) return -1; }
  else { if (make_map(&tmB, B, N, K, ldb, 64, BLOCK_K)) return -1; }

  long long total = (long long)num_m * num_n * sk;
  int grid = (int)(total < sms ? total : sms);
  switch (best_bn) {
    case 256: return launch_gemm<256, 4>(tmA, tmB, p, grid);
    case 192: return launch_gemm<192, 4>(tmA, tmB, p, grid);
    case 128: return launch_gemm<128, 6>(tmA, tmB, p, grid);
    case 64: return launch_gemm<64, 8>(tmA, tmB, p, grid);
  }
  return NFB_FAIL("gemm_bf16: unsupported tile N %d", best_bn);
}
