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
//
// CG = 2 (cta_group::2): two CTAs of a cluster (an SM pair) share one 256 x BLOCK_N tile.  Each CTA
// stages its own 128 rows of A and HALF of the B tile; the leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2, which reads both halves over the pair interconnect and writes rows
// 0..127 of the accumulator into its own TMEM and rows 128..255 into the peer's.  Per CTA and
// k-block the smem fill drops from 16 KB + BLOCK_N*128 B to 16 KB + BLOCK_N*64 B, which is what
// lifts the kernel off the 64 B/clk/SM L2->smem port limit (a 128x256 single-CTA tile needs 768
// clk of fill for 512 clk of MMA; the paired tile needs 512).  Barrier protocol in pair mode:
// both producers' TMA loads complete_tx on the LEADER's full barrier; tcgen05.commit multicasts
// the "slot free" and "accumulator ready" arrivals to both CTAs; the peer's epilogue warps arrive
// on the leader's tmem_empty barrier through its shared::cluster address.
//
// Programmatic dependent launch: the kernel is launched with the programmatic-stream-serialization
// attribute and executes griddepcontrol.wait after its prologue (barrier init, TMEM alloc,
// tensor-map prefetch), so that prologue overlaps the tail of the preceding kernel in the stream.
#include "common.cuh"
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <tuple>

#define BLOCK_M 128
#define BLOCK_K 64
#define UMMA_K 16
#define GEMM_THREADS 320

#define GEMM_MAX_GROUP 4

// One contraction of a (possibly grouped) launch.  A grouped launch runs up to GEMM_MAX_GROUP independent problems with the
// same operand layouts / output type in ONE persistent kernel (the four weight-gradient GEMMs of a Transformer layer).
struct GemmProblem {
  int M, N, K;
  int num_m, num_n, k_blocks;   // num_m counts (128 * CG)-row tiles
  int tile_base;                // index of this problem's first tile in the launch-wide tile order
  long long unit_base;          // stream-K: index of this problem's first (tile, k-block) unit in the launch-wide order
  void* C;
  long long ldc;
};

struct GemmParams {
  int nprob;
  int mode;             // 0 = static persistent schedule (tile t -> cluster t % clusters, optional split-K over ks); 1 = stream-K
  GemmProblem pr[GEMM_MAX_GROUP];
  int a_mn, b_mn;       // operand major-ness (1 = MN-major)
  int split_k, k_blocks_per_split;   // static mode (single problem)
  int slice_rows;       // > 0: k-split ks stores (no add) its partial product at row offset ks * slice_rows of C (deterministic split-K)
  int n_fast;           // tile order: 0 = m fastest (concurrent tiles share a B panel), 1 = n fastest (share an A panel)
  // stream-K (hybrid): the LAST tiles of the launch-wide order -- the ones that do not fill a whole wave -- are pooled: their
  // (tile, k-block) units [sk_unit_base, total_units) are cut into equal contiguous ranges, one per cluster; a tile cut by a
  // range boundary is finished by the cluster that holds its LAST k-block (the owner), which adds the fp32 partial tiles the
  // earlier clusters left in `ws` (one tile-sized slot per cluster) in cluster order -- deterministic, no atomics on the
  // data.  flags[c] counts the epilogue warps that have delivered a partial for cluster c's owned tile.  The first dp_tiles
  // tiles are whole-tile work (tile t -> cluster t % clusters) done AFTER the pooled part, so the fix-up traffic of the
  // pooled part overlaps their main loops and the launch ends on a plain epilogue.
  long long total_units, sk_unit_base;
  int units_per_cluster;
  int dp_tiles;
  float4* ws;
  int* flags;
  int c_bf16;
  const float* bias;
  const float* residual;
  long long ldr;
  void* aux;            // optional bf16 pre-activation output (same shape as C)
  long long ldaux;
  int epilogue;         // 0 none, 1 gelu_tanh, 2 relu
  int accumulate;       // C += (fp32 only: TMA reduce-add / red.global.add of the finished tile)
  float clip;           // > 0: clamp the result to [-clip, clip]
  int vec_ok;           // C / residual / aux pitches and bases allow 128-bit (fp32) / 64-bit (bf16) accesses
  int tma_store;        // epilogue through shared memory + TMA store / reduce-add (C 16-byte aligned rows, no residual, no aux)
  // rotary position embedding fused into the bf16 TMA-store epilogue (the qkv projection of models/language/gpt2.py:78-95,
  // 228-234): output columns [0, rope_cols) are heads of 64 columns rotated by (x1, x2) -> (x1 c - x2 s, x2 c + x1 s) with
  // x2 = the column 32 further, position = row % rope_S, tables (rope_S, 32) fp32 -- before the rounding to bf16
  const float* rope_cos;
  const float* rope_sin;
  int rope_S, rope_cols;
  unsigned long long* timeline;  // debug: per-CTA clock64 stamps of the pipeline events (NULL in production)
  unsigned long long* span;      // optional: span[0] = min %globaltimer at CTA entry, span[1] = ~max at CTA exit (bench roofline)
};

struct GemmMaps { CUtensorMap a[GEMM_MAX_GROUP], b[GEMM_MAX_GROUP], c[GEMM_MAX_GROUP]; };

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

// ---- cluster (CTA pair) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address) in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pair-mode TMA load: data lands in THIS CTA's smem, the bytes are counted on the barrier at `bar_cluster_addr`
// (the leader CTA's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// TMA store / reduce-add of a shared-memory box into global memory (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]   (bf16 x bf16 -> fp32)
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrives once every previously issued tcgen05.mma has completed (implies fence::before_thread_sync);
// in pair mode the arrival is multicast to the barrier at the same offset in both CTAs
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
  }
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

// explicit shared-space 128-bit accesses: the staging pointer is derived from an aligned uintptr_t, so plain C++
// dereferences compile to GENERIC ld/st (LD.E/ST.E) instead of LDS/STS
__device__ __forceinline__ void sts128(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  float inner = 0.79788456080286535588f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(inner));
}

// debug timeline: slot layout per CTA (64 x u64): 0 entry, 1 setup done, 2 first TMA issued, 3 end;
// 8+4t first data landed (tile t), 9+4t last MMA issued, 10+4t accumulator ready, 11+4t epilogue done
__device__ __forceinline__ void tl_stamp(const GemmParams& p, int slot) {
  if (p.timeline && blockIdx.x < 8 && slot < 64) p.timeline[blockIdx.x * 64 + slot] = (unsigned long long)clock64();
}

template <int BLOCK_N, int STAGES, int CG>
struct GemmSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB: this CTA's 128 rows of A
  static constexpr int B_ROWS = BLOCK_N / CG;             // pair mode: each CTA stages half of the B tile
  // an MN-major B operand is staged in 64-column chunks of 8 KB (one TMA box each); a half-tile that is not a multiple of
  // 64 columns (the 96-column halves of a 192-wide pair tile) loads -- and reserves -- a whole last chunk, of which the MMA
  // reads the first 32 columns
  static constexpr int B_CHUNKS = (B_ROWS + 63) / 64;
  static constexpr int B_BYTES_KMAJOR = B_ROWS * BLOCK_K * 2;
  static constexpr int B_BYTES_MNMAJOR = B_CHUNKS * 8192;
  static constexpr int B_BYTES = B_BYTES_MNMAJOR > B_BYTES_KMAJOR ? B_BYTES_MNMAJOR : B_BYTES_KMAJOR;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 8 * 32 * 32 * 4;   // per-epilogue-warp 32x32 fp32 transpose buffers (XOR swizzled)
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;  // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget (227 KB per CTA)");
  static_assert(B_ROWS % 8 == 0, "B half-tile rows");
};

// ---------------------------------------------------------------- work schedule
// What a cluster does next: k-blocks [kb0, kb1) of tile (m_blk, n_blk) of problem g.
//   type 0: the whole K range of the tile (or, in static split-K mode, k-split `ks` of it) -> normal epilogue
//   type 1: stream-K contributor -- a partial sum that does not contain the tile's last k-block: dumped into this
//           cluster's workspace slot and announced on the owner's arrival counter
//   type 2: stream-K owner -- contains the tile's last k-block but not its first: adds the partials of clusters
//           first_contrib .. cluster-1 (in that order) to its accumulator, then runs the normal epilogue
struct Seg {
  int g, m_blk, n_blk, kb0, kb1, ks, type;
  int first_contrib;   // type 2: the lowest cluster that holds k-blocks of this tile
  int owner;           // type 1: the cluster that finishes this tile
};

// All three warp roles walk the same list.  Stream-K ranges are walked from the TOP down: a cluster first computes the
// partial it hands to its successor (type 1), then its whole tiles, and LAST the tile it finishes for its predecessor
// (type 2) -- whose partial was therefore written at the very start of the launch and is never waited for in practice.
struct SegIter {
  int tile, stride, total_tiles;   // whole-tile work: static mode, and the data-parallel part of the hybrid schedule
  long long lo, hi;                // stream-K: pooled units [lo, hi) still to do
  __device__ __forceinline__ SegIter(const GemmParams& p, int cluster_id, int num_clusters) {
    tile = cluster_id;
    stride = num_clusters;
    total_tiles = p.mode == 0 ? p.pr[0].num_m * p.pr[0].num_n * p.split_k : p.dp_tiles;
    lo = p.sk_unit_base + (long long)cluster_id * p.units_per_cluster;
    hi = lo + p.units_per_cluster;
    if (hi > p.total_units) hi = p.total_units;
    if (p.mode == 0) hi = lo;
  }
  __device__ __forceinline__ bool next(const GemmParams& p, Seg& s) {
    int g = 0, t;
    if (hi > lo) {
      // ---- pooled (stream-K) part, walked from the top down
      const long long u = hi - 1;
#pragma unroll
      for (int i = 1; i < GEMM_MAX_GROUP; ++i)
        if (i < p.nprob && u >= p.pr[i].unit_base) g = i;
      const int kbt = p.pr[g].k_blocks;
      const long long rel = u - p.pr[g].unit_base;
      t = (int)(rel / kbt);
      const long long tile_u0 = p.pr[g].unit_base + (long long)t * kbt;
      const long long seg_lo = tile_u0 > lo ? tile_u0 : lo;
      s.ks = 0;
      s.kb0 = (int)(seg_lo - tile_u0);
      s.kb1 = (int)(u - tile_u0) + 1;
      s.type = (s.kb1 != kbt) ? 1 : (s.kb0 == 0 ? 0 : 2);
      s.first_contrib = (int)((tile_u0 - p.sk_unit_base) / p.units_per_cluster);
      s.owner = (int)((tile_u0 + kbt - 1 - p.sk_unit_base) / p.units_per_cluster);
      hi = seg_lo;
    } else {
      if (tile >= total_tiles) return false;
      if (p.mode == 0) {
        const int tiles_mn = p.pr[0].num_m * p.pr[0].num_n;
        s.ks = tile / tiles_mn;
        t = tile - s.ks * tiles_mn;
        s.kb0 = s.ks * p.k_blocks_per_split;
        s.kb1 = min(s.kb0 + p.k_blocks_per_split, p.pr[0].k_blocks);
      } else {
#pragma unroll
        for (int i = 1; i < GEMM_MAX_GROUP; ++i)
          if (i < p.nprob && tile >= p.pr[i].tile_base) g = i;
        t = tile - p.pr[g].tile_base;
        s.ks = 0;
        s.kb0 = 0;
        s.kb1 = p.pr[g].k_blocks;
      }
      s.type = 0;
      s.first_contrib = s.owner = 0;
      tile += stride;
    }
    s.g = g;
    const int nm = p.pr[g].num_m, nn = p.pr[g].num_n;
    s.m_blk = p.n_fast ? t / nn : t % nm;
    s.n_blk = p.n_fast ? t % nn : t / nm;
    return true;
  }
};

__device__ __forceinline__ float4 ldcg_f4(const float4* ptr) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void stcg_f4(float4* ptr, float a, float b, float c, float d) {
  asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* ptr) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int BLOCK_N, int STAGES, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tc(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmParams p) {
  using L = GemmSmem<BLOCK_N, STAGES, CG>;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  constexpr int WS_SLOT_F4 = CG * (BLOCK_N / 4) * BLOCK_M;   // float4s of one cluster's partial-tile slot
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-B alignment; the dynamic-smem base offset is identical in both CTAs of a pair, so the
  // aligned layout (and with it every UMMA descriptor / barrier offset) is identical too
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  float* staging = (float*)(smem + STAGES * L::STAGE_BYTES);
  uint64_t* full_bar = (uint64_t*)((uint8_t*)staging + L::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
  const int cluster_id = (int)blockIdx.x / CG, num_clusters = (int)gridDim.x / CG;
  if (threadIdx.x == 0) {
    tl_stamp(p, 0);
    if (p.span) atomicMin(p.span, globaltimer_ns());   // kernel entry (prologue included), like ncu's gpu__time_duration
  }

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.nprob; ++g) {
      tma_prefetch_desc(&maps.a[g]);
      tma_prefetch_desc(&maps.b[g]);
      if (p.tma_store) tma_prefetch_desc(&maps.c[g]);
    }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8 * CG); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot, TMEM_COLS);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlapped the previous kernel's tail; its results are visible only after this point
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) tl_stamp(p, 1);

  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      SegIter it(p, cluster_id, num_clusters);
      Seg s;
      while (it.next(p, s)) {
        const CUtensorMap* tmA = &maps.a[s.g];
        const CUtensorMap* tmB = &maps.b[s.g];
        const int m0 = (s.m_blk * CG + (int)cta_rank) * BLOCK_M;
        const int n0 = s.n_blk * BLOCK_N + (int)cta_rank * L::B_ROWS;
        const uint32_t tx_bytes = (uint32_t)(L::A_BYTES + (p.b_mn ? L::B_BYTES_MNMAJOR : L::B_BYTES_KMAJOR));   // what the TMA boxes of one CTA deliver per stage
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          int k0 = kb * BLOCK_K;
          if (CG == 1) {
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            if (!p.a_mn) {
              tma_load_2d(sa, tmA, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * 8192, tmA, &full_bar[stage], m0 + j * 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d(sb, tmB, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < L::B_CHUNKS; ++j) tma_load_2d(sb + j * 8192, tmB, &full_bar[stage], n0 + j * 64, k0);
            }
          } else {
            // the leader's barrier counts the bytes of BOTH CTAs (the peer's complete_tx may land before the
            // leader's expect_tx: the phase cannot complete until the leader's own arrival)
            const uint32_t lead_bar = mapa_rank(smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], CG * tx_bytes);
            if (!p.a_mn) {
              tma_load_2d_pair(sa, tmA, lead_bar, k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d_pair(sa + j * 8192, tmA, lead_bar, m0 + j * 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d_pair(sb, tmB, lead_bar, k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < L::B_CHUNKS; ++j) tma_load_2d_pair(sb + j * 8192, tmB, lead_bar, n0 + j * 64, k0);
            }
          }
          if (first) { tl_stamp(p, 2); first = false; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only in pair mode) =====================
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = make_idesc(p.a_mn, p.b_mn, BLOCK_M * CG, BLOCK_N);
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;  // bytes per UMMA_K
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tcount = 0;
      SegIter it(p, cluster_id, num_clusters);
      Seg s;
      for (; it.next(p, s); ++tcount) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (kb == s.kb0) tl_stamp(p, 8 + 4 * tcount);
          tc_fence_after();
          uint32_t sa = smem_u32(stage_base + stage * L::STAGE_BYTES);
          uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            umma_bf16<CG>(d_tmem, da, db, idesc, (kb > s.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit<CG>(&empty_bar[stage]);  // frees the smem slot (in both CTAs) once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(&tmem_full[acc]);  // accumulator ready for the epilogue warps (of both CTAs)
        tl_stamp(p, 9 + 4 * tcount);
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
    const uint32_t st4 = smem_u32(staging) + (uint32_t)(warp - 2) * 4096u;  // this warp's 32x32 fp32 transpose buffer (float4 slots of 16 B)
    const int g = lane & 7, rsub = lane >> 3;
    // stream-K partial tiles: float4 (col4, row) of CTA `rank` of cluster c at ws[(c*CG + rank) * (BLOCK_N/4) * 128 + col4 * 128 + row]
    // -- consecutive lanes (rows) touch consecutive 16-byte slots: coalesced both ways
    const int ws_row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tcount = 0;
    SegIter it(p, cluster_id, num_clusters);
    Seg s;
    for (; it.next(p, s); ++tcount) {
      const GemmProblem& pb = p.pr[s.g];
      const int ks = s.ks, n_blk = s.n_blk;
      mbar_wait(&tmem_full[acc], acc_phase);
      if (warp == 2 && lane == 0) tl_stamp(p, 10 + 4 * tcount);
      tc_fence_after();
      const int row0 = (s.m_blk * CG + (int)cta_rank) * BLOCK_M + q * 32;
      const bool add_bias = p.bias != nullptr && ks == 0;
      const bool add_res = p.residual != nullptr && ks == 0;
      const int n_contrib = (s.type == 2) ? cluster_id - s.first_contrib : 0;
      if (s.type == 2) {
        // the partials of this tile were written at the start of their clusters' ranges; wait for all of them
        if (lane == 0) {
          const int expect = n_contrib * 8 * CG;
          uint32_t spins = 0;
          while (ld_acquire_gpu(p.flags + cluster_id) < expect) {
            __nanosleep(40);
            if (++spins > 20000000u) {
              printf("nfb gemm_tc: stream-K partial wait timed out (cluster %d)\n", cluster_id);
              __trap();
            }
          }
        }
        __syncwarp();
      }
      // adds the partial sums of the contributing clusters (fixed order) to 32 accumulator columns starting at tile column c
      auto add_partials = [&](float* v, int c) {
        for (int cc = s.first_contrib; cc < cluster_id; ++cc) {
          const float4* src = p.ws + (size_t)(cc * CG + (int)cta_rank) * (size_t)(WS_SLOT_F4 / CG) + (size_t)(c >> 2) * BLOCK_M + ws_row;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = ldcg_f4(src + (size_t)j4 * BLOCK_M);
            v[4 * j4] += t.x; v[4 * j4 + 1] += t.y; v[4 * j4 + 2] += t.z; v[4 * j4 + 3] += t.w;
          }
        }
      };
      if (s.type == 1) {
        // ---- stream-K contributor: raw fp32 partial -> this cluster's workspace slot, then one arrival per warp
#pragma unroll 1
        for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
          if (n_blk * BLOCK_N + c0 >= pb.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
          tmem_ld_wait();
          float4* dst = p.ws + (size_t)(cluster_id * CG + (int)cta_rank) * (size_t)(WS_SLOT_F4 / CG) + (size_t)(c0 >> 2) * BLOCK_M + ws_row;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            stcg_f4(dst + (size_t)j4 * BLOCK_M, __uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                    __uint_as_float(r[4 * j4 + 3]));
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          atomicAdd(p.flags + s.owner, 1);
        }
      } else if (p.tma_store) {
        // ---- TMA-store epilogue.  tcgen05.ld leaves thread = row holding consecutive columns; the row segment is
        // written straight into this warp's 32-row x 128-byte staging box in the SWIZZLE_128B pattern (16-byte slot c
        // of row r at c ^ (r & 7): conflict-free STS.128) and ONE elected lane hands the box to the TMA unit, which
        // writes (or reduce-adds, for wgrad) full lines to global memory and clips the ragged M / N edges.
        // No LDS, no per-thread global stores.  bf16: 64 columns per step, fp32: 32 columns per step.
        const CUtensorMap* tmC = &maps.c[s.g];
        const uint32_t sbox = smem_u32(staging) + (uint32_t)(warp - 2) * 4096u;
        const uint32_t srow = sbox + (uint32_t)lane * 128u;
        const int sw = lane & 7;
        const int step = p.c_bf16 ? 64 : 32;
#pragma unroll 1
        for (int c0 = half * step; c0 < BLOCK_N; c0 += 2 * step) {
          const int col0 = n_blk * BLOCK_N + c0;
          if (col0 >= pb.N) break;  // warp-uniform
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
          if (lane == 0) tma_store_wait_read();  // the previous box of this warp has been read out of shared memory
          __syncwarp();
          if (p.c_bf16 && col0 < p.rope_cols) {
            // ---- one head (64 columns) per step, both halves in registers: bias, rotation, clip, bf16
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(taddr, r0);
            tmem_ld_32x32b_x32(taddr + 32u, r1);
            tmem_ld_wait();
            float v0[32], v1[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { v0[j] = __uint_as_float(r0[j]); v1[j] = __uint_as_float(r1[j]); }
            if (n_contrib) { add_partials(v0, c0); add_partials(v1, c0 + 32); }
            if (add_bias) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j4 * 4));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 32 + j4 * 4));
                v0[4 * j4] += b0.x; v0[4 * j4 + 1] += b0.y; v0[4 * j4 + 2] += b0.z; v0[4 * j4 + 3] += b0.w;
                v1[4 * j4] += b1.x; v1[4 * j4 + 1] += b1.y; v1[4 * j4 + 2] += b1.z; v1[4 * j4 + 3] += b1.w;
              }
            }
            {
              const int pos = (row0 + lane) % p.rope_S;
              const float4* ct = reinterpret_cast<const float4*>(p.rope_cos + (size_t)pos * 32);
              const float4* st = reinterpret_cast<const float4*>(p.rope_sin + (size_t)pos * 32);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 c4 = __ldg(ct + j4), s4 = __ldg(st + j4);
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = v0[4 * j4 + e], b = v1[4 * j4 + e];
                  v0[4 * j4 + e] = a * cc[e] - b * ss[e];
                  v1[4 * j4 + e] = b * cc[e] + a * ss[e];
                }
              }
            }
            if (p.clip > 0.f) {
              const float lo = -p.clip, hi = p.clip;
#pragma unroll
              for (int j = 0; j < 32; ++j) { v0[j] = fminf(fmaxf(v0[j], lo), hi); v1[j] = fminf(fmaxf(v1[j], lo), hi); }
            }
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) {   // eight 16-byte slots (8 bf16 each): slots 0..3 = first half, 4..7 = second half
              const float* src = sl < 4 ? v0 + 8 * sl : v1 + 8 * (sl - 4);
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(src[2 * e], src[2 * e + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&b2);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)((sl ^ sw) * 16)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            }
          } else if (p.c_bf16) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(taddr + (uint32_t)(hh * 32), r);
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
              if (n_contrib) add_partials(v, c0 + hh * 32);
              if (add_bias) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const int c = col0 + hh * 32 + j4 * 4;
                  if (c + 3 < pb.N) {
                    float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
                    v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
                  } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (c + e < pb.N) v[4 * j4 + e] += __ldg(p.bias + c + e);
                  }
                }
              }
              if (p.epilogue == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
              } else if (p.epilogue == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if (p.clip > 0.f) {
                const float lo = -p.clip, hi = p.clip;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], lo), hi);
              }
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) {   // four 16-byte slots (8 bf16 each) per 32 columns
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(v[8 * sl + 2 * e], v[8 * sl + 2 * e + 1]);
                  w[e] = *reinterpret_cast<uint32_t*>(&b2);
                }
                const int slot = hh * 4 + sl;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)((slot ^ sw) * 16)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              }
            }
          } else {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (n_contrib) add_partials(v, c0);
            if (add_bias) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const int c = col0 + j4 * 4;
                if (c + 3 < pb.N) {
                  float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
                  v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) if (c + e < pb.N) v[4 * j4 + e] += __ldg(p.bias + c + e);
                }
              }
            }
            if (p.epilogue == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
            } else if (p.epilogue == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.clip > 0.f) {
              const float lo = -p.clip, hi = p.clip;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], lo), hi);
            }
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
              sts128(srow + (uint32_t)((sl ^ sw) * 16), v[4 * sl], v[4 * sl + 1], v[4 * sl + 2], v[4 * sl + 3]);
          }
          fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            if (p.accumulate) tma_reduce_add_2d(tmC, sbox, col0, row0);
            else tma_store_2d(tmC, sbox, col0, row0 + ks * p.slice_rows);
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
        const int col0 = n_blk * BLOCK_N + c0;
        if (col0 >= pb.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
        tmem_ld_wait();
        if (n_contrib) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          add_partials(v, c0);
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j]);
        }
#pragma unroll
        for (int gg = 0; gg < 8; ++gg)
          sts128(st4 + (uint32_t)(lane * 8 + (gg ^ (lane & 7))) * 16u, __uint_as_float(r[4 * gg]), __uint_as_float(r[4 * gg + 1]),
                 __uint_as_float(r[4 * gg + 2]), __uint_as_float(r[4 * gg + 3]));
        __syncwarp();
        const int col = col0 + g * 4;
        if (p.vec_ok && row0 + 32 <= pb.M && col0 + 32 <= pb.N) {
          // ---- fast path: the whole 32x32 chunk is in bounds and 16-byte addressable.  Every stage is its own
          // unrolled loop over the thread's 8 rows (conditions are kernel-uniform), so the 8 LDS / 8 residual LDG /
          // 8 STG of a stage are in flight together instead of one row at a time.
          float4 t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + rsub;
            t[i] = lds128(st4 + (uint32_t)(rr * 8 + (g ^ (rr & 7))) * 16u);
          }
          const long long rbase = (long long)(row0 + rsub);
          if (add_bias) {
            const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) { t[i].x += b4.x; t[i].y += b4.y; t[i].z += b4.z; t[i].w += b4.w; }
          }
          if (p.aux) {
            __nv_bfloat16* ap = (__nv_bfloat16*)p.aux + rbase * p.ldaux + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 a0 = __floats2bfloat162_rn(t[i].x, t[i].y), a1 = __floats2bfloat162_rn(t[i].z, t[i].w);
              *reinterpret_cast<uint2*>(ap + (long long)i * 4 * p.ldaux) = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
            }
          }
          if (p.epilogue == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { t[i].x = gelu_tanh_f(t[i].x); t[i].y = gelu_tanh_f(t[i].y); t[i].z = gelu_tanh_f(t[i].z); t[i].w = gelu_tanh_f(t[i].w); }
          } else if (p.epilogue == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { t[i].x = fmaxf(t[i].x, 0.f); t[i].y = fmaxf(t[i].y, 0.f); t[i].z = fmaxf(t[i].z, 0.f); t[i].w = fmaxf(t[i].w, 0.f); }
          }
          if (add_res) {
            const float* rp = p.residual + rbase * p.ldr + col;
            float4 rv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) rv[i] = *reinterpret_cast<const float4*>(rp + (long long)i * 4 * p.ldr);
#pragma unroll
            for (int i = 0; i < 8; ++i) { t[i].x += rv[i].x; t[i].y += rv[i].y; t[i].z += rv[i].z; t[i].w += rv[i].w; }
          }
          if (p.clip > 0.f) {
            const float lo = -p.clip, hi = p.clip;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              t[i].x = fminf(fmaxf(t[i].x, lo), hi); t[i].y = fminf(fmaxf(t[i].y, lo), hi);
              t[i].z = fminf(fmaxf(t[i].z, lo), hi); t[i].w = fminf(fmaxf(t[i].w, lo), hi);
            }
          }
          if (p.c_bf16) {
            __nv_bfloat16* cp = (__nv_bfloat16*)pb.C + rbase * pb.ldc + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 a0 = __floats2bfloat162_rn(t[i].x, t[i].y), a1 = __floats2bfloat162_rn(t[i].z, t[i].w);
              *reinterpret_cast<uint2*>(cp + (long long)i * 4 * pb.ldc) = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
            }
          } else if (p.accumulate) {
            float* cp = (float*)pb.C + rbase * pb.ldc + col;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + (long long)i * 4 * pb.ldc), "f"(t[i].x), "f"(t[i].y), "f"(t[i].z), "f"(t[i].w) : "memory");
          } else {
            float* cp = (float*)pb.C + rbase * pb.ldc + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(cp + (long long)i * 4 * pb.ldc) = t[i];
          }
          __syncwarp();
          continue;
        }
        // ---- generic path: ragged edges of M / N or unaligned pitches
        const bool full4 = (col + 3 < pb.N) && p.vec_ok;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (add_bias) {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (col + e < pb.N) bv[e] = p.bias[col + e];
        }
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rsub;
          const int row = row0 + rr;
          float4 t4 = lds128(st4 + (uint32_t)(rr * 8 + (g ^ (rr & 7))) * 16u);
          if (row < pb.M && col < pb.N) {
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
                *reinterpret_cast<uint2*>((__nv_bfloat16*)pb.C + (long long)row * pb.ldc + col) = u;
              } else {
                float* dst = (float*)pb.C + (long long)row * pb.ldc + col;
                if (p.accumulate) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                } else {
                  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (col + e >= pb.N) break;
                float x = v[e];
                if (p.aux) ((__nv_bfloat16*)p.aux)[(long long)row * p.ldaux + col + e] = __float2bfloat16_rn(x);
                if (p.epilogue == 1) x = gelu_tanh_f(x);
                else if (p.epilogue == 2) x = fmaxf(x, 0.f);
                if (add_res) x += p.residual[(long long)row * p.ldr + col + e];
                if (p.clip > 0.f) x = fminf(fmaxf(x, -p.clip), p.clip);
                if (p.c_bf16) ((__nv_bfloat16*)pb.C)[(long long)row * pb.ldc + col + e] = __float2bfloat16_rn(x);
                else {
                  float* dst = (float*)pb.C + (long long)row * pb.ldc + col + e;
                  if (p.accumulate) atomicAdd(dst, x);
                  else *dst = x;
                }
              }
            }
          }
        }
        __syncwarp();
      }
      }  // !tma_store
      if (s.type == 2) {
        // every warp of the cluster has consumed the partials: the last one re-arms the counters for the next launch
        __syncwarp();
        if (lane == 0) {
          const int old = atomicAdd(p.flags + 256 + cluster_id, 1);
          if (old == 8 * CG - 1) {
            p.flags[256 + cluster_id] = 0;
            p.flags[cluster_id] = 0;
            __threadfence();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || cta_rank == 0) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty[acc]), 0));  // the leader's MMA warp owns the accumulator hand-off
        if (warp == 2) tl_stamp(p, 11 + 4 * tcount);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  // the staging boxes must have been read out before the CTA's shared memory goes away (the stores complete with the grid)
  if (warp >= 2 && lane == 0 && p.tma_store) tma_store_wait_read();
  tc_fence_before();
  if (threadIdx.x == 0) tl_stamp(p, 3);
  if (CG == 2) cluster_sync_all(); else __syncthreads();  // pair mode: neither CTA may exit while the other still reads its smem / signals its barriers
  if (threadIdx.x == 0 && p.span) atomicMin(p.span + 1, ~globaltimer_ns());   // stored complemented: one 0xFF memset arms both words
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, TMEM_COLS);
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

typedef std::tuple<const void*, long long, long long, long long, int, int, int> MapKey;
static std::map<MapKey, CUtensorMap> g_map_cache;
static std::mutex g_map_mu;

// 2-D bf16 tensor map: dim0 (contiguous) x dim1, row pitch ld elements, box {box0, box1}, 128-B swizzle
static int make_map(CUtensorMap* out, const void* ptr, long long dim0, long long dim1, long long ld, int box0, int box1,
                    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  const int esz = (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32) ? 4 : 2;
  MapKey key(ptr, dim0, dim1, ld, box0, box1, (int)dt);
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *out = it->second; return 0; }
  }
  if (get_encode_fn()) return -1;
  if (((uintptr_t)ptr & 15) != 0) return NFB_FAIL("gemm_bf16: operand base %p is not 16-byte aligned", ptr);
  if ((ld * esz) % 16 != 0) return NFB_FAIL("gemm_bf16: leading dimension %lld is not a multiple of 16 bytes", ld);
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return NFB_FAIL("cuTensorMapEncodeTiled failed (%d) dims=(%lld,%lld) ld=%lld box=(%d,%d)", (int)r, dim0, dim1, ld, box0, box1);
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}


template <int BLOCK_N, int STAGES, int CG>
static int launch_gemm(const GemmMaps& maps, const GemmParams& p, int grid) {
  using L = GemmSmem<BLOCK_N, STAGES, CG>;
  static bool configured = false;
  if (!configured) {
    NFB_CUDA(cudaFuncSetAttribute(k_gemm_tc<BLOCK_N, STAGES, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = nfb_cur_stream();
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (CG == 2) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 2;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  if (nfb_pdl_enabled()) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  NFB_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_tc<BLOCK_N, STAGES, CG>, maps, p));
  NFB_LAUNCH_CHECK();
  return 0;
}

static int dispatch_gemm(int bn, int cg, const GemmMaps& maps, const GemmParams& p, int grid) {
  if (cg == 1) {
    switch (bn) {
      case 256: return launch_gemm<256, 4, 1>(maps, p, grid);
      case 192: return launch_gemm<192, 4, 1>(maps, p, grid);
      case 128: return launch_gemm<128, 6, 1>(maps, p, grid);
      case 64: return launch_gemm<64, 8, 1>(maps, p, grid);
    }
  } else {
    switch (bn) {
      case 256: return launch_gemm<256, 6, 2>(maps, p, grid);
      case 192: return launch_gemm<192, 6, 2>(maps, p, grid);
      case 128: return launch_gemm<128, 8, 2>(maps, p, grid);
      case 64: return launch_gemm<64, 8, 2>(maps, p, grid);
    }
  }
  return NFB_FAIL("gemm_bf16: unsupported tile N %d (cta group %d)", bn, cg);
}

static int g_force_block_n = 0, g_force_cg = 0, g_tma_store = 1, g_force_raster = 0;
NFB_API int nfb_gemm_bf16_set_tma_store(int on) { g_tma_store = on ? 1 : 0; return 0; }  // test hook: 0 = register epilogue everywhere
static unsigned long long* g_timeline = nullptr;
// debug: device buffer of 8 x 64 u64 that the next launches fill with clock64 stamps (NULL = off)
NFB_API int nfb_gemm_bf16_set_timeline(void* dev_buf) { g_timeline = (unsigned long long*)dev_buf; return 0; }
static unsigned long long* g_span = nullptr;
// measurement: two device u64 words (armed to all-ones) that the NEXT launch fills with min(%globaltimer at CTA entry) and
// ~max(%globaltimer at CTA exit); NULL = off.  The pointer is consumed by one launch.
NFB_API int nfb_gemm_bf16_set_span(void* dev_two_u64) { g_span = (unsigned long long*)dev_two_u64; return 0; }
static const float* g_rope_cos = nullptr;
static const float* g_rope_sin = nullptr;
static int g_rope_S = 0, g_rope_cols = 0;
// The NEXT nfb_gemm_bf16 launch rotates output columns [0, cols) (heads of 64 columns; position = row % S; tables (S, 32) fp32)
// in its epilogue.  Needs a bf16 C with 16-byte rows and no residual / aux (the TMA-store epilogue), cols % 64 == 0, no epilogue op.
NFB_API int nfb_gemm_bf16_set_rope(const float* cos_t, const float* sin_t, int S, int cols) {
  g_rope_cos = cos_t; g_rope_sin = sin_t; g_rope_S = S; g_rope_cols = cols;
  return 0;
}
NFB_API int nfb_gemm_bf16_force_tile(int block_n) { g_force_block_n = block_n; return 0; }
NFB_API int nfb_gemm_bf16_force_cta_group(int cg) { g_force_cg = cg; return 0; }
NFB_API int nfb_gemm_bf16_force_raster(int order) { g_force_raster = order; return 0; }   // test hook: 0 = auto, 1 = m fastest, 2 = n fastest
static int g_auto_split = 1;
// internal split-K of non-accumulating launches: 0 = off, 1 = automatic (cost model), 2 = whenever eligible (test hook)
NFB_API int nfb_gemm_bf16_set_auto_split(int mode) { g_auto_split = mode; return 0; }
static int g_stream_k = 1;
// stream-K schedule of single launches: 0 / 1 = classic schedule (grouped launches always use stream-K), 2 = whenever possible (test hook)
NFB_API int nfb_gemm_bf16_set_stream_k(int mode) { g_stream_k = mode; return 0; }
static int g_sm_limit = 0;
// number of SMs the persistent grids may occupy (0 = all): a data-parallel step leaves room for NCCL's channel CTAs while a
// gradient bucket is on the wire, so that a one-wave GEMM does not need a second round on the SMs NCCL holds
NFB_API int nfb_gemm_bf16_set_sm_limit(int sms) { g_sm_limit = sms < 0 ? 0 : sms; return 0; }
static int gemm_sms() {
  int sms = nfb_sm_count();
  if (g_sm_limit > 0 && g_sm_limit < sms) sms = g_sm_limit;
  return sms < 2 ? 2 : sms;
}

// Tile choice: a small cost model in SM clocks.  Per CTA and 64-deep k-block the MMA needs 2*BLOCK_N clk (tcgen05
// floor, 128 x BLOCK_N x 16 per BLOCK_N/2 clk) and the smem fill needs (16 KB + BLOCK_N/CG * 128 B) / 64 B/clk
// (the measured L2->smem port rate); a tile costs kblocks * max(of the two) and every launch pays a fixed
// prologue + first-load + last-epilogue cost.  Whole-launch cost = waves * tile + fixed.
struct TileChoice { int bn, cg, sk; double cost; };
static double kb_cost(int bn, int cg, int b_mn = 0) {
  const double b_bytes = b_mn ? (double)(((bn / cg) + 63) / 64) * 8192.0 : (double)(bn / cg) * 128.0;   // MN-major B: whole 64-column chunks
  double mma = 2.0 * bn, fill = (16384.0 + b_bytes) / 64.0;
  return mma > fill ? mma : fill;
}
static double epi_cost(int bn) { return 150.0 * (bn / 32) / 2.0; }   // drain of one accumulator (two warps per lane quarter)
static double tile_cost(int M, int N, int k_blocks, int bn, int cg, int sk, int sms, bool reduce_add = true, int b_mn = 0) {
  long long num_m = nfb_cdiv(M, BLOCK_M * cg), num_n = nfb_cdiv(N, bn);
  long long units = sms / cg;                     // CTAs (or CTA pairs) running concurrently
  long long work = num_m * num_n * sk;
  long long waves = nfb_cdiv(work, units);
  double kb = (double)nfb_cdiv(k_blocks, sk);
  // (a reduce-add penalty of ~10 B/clk/SM for split-K partials was tried here and picked slower tiles: measured 18.0 / 28.9 /
  // 12.4 / 28.2 us against 14.5 / 20.4 / 11.2 / 21.9 us without it on the four layer weight gradients)
  (void)reduce_add;
  return (double)waves * (kb * kb_cost(bn, cg, b_mn) + 250.0) + epi_cost(bn) + (sk > 1 ? 100.0 * (bn / 32) / 2.0 : 0.0) + 2500.0;
}
// Hybrid stream-K: how many of the `tiles` tiles (the last ones of the launch-wide order) are pooled and cut into equal
// k-block ranges over `clusters` clusters; the others run as whole tiles.  Whole waves stay whole tiles (no fix-up, and
// concurrently running tiles stay neighbours, which is what keeps their operand panels shared in L2); the remainder of a
// wave is pooled, together with one more full wave when it is so small that a tile would be cut more than ~6 ways.
static long long stream_k_pool(long long tiles, long long clusters) {
  if (tiles <= clusters) return tiles;
  const long long rem = tiles % clusters;
  if (rem == 0) return 0;
  if (rem * 6 >= clusters) return rem;
  return rem + clusters;
}
// cost of the hybrid schedule: whole-tile waves + the pooled share of every cluster; the pooled part pays one extra
// accumulator drain per cluster (partial out / partial in), exposed only when no whole tile follows it
static double stream_k_cost(long long tiles, int k_blocks, int bn, int cg, int sms, int b_mn = 0) {
  long long clusters = sms / cg;
  const long long pool = stream_k_pool(tiles, clusters);
  const long long dp_waves = (tiles - pool) / clusters;
  long long units = pool * k_blocks;
  long long active = clusters > units ? units : clusters;
  const long long per = units > 0 ? nfb_cdiv(units, active) : 0;
  const double cuts = pool > 0 ? (double)active / (double)pool : 0.0;   // clusters sharing one pooled tile
  const double fix = pool > 0 ? epi_cost(bn) + 900.0 * (cuts > 1.0 ? cuts : 1.0) : 0.0;
  const double exposed = dp_waves > 0 ? 0.3 * fix : fix;
  return (double)dp_waves * (k_blocks * kb_cost(bn, cg, b_mn) + 250.0) + (double)per * kb_cost(bn, cg, b_mn) + 500.0 + exposed + epi_cost(bn) + 2500.0;
}

// Best (tile N, CTA group, split-K) for a launch.  may_split: the caller accumulates into a pre-initialised fp32 C, so K may
// be cut (split_k >= 1 pins the factor, 0 lets the model choose).
static TileChoice choose_tiles(int M, int N, int k_blocks, int b_mn, bool may_split, int split_k, int sms) {
  TileChoice best = {128, 1, 1, 1e30};
  const int cand_bn[4] = {256, 192, 128, 64};
  for (int cg = 1; cg <= 2; ++cg) {
    if (g_force_cg && cg != g_force_cg) continue;
    if (cg == 2 && (sms % 2 != 0 || M <= BLOCK_M)) continue;
    for (int i = 0; i < 4; ++i) {
      int bn = cand_bn[i];
      if (g_force_block_n && bn != g_force_block_n) continue;
      if (cg == 2 && b_mn && (bn / 2) % 32 != 0) continue;   // MN-major B halves: whole 64-column chunks + an optional 32-column one
      int sk_lo = 1, sk_hi = 1;
      if (may_split) {
        if (split_k >= 1) sk_lo = sk_hi = split_k;
        else sk_hi = k_blocks >= 8 ? (k_blocks / 4 < 16 ? k_blocks / 4 : 16) : 1;
      }
      for (int sk = sk_lo; sk <= sk_hi; ++sk) {
        double c = tile_cost(M, N, k_blocks, bn, cg, sk, sms, true, b_mn);
        if (c < best.cost * 0.98) best = {bn, cg, sk, c};
      }
    }
  }
  return best;
}
// Best (tile N, CTA group) under the stream-K schedule for a group of problems sharing the operand layouts.
static TileChoice choose_tiles_stream_k(int nprob, const int* Ms, const int* Ns, const int* Ks, int b_mn, int sms) {
  TileChoice best = {128, 1, 1, 1e30};
  const int cand_bn[4] = {256, 192, 128, 64};
  int min_m = Ms[0];
  for (int g = 1; g < nprob; ++g) if (Ms[g] < min_m) min_m = Ms[g];
  for (int cg = 1; cg <= 2; ++cg) {
    if (g_force_cg && cg != g_force_cg) continue;
    if (cg == 2 && (sms % 2 != 0 || min_m <= BLOCK_M)) continue;
    for (int i = 0; i < 4; ++i) {
      int bn = cand_bn[i];
      if (g_force_block_n && bn != g_force_block_n) continue;
      if (cg == 2 && b_mn && (bn / 2) % 32 != 0) continue;
      // a group is costed as one problem with the summed units (all k-block counts enter through the unit total)
      double units = 0;
      long long tiles = 0;
      for (int g = 0; g < nprob; ++g) {
        long long t = nfb_cdiv(Ms[g], BLOCK_M * cg) * nfb_cdiv(Ns[g], bn);
        tiles += t;
        units += (double)t * (double)nfb_cdiv(Ks[g], BLOCK_K);
      }
      int kb_avg = (int)(units / (double)tiles + 0.5);
      if (kb_avg < 1) kb_avg = 1;
      double c = stream_k_cost(tiles, kb_avg, bn, cg, sms, b_mn);
      if (c < best.cost * 0.98) best = {bn, cg, 1, c};
    }
  }
  return best;
}

// Everything the host decides about a launch, in one place (also what nfb_gemm_bf16_plan reports to the CPU tests).
struct GemmPlan {
  int bn, cg, sk;        // tile N, CTA group, k-splits of the (possibly accumulating) launch
  int n_fast;            // tile order (see below)
  int internal_split;    // > 1: cut K into this many slices through the fp32 workspace + finishing pass instead
  int stream_k;          // 1: balanced (tile, k-block) ranges per cluster with the in-kernel fix-up
};
// may_split / split_k as in choose_tiles.  split_eligible: the launch does not accumulate, has no fused epilogue operands and
// an aligned C, so an under-filled grid may be traded for an internal split-K.  stream_k_ok: the stream-K workspace of the
// current stream is available.
static GemmPlan plan_gemm(int M, int N, int K, int b_mn, bool may_split, int split_k, bool split_eligible, bool stream_k_ok, int sms) {
  const int k_blocks = (int)nfb_cdiv(K, BLOCK_K);
  GemmPlan g = {128, 1, 1, 0, 0, 0};
  if (split_eligible && g_auto_split && k_blocks >= 16 && N % 4 == 0) {
    const TileChoice one = choose_tiles(M, N, k_blocks, b_mn, false, 0, sms);
    const TileChoice many = choose_tiles(M, N, k_blocks, b_mn, true, g_auto_split == 2 ? 3 : 0, sms);
    const int sk = (int)nfb_cdiv(k_blocks, nfb_cdiv(k_blocks, many.sk));   // the factor after rounding the k-blocks per split
    // the extra launch + the slices written once and read once + C written, at ~3400 B/clk of HBM
    // (the reduce-add penalty inside `many.cost` does not apply: the slices are stored, not added)
    const double many_cost = tile_cost(M, N, k_blocks, many.bn, many.cg, sk, sms, false, b_mn);
    const double overhead = 3000.0 + (double)M * (double)N * (8.0 * sk + 2.0) / 3400.0;
    if (sk > 1 && many.cost < 1e30 && one.cost < 1e30 && (g_auto_split == 2 || many_cost + overhead < 0.85 * one.cost)) {
      g.internal_split = sk;
    }
  }
  const bool slices = g.internal_split > 1;
  TileChoice best = slices ? choose_tiles(M, N, k_blocks, b_mn, true, g.internal_split, sms) : choose_tiles(M, N, k_blocks, b_mn, may_split, split_k, sms);
  if (best.cost >= 1e30) {   // a forced combination that is not available: fall back to the forced tile on one CTA
    best = {g_force_block_n ? g_force_block_n : 128, 1, slices ? g.internal_split : ((may_split && split_k >= 1) ? split_k : 1), 0.0};
  }
  if (stream_k_ok && g_stream_k && !(g_auto_split == 2 && slices)) {
    const TileChoice sk = choose_tiles_stream_k(1, &M, &N, &K, b_mn, sms);
    // measured (profiles/r02_gemm_streamk.txt): for ONE contraction of the step's sizes the fix-up (partial tiles out and back
    // through L2, latency-bound) costs what the balance wins -- classic 15.5 us vs 20.1 us on the 96-tile dgrad, 11.4 vs 16.9 us
    // on the 48-tile ones -- so single launches keep the classic schedule unless the test hook forces stream-K; the
    // schedule pays in GROUPED launches (nfb_gemm_bf16_grouped: 88 -> 46 us for the four weight gradients of a block)
    if (sk.cost < 1e30 && g_stream_k == 2) {
      g.stream_k = 1;
      g.internal_split = 0;
      best = sk;
    }
  }
  g.bn = best.bn; g.cg = best.cg; g.sk = g.stream_k ? 1 : best.sk;
  if (!g.stream_k && !(g.internal_split > 1)) g.sk = (int)nfb_cdiv(k_blocks, nfb_cdiv(k_blocks, g.sk));   // (slices: already the rounded factor, every split non-empty)
  // Tile order for L2 reuse.  The tiles of one wave run concurrently and walk K roughly in lockstep, so a panel they
  // share is fetched from HBM once; a panel they do not share is fetched again for every tile that needs it unless the
  // whole operand stays in L2.  m fastest: a wave shares B panels, A is read num_n times if it does not fit; n fastest:
  // the other way round.  Only the LM-head weight gradient (A = dlogits^T, 412 MB, num_n = 3) cares: 1.56 -> 0.74 GB.
  const double num_m = (double)nfb_cdiv(M, BLOCK_M * g.cg), num_n = (double)nfb_cdiv(N, g.bn);
  const double l2_keep = 48e6;   // what an operand may occupy and still survive in the 126 MB L2 next to the other traffic
  const double a_bytes = 2.0 * (double)M * K, b_bytes = 2.0 * (double)N * K;
  const double m_fast_cost = (a_bytes > l2_keep ? num_n * a_bytes : a_bytes) + b_bytes;
  const double n_fast_cost = a_bytes + (b_bytes > l2_keep ? num_m * b_bytes : b_bytes);
  g.n_fast = g_force_raster ? (g_force_raster == 2) : (n_fast_cost < m_fast_cost);
  return g;
}
// What nfb_gemm_bf16 would do with a launch of this shape on `sms` SMs (aligned operands, no bias / residual / aux):
// out = {tile N, CTA group, k-splits, n-fastest tile order, internal split-K slices, stream-K}.  Host logic only, no device needed.
NFB_API int nfb_gemm_bf16_plan(int transB, int M, int N, int K, int accumulate, int epilogue, int split_k, int sms, int* out) {
  if (M <= 0 || N <= 0 || K <= 0 || sms <= 0 || !out) return NFB_FAIL("gemm_bf16_plan: bad arguments");
  const bool eligible = !accumulate && split_k <= 1 && epilogue == 0 && g_tma_store;
  const GemmPlan g = plan_gemm(M, N, K, transB ? 0 : 1, accumulate && epilogue == 0, split_k, eligible, split_k <= 1, sms);
  out[0] = g.bn; out[1] = g.cg; out[2] = g.sk; out[3] = g.n_fast; out[4] = g.internal_split; out[5] = g.stream_k;
  return 0;
}

// finishing pass of an internal split-K: C = clip(slice 0 + slice 1 + ... in that fixed order), cast to C's type
// (4 elements per thread, 64-bit indexing; the slices are (slice_rows, N) fp32, stacked)
template <typename TC>
__global__ void __launch_bounds__(256)
k_splitk_finish(const float* __restrict__ ws, int slices, int64_t slice_elems, TC* __restrict__ C, int64_t M, int N, int64_t ldc, float clip) {
  nfb_pdl_enter();
  const int nv = N >> 2;
  const int64_t total = M * nv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nv;
    const int c = (int)(i - r * nv) << 2;
    const float* src = ws + r * N + c;
    f4 v = ld4<float>(src);
    for (int k = 1; k < slices; ++k) {
      const f4 u = ld4<float>(src + k * slice_elems);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (clip > 0.f) {
      v.x = fminf(fmaxf(v.x, -clip), clip); v.y = fminf(fmaxf(v.y, -clip), clip);
      v.z = fminf(fmaxf(v.z, -clip), clip); v.w = fminf(fmaxf(v.w, -clip), clip);
    }
    st4<TC>(C + r * ldc + c, v);
  }
}

// ---- stream-K fix-up workspace: one partial-tile slot (128 x 256 fp32) per CTA and 2 x 256 arrival counters, PER STREAM
// (the weight-gradient side stream runs its own stream-K launches next to the compute stream's)
struct StreamKWs { float4* ws; int* flags; };
static StreamKWs g_skws[2] = {{nullptr, nullptr}, {nullptr, nullptr}};
static int stream_k_ws(StreamKWs* out) {
  const int which = (nfb_aux_stream() && nfb_cur_stream() == nfb_aux_stream()) ? 1 : 0;
  StreamKWs& w = g_skws[which];
  if (!w.ws) {
    const size_t bytes = (size_t)nfb_sm_count() * BLOCK_M * 256 * sizeof(float);
    NFB_CUDA(cudaMalloc((void**)&w.ws, bytes));
    NFB_CUDA(cudaMalloc((void**)&w.flags, 512 * sizeof(int)));
    NFB_CUDA(cudaMemsetAsync(w.flags, 0, 512 * sizeof(int), nfb_cur_stream()));
  }
  *out = w;
  return 0;
}
// stream-K needs its workspace; a third concurrent stream would race on it, so only the compute stream and the registered
// auxiliary stream qualify
static bool stream_k_available() { return true; }

struct GemmOperands {
  int M, N, K;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
};

// Fills the tensor maps + problem table of `n` problems for tile (bn, cg) and launches.  Common to single and grouped launches.
static int launch_problems(int n, const GemmOperands* ops, int transA, int transB, int c_dtype, GemmParams& p, int bn, int cg,
                           int slice_split, int slice_rows, int sms) {
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.nprob = n;
  p.a_mn = transA ? 1 : 0;
  p.b_mn = transB ? 0 : 1;
  p.c_bf16 = (c_dtype == NFB_BF16);
  long long units = 0, tiles = 0;
  const int b_rows = bn / cg;
  for (int g = 0; g < n; ++g) {
    const GemmOperands& o = ops[g];
    GemmProblem& q = p.pr[g];
    q.M = o.M; q.N = o.N; q.K = o.K;
    q.num_m = (int)nfb_cdiv(o.M, BLOCK_M * cg);
    q.num_n = (int)nfb_cdiv(o.N, bn);
    q.k_blocks = (int)nfb_cdiv(o.K, BLOCK_K);
    q.unit_base = units;
    q.tile_base = (int)tiles;
    q.C = o.C; q.ldc = o.ldc;
    units += (long long)q.num_m * q.num_n * q.k_blocks;
    tiles += (long long)q.num_m * q.num_n;
    if (p.tma_store) {
      const long long c_rows = slice_split > 1 ? (long long)slice_split * slice_rows : o.M;
      if (make_map(&maps.c[g], o.C, o.N, c_rows, o.ldc, p.c_bf16 ? 64 : 32, 32, p.c_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return -1;
    }
    if (!p.a_mn) { if (make_map(&maps.a[g], o.A, o.K, o.M, o.lda, BLOCK_K, BLOCK_M)) return -1; }
    else { if (make_map(&maps.a[g], o.A, o.M, o.K, o.lda, 64, BLOCK_K)) return -1; }
    if (!p.b_mn) { if (make_map(&maps.b[g], o.B, o.K, o.N, o.ldb, BLOCK_K, b_rows)) return -1; }
    else { if (make_map(&maps.b[g], o.B, o.N, o.K, o.ldb, 64, BLOCK_K)) return -1; }
  }
  for (int g = n; g < GEMM_MAX_GROUP; ++g) { p.pr[g] = p.pr[0]; p.pr[g].unit_base = units; p.pr[g].tile_base = (int)tiles; }
  p.total_units = units;
  p.sk_unit_base = units;
  p.dp_tiles = 0;
  p.units_per_cluster = 1;
  long long clusters = sms / cg;
  if (clusters < 1) clusters = 1;
  if (p.mode == 1) {
    const long long sk_tiles = stream_k_pool(tiles, clusters);
    p.dp_tiles = (int)(tiles - sk_tiles);
    // unit index of the first pooled tile
    long long u0 = units;
    for (int g = 0; g < n; ++g) {
      const GemmProblem& q = p.pr[g];
      const long long tg = (long long)q.num_m * q.num_n;
      if (p.dp_tiles < q.tile_base + tg) { u0 = q.unit_base + (long long)(p.dp_tiles - q.tile_base) * q.k_blocks; break; }
    }
    p.sk_unit_base = u0;
    const long long sk_units = units - u0;
    long long active = clusters;
    if (sk_units > 0) {
      if (active > sk_units) active = sk_units;
      p.units_per_cluster = (int)nfb_cdiv(sk_units, active);
      active = nfb_cdiv(sk_units, p.units_per_cluster);   // no empty range in the middle
    } else {
      active = 0;
    }
    if (p.dp_tiles == 0) clusters = active;
    else if (clusters > p.dp_tiles && clusters > active) clusters = p.dp_tiles > active ? p.dp_tiles : active;
    if (clusters > 256) return NFB_FAIL("gemm_bf16: stream-K on %lld clusters (arrival counters hold 256)", clusters);
    StreamKWs w;
    if (stream_k_ws(&w)) return -1;
    p.ws = w.ws;
    p.flags = w.flags;
  } else {
    long long work = tiles * p.split_k;
    if (clusters > work) clusters = work;
    p.ws = nullptr;
    p.flags = nullptr;
  }
  p.timeline = g_timeline;
  p.span = g_span;
  g_span = nullptr;
  return dispatch_gemm(bn, cg, maps, p, (int)clusters * cg);
}

// C (+)= epi(opA . opB + bias) + residual.  A, B bf16; C fp32 (c_dtype = NFB_F32) or bf16.
// split_k: 0 = choose automatically (only used when accumulate != 0), >= 1 = explicit.
// slice_split > 1 (internal): C is a stack of slice_split (slice_rows, N) fp32 slices; k-split ks STORES its partial product
// into slice ks (no pre-zeroing, no atomics); the caller sums the slices in a fixed order.
static int gemm_bf16_impl(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
                          void* C, int c_dtype, int64_t ldc, const float* bias, const float* residual, int64_t ldr, void* aux,
                          int64_t ldaux, int epilogue, int accumulate, int split_k, float clip, int slice_split, int slice_rows) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0) return NFB_FAIL("gemm_bf16: K must be positive");
  if (c_dtype != NFB_F32 && c_dtype != NFB_BF16) return NFB_FAIL("gemm_bf16: C dtype %d", c_dtype);
  if (accumulate && c_dtype != NFB_F32) return NFB_FAIL("gemm_bf16: accumulate needs an fp32 C");
  if (!accumulate && split_k > 1) return NFB_FAIL("gemm_bf16: split_k > 1 requires accumulate (C pre-initialised)");
  if (split_k > 1 && epilogue != 0) return NFB_FAIL("gemm_bf16: split-K cannot be combined with an activation epilogue");
  const int sms = gemm_sms();
  const int b_mn = transB ? 0 : 1;
  const int k_blocks = (int)nfb_cdiv(K, BLOCK_K);

  // ---- internal split-K: a launch whose tiles fill well under one wave and whose K is long is cut along K; every k-split
  // stores its partial product into its own fp32 slice of a workspace and one finishing pass sums the slices in a fixed
  // order, clips and casts -- deterministic, unlike a reduce-add.  Two launches instead of one; superseded by the stream-K
  // schedule wherever that is available (kept as the fallback and as a second implementation for the tests).
  const bool eligible = !slice_split && !accumulate && split_k <= 1 && epilogue == 0 && !bias && !residual && !aux && g_tma_store &&
                        ldc % 4 == 0 && (uintptr_t)C % 16 == 0 && !(nfb_aux_stream() && nfb_cur_stream() == nfb_aux_stream());
  const bool stream_k_ok = !slice_split && split_k <= 1 && stream_k_available();
  const GemmPlan plan = slice_split > 1 ? plan_gemm(M, N, K, b_mn, true, slice_split, false, false, sms)
                                        : plan_gemm(M, N, K, b_mn, accumulate && epilogue == 0, split_k, eligible, stream_k_ok, sms);
  if (plan.internal_split > 1) {
    const int sk = plan.internal_split;
    const int64_t rows = nfb_cdiv(M, 2 * BLOCK_M) * 2 * BLOCK_M;           // whole tiles: an edge tile never reaches the next slice
    float* ws = nullptr;
    if (nfb_malloc(sizeof(float) * (size_t)sk * rows * N, (void**)&ws)) return -1;
    int rc = gemm_bf16_impl(transA, transB, M, N, K, A, lda, B, ldb, ws, NFB_F32, N, nullptr, nullptr, 0, nullptr, 0, 0, 0, 0, 0.f, sk, (int)rows);
    if (rc == 0) {
      const int grid = nfb_grid_for((int64_t)M * (N / 4), 256, 8);
      if (c_dtype == NFB_BF16) nfb_launch_pdl(k_splitk_finish<__nv_bfloat16>, dim3(grid), dim3(256), 0, nfb_cur_stream(), (const float*)ws, sk, rows * (int64_t)N, (__nv_bfloat16*)C, (int64_t)M, N, (int64_t)ldc, clip);
      else nfb_launch_pdl(k_splitk_finish<float>, dim3(grid), dim3(256), 0, nfb_cur_stream(), (const float*)ws, sk, rows * (int64_t)N, (float*)C, (int64_t)M, N, (int64_t)ldc, clip);
    }
    nfb_free(ws);
    if (rc) return rc;
    NFB_LAUNCH_CHECK();
    return 0;
  }

  const int best_bn = plan.bn, cg = plan.cg;
  int sk = slice_split > 1 ? slice_split : plan.sk;
  const int kbps = (int)nfb_cdiv(k_blocks, sk);

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.mode = plan.stream_k ? 1 : 0;
  p.split_k = sk; p.k_blocks_per_split = kbps;
  p.slice_rows = slice_split > 1 ? slice_rows : 0;
  if (slice_split > 1 && (sk != slice_split || (long long)(sk - 1) * kbps >= k_blocks)) return NFB_FAIL("gemm_bf16: internal split-K factor %d does not fit %d k-blocks", slice_split, k_blocks);
  p.n_fast = plan.n_fast;
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.aux = aux; p.ldaux = ldaux;
  p.epilogue = epilogue; p.accumulate = accumulate; p.clip = clip;
  {
    bool ok = (ldc % 4 == 0) && ((uintptr_t)C % 16 == 0);
    if (residual) ok = ok && (ldr % 4 == 0) && ((uintptr_t)residual % 16 == 0);
    if (aux) ok = ok && (ldaux % 4 == 0) && ((uintptr_t)aux % 8 == 0);
    p.vec_ok = ok ? 1 : 0;
  }
  // TMA-store epilogue: C rows must be 16-byte addressable; residual / aux outputs keep the register epilogue
  p.tma_store = (g_tma_store && !residual && !aux && ((uintptr_t)C % 16 == 0) && ((ldc * (c_dtype == NFB_BF16 ? 2 : 4)) % 16 == 0)) ? 1 : 0;
  if (!p.tma_store && slice_split > 1) return NFB_FAIL("gemm_bf16: the internal split-K needs the TMA-store epilogue");
  if (g_rope_cols > 0 && slice_split <= 1) {
    const int cols = g_rope_cols;
    p.rope_cos = g_rope_cos; p.rope_sin = g_rope_sin; p.rope_S = g_rope_S; p.rope_cols = cols;
    g_rope_cols = 0;   // consumed
    if (!p.tma_store || c_dtype != NFB_BF16 || epilogue != 0 || cols % 64 != 0 || cols > N || best_bn % 64 != 0 || p.rope_S <= 0 || !p.rope_cos || !p.rope_sin ||
        (bias && ((uintptr_t)bias % 16 != 0)))
      return NFB_FAIL("gemm_bf16: the fused rotary epilogue needs a bf16 C with 16-byte rows, no residual / aux / activation, cols %% 64 == 0 (got cols=%d N=%d tile %d)", cols, N, best_bn);
  }
  GemmOperands op = {M, N, K, A, lda, B, ldb, C, ldc};
  return launch_problems(1, &op, transA, transB, c_dtype, p, best_bn, cg, slice_split, slice_rows, sms);
}

NFB_API int nfb_gemm_bf16(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
                          void* C, int c_dtype, int64_t ldc, const float* bias, const float* residual, int64_t ldr, void* aux,
                          int64_t ldaux, int epilogue, int accumulate, int split_k, float clip) {
  return gemm_bf16_impl(transA, transB, M, N, K, A, lda, B, ldb, C, c_dtype, ldc, bias, residual, ldr, aux, ldaux, epilogue, accumulate,
                        split_k, clip, 0, 0);
}

// Grouped launch: n <= 4 independent contractions C_g (+)= opA_g . opB_g with the same operand layouts and output type in
// ONE persistent stream-K kernel (no epilogue operands).  This is how the four weight-gradient GEMMs of a Transformer layer
// run: 90 tiles x 64 k-blocks are cut into 74 equal ranges instead of four launches that each split K eight ways and
// reduce-add 8x their output through L2.  Each output element receives exactly one store / one reduce-add: deterministic.
NFB_API int nfb_gemm_bf16_grouped(int n, int transA, int transB, const int* Ms, const int* Ns, const int* Ks, const void* const* As,
                                  const int64_t* ldas, const void* const* Bs, const int64_t* ldbs, void* const* Cs, const int64_t* ldcs,
                                  int c_dtype, int accumulate) {
  if (n <= 0) return 0;
  if (n > GEMM_MAX_GROUP) return NFB_FAIL("gemm_bf16_grouped: at most %d problems per launch (got %d)", GEMM_MAX_GROUP, n);
  if (c_dtype != NFB_F32 && c_dtype != NFB_BF16) return NFB_FAIL("gemm_bf16_grouped: C dtype %d", c_dtype);
  if (accumulate && c_dtype != NFB_F32) return NFB_FAIL("gemm_bf16_grouped: accumulate needs an fp32 C");
  GemmOperands ops[GEMM_MAX_GROUP];
  int m = 0;
  bool tma_ok = g_tma_store != 0, vec_ok = true;
  for (int g = 0; g < n; ++g) {
    if (Ms[g] <= 0 || Ns[g] <= 0) continue;
    if (Ks[g] <= 0) return NFB_FAIL("gemm_bf16_grouped: K must be positive");
    ops[m] = {Ms[g], Ns[g], Ks[g], As[g], ldas[g], Bs[g], ldbs[g], Cs[g], ldcs[g]};
    tma_ok = tma_ok && ((uintptr_t)Cs[g] % 16 == 0) && ((ldcs[g] * (c_dtype == NFB_BF16 ? 2 : 4)) % 16 == 0);
    vec_ok = vec_ok && (ldcs[g] % 4 == 0) && ((uintptr_t)Cs[g] % 16 == 0);
    ++m;
  }
  if (m == 0) return 0;
  const int sms = gemm_sms();
  int Mv[GEMM_MAX_GROUP], Nv[GEMM_MAX_GROUP], Kv[GEMM_MAX_GROUP];
  for (int g = 0; g < m; ++g) { Mv[g] = ops[g].M; Nv[g] = ops[g].N; Kv[g] = ops[g].K; }
  TileChoice t = choose_tiles_stream_k(m, Mv, Nv, Kv, transB ? 0 : 1, sms);
  if (t.cost >= 1e30) t = {g_force_block_n ? g_force_block_n : 128, 1, 1, 0.0};
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.mode = 1;
  p.split_k = 1;
  p.k_blocks_per_split = 0;
  p.accumulate = accumulate;
  p.vec_ok = vec_ok ? 1 : 0;
  p.tma_store = tma_ok ? 1 : 0;
  return launch_problems(m, ops, transA, transB, c_dtype, p, t.bn, t.cg, 0, 0, sms);
}
