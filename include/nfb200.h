/* nfb200.h -- C-ABI of libnfb200.so: the B200 (sm_100a) compute library behind the Neural Forge
 * `neural_arch.backends.Backend` plugin (reference: src/neural_arch/backends/backend.py:9-375).
 *
 * The reference has NO native code: its plugin boundary is a Python class whose methods
 * call NumPy / CuPy.  A B200 backend needs a native boundary below that class; this header is it.
 * Each entry point below cites the reference interface (file:line, relative to
 * /root/reference/src/neural_arch unless noted) whose work it performs.  The Python host binds them
 * with ctypes (neural-network-from-scratch_b200/nfb200/_lib.py); INTEGRATION.md shows the stub a
 * Neural Forge maintainer adds.
 *
 * Conventions: every function returns 0 on success and a negative value on failure (message from
 * nfb_last_error()), never throws, never allocates host memory the caller must free.  Pointers are
 * DEVICE pointers unless named host/pinned.  All work is issued on the library's current compute stream
 * (nfb_set_stream) except where an explicit stream handle is taken.  dtype codes: 0 f32, 1 f64, 2 i32,
 * 3 i64, 4 bool(u8), 5 bf16.  Strides / leading dimensions are in ELEMENTS.  sm_100a only.
 */
#ifndef NFB200_H
#define NFB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { NFB_F32 = 0, NFB_F64 = 1, NFB_I32 = 2, NFB_I64 = 3, NFB_BOOL = 4, NFB_BF16 = 5 };

/* ---- runtime: device, caching allocator, copies, streams/events, CUDA graphs ----------------------
 * replaces CuPy's device/memory-pool setup in backends/cuda_backend.py:24-58 and the CUDA-event timing of
 * backends/cuda_kernels.py:490-526 (CUDAKernelManager.benchmark_kernel). */
const char* nfb_last_error(void);
const char* nfb_version(void);
unsigned long long nfb_launch_count(void);           /* kernels launched by this library so far */
int nfb_device_count(int* out);
int nfb_init(int device);                            /* one process per GPU: LOCAL_RANK (distributed/launcher.py:145-162) */
int nfb_current_device(void);
int nfb_get_sm_count(void);
int nfb_device_name(char* buf, int len);
int nfb_malloc(size_t nbytes, void** out);           /* stream-ordered caching allocator */
int nfb_free(void* p);
int nfb_empty_cache(void);
int nfb_mem_stats(unsigned long long* stats);        /* [live, cached, peak, #cudaMalloc] */
int nfb_host_alloc(size_t nbytes, void** out);       /* pinned host memory */
int nfb_host_free(void* p);
int nfb_memcpy_h2d(void* dst, const void* src, size_t n);        /* Backend.from_numpy  (backends/backend.py:272) */
int nfb_memcpy_h2d_async(void* dst, const void* src, size_t n);
int nfb_memcpy_d2h(void* dst, const void* src, size_t n);        /* Backend.to_numpy    (backends/backend.py:267) */
int nfb_memcpy_d2h_async(void* dst, const void* src, size_t n);
int nfb_memcpy_d2d(void* dst, const void* src, size_t n);
int nfb_memset(void* dst, int value, size_t n);
int nfb_stream_create(void** out);
int nfb_stream_destroy(void* s);
void* nfb_get_stream(void);
int nfb_set_stream(void* s);
int nfb_stream_sync(void* s);
int nfb_device_sync(void);
int nfb_event_create(void** out);
int nfb_event_create_notiming(void** out);
int nfb_event_destroy(void* e);
int nfb_event_record(void* e, void* stream);
int nfb_event_record_external(void* event, void* stream);   /* like nfb_event_record; inside a stream capture it becomes an external event node whose time of the last replay is readable */
int nfb_event_sync(void* e);
int nfb_event_elapsed_ms(void* a, void* b, float* ms);
int nfb_stream_wait_event(void* stream, void* e);
int nfb_graph_begin_capture(void);
int nfb_graph_end_capture(void** out_exec);
int nfb_graph_launch(void* exec);
int nfb_graph_destroy(void* exec);
int nfb_flush_l2(void);

/* ---- elementwise: Backend.add/subtract/multiply/divide/power/maximum/minimum + comparisons
 * (backends/backend.py:107-131,216-244; numpy_backend.py:115-131,183-199), unary exp/log/sqrt/abs/sign/clip/tanh
 * (backends/backend.py:183-212) and the activation math of functional/activation.py:14-676.
 * op codes: binary 0 add 1 sub 2 mul 3 div 4 pow 5 max 6 min 10 eq 11 ne 12 lt 13 le 14 gt 15 ge;
 * unary 0 neg 1 exp 2 log 3 sqrt 4 abs 5 sign 6 tanh 7 sigmoid 8 relu 9 gelu_erf 10 gelu_tanh 11 silu 12 square
 * 13 recip 14 clip(p0,p1) 15 affine(x*p0+p1) 16 pow(x,p0) 17 rsqrt 18 nan_to_num 19 sin 20 cos 21 floor. */
int nfb_binary(int op, int dtype, const void* a, const void* b, void* out, int ndim, const int64_t* shape, const int64_t* sa, const int64_t* sb);
int nfb_binary_scalar(int op, int dtype, const void* a, double scalar, void* out, int64_t n, int reverse);
int nfb_unary(int op, int dtype, const void* x, void* out, int64_t n, double p0, double p1);
int nfb_unary_bwd(int op, int dtype, const void* src, const void* g, void* out, int64_t n);   /* functional/activation.py:389-462 */
int nfb_cast(int sd, int dd, const void* src, void* dst, int64_t n);                         /* Backend.astype (backends/backend.py:262) */
int nfb_where(int dtype, const void* cond, const int64_t* scond, const void* x, const int64_t* sx, const void* y, const int64_t* sy,
              void* out, int ndim, const int64_t* shape);                                    /* Backend.where (backends/backend.py:313) */
int nfb_copy_strided(int elem_size, const void* src, void* dst, int ndim, const int64_t* shape, const int64_t* src_strides,
                     const int64_t* dst_strides);                                            /* transpose/concatenate/stack/split (backends/backend.py:91-104,247-259) */
int nfb_fill(int dtype, void* p, int64_t n, double value);                                   /* zeros/ones/full (backends/backend.py:40-54) */
int nfb_arange(int dtype, void* p, int64_t n, double start, double step);                    /* arange (backends/backend.py:57) */
int nfb_swiglu_fwd(int dtype, const void* gate, const void* up, void* out, int64_t rows, int64_t C, int64_t in_ld);   /* models/language/gpt2.py:119-139 */
int nfb_swiglu_bwd(int dtype, const void* gate, const void* up, const void* dout, void* dgate, void* dup, int64_t rows, int64_t C, int64_t in_ld);
int nfb_add3(const float* a, const float* b, const float* c, float* out, int64_t n);

/* ---- reductions: Backend.sum/mean/max/min/argmax/argmin (backends/backend.py:144-180) over an [outer, R, inner] view;
 * nfb_colsum = reduce_gradient of a (T,n) -> (n) bias broadcast (functional/utils.py:33-96).  op: 0 sum 1 mean 2 max 3 min 4 prod. */
int nfb_reduce(int op, int dtype, const void* x, void* out, int64_t outer, int64_t R, int64_t inner);
int nfb_argreduce(int is_max, int dtype, const void* x, int64_t* out, int64_t outer, int64_t R, int64_t inner);
int nfb_colsum(int dtype, const void* x, int64_t R, int C, int64_t ld, float* out, int accumulate);

/* ---- normalisation: LayerNorm / RMSNorm forward + backward (nn/normalization.py:86-221,269-359;
 * models/language/gpt2.py:98-116; CuPy precedent CudaBackend.layernorm backends/cuda_backend.py:334-352).  kind 0 LN, 1 RMS. */
int nfb_norm_fwd(int kind, int xdt, int ydt, const void* x, const void* res, void* sum_out, const float* gamma, const float* beta,
                 void* y, float* mean, float* rstd, int64_t rows, int cols, float eps);
int nfb_norm_bwd(int kind, int xdt, int ydt, const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                 const void* dres, void* dx, float* dgamma, float* dbeta, int accumulate, int64_t rows, int cols, float clip, void* dx_twin);

/* ---- softmax and cross entropy (functional/activation.py:65-147; functional/loss.py:15-112) */
int nfb_softmax_fwd(const float* x, float* y, int64_t rows, int cols);
int nfb_softmax_bwd(const float* p, const float* g, float* dx, int64_t rows, int cols);
int nfb_cross_entropy(const float* logits, const void* targets, int tgt_is_i64, float* loss, void* dlogits, int ddt, int64_t rows, int V,
                      int64_t ld, int64_t dld, float grad_scale);
/* same with the logits dtype stated (NFB_F32 or NFB_BF16; bf16 logits need the wide-row kernel: 8192 <= V <= 53248) */
int nfb_cross_entropy_set_bulk(int on);   /* test hook: 0 = bf16 logits through the register-resident kernel instead of the bulk-copy one */
int nfb_cross_entropy_ex(int ldt, const void* logits, const void* targets, int tgt_is_i64, float* loss, void* dlogits, int ddt, int64_t rows,
                         int V, int64_t ld, int64_t dld, float grad_scale);

/* ---- embedding gather / bit-exact ordered scatter-add (nn/embedding.py:27-63) */
int nfb_gather_rows(int idx_dtype, int wdt, int odt, const void* W, const void* idx, void* out, int64_t n, int d, int64_t V);
int nfb_scatter_add_rows(int idx_dtype, int gdt, const void* grad, const void* idx, float* dW, int64_t n, int d, int64_t V, int accumulate);
int nfb_index_error_check(int* flag);

/* ---- contractions: Backend.matmul / dot (backends/backend.py:133-140; functional/arithmetic.py:441-575) and the
 * Linear forward / dgrad / wgrad of nn/optimized.py:149-317 and nn/linear.py:202-205 (fused bias + tanh-GELU/ReLU:
 * optimization/fusion.py:55-80; CuPy precedent CudaBackend.fused_linear_gelu backends/cuda_backend.py:321-332).
 * nfb_gemm_f32: SIMT fp32 (1e-5 contract); nfb_gemm_bf16: tcgen05/TMEM + TMA (throughput path). */
int nfb_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int64_t lda, int64_t strideA, const float* B,
                 int64_t ldb, int64_t strideB, float beta, float* C, int64_t ldc, int64_t strideC, int batch, const float* bias);
int nfb_gemm_f64(int transA, int transB, int M, int N, int K, const double* A, int64_t lda, int64_t strideA, const double* B, int64_t ldb,
                 int64_t strideB, double* C, int64_t ldc, int64_t strideC, int batch);   /* Backend.matmul on float64 arrays (not on the training path) */
/* fp32-accurate contraction on the tensor cores: x (R, C) fp32 (row pitch ld) is split three ways into bf16 (x = x1 + x2 + x3,
 * 24 significand bits) and written with its K axis concatenated six times -- role 0 (A operand): parts 2 1 3 1 2 1, role 1 (B operand):
 * 2 3 1 2 1 1 (smallest products first) -- as (R, 6*C) when k_axis = 1 (K along the columns) or (6*R, C) when k_axis = 0.  nfb_gemm_bf16 over the two outputs
 * with K' = 6K and an fp32 C then equals the fp32 product to O(2^-24) (functional/arithmetic.py:441-575 at the 1e-5 tolerance). */
int nfb_split_bf16x3(const float* x, int64_t R, int64_t C, int64_t ld, int role, int k_axis, void* out);
int nfb_gemm_bf16(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int c_dtype,
                  int64_t ldc, const float* bias, const float* residual, int64_t ldr, void* aux, int64_t ldaux, int epilogue, int accumulate,
                  int split_k, float clip);
int nfb_gemm_bf16_force_tile(int block_n);           /* test hook: pin the N tile (0 = automatic) */
int nfb_gemm_bf16_force_cta_group(int cg);           /* test hook: 1 = single-CTA tiles, 2 = cta_group::2 SM pairs, 0 = automatic */
int nfb_gemm_bf16_force_raster(int order);           /* test hook: tile order 1 = m fastest, 2 = n fastest, 0 = automatic (L2-reuse estimate) */
int nfb_gemm_bf16_set_auto_split(int mode);          /* internal split-K of under-filled non-accumulating launches: 0 = off, 1 = automatic, 2 = whenever eligible (test hook) */
/* What nfb_gemm_bf16 would do with a launch of this shape on `sms` SMs (aligned operands, no bias / residual / aux):
   out[6] = {tile N, CTA group, k-splits, n-fastest tile order, internal split-K slices, stream-K}.  Host logic only, no device needed. */
int nfb_gemm_bf16_plan(int transB, int M, int N, int K, int accumulate, int epilogue, int split_k, int sms, int* out);
/* Grouped launch: n <= 4 independent contractions C_g (+)= opA_g . opB_g with the same operand layouts / output type in ONE
   persistent stream-K kernel (no epilogue operands) -- the four weight-gradient GEMMs dW = dY^T X of a Transformer layer
   (nn/optimized.py:286-317 computes them one np.matmul at a time).  Every output element gets exactly one store / reduce-add. */
int nfb_gemm_bf16_grouped(int n, int transA, int transB, const int* Ms, const int* Ns, const int* Ks, const void* const* As, const int64_t* ldas,
                          const void* const* Bs, const int64_t* ldbs, void* const* Cs, const int64_t* ldcs, int c_dtype, int accumulate);
/* Fused rotary embedding (models/language/gpt2.py:78-95 apply_rotary_pos_emb, rotate_half form): the NEXT nfb_gemm_bf16 launch rotates
   output columns [0, cols) -- heads of 64 columns, position = row % S, tables (S, 32) fp32 -- in its bf16 TMA-store epilogue, before
   the rounding to bf16 (the qkv projection: cols = 2 * d covers the q and k parts of the packed (T, 3, H, 64) output). */
int nfb_gemm_bf16_set_rope(const float* cos_t, const float* sin_t, int S, int cols);
int nfb_gemm_bf16_set_stream_k(int mode);            /* stream-K schedule (balanced (tile, k-block) ranges per SM pair + in-kernel fix-up) of SINGLE launches: 0 / 1 = classic (grouped launches always use it), 2 = whenever possible (test hook) */
int nfb_gemm_bf16_set_sm_limit(int sms);             /* SMs the persistent GEMM grids may occupy (0 = all): data-parallel steps leave room for NCCL's channel CTAs */
int nfb_gemm_bf16_set_span(void* dev_two_u64);      /* measurement: the NEXT launch writes min(%globaltimer at CTA entry) and ~max(%globaltimer at CTA exit) into two u64 words armed to all-ones */
int nfb_gemm_bf16_set_timeline(void* dev_buf);      /* debug: 8 x 64 u64 clock64 stamps of the GEMM pipeline events (NULL = off) */
int nfb_gemm_bf16_set_tma_store(int on);             /* test hook: 0 = register epilogue everywhere (default 1: TMA store when C rows are 16-byte addressable) */
int nfb_set_aux_stream(void* stream);   /* side stream for optimizer-only work (norm-parameter reductions; the host also issues wgrad GEMMs / bias sums there); NULL = none */
int nfb_norm_flush(void);                /* with an aux stream, nfb_norm_bwd QUEUES its dgamma / dbeta block-partial reductions (up to 8 go out as one launch); this launches whatever is queued -- call before joining / fencing the aux stream */
int nfb_set_pdl(int on);                             /* programmatic dependent launch of the GEMM kernels (default on; env NFB_PDL=0) */

/* ---- attention: RoPE (models/language/gpt2.py:25-95) and the fused QK^T -> softmax -> V core of
 * models/language/gpt2.py:183-215, models/language/bert.py:155-231, models/vision/vision_transformer.py:100-134
 * (CuPy precedent CudaBackend.flash_attention backends/cuda_backend.py:354-375).  mask_mode bit0 causal(-1e4), bit1 key mask. */
int nfb_rope(int dtype, void* x, const float* cos_t, const float* sin_t, int n_parts, int64_t part_stride, int B, int S, int H, int head_dim,
             int64_t sb, int64_t ss, int64_t sh, int inverse);
/* RotaryPositionalEmbedding.apply_rope (nn/positional.py:212-270): interleaved pairs (2j, 2j+1), tables (S, head_dim/2), in place on a
 * (b, s, h, d)-strided tensor; inverse = 1 applies the transpose rotation (the gradient of the forward rotation). */
int nfb_rope_interleaved(int dtype, void* x, const float* cos_t, const float* sin_t, int B, int S, int H, int head_dim, int64_t sb, int64_t ss,
                         int64_t sh, int inverse);
int nfb_flash_attn_set_timeline(void* dev_buf);   /* debug: 256 x u64 clock64 stamps of the tcgen05 forward (NULL = off) */
int nfb_flash_attn_set_poly(int fwd, int bwd);   /* tcgen05 attention: n of every 4 softmax exponentials go through a degree-4 polynomial on the FMA pipe instead of MUFU.EX2 (0, 1 or 2; the kernels are MUFU-bound at head_dim 64) */
int nfb_flash_attn_set_tc(int on);   /* test hook: 0 = mma.sync flash kernels only, 1 = tcgen05 forward only, 3 = tcgen05 forward + backward (default) when the layout is TMA-addressable */
int nfb_flash_attn_fwd(int dtype, const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int S, int head_dim,
                       int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss,
                       int64_t o_sh, int Skv, float scale, int mask_mode, const float* key_mask);
int nfb_flash_attn_bwd(int dtype, const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq,
                       void* dk, void* dv, int B, int H, int S, int head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t kv_sb,
                       int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh, int Skv, float scale, int mask_mode,
                       const float* key_mask);

/* Grouped-query / multi-query attention (nn/modern_attention.py:26-358: GroupedQueryAttention, MultiQueryAttention): q / o / lse / dq
 * have H heads, k / v / dk / dv have Hkv heads, H % Hkv == 0.  Query head h reads key/value head h / (H / Hkv) in place (the
 * repeat_kv copy of modern_attention.py:137-161 is never materialised); dk / dv come back already summed over each group. */
int nfb_flash_attn_fwd_gqa(int dtype, const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv, int S, int head_dim,
                           int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss,
                           int64_t o_sh, int Skv, float scale, int mask_mode, const float* key_mask);
int nfb_flash_attn_bwd_gqa(int dtype, const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq,
                           void* dk, void* dv, int B, int H, int Hkv, int S, int head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh,
                           int64_t kv_sb, int64_t kv_ss, int64_t kv_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh, int Skv, float scale,
                           int mask_mode, const float* key_mask);

int nfb_flash_attn_set_dq_f32_blocks(int n); /* test hook: dQ partials of more than n key blocks are summed in fp32 (default 4) */

/* ---- optimizers: Adam.step / AdamW.step / SGD.step (optim/adam.py:165-252, optim/adamw.py:278-372, optim/sgd.py:87-115)
 * and global-norm clipping (examples/translation/train_conversational.py:38-54; distributed/advanced.py:362-388). */
int nfb_adam_step(int variant, int moment_dtype, float* p, const float* g, void* m, void* v, void* p_bf16, int64_t n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int step, int maximize, const float* grad_scale, const int* step_dev);
/* lr_dev (device double, optional): the learning rate is read on the device, so the schedulers of optim/lr_scheduler.py:19-606
 * keep acting on a training step that was captured into a CUDA graph. */
int nfb_adam_step_ex(int variant, int moment_dtype, float* p, const float* g, void* m, void* v, void* p_bf16, int64_t n, double lr,
                     double beta1, double beta2, double eps, double weight_decay, int step, int maximize, const float* grad_scale,
                     const int* step_dev, const double* lr_dev);
/* Lion.step (optim/lion.py:129-175): sign-of-interpolation update clipped to +-1, decoupled lr*wd*p term, EMA momentum. */
int nfb_lion_step(int moment_dtype, float* p, const float* g, void* m, void* p_bf16, int64_t n, double lr, double beta1, double beta2,
                  double weight_decay, int maximize, const float* grad_scale, const double* lr_dev);
int nfb_add_i32(int* counter, int delta);
int nfb_sgd_step(float* p, const float* g, float* buf, void* p_bf16, int64_t n, double lr, double momentum, double dampening,
                 double weight_decay, int nesterov, int first_step, int maximize);
int nfb_sgd_step_ex(float* p, const float* g, float* buf, void* p_bf16, int64_t n, double lr, double momentum, double dampening,
                    double weight_decay, int nesterov, int first_step, int maximize, const double* lr_dev);   /* lr_dev: as nfb_adam_step_ex */
int nfb_sumsq(const float* g, int64_t n, double* total, int accumulate);
int nfb_clip_coef(const double* total, float max_norm, float eps, float* coef, float* norm_out);
int nfb_scale_by_device_scalar(float* g, int64_t n, const float* coef);
/* AdvancedGradScaler.unscale_ (optimization/grad_scaler.py:131-196): g /= scale per parameter, stop at the first parameter with a
 * non-finite value (*first_bad, nseg if none), clip each parameter's own L2 norm to clip_threshold (0 = off).  chunks: nchunks
 * records {int32 seg; int32 len; int64 start} cutting g into per-parameter pieces; norms2[nseg] = sum((g/scale)^2). */
int nfb_grad_unscale(float* g, const void* chunks, int nchunks, int nseg, float scale, float clip_threshold, double* norms2,
                     int* first_bad);

/* ---- collectives: NCCL over NVLink 5 / NVSwitch (distributed/communication.py:115-347 NCCLBackend; ReduceOp :25-32;
 * call sites distributed/data_parallel.py:88-99,250-287).  op: 0 sum 1 average 2 max 3 min 4 product. */
int nfb_comm_load(const char* path);
int nfb_comm_unique_id(char* out, int len);          /* 128 bytes, created on rank 0, shipped by the host rendezvous */
int nfb_comm_init(int rank, int world, const char* id_bytes, int len, void** comm_out);
int nfb_comm_destroy(void* comm);
int nfb_comm_allreduce(void* comm, const void* send, void* recv, int64_t count, int dtype, int op, void* stream);
int nfb_comm_broadcast(void* comm, void* buf, int64_t count, int dtype, int root, void* stream);
int nfb_comm_allgather(void* comm, const void* send, void* recv, int64_t sendcount, int dtype, void* stream);
int nfb_comm_reduce_scatter(void* comm, const void* send, void* recv, int64_t recvcount, int dtype, int op, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NFB200_H */
