"""HBM-roofline microbench of the bandwidth-bound kernels at the GPT-2 small shapes (T = 4096 tokens, d = 768, V = 50257,
109.4 M parameters).  Each kernel is captured in a CUDA graph as REP x (L2 flush + kernel); the flush-only graph is timed
too and subtracted, so the figure is the cold-L2 kernel time without host launch overhead.  GB/s = algorithmic bytes / time."""
import sys, json, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R, _lib
from nfb200.array import B200Array, bfloat16, dtype_code
be = nfb200.get_backend("b200")
call = _lib.call
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if len(sys.argv) < 2 else float(sys.argv[1])
REP = 10
rng = np.random.default_rng(0)
T, d, V, NP = 4096, 768, 50257, 109_432_320
def f32(*shape, scale=1.0): return B200Array.from_numpy((rng.standard_normal(shape) * scale).astype(np.float32))
def bf(*shape, scale=1.0): return Kn.to_bf16(f32(*shape, scale=scale))

def timed(fn):
    def run(with_fn):
        for _ in range(REP):
            R.flush_l2()
            if with_fn: fn()
    fn(); R.synchronize()
    ga = R.Graph()
    with ga: run(True)
    gb = R.Graph()
    with gb: run(False)
    R.synchronize()
    ta = min(R.time_ms(ga.replay, iters=3, warmup=1) for _ in range(3))
    tb = min(R.time_ms(gb.replay, iters=3, warmup=1) for _ in range(3))
    return (ta - tb) / REP

rows = []
def report(name, nbytes, fn):
    ms = timed(fn)
    gbs = nbytes / ms / 1e6
    rows.append((name, nbytes, ms, gbs))
    print(f"{name:34s} {nbytes/1e6:9.1f} MB {ms*1e3:9.1f} us {gbs:8.0f} GB/s  {gbs/PEAK*100:5.1f}% of {PEAK:.0f}", flush=True)

_w = f32(T, d)
timed(lambda: be.norm_forward(_w, None, None, 1e-5, rms=True, out_dtype=bfloat16))   # warm-up (clocks, allocator)
# ---- RMSNorm fwd (x fp32 -> y bf16) and bwd (dy bf16, x fp32, dres fp32 -> dx fp32 + bf16 twin)
x = f32(T, d); w = f32(d); dres = f32(T, d); dyb = bf(T, d)
y, mean, rstd = be.norm_forward(x, w, None, 1e-5, rms=True, out_dtype=bfloat16)
report("rmsnorm fwd  f32->bf16", T*d*(4+2), lambda: be.norm_forward(x, w, None, 1e-5, rms=True, out_dtype=bfloat16))
dw = f32(d)
report("rmsnorm bwd  (+dres, twin, dgamma)", T*d*(2+4+4+4+2), lambda: be.norm_backward(dyb, x, w, mean, rstd, rms=True, dresidual=dres, dweight_out=dw, accumulate=True, clip=10.0, bf16_twin=True))
# ---- LayerNorm (BERT / ViT shapes: same T x 768)
b = f32(d); yl, meanl, rstdl = be.norm_forward(x, w, b, 1e-12, rms=False, out_dtype=bfloat16)
db = f32(d)
report("layernorm fwd f32->bf16", T*d*(4+2), lambda: be.norm_forward(x, w, b, 1e-12, rms=False, out_dtype=bfloat16))
report("layernorm bwd (+dres, twin, dg, db)", T*d*(2+4+4+4+2), lambda: be.norm_backward(dyb, x, w, meanl, rstdl, rms=False, dresidual=dres, dweight_out=dw, dbias_out=db, accumulate=True, clip=10.0, bf16_twin=True))
# ---- SwiGLU on the packed (T, 2*1536) bf16 projection
C2 = 1536
gu = bf(T, 2 * C2); dout = bf(T, C2); out = B200Array.empty((T, C2), bfloat16); dgu = B200Array.empty((T, 2 * C2), bfloat16)
report("swiglu fwd bf16", T*C2*2*3, lambda: call("nfb_swiglu_fwd", dtype_code(bfloat16), gu.ptr, gu.ptr + C2*2, out.ptr, T, C2, 2*C2))
report("swiglu bwd bf16", T*C2*2*5, lambda: call("nfb_swiglu_bwd", dtype_code(bfloat16), gu.ptr, gu.ptr + C2*2, dout.ptr, dgu.ptr, dgu.ptr + C2*2, T, C2, 2*C2))
# ---- bias gradient (column sums of dY, bf16 (T, 2304))
dy2 = bf(T, 2304); dbias = f32(2304)
report("colsum bf16 (T,2304)", T*2304*2, lambda: call("nfb_colsum", dtype_code(bfloat16), dy2.ptr, T, 2304, 2304, dbias.ptr, 1))
# ---- cross entropy fwd+bwd: logits (T, V) fp32 or bf16, dlogits bf16
ldv = (V + 7) // 8 * 8
tg = B200Array.from_numpy(rng.integers(0, V, (T,)).astype(np.int64))
lg32 = B200Array.from_numpy((rng.standard_normal((T, ldv))).astype(np.float32))[:, :V]
report("cross-entropy f32 logits -> bf16 d", T*V*(4+2), lambda: be.cross_entropy(lg32, tg, "mean", True, bfloat16))
lg16 = Kn.to_bf16(lg32.contiguous() if False else B200Array.from_numpy((rng.standard_normal((T, ldv))).astype(np.float32)))[:, :V]
report("cross-entropy bf16 logits -> bf16 d", T*V*(2+2), lambda: be.cross_entropy(lg16, tg, "mean", True, bfloat16))
del lg32, lg16
# ---- AdamW over the flat 109.4 M-parameter arena (fp32 p/g/m/v + bf16 shadow)
p = B200Array.full((NP,), 0.01, np.float32); g = B200Array.full((NP,), 0.001, np.float32)
m = B200Array.full((NP,), 0.0, np.float32); v = B200Array.full((NP,), 0.0, np.float32); pb = B200Array.empty((NP,), bfloat16)
report("adamw fp32 moments (+bf16 shadow)", NP*30, lambda: call("nfb_adam_step", 1, 0, p.ptr, g.ptr, m.ptr, v.ptr, pb.ptr, NP, 5e-4, 0.9, 0.999, 1e-6, 0.1, 3, 0, None, None))
del p, g, m, v, pb
# ---- embedding gather and RoPE
W = f32(V, d, scale=0.02); ids = B200Array.from_numpy(rng.integers(0, V, (T,)).astype(np.int64))
report("embedding gather (T rows of 768)", T*d*(4+4), lambda: be.embedding_forward(W, ids))
qkv = bf(8, 512, 3, 12, 64)
report("rope q,k in place (packed qkv)", 8*512*2*12*64*2*2, lambda: Kn.rope_packed_qk_inplace(qkv, 12))
# ---- loss-scaler unscale + finite check + per-parameter norm clip over a 109.4 M-element gradient arena cut like GPT-2 small
# (two passes: statistics reads g once, apply reads and writes it: 12 B / element)
from nfb200.core import Parameter
from nfb200.optim import SGD
from nfb200.optimization import AdvancedGradScaler, ScalerConfig
shapes = [(V, d)] + [s_ for _ in range(12) for s_ in ((d,), (2304, d), (d, d), (d,), (3072, d), (d, 1536))] + [(d,)]
params = {f"p{i}": Parameter(B200Array.full(s_, 0.01, np.float32)) for i, s_ in enumerate(shapes)}
sgd = SGD(params, lr=0.0)
for p_ in params.values():
    p_.grad = B200Array.full(p_.shape, 1e-3, np.float32)
sc = AdvancedGradScaler(ScalerConfig(init_scale=1.0, gradient_clip_threshold=1e9))
sc.unscale_(sgd)
tab = sc._arena_table(sgd.arena, list(params.values()))
ntot = sum(int(np.prod(s_)) for s_ in shapes)
report("grad unscale+check+clip (arena)", ntot*12, lambda: call("nfb_grad_unscale", sgd.arena.flat_g.ptr, tab[0].ptr, tab[1], len(shapes), 1.0, 1e9, tab[2].ptr, tab[3].ptr))
del params, sgd
# ---- interleaved-pair RoPE (nn/positional.py) in place on bf16 (8, 12, 512, 64)
from nfb200.nn import RotaryPositionalEmbedding
rp = RotaryPositionalEmbedding(64, 512)
xq = bf(8, 512, 12, 64).transpose(0, 2, 1, 3)
ct, st_ = B200Array.from_numpy(rp.cos_cached), B200Array.from_numpy(rp.sin_cached)
report("rope interleaved in place bf16", 8*512*12*64*2*2, lambda: Kn.rope_interleaved_inplace(xq, ct, st_))
json.dump([{"kernel": n, "bytes": b_, "ms": ms, "gbs": gb, "frac": gb / PEAK} for n, b_, ms, gb in rows], open("gpurun_out/bw_kernels.json", "w"), indent=1)
