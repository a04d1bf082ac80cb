"""Stream-K / grouped-launch calibration of the tcgen05 GEMM on the GPT-2 small step shapes: every shape under the classic
schedule (stream_k = 0), the forced stream-K schedule (2) and the library's own choice (1), then the four weight gradients
of a block as four launches vs ONE grouped launch.  Graph replay over rotating operand sets, as bench_gemm.py."""
import sys, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R, _lib
from nfb200.array import B200Array, bfloat16
be = nfb200.get_backend("b200")
shapes = [
 ("fwd qkv", 4096, 2304, 768, False, True, "bf16", False),
 ("fwd proj+res", 4096, 768, 768, False, True, "f32", False),
 ("fwd w12", 4096, 3072, 768, False, True, "bf16", False),
 ("fwd w3", 4096, 768, 1536, False, True, "f32", False),
 ("dgrad w3", 4096, 1536, 768, False, False, "bf16", False),
 ("dgrad w12", 4096, 768, 3072, False, False, "bf16", False),
 ("dgrad qkv", 4096, 768, 2304, False, False, "bf16", False),
 ("dgrad proj", 4096, 768, 768, False, False, "bf16", False),
 ("wgrad w3", 768, 1536, 4096, True, False, "f32", True),
 ("wgrad w12", 3072, 768, 4096, True, False, "f32", True),
 ("wgrad proj", 768, 768, 4096, True, False, "f32", True),
 ("wgrad qkv", 2304, 768, 4096, True, False, "f32", True),
 ("fwd logits", 4096, 50257, 768, False, True, "bf16", False),
 ("dgrad logits", 4096, 768, 50257, False, False, "bf16", False),
 ("wgrad logits", 50257, 768, 4096, True, False, "f32", True),
 ("bert ffn1", 16384, 3072, 768, False, True, "bf16", False),
 ("bert dgrad ffn2", 16384, 3072, 768, False, False, "bf16", False),
 ("g2m dgrad proj", 8192, 1024, 1024, False, False, "bf16", False),
 ("g2m fwd w3", 8192, 1024, 2048, False, True, "f32", False),
 ("square 8192", 8192, 8192, 8192, False, True, "bf16", False),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
rng = np.random.default_rng(0)
def rand_bf16(shape):
    x = B200Array.from_numpy((rng.standard_normal(shape) * 0.5).astype(np.float32))
    return Kn.to_bf16(x)
def timed(run, n):
    run(); R.synchronize()
    g = R.Graph()
    with g: run()
    R.synchronize()
    return R.time_ms(g.replay, iters=5, warmup=2) / n
def mk(M, N, K, ta, tb, odt, acc):
    Kp = (K + 7)//8*8
    a = rand_bf16((K, (M+7)//8*8) if ta else (M, Kp)); a = a[:, :M] if ta else a[:, :K]
    b = rand_bf16((N, Kp) if tb else (K, (N+7)//8*8)); b = b[:, :K] if tb else b[:, :N]
    ldc = (N + 7) // 8 * 8
    out = B200Array.full((M, ldc), 0.0, np.float32 if (acc or odt == "f32") else bfloat16)[:, :N]
    return a, b, out
for name, M, N, K, ta, tb, odt, acc in shapes:
    if only and only not in name: continue
    big = M * N * K > 1e11
    NSET, REP = (2, 2) if big else (4, 5)
    sets = [mk(M, N, K, ta, tb, odt, acc) for _ in range(NSET)]
    res = []
    for label, mode, bn, cg in [("classic", 0, 0, 0), ("sk", 2, 0, 0), ("sk2x256", 2, 256, 2), ("sk2x128", 2, 128, 2), ("sk1x256", 2, 256, 1), ("auto", 1, 0, 0)]:
        _lib.call("nfb_gemm_bf16_set_stream_k", mode)
        _lib.call("nfb_gemm_bf16_force_tile", bn)
        _lib.call("nfb_gemm_bf16_force_cta_group", cg)
        def run():
            for _ in range(REP):
                for a, b, out in sets:
                    Kn.gemm_bf16(a, b, ta, tb, out=out, out_dtype=(bfloat16 if odt == "bf16" else np.float32), accumulate=acc)
        try:
            ms = timed(run, NSET * REP)
            res.append(f"{label}:{ms*1e3:6.1f}us {2*M*N*K/ms/1e9:5.0f}TF")
        except Exception as e:
            res.append(f"{label}: ERR {str(e)[:40]}")
    _lib.call("nfb_gemm_bf16_set_stream_k", 1)
    _lib.call("nfb_gemm_bf16_force_tile", 0)
    _lib.call("nfb_gemm_bf16_force_cta_group", 0)
    print(f"{name:15s} M={M:5d} N={N:5d} K={K:5d} | " + " | ".join(res), flush=True)

if not only or "group" in only:
    for tag, T, d in (("gpt2-small block", 4096, 768), ("gpt2-medium block", 8192, 1024)):
        wg = [(3 * d, d), (d, d), (4 * d, d), (d, 2 * d)]   # qkv, proj, w1|w2, w3: dW (out,in) = dZ^T (out,T) . X (T,in)
        NSET = 3
        sets = []
        for _ in range(NSET):
            sets.append([(rand_bf16((T, m)), rand_bf16((T, n)), B200Array.full((m, n), 0.0, np.float32)) for m, n in wg])
        flop = sum(2.0 * T * m * n for m, n in wg)
        out = []
        for label, mode in (("4 launches classic", 0), ("4 launches auto", 1)):
            _lib.call("nfb_gemm_bf16_set_stream_k", mode)
            def run():
                for st in sets:
                    for a, b, o in st:
                        Kn.gemm_bf16(a, b, True, False, out=o, accumulate=True)
            ms = timed(run, NSET)
            out.append(f"{label}: {ms*1e3:6.1f}us {flop/ms/1e9:5.0f}TF")
        _lib.call("nfb_gemm_bf16_set_stream_k", 1)
        for bn, cg in ((0, 0), (256, 2), (128, 2), (256, 1)):
            _lib.call("nfb_gemm_bf16_force_tile", bn)
            _lib.call("nfb_gemm_bf16_force_cta_group", cg)
            def run():
                for st in sets:
                    Kn.gemm_bf16_grouped(st, True, False, accumulate=True)
            try:
                ms = timed(run, NSET)
                out.append(f"grouped {cg}x{bn}: {ms*1e3:6.1f}us {flop/ms/1e9:5.0f}TF")
            except Exception as e:
                out.append(f"grouped {cg}x{bn}: ERR {str(e)[:40]}")
        _lib.call("nfb_gemm_bf16_force_tile", 0)
        _lib.call("nfb_gemm_bf16_force_cta_group", 0)
        print(f"wgrad x4 {tag}: " + " | ".join(out), flush=True)
