"""GPU-bound GEMM microbench: every (shape, forced tile-N) is captured as a CUDA graph of NSET*REP launches rotating over
NSET operand sets (weights/activations of different launches do not alias), random bf16 data, timed by graph replay."""
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
 ("square 8192", 8192, 8192, 8192, False, True, "bf16", False),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
tiles = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 64, 128, 192, 256]
cgs = [int(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2]
if len(sys.argv) > 4: _lib.call("nfb_gemm_bf16_force_raster", int(sys.argv[4]))   # 1 = m fastest, 2 = n fastest tile order
rng = np.random.default_rng(0)
def rand_bf16(shape):
    x = B200Array.from_numpy((rng.standard_normal(shape) * 0.5).astype(np.float32))
    return Kn.to_bf16(x)
for name, M, N, K, ta, tb, odt, acc in shapes:
    if only and only not in name: continue
    big = M * N * K > 1e11
    NSET, REP = (2, 2) if big else (4, 5)
    Kp = (K + 7)//8*8
    sets = []
    for _ in range(NSET):
        a = rand_bf16((K, (M+7)//8*8) if ta else (M, Kp)); a = a[:, :M] if ta else a[:, :K]
        b = rand_bf16((N, Kp) if tb else (K, (N+7)//8*8)); b = b[:, :K] if tb else b[:, :N]
        ldc = (N + 7) // 8 * 8   # the library pads wide ragged outputs (logits) to a 16-byte row pitch; do the same here
        out = B200Array.full((M, ldc), 0.0, np.float32 if (acc or odt == "f32") else bfloat16)[:, :N]
        sets.append((a, b, out))
    res = []
    for cg, bn in [(0, 0)] + [(c, b) for c in cgs for b in tiles if b]:
        _lib.call("nfb_gemm_bf16_force_tile", bn)
        _lib.call("nfb_gemm_bf16_force_cta_group", cg)
        def run():
            for _ in range(REP):
                for a, b, out in sets:
                    Kn.gemm_bf16(a, b, ta, tb, out=out, out_dtype=(bfloat16 if odt == "bf16" else np.float32), accumulate=acc)
        try:
            run(); R.synchronize()
            g = R.Graph()
            with g: run()
            R.synchronize()
            ms = R.time_ms(g.replay, iters=5, warmup=2) / (NSET * REP)
            res.append(f"{cg}x{bn}:{ms*1e3:6.1f}us {2*M*N*K/ms/1e9:5.0f}TF")
        except Exception as e:
            res.append(f"{cg}x{bn}: ERR {str(e)[:30]}")
    _lib.call("nfb_gemm_bf16_force_tile", 0)
    _lib.call("nfb_gemm_bf16_force_cta_group", 0)
    print(f"{name:13s} M={M:5d} N={N:5d} K={K:5d} | " + " | ".join(res), flush=True)
