"""Per-CTA pipeline timeline (clock64 stamps) of one GEMM launch: python scratch/gemm_timeline.py M N K ta tb bn cg outdt"""
import sys, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R, _lib
from nfb200.array import B200Array, bfloat16
be = nfb200.get_backend("b200")
M, N, K, ta, tb, bn, cg = [int(x) for x in sys.argv[1:8]]
odt = sys.argv[8] if len(sys.argv) > 8 else "bf16"
rng = np.random.default_rng(0)
def rb(shape): return Kn.to_bf16(B200Array.from_numpy((rng.standard_normal(shape) * 0.5).astype(np.float32)))
a = rb((K, M) if ta else (M, K)); b = rb((N, K) if tb else (K, N))
out = B200Array.full((M, N), 0.0, np.float32 if odt == "f32" else bfloat16)
_lib.call("nfb_gemm_bf16_force_tile", bn); _lib.call("nfb_gemm_bf16_force_cta_group", cg)
f = lambda: Kn.gemm_bf16(a, b, bool(ta), bool(tb), out=out, out_dtype=(bfloat16 if odt == "bf16" else np.float32))
for _ in range(3): f()
tl = B200Array.full((8 * 64,), 0, np.int64)
_lib.call("nfb_gemm_bf16_set_timeline", tl.ptr)
f(); R.synchronize()
_lib.call("nfb_gemm_bf16_set_timeline", None)
t = tl.get().reshape(8, 64)
print(sys.argv[1:])
for blk in (0, 1, 5):
    r = t[blk]; t0 = r[0]
    print(f"block {blk}: setup {r[1]-t0}  firstTMA {r[2]-t0}  end {r[3]-t0}")
    for i in range(14):
        s = r[8 + 4*i: 12 + 4*i]
        if s[2] == 0: break
        print(f"   tile {i}: data {s[0]-t0}  mma_issued {s[1]-t0}  acc_ready {s[2]-t0}  epi_done {s[3]-t0}")
print("ms", R.time_ms(f, iters=20, warmup=3))
