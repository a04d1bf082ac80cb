"""Accuracy and speed of the fp32 contraction paths: FFMA kernel vs the bf16x3 split on tcgen05, against the float64 product.
Also the pure accumulation error of the tensor-core pipe (bf16 GEMM vs float64 product of the SAME bf16-rounded operands)."""
import sys, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R
from nfb200.array import B200Array
be = nfb200.get_backend("b200")
rng = np.random.default_rng(0)
def rel(got, want, scale): return float(np.max(np.abs(got - want) / scale)), float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
for M, N, K, ta, tb in [(4096, 2304, 768, False, True), (4096, 768, 3072, False, True), (768, 2304, 4096, True, False), (4096, 768, 768, False, False), (2048, 2048, 8192, False, True)]:
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = (rng.standard_normal((N, K) if tb else (K, N)) * 0.02).astype(np.float32)
    A64, B64 = (a.T if ta else a).astype(np.float64), (b.T if tb else b).astype(np.float64)
    want = A64 @ B64
    scale = np.abs(A64) @ np.abs(B64)
    da, db = B200Array.from_numpy(a), B200Array.from_numpy(b)
    out = {}
    for name, on in (("ffma", False), ("bf16x3", True)):
        Kn.set_fp32_tensor_core(on)
        got = Kn.gemm_f32(da, db, ta, tb).get().astype(np.float64)
        def fn():
            for _ in range(4): Kn.gemm_f32(da, db, ta, tb)
        fn(); R.synchronize()
        g = R.Graph()
        with g: fn()
        R.synchronize()
        ms = R.time_ms(g.replay, iters=5, warmup=2) / 4
        e_scale, e_max = rel(got, want, scale)
        # elementwise relative error where the result is not a near-cancellation
        big = np.abs(want) > 0.1 * np.abs(want).mean()
        e_rel = float(np.max(np.abs(got - want)[big] / np.abs(want)[big]))
        out[name] = (ms, e_scale, e_max, e_rel)
    # accumulation error of the tensor pipe alone
    ab, bb = Kn.to_bf16(da), Kn.to_bf16(db)
    gotb = Kn.gemm_bf16(ab, bb, ta, tb, split_k=1).get().astype(np.float64)
    A16, B16 = ab.get().astype(np.float64), bb.get().astype(np.float64)
    wantb = (A16.T if ta else A16) @ (B16.T if tb else B16)
    acc_scale = float(np.max(np.abs(gotb - wantb) / scale))
    print(f"M={M} N={N} K={K} ta={ta} tb={tb}: " + " | ".join(f"{n}: {v[0]*1e3:7.1f} us {2.0*M*N*K/v[0]/1e9:6.1f} TF/s  err/(|A||B|) {v[1]:.1e}  err/max|C| {v[2]:.1e}  max rel (|C|>0.1 mean) {v[3]:.1e}" for n, v in out.items())
          + f" | tensor-pipe accumulation alone err/(|A||B|) {acc_scale:.1e}", flush=True)
Kn.set_fp32_tensor_core(True)
