"""Polynomial-exp2 offload in the tcgen05 attention kernels: time forward / backward for 0, 1 or 2 of every 4 exponentials on the
FMA pipe (nfb_flash_attn_set_poly), and the deviation of the outputs / gradients from the all-MUFU result."""
import sys, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R, _lib
from nfb200.array import B200Array, bfloat16
be = nfb200.get_backend("b200")
rng = np.random.default_rng(0)
def bf(*shape): return Kn.to_bf16(B200Array.from_numpy(rng.standard_normal(shape).astype(np.float32)))
for B, H, S, causal in [(8, 12, 512, True), (32, 12, 512, False), (64, 12, 197, False), (8, 16, 1024, True)]:
    D = 64
    sets = []
    for _ in range(3):
        qkv = bf(B, S, 3, H, D)
        sets.append((qkv[:, :, 0].transpose(0, 2, 1, 3), qkv[:, :, 1].transpose(0, 2, 1, 3), qkv[:, :, 2].transpose(0, 2, 1, 3), bf(B, S, H, D).transpose(0, 2, 1, 3)))
    flop_fwd = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
    base = None
    for pf, pb in ((0, 0), (1, 1), (2, 2), (1, 0), (2, 0), (0, 1), (0, 2)):
        _lib.call("nfb_flash_attn_set_poly", pf, pb)
        outs = [Kn.attention_forward(q, k, v, 0.125, causal=causal) for q, k, v, _ in sets]
        def fwd():
            for q, k, v, _ in sets: Kn.attention_forward(q, k, v, 0.125, causal=causal)
        def bwd():
            for (q, k, v, go), (o, lse) in zip(sets, outs): Kn.attention_backward(go, q, k, v, o, lse, 0.125, causal=causal, packed=True)
        res = []
        for name, fn, fl in (("fwd", fwd, flop_fwd), ("bwd", bwd, 2.0 * flop_fwd)):
            fn(); R.synchronize()
            g = R.Graph()
            with g: fn()
            R.synchronize()
            ms = R.time_ms(g.replay, iters=10, warmup=2) / len(sets)
            res.append(f"{name} {ms*1e3:7.1f} us {fl/ms/1e9:6.0f} TF/s")
        q, k, v, go = sets[0]
        o, lse = outs[0]
        dq, dk, dv, _ = Kn.attention_backward(go, q, k, v, o, lse, 0.125, causal=causal, packed=True)
        cur = [x.get().astype(np.float64) for x in (o, dq, dk, dv)]
        if base is None:
            base = cur
            dev = ""
        else:
            dev = "  max |d| / max |ref| vs all-MUFU: " + " ".join(f"{n} {np.abs(a - b).max() / np.abs(b).max():.1e}" for n, a, b in zip(("o", "dq", "dk", "dv"), cur, base))
        print(f"B={B} H={H} S={S} causal={causal} poly fwd/bwd {pf}/{pb}: " + " | ".join(res) + dev, flush=True)
_lib.call("nfb_flash_attn_set_poly", 0, 0)
