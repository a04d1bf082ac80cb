"""Attention microbench at the GPT-2 small shape (B=8, H=12, S=512, D=64, packed qkv, causal): graph-captured, GPU-bound."""
import sys, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R, _lib
from nfb200.array import B200Array, bfloat16
be = nfb200.get_backend("b200")
rng = np.random.default_rng(0)
def bf(*shape): return Kn.to_bf16(B200Array.from_numpy(rng.standard_normal(shape).astype(np.float32)))
cfgs = [(8, 12, 512, True), (32, 12, 512, False), (64, 12, 197, False), (8, 16, 1024, True)]
for B, H, S, causal in cfgs:
    D = 64
    sets = []
    for _ in range(3):
        qkv = bf(B, S, 3, H, D)
        q = qkv[:, :, 0].transpose(0, 2, 1, 3); k = qkv[:, :, 1].transpose(0, 2, 1, 3); v = qkv[:, :, 2].transpose(0, 2, 1, 3)
        go = bf(B, S, H, D).transpose(0, 2, 1, 3)
        sets.append((q, k, v, go))
    scale = 0.125
    flop_fwd = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
    for impl in (0, 3):
        _lib.call("nfb_flash_attn_set_tc", impl)
        outs = [Kn.attention_forward(q, k, v, scale, causal=causal) for q, k, v, _ in sets]
        def fwd():
            for q, k, v, _ in sets: Kn.attention_forward(q, k, v, scale, causal=causal)
        def bwd():
            for (q, k, v, go), (o, lse) in zip(sets, outs): Kn.attention_backward(go, q, k, v, o, lse, scale, causal=causal, packed=True)
        res = []
        for name, fn, fl in (("fwd", fwd, flop_fwd), ("bwd", bwd, 2.5 * flop_fwd)):
            fn(); R.synchronize()
            g = R.Graph()
            with g: fn()
            R.synchronize()
            ms = R.time_ms(g.replay, iters=10, warmup=2) / len(sets)
            res.append(f"{name} {ms*1e3:7.1f} us {fl/ms/1e9:6.0f} TF/s")
        print(f"B={B} H={H} S={S} causal={causal} impl={'tcgen05' if impl else 'mma.sync'}: " + " | ".join(res), flush=True)
_lib.call("nfb_flash_attn_set_tc", 3)
# ---- grouped-query attention: K/V heads read in place; same algorithmic flops as MHA with H query heads
for B, H, Hkv, S, causal in [(8, 12, 12, 512, True), (8, 12, 4, 512, True), (8, 12, 1, 512, True), (8, 16, 4, 1024, True), (8, 32, 8, 2048, True)]:
    D = 64
    sets = []
    for _ in range(3):
        q = bf(B, S, H, D).transpose(0, 2, 1, 3); k = bf(B, S, Hkv, D).transpose(0, 2, 1, 3); v = bf(B, S, Hkv, D).transpose(0, 2, 1, 3)
        go = bf(B, S, H, D).transpose(0, 2, 1, 3)
        sets.append((q, k, v, go))
    flop_fwd = 4.0 * B * H * S * S * D * 0.5
    outs = [Kn.attention_forward(q, k, v, 0.125, causal=causal) for q, k, v, _ in sets]
    def fwd():
        for q, k, v, _ in sets: Kn.attention_forward(q, k, v, 0.125, causal=causal)
    def bwd():
        for (q, k, v, go), (o, lse) in zip(sets, outs): Kn.attention_backward(go, q, k, v, o, lse, 0.125, causal=causal)
    res = []
    for name, fn, fl in (("fwd", fwd, flop_fwd), ("bwd", bwd, 2.5 * flop_fwd)):
        fn(); R.synchronize()
        g = R.Graph()
        with g: fn()
        R.synchronize()
        ms = R.time_ms(g.replay, iters=10, warmup=2) / len(sets)
        res.append(f"{name} {ms*1e3:7.1f} us {fl/ms/1e9:6.0f} TF/s")
    print(f"GQA B={B} H={H} Hkv={Hkv} S={S} causal: " + " | ".join(res), flush=True)
