import sys, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R, _lib
from nfb200.array import B200Array, bfloat16
be = nfb200.get_backend("b200")
B, H, S, causal = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rng = np.random.default_rng(0)
qkv = Kn.to_bf16(B200Array.from_numpy(rng.standard_normal((B, S, 3, H, 64)).astype(np.float32)))
q = qkv[:, :, 0].transpose(0, 2, 1, 3); k = qkv[:, :, 1].transpose(0, 2, 1, 3); v = qkv[:, :, 2].transpose(0, 2, 1, 3)
for _ in range(3): Kn.attention_forward(q, k, v, 0.125, causal=bool(causal))
tlb = B200Array.full((256 + 3 * 256,), 0, np.int64)
_lib.call("nfb_flash_attn_set_timeline", tlb.ptr)
Kn.attention_forward(q, k, v, 0.125, causal=bool(causal)); R.synchronize()
_lib.call("nfb_flash_attn_set_timeline", None)
t = tlb.get(); t0 = t[0]
print("end", t[1] - t0)
for i in range(8):
    m = t[16 + 4 * i: 20 + 4 * i]
    if m[3] == 0: break
    print(f"mma blk {i}: p0_ready {m[0]-t0} | mid {m[1]-t0} | p1_ready {m[2]-t0} | issued {m[3]-t0}")
for g in (0, 1):
    for i in range(8):
        s = t[64 + g * 64 + 8 * i: 64 + g * 64 + 8 * i + 5]
        if s[4] == 0: break
        print(f"softmax g{g} blk {i}: s_full {s[0]-t0} | ld {s[1]-s[0]} | max {s[2]-s[1]} | exp+sts {s[3]-s[2]} | arrive {s[4]-s[3]}")
c = t[256:256 + 3 * 192].reshape(-1, 3)
base = c[:, 0].min()
print("CTA schedule (us from first start; start = behind the dependency wait): cta start end duration sm")
order = np.argsort(c[:, 0])
for idx in list(order[:4]) + list(order[96:100]) + list(order[-50::6]):
    print(idx, (c[idx, 0] - base) / 1e3, (c[idx, 1] - base) / 1e3, round((c[idx, 1] - c[idx, 0]) / 1e3, 2), c[idx, 2])
dur = (c[:, 1] - c[:, 0]) / 1e3
print("durations us: heavy CTAs (0..95) mean %.2f min %.2f max %.2f | light, first wave (96..147) mean %.2f | light, second wave (148..191) mean %.2f" % (
    dur[:96].mean(), dur[:96].min(), dur[:96].max(), dur[96:148].mean(), dur[148:192].mean()))
print("last end us:", (c[:, 1].max() - base) / 1e3, " distinct SMs:", len(set(c[:, 2])))
