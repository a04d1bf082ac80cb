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
go = Kn.to_bf16(B200Array.from_numpy(rng.standard_normal((B, S, H, 64)).astype(np.float32))).transpose(0, 2, 1, 3)
o, lse = Kn.attention_forward(q, k, v, 0.125, causal=bool(causal))
for _ in range(2): Kn.attention_backward(go, q, k, v, o, lse, 0.125, causal=bool(causal), packed=True)
tlbuf = B200Array.full((256 + 3 * 256,), 0, np.int64)
_lib.call("nfb_flash_attn_set_timeline", tlbuf.ptr)
Kn.attention_backward(go, q, k, v, o, lse, 0.125, causal=bool(causal), packed=True); R.synchronize()
_lib.call("nfb_flash_attn_set_timeline", None)
t = tlbuf.get(); t0 = t[0]
print("loop end", t[1] - t0, "after reduce wait", t[2] - t0)
for it in range(8):
    s = t[16 + 8 * it: 16 + 8 * it + 7]
    if s[0] == 0: break
    print(f"it {it}: s_full {s[0]-t0} | P {s[1]-s[0]} | drain(prev) {s[2]-s[1]} | wait delta {s[5]-s[2]} | wait dP {s[3]-s[5]} | dS {s[4]-s[3]}")
