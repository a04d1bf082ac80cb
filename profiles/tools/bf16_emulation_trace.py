"""Where does the CUDA bf16 path leave the bf16-emulating oracle?  Runs the reference-architecture GPT-2 block by block on
the GPU and compares every intermediate with oracle/np_model.GPT2OracleBF16's (same seeded weights, same batch)."""
import sys, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as K, functional as F
from nfb200.array import bfloat16
from nfb200.core import Tensor, set_default_device
from nfb200.models import GPT2
from oracle.np_model import GPT2OracleBF16

cfg = {"vocab_size": 50257, "n_positions": 128, "n_embd": 256, "n_layer": 2, "n_head": 4}
rng = np.random.default_rng(1234)
ids = rng.integers(0, cfg["vocab_size"], (2, 128))
labels = np.concatenate([ids[:, 1:], np.zeros((2, 1), dtype=ids.dtype)], axis=1)
np.random.seed(0); orc = GPT2OracleBF16(cfg)
orc.forward(ids, labels)
caches = orc._cache[1]

def cmp(name, got, want):
    got = np.asarray(got.get() if hasattr(got, "get") else got, np.float64).reshape(-1)
    want = np.asarray(want, np.float64).reshape(-1)
    d = np.abs(got - want)
    print(f"{name:28s} max|d|/max|w| {d.max() / (np.abs(want).max() + 1e-30):.3e}  relL2 {np.linalg.norm(d) / (np.linalg.norm(want) + 1e-30):.3e}  differing {np.mean(d > 0):.4f}")

set_default_device("cuda"); K.set_precision("bf16")
np.random.seed(0); model = GPT2(cfg)
x = model.transformer.wte(Tensor(ids, device="cuda"))
B, S, d = x.shape
for i, blk in enumerate(model.transformer.h):
    c = caches[i]
    cmp(f"L{i} x in", x.data, c["x"])
    h1 = blk.ln_1(x, out_dtype=bfloat16); cmp(f"L{i} ln_1 out", h1.data, c["h1"])
    qkv = blk.attn.c_attn(h1, rope=(S, 2 * d, blk.attn.head_dim))     # q, k rotated in the projection's epilogue
    cmp(f"L{i} v part of qkv", qkv.data.reshape(B, S, 3, -1)[:, :, 2], c["qkv"].reshape(B, S, 3, -1)[:, :, 2])
    a2 = F.packed_qkv_attention(qkv, blk.attn.num_heads, causal=True, use_rope=True, scale=float(1.0 / np.sqrt(blk.attn.head_dim)), rope_applied=True)
    cmp(f"L{i} attention out", a2.data, c["a2"])
    x2 = blk.attn.c_proj(a2, residual=x); cmp(f"L{i} x2 (proj + res)", x2.data, c["x2"])
    h2 = blk.ln_2(x2, out_dtype=bfloat16); cmp(f"L{i} ln_2 out", h2.data, c["h2"])
    x = blk.mlp(h2, residual=x2); cmp(f"L{i} block out", x.data, c["out"])
hf = model.transformer.ln_f(x, out_dtype=bfloat16); cmp("ln_f out", hf.data, orc._dbg["hf"].reshape(B, S, d))
logits = model.lm_head(F.reshape(hf, (B * S, d)), out_dtype=bfloat16); cmp("logits", logits.data, orc._dbg["logits"])
