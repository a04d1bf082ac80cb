"""Full-size training-step timing of the other BASELINE.json configs on one B200 (graph-captured, device-timed):
BERT-base MLM B32xS512, ViT-B/16 B64 (224x224), GPT-2 medium B8xS1024 (the per-GPU shard of the 8-GPU config)."""
import sys, time, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as K, runtime as R, _lib
from nfb200.core import Tensor, set_default_device
from nfb200.optim import AdamW
from nfb200.training import GraphedStep
_lib.init(0); set_default_device("cuda:0"); K.set_precision("bf16"); K.set_wgrad_overlap(True)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(1234)

def run(name, build, make_batch, units, unit_name, steps=10):
    np.random.seed(0)
    model = build()
    opt = AdamW(model.parameters(), lr=2e-5, weight_decay=0.01, moment_dtype="float32")
    args, kwargs = make_batch()
    def step():
        opt.zero_grad()
        loss = model(*args, **kwargs)["loss"]
        loss.backward(); opt.step()
        return loss
    for _ in range(3): l = step()
    R.synchronize()
    l0 = float(l.item())
    gs = GraphedStep(step, warmup=1, optimizers=[opt])
    for _ in range(2): gs()
    R.synchronize()
    a, b = R.Event(), R.Event()
    a.record()
    for _ in range(steps): l = gs()
    b.record(); b.synchronize()
    ms = a.elapsed_ms(b) / steps
    st = R.mem_stats()
    print(f"{name}: {ms:.2f} ms/step  {units/ms*1e3:,.0f} {unit_name}/s  loss {l0:.3f}->{float(l.item()):.3f}  peak HBM {st['peak_bytes']/2**30:.1f} GiB", flush=True)

if which in ("all", "bert"):
    from nfb200.models import BERTForMaskedLM, BERTConfig
    def mk():
        ids = rng.integers(0, 30522, (32, 512))
        return (Tensor(ids, device="cuda"),), dict(attention_mask=Tensor(np.ones((32, 512), np.float32), device="cuda"), labels=Tensor(ids, device="cuda"))
    run("BERT-base MLM B32xS512", lambda: BERTForMaskedLM(BERTConfig()), mk, 32 * 512, "tokens")
if which in ("all", "vit"):
    from nfb200.models import vit_b_16
    def mk():
        img = rng.standard_normal((64, 3, 224, 224)).astype(np.float32)
        lab = rng.integers(0, 1000, (64,))
        return (Tensor(img, device="cuda"),), dict(labels=Tensor(lab, device="cuda"))
    run("ViT-B/16 B64 224x224", lambda: vit_b_16(num_classes=1000), mk, 64, "images")
if which in ("all", "gpt2m"):
    from nfb200.models import gpt2_medium
    def mk():
        ids = rng.integers(0, 50257, (8, 1024))
        lab = np.concatenate([ids[:, 1:], np.zeros((8, 1), dtype=ids.dtype)], axis=1)
        return (Tensor(ids, device="cuda"),), dict(labels=Tensor(lab, device="cuda"))
    run("GPT-2 medium B8xS1024 (per-GPU shard of config 5)", gpt2_medium, mk, 8 * 1024, "tokens")
if which in ("all", "zoo", "prenorm"):
    from nfb200.models import PreNormTransformer
    def mk():
        ids = rng.integers(0, 30000, (8, 512))
        return (Tensor(ids, device="cuda"),), dict(labels=Tensor(ids, device="cuda"))
    run("PreNormTransformer base (d768 L12 H12, swiglu+rmsnorm, interleaved RoPE) B8xS512",
        lambda: PreNormTransformer(d_model=768, num_layers=12, num_heads=12, d_ff=3072, max_seq_len=512, vocab_size=30000, activation="swiglu",
                                   normalization="rmsnorm"), mk, 8 * 512, "tokens")
if which in ("all", "zoo", "roberta"):
    from nfb200.models import RoBERTaForMaskedLM
    def mk():
        ids = rng.integers(3, 50265, (32, 512))
        return (Tensor(ids, device="cuda"),), dict(attention_mask=Tensor(np.ones((32, 512), np.float32), device="cuda"), labels=Tensor(ids, device="cuda"))
    run("RoBERTa-base MLM B32xS512", lambda: RoBERTaForMaskedLM(), mk, 32 * 512, "tokens")
if which in ("all", "zoo", "clip"):
    from nfb200.models import clip_base
    def mk():
        img = rng.standard_normal((64, 3, 224, 224)).astype(np.float32)
        txt = rng.integers(0, 49408, (64, 77))
        return (Tensor(img, device="cuda"), Tensor(txt, device="cuda")), {}
    run("CLIP base (ViT-B/32 + 12-layer text tower) B64", clip_base, mk, 64, "pairs")
if which in ("all", "zoo", "diff"):
    from nfb200.nn import DifferentialTransformer
    from nfb200 import functional as Fn
    class _LM:
        def __init__(self): self.net = DifferentialTransformer(n_layers=12, d_model=768, n_heads=12, d_ff=2048, vocab_size=50257, max_seq_len=512)
        def parameters(self): return self.net.parameters()
        def __call__(self, ids, labels):
            lg = self.net(ids)
            B, S, V = lg.shape
            return {"loss": Fn.cross_entropy_loss(Fn.reshape(lg, (B * S, V)), labels.data.reshape(-1))}
    def mk():
        ids = rng.integers(0, 50257, (8, 512))
        return (Tensor(ids, device="cuda"),), dict(labels=Tensor(ids, device="cuda"))
    run("DifferentialTransformer (d768 L12 H12) B8xS512", _LM, mk, 8 * 512, "tokens")
