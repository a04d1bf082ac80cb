#!/usr/bin/env python
"""ORACLE INFRASTRUCTURE: generate golden input/output vectors by running the UNMODIFIED reference
(/root/reference/src/neural_arch: functional/, nn/, optim/, models/, distributed/) in this container, over the
reconstructed `core` (oracle/compat).  The reference cannot travel to the GPU box, so the vectors are
committed under tests/golden/ together with this script; tests/test_oracle_golden.py pins oracle/np_ops.py against
them on CPU and tests/test_gpu_golden.py pins the CUDA path against them on the B200.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz (needs /root/reference)

Every fixture stores the seeded inputs and what the reference computed from them (forward values and, where the
reference has a working backward, gradients).  Sizes are tiny on purpose (total < 1 MB).
"""
import os
import sys

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")  # the reference's Numba kernels then run as the plain Python/NumPy they wrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle", "compat"), ROOT, os.path.join(ROOT, "neural-network-from-scratch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _ref():
    import neural_arch  # noqa: F401  (the shim in oracle/compat + the reference tree)
    from neural_arch.optimization_config import configure
    configure(enable_jit=False, enable_fusion=False)  # pin the plain NumPy Linear path
    return neural_arch


def gen_functional(g):
    import neural_arch.functional as RF
    from neural_arch.core import Tensor
    rng = np.random.default_rng(100)
    a = rng.standard_normal((4, 6)).astype(np.float32)
    b = (rng.standard_normal((6,)) + 3).astype(np.float32)
    up = rng.standard_normal((4, 6)).astype(np.float32)
    for name, fn in (("add", RF.add), ("sub", RF.sub), ("mul", RF.mul), ("div", RF.div)):
        ta, tb = Tensor(a, requires_grad=True), Tensor(b, requires_grad=True)
        out = fn(ta, tb)
        out.backward(up)
        g[f"{name}_out"], g[f"{name}_da"], g[f"{name}_db"] = out.data, np.asarray(ta.grad), np.asarray(tb.grad)
    g["ab_a"], g["ab_b"], g["ab_up"] = a, b, up
    m1 = rng.standard_normal((3, 5, 4)).astype(np.float32)
    m2 = rng.standard_normal((3, 4, 6)).astype(np.float32)
    mu = rng.standard_normal((3, 5, 6)).astype(np.float32)
    t1, t2 = Tensor(m1, requires_grad=True), Tensor(m2, requires_grad=True)
    out = RF.matmul(t1, t2)
    out.backward(mu)
    g["mm_a"], g["mm_b"], g["mm_up"], g["mm_out"], g["mm_da"], g["mm_db"] = m1, m2, mu, out.data, np.asarray(t1.grad), np.asarray(t2.grad)
    x = (rng.standard_normal((5, 9)) * 2.5).astype(np.float32)
    x.flat[:4] = [0.0, 30.0, -30.0, 1e-3]
    ux = rng.standard_normal((5, 9)).astype(np.float32)
    g["act_x"], g["act_up"] = x, ux
    for name, fn in (("relu", RF.relu), ("sigmoid", RF.sigmoid), ("tanh", RF.tanh), ("gelu", RF.gelu), ("silu", RF.silu),
                     ("gelu_tanh", lambda t: RF.gelu(t, approximate=True)), ("softmax", lambda t: RF.softmax(t, axis=-1)),
                     ("softmax0", lambda t: RF.softmax(t, axis=0))):
        t = Tensor(x, requires_grad=True)
        out = fn(t)
        out.backward(ux)
        g[f"{name}_out"], g[f"{name}_dx"] = out.data, np.asarray(t.grad)
    z = (rng.standard_normal((7, 11)) * 3).astype(np.float32)
    tg = rng.integers(0, 11, 7)
    g["ce_z"], g["ce_t"] = z, tg
    for red in ("mean", "sum"):
        t = Tensor(z, requires_grad=True)
        loss = RF.cross_entropy_loss(t, Tensor(tg), reduction=red)
        loss.backward()
        g[f"ce_{red}_loss"], g[f"ce_{red}_dz"] = np.asarray(loss.data), np.asarray(t.grad)


def gen_layers(g):
    from neural_arch.core import Tensor
    from neural_arch.nn import Embedding, LayerNorm, Linear, RMSNorm
    from neural_arch.nn.linear import Linear as StandardLinear
    rng = np.random.default_rng(200)
    x = (rng.standard_normal((3, 5, 16)) * 2 + 0.5).astype(np.float32)
    up = rng.standard_normal((3, 5, 16)).astype(np.float32)
    g["norm_x"], g["norm_up"] = x, up
    w = rng.standard_normal(16).astype(np.float32)
    b = rng.standard_normal(16).astype(np.float32)
    g["norm_w"], g["norm_b"] = w, b
    for name, layer in (("ln", LayerNorm(16, eps=1e-5)), ("ln12", LayerNorm(16, eps=1e-12)), ("rms", RMSNorm(16, eps=1e-8))):
        layer.weight.data = w.copy()
        if getattr(layer, "bias", None) is not None:
            layer.bias.data = b.copy()
        t = Tensor(x, requires_grad=True)
        out = layer(t)
        out.backward(up)
        g[f"{name}_out"], g[f"{name}_dx"], g[f"{name}_dw"] = out.data, np.asarray(t.grad), np.asarray(layer.weight.grad)
        if getattr(layer, "bias", None) is not None:
            g[f"{name}_db"] = np.asarray(layer.bias.grad)
    # Linear: OptimizedLinear (out,in) standard path (2-D input so the reference backward is valid) + StandardLinear (in,out)
    np.random.seed(5)
    lin = Linear(16, 8)
    x2 = rng.standard_normal((6, 16)).astype(np.float32)
    u2 = rng.standard_normal((6, 8)).astype(np.float32)
    t = Tensor(x2, requires_grad=True)
    out = lin(t)
    out.backward(u2)
    g["lin_W"], g["lin_b"], g["lin_x"], g["lin_up"] = lin.weight.data, lin.bias.data, x2, u2
    g["lin_out"], g["lin_dx"], g["lin_dW"], g["lin_db"] = out.data, np.asarray(t.grad), np.asarray(lin.weight.grad), np.asarray(lin.bias.grad)
    np.random.seed(6)
    sl = StandardLinear(16, 8)
    t = Tensor(x2, requires_grad=True)
    out = sl(t)
    out.backward(u2)
    g["slin_W"], g["slin_b"] = sl.weight.data, sl.bias.data
    g["slin_out"], g["slin_dx"], g["slin_dW"], g["slin_db"] = out.data, np.asarray(t.grad), np.asarray(sl.weight.grad), np.asarray(sl.bias.grad)
    # fused Linear+GELU (tanh form): forward only via the reference's own fused NumPy/Numba routine
    from neural_arch.optimization.fusion import fuse_linear_activation
    g["lin_gelu_out"] = np.asarray(fuse_linear_activation(x2, lin.weight.data, lin.bias.data, "gelu"))
    # Embedding: gather + the sequential scatter-add backward, with duplicate indices
    np.random.seed(7)
    emb = Embedding(13, 8)
    idx = np.array([[0, 0, 1, 12], [5, 1, 0, 5]])
    ue = rng.standard_normal((2, 4, 8)).astype(np.float32)
    out = emb(Tensor(idx))
    out.backward(ue)
    out._backward()
    g["emb_W"], g["emb_idx"], g["emb_up"], g["emb_out"], g["emb_dW"] = emb.weight.data, idx, ue, out.data, np.asarray(emb.weight.grad)


def gen_optim(g):
    from neural_arch.core import Parameter
    from neural_arch.optim import SGD, Adam, AdamW
    rng = np.random.default_rng(300)
    p0 = rng.standard_normal((5, 7)).astype(np.float32)
    grads = [(rng.standard_normal((5, 7)) * (10.0 ** rng.integers(-3, 1))).astype(np.float32) for _ in range(6)]
    g["opt_p0"], g["opt_grads"] = p0, np.stack(grads)
    for name, make in (("adam", lambda p: Adam({"w": p}, lr=1e-3)), ("adam_wd", lambda p: Adam({"w": p}, lr=0.05, weight_decay=0.01)),
                       ("adamw", lambda p: AdamW({"w": p}, lr=5e-4, weight_decay=0.1)), ("adamw_default", lambda p: AdamW({"w": p})),
                       ("sgd", lambda p: SGD({"w": p}, lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=True))):
        p = Parameter(p0.copy())
        opt = make(p)
        traj = []
        for gr in grads:
            p.grad = gr.copy()
            opt.step()
            traj.append(np.asarray(p.data).copy())
        g[f"{name}_traj"] = np.stack(traj)
        if hasattr(opt, "state"):
            g[f"{name}_m"], g[f"{name}_v"] = opt.state["w"]["exp_avg"], opt.state["w"]["exp_avg_sq"]


def gen_gpt2_pieces(g):
    """Forward values of the GPT-2 building blocks exactly as models/language/gpt2.py computes them (its backward graph is
    severed -- SURVEY F6 -- so only forwards can be pinned at this level)."""
    from neural_arch.core import Tensor
    from neural_arch.models.language.gpt2 import GPT2, GPT2Attention, RMSNorm, RotaryEmbedding, SwiGLU, apply_rotary_pos_emb
    rng = np.random.default_rng(400)
    cfg = {"vocab_size": 61, "n_positions": 32, "n_embd": 32, "n_layer": 2, "n_head": 2}
    rot = RotaryEmbedding(16)
    q = rng.standard_normal((2, 2, 10, 16)).astype(np.float32)
    k = rng.standard_normal((2, 2, 10, 16)).astype(np.float32)
    cos, sin = rot(Tensor(q), seq_len=10)
    qr, kr = apply_rotary_pos_emb(Tensor(q), Tensor(k), cos, sin)
    g["rope_q"], g["rope_k"], g["rope_qr"], g["rope_kr"] = q, k, qr.data, kr.data
    np.random.seed(11)
    attn = GPT2Attention(cfg)
    v = rng.standard_normal((2, 2, 10, 16)).astype(np.float32)
    g["attn_v"], g["attn_out"] = v, attn._attn(qr, kr, Tensor(v)).data          # causal -1e4 replace, softmax, @ v
    x = rng.standard_normal((2, 10, 32)).astype(np.float32)
    g["blk_x"] = x
    a_out, _ = attn(Tensor(x))
    g["attn_Wqkv"], g["attn_bqkv"], g["attn_Wo"], g["attn_bo"] = attn.c_attn.weight.data, attn.c_attn.bias.data, attn.c_proj.weight.data, attn.c_proj.bias.data
    g["attn_block_out"] = a_out.data
    np.random.seed(12)
    sw = SwiGLU(32)
    g["sw_w1"], g["sw_w2"], g["sw_w3"], g["sw_out"] = sw.w1.weight.data, sw.w2.weight.data, sw.w3.weight.data, sw(Tensor(x)).data
    rn = RMSNorm(32, eps=1e-5)
    rn.weight.data = rng.standard_normal(32).astype(np.float32)
    g["grms_w"], g["grms_out"] = rn.weight.data, rn(Tensor(x)).data
    # whole model forward (logits) from a seed: pins construction order / init RNG calls / tying / block wiring
    np.random.seed(0)
    model = GPT2(cfg)
    ids = rng.integers(0, 61, (2, 12))
    g["gpt2_ids"] = ids
    g["gpt2_logits"] = model(Tensor(ids))["logits"].data
    g["gpt2_wte_sample"] = model.transformer.wte.weight.data[:4]
    g["gpt2_cattn0_sample"] = model.transformer.h[0].attn.c_attn.weight.data[:2]


def gen_bert_vit(g):
    """Seeded BERT-MLM and ViT forwards of the reference model zoo (tiny configs): pins init order, wiring, the additive
    attention mask, the fused tanh-GELU MLP and LayerNorm epsilons for the other two named model families."""
    from neural_arch.core import Tensor
    from neural_arch.optimization_config import configure
    configure(enable_jit=False, enable_fusion=True)   # BERT/ViT MLPs use the fused Linear+GELU path (tanh form)
    from neural_arch.models.language.bert import BERTConfig, BERTForMaskedLM
    from neural_arch.models.vision.vision_transformer import VisionTransformer
    rng = np.random.default_rng(500)
    np.random.seed(0)
    cfg = BERTConfig(vocab_size=97, hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64, max_position_embeddings=16)
    bert = BERTForMaskedLM(cfg)
    ids = rng.integers(0, 97, (2, 8))
    mask = np.ones((2, 8), dtype=np.float32)
    mask[1, 5:] = 0
    out = bert(Tensor(ids), attention_mask=Tensor(mask))
    g["bert_ids"], g["bert_mask"], g["bert_logits"] = ids, mask, out["logits"].data
    g["bert_word_emb_sample"] = bert.bert.embeddings.word_embeddings.weight.data[:3]
    g["bert_cls_w_sample"] = bert.cls.weight.data[:2]
    np.random.seed(0)
    vit = VisionTransformer(img_size=32, patch_size=8, num_classes=10, embed_dim=32, depth=2, num_heads=2)
    img = rng.standard_normal((2, 3, 32, 32)).astype(np.float32)
    g["vit_img"], g["vit_logits"] = img, vit(Tensor(img)).data
    g["vit_pos_sample"] = vit.pos_embed.data[0, :2]
    configure(enable_jit=False, enable_fusion=False)


def gen_sampler(g):
    from neural_arch.distributed.data_parallel import DistributedSampler
    for tag, kw in (("a", dict(dataset_size=10, num_replicas=4, rank=3, shuffle=False)), ("b", dict(dataset_size=23, num_replicas=3, rank=1, shuffle=True, seed=5)),
                    ("c", dict(dataset_size=7, num_replicas=2, rank=1, shuffle=True, seed=5, drop_last=True))):
        s = DistributedSampler(**kw)
        s.set_epoch(2 if tag == "b" else 0)
        g[f"sampler_{tag}"] = np.array(list(s))


SCHEDULE_CASES = [  # (tag, class or factory name, kwargs, initial lr, epochs)
    ("step", "StepLR", dict(step_size=3, gamma=0.5), 0.01, 12),
    ("step_resume", "StepLR", dict(step_size=4, gamma=0.1, last_epoch=5), 0.2, 8),
    ("exp", "ExponentialLR", dict(gamma=0.9), 5e-4, 10),
    ("cosine", "CosineAnnealingLR", dict(T_max=8, eta_min=1e-5), 1e-3, 20),
    ("linear", "LinearLR", dict(start_factor=0.25, end_factor=1.0, total_iters=6), 2e-3, 10),
    ("warmup", "WarmupLR", dict(warmup_epochs=5, warmup_factor=0.2), 3e-4, 9),
    ("poly", "PolynomialLR", dict(total_iters=9, power=2.0, end_lr=1e-6), 1e-3, 12),
    ("lin_warm", "get_linear_schedule_with_warmup", dict(num_warmup_steps=4, num_training_steps=16), 5e-4, 20),
    ("cos_warm", "get_cosine_schedule_with_warmup", dict(num_warmup_steps=3, num_training_steps=13), 1e-3, 16),
]
PLATEAU_CASES = [
    ("plateau_min", dict(mode="min", factor=0.5, patience=2, threshold=1e-2, cooldown=1, min_lr=1e-5), 1e-2,
     [1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]),
    ("plateau_max_abs", dict(mode="max", factor=0.1, patience=0, threshold=0.05, threshold_mode="abs"), 1.0,
     [0.1, 0.2, 0.22, 0.3, 0.3, 0.31, 0.5]),
]


def gen_train_glue(g):
    """Training-loop glue: learning-rate traces of every scheduler in optim/lr_scheduler.py driven through the
    reference AdamW, and step-by-step traces of optimization/grad_scaler.py's AdvancedGradScaler (scale, skipped steps, the
    unscaled + per-parameter-clipped gradients) over a gradient sequence with injected overflows."""
    import json
    from neural_arch.core import Parameter, Tensor
    from neural_arch.optim import AdamW
    from neural_arch.optim import lr_scheduler as S
    from neural_arch.optimization import grad_scaler as GS

    def opt_of(lr):
        return AdamW({"w": Parameter(np.ones(3, dtype=np.float32))}, lr=lr)
    for tag, name, kw, lr0, epochs in SCHEDULE_CASES:
        opt = opt_of(lr0)
        sch = getattr(S, name)(opt, **kw)
        trace = [opt.lr]                      # the constructor already stepped once
        for _ in range(epochs):
            sch.step()
            trace.append(opt.lr)
        g[f"sched_{tag}"] = np.asarray(trace, dtype=np.float64)
    for tag, kw, lr0, metrics in PLATEAU_CASES:
        opt = opt_of(lr0)
        sch = S.ReduceLROnPlateau(opt, **kw)
        trace = []
        for m in metrics:
            sch.step(m)
            trace.append(opt.lr)
        g[f"sched_{tag}"] = np.asarray(trace, dtype=np.float64)
    g["sched_cases"] = np.frombuffer(json.dumps({"sched": SCHEDULE_CASES, "plateau": PLATEAU_CASES}).encode(), dtype=np.uint8)

    # --- AdvancedGradScaler: the reference reads `param.grad.data`, i.e. expects Tensor-valued grads (its tests assign Tensors)
    class P:      # duck-typed parameter / optimizer: the scaler touches only .grad.data, .parameters and .step()
        def __init__(self, shape):
            self.grad, self.shape = None, shape

    class O:
        def __init__(self, params):
            self.parameters, self.steps = params, 0

        def step(self):
            self.steps += 1
    rng = np.random.default_rng(77)
    shapes = [(5, 7), (11,), (3, 4, 2), (9,)]
    nstep = 14
    base = [[(rng.standard_normal(sh) * 10.0 ** rng.integers(-3, 1)).astype(np.float32) for sh in shapes] for _ in range(nstep)]
    bad = {3: (1, np.inf), 4: (2, np.nan), 9: (0, -np.inf)}       # step -> (parameter index, value) injected overflow
    for tag, cfg in (("dyn", dict(init_scale=1024.0, growth_interval=3, gradient_clip_threshold=1.0)),
                     ("noclip", dict(init_scale=96.0, growth_interval=2, growth_factor=1.5, gradient_clip_threshold=0.0, max_loss_scale=200.0)),
                     ("fixed", dict(init_scale=8.0, strategy=GS.ScalingStrategy.FIXED, gradient_clip_threshold=0.5)),
                     ("conservative", dict(init_scale=4.0, strategy=GS.ScalingStrategy.CONSERVATIVE, consecutive_success_threshold=2,
                                           overflow_cooldown=3, gradient_clip_threshold=1.0, min_loss_scale=2.0))):
        sc = GS.AdvancedGradScaler(GS.ScalerConfig(**cfg))
        params = {f"p{i}": P(sh) for i, sh in enumerate(shapes)}
        opt = O(params)
        scales, taken, scaled_in, outs = [], [], [], []
        for t in range(nstep):
            s_now = sc.get_scale()
            ins = []
            for i, p in enumerate(params.values()):
                gr = (base[t][i] * np.float32(s_now)).astype(np.float32)   # what backward of the scaled loss would have produced
                if t in bad and bad[t][0] == i:
                    gr.flat[1] = bad[t][1]
                ins.append(gr.copy())
                p.grad = Tensor(gr)
            ok = sc.step(opt)
            scales.append(sc.get_scale())
            taken.append(bool(ok))
            scaled_in.append(np.concatenate([a.ravel() for a in ins]))
            outs.append(np.concatenate([np.asarray(p.grad.data, dtype=np.float32).ravel() for p in params.values()]))
        g[f"scaler_{tag}_scale"], g[f"scaler_{tag}_taken"] = np.asarray(scales, dtype=np.float64), np.asarray(taken)
        g[f"scaler_{tag}_in"], g[f"scaler_{tag}_out"] = np.stack(scaled_in), np.stack(outs)
        g[f"scaler_{tag}_optsteps"] = np.asarray(opt.steps)
        st = sc.get_statistics()
        g[f"scaler_{tag}_stats"] = np.asarray([st["total_overflows"], st["successful_steps"], st["growth_tracker"]], dtype=np.float64)
        g[f"scaler_{tag}_norm_hist"] = np.asarray(sc._gradient_norms_history, dtype=np.float64)
    g["scaler_shapes"] = np.frombuffer(json.dumps(shapes).encode(), dtype=np.uint8)
    # loss scaling itself: scale(loss) multiplies the value (grad_scaler.py:90-129)
    sc = GS.AdvancedGradScaler(GS.ScalerConfig(init_scale=512.0))
    g["scaler_scaled_loss"] = np.asarray(sc.scale(Tensor(np.asarray(0.37, dtype=np.float32), requires_grad=True)).data)
    # clip_gradients_by_norm / check_gradients_finite helpers (grad_scaler.py:501-597)
    params = {f"p{i}": P(sh) for i, sh in enumerate(shapes)}
    for i, p in enumerate(params.values()):
        p.grad = Tensor(base[0][i] * np.float32(3.0))
    opt = O(params)
    fin, stats = GS.check_gradients_finite(opt)
    g["helper_in"] = np.concatenate([np.asarray(p.grad.data).ravel() for p in params.values()])
    g["helper_stats"] = np.asarray([float(fin), stats["params_with_grad"], stats["finite_grads"], stats["max_grad_norm"], stats["avg_grad_norm"]])
    g["helper_total_norm"] = np.asarray(GS.clip_gradients_by_norm(opt, 0.75))
    g["helper_clipped"] = np.concatenate([np.asarray(p.grad.data, dtype=np.float32).ravel() for p in params.values()])


def gen_modern(g):
    """Zoo blocks on the same kernels: nn/modern_attention.py GroupedQueryAttention / MultiQueryAttention forwards
    (weights stored, so the fixtures do not depend on RNG call order) and nn/positional.py forwards (+ the learned table's
    gradients; the reference's GQA and RoPE backwards are placeholders / wrong -- PARITY.md #17, #18 -- and are not pinned)."""
    from neural_arch.core import Tensor
    from neural_arch.nn.modern_attention import GroupedQueryAttention, MultiQueryAttention
    from neural_arch.nn.positional import LearnedPositionalEmbedding, RotaryPositionalEmbedding, SinusoidalPositionalEncoding
    rng = np.random.default_rng(500)
    for tag, kw, B, S, mask_kind in (("gqa", dict(d_model=64, n_heads=4, n_kv_heads=2), 2, 9, None),
                                     ("gqa_bias_keymask", dict(d_model=96, n_heads=6, n_kv_heads=3, bias=True), 3, 7, "key"),
                                     ("gqa_full_mask", dict(d_model=64, n_heads=4, n_kv_heads=1), 2, 6, "causal"),
                                     ("mha", dict(d_model=32, n_heads=2), 1, 5, None)):
        np.random.seed(7)
        m = GroupedQueryAttention(**kw)
        for n in ("b_q", "b_k", "b_v", "b_o"):
            if getattr(m, n) is not None:
                getattr(m, n).data = (rng.standard_normal(getattr(m, n).shape) * 0.1).astype(np.float32)
        x = rng.standard_normal((B, S, kw["d_model"])).astype(np.float32)
        mask = None
        if mask_kind == "key":
            mask = np.where(rng.random((B, 1, 1, S)) < 0.3, -1e4, 0.0).astype(np.float32)
            mask[..., 0] = 0.0
        elif mask_kind == "causal":
            mask = np.triu(np.full((S, S), -1e9, dtype=np.float32), k=1)
        out = m(Tensor(x), None if mask is None else Tensor(mask))
        g[f"{tag}_x"], g[f"{tag}_out"] = x, np.asarray(out.data)
        if mask is not None:
            g[f"{tag}_mask"] = mask
        for n in ("W_q", "W_k", "W_v", "W_o", "b_q", "b_k", "b_v", "b_o"):
            if getattr(m, n) is not None:
                g[f"{tag}_{n}"] = np.asarray(getattr(m, n).data)
    np.random.seed(8)
    mq = MultiQueryAttention(d_model=64, n_heads=4)
    x = rng.standard_normal((2, 5, 64)).astype(np.float32)
    g["mqa_x"], g["mqa_out"] = x, np.asarray(mq(Tensor(x)).data)
    for n in ("W_q", "W_k", "W_v", "W_o"):
        g[f"mqa_{n}"] = np.asarray(getattr(mq.gqa, n).data)
    kv = rng.standard_normal((2, 3, 2, 4)).astype(np.float32)
    g["repeat_in"], g["repeat_out"] = kv, GroupedQueryAttention(d_model=32, n_heads=8, n_kv_heads=2).repeat_kv(kv)
    # positional
    x = rng.standard_normal((2, 7, 16)).astype(np.float32)
    g["pos_x"] = x
    g["sin_out"] = np.asarray(SinusoidalPositionalEncoding(16, max_len=64)(Tensor(x), start_pos=3).data)
    g["sin_out_base"] = np.asarray(SinusoidalPositionalEncoding(16, max_len=32, base=500.0)(Tensor(x)).data)
    q4, k4 = rng.standard_normal((2, 3, 7, 16)).astype(np.float32), rng.standard_normal((2, 3, 7, 16)).astype(np.float32)
    rope = RotaryPositionalEmbedding(16, max_seq_len=32)
    rq, rk = rope(Tensor(q4), Tensor(k4), start_pos=5)
    g["rope_q"], g["rope_k"], g["rope_q_out"], g["rope_k_out"] = q4, k4, np.asarray(rq.data), np.asarray(rk.data)
    r3q, _ = RotaryPositionalEmbedding(16, max_seq_len=8, base=100.0)(Tensor(x), Tensor(x), start_pos=4)   # grows its tables
    g["rope3_out"] = np.asarray(r3q.data)
    np.random.seed(9)
    lpe = LearnedPositionalEmbedding(12, 16)
    xt = Tensor(x, requires_grad=True)
    out = lpe(xt, start_pos=2)
    up = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(up)
    g["lpe_table"], g["lpe_out"], g["lpe_up"] = np.asarray(lpe.embedding.data), np.asarray(out.data), up
    g["lpe_dx"], g["lpe_dtable"] = np.asarray(xt.grad), np.asarray(lpe.embedding.grad)
    # gated feed-forward blocks (nn/modern_activations.py): the reference's SwiGLU has a complete backward -> gradients pinned too
    from neural_arch.nn.modern_activations import GeGLU, SwiGLU, Swish
    np.random.seed(21)
    ff = SwiGLU(24, 40)
    for n in ("b_gate", "b_up", "b_down"):
        getattr(ff, n).data = (rng.standard_normal(getattr(ff, n).shape) * 0.1).astype(np.float32)
    x = rng.standard_normal((2, 5, 24)).astype(np.float32)
    up = (rng.standard_normal((2, 5, 24)) * 0.1).astype(np.float32)
    xt = Tensor(x, requires_grad=True)
    out = ff(xt)
    out.backward(up)
    g["swiglu_x"], g["swiglu_up"], g["swiglu_out"], g["swiglu_dx"] = x, up, np.asarray(out.data), np.asarray(xt.grad)
    for n in ("W_gate", "W_up", "W_down", "b_gate", "b_up", "b_down"):
        g[f"swiglu_{n}"], g[f"swiglu_d{n}"] = np.asarray(getattr(ff, n).data), np.asarray(getattr(ff, n).grad)
    g["swiglu_default_hidden"] = np.asarray(SwiGLU(30).hidden_dim)
    np.random.seed(22)
    ge = GeGLU(24, 40, bias=False)
    g["geglu_out"] = np.asarray(ge(Tensor(x)).data)
    for n in ("W_gate", "W_up", "W_down"):
        g[f"geglu_{n}"] = np.asarray(getattr(ge, n).data)
    xt = Tensor(x, requires_grad=True)
    sw = Swish()(xt)
    sw.backward(up)
    g["swish_out"], g["swish_dx"] = np.asarray(sw.data), np.asarray(xt.grad)
    # differential attention (nn/differential_attention.py): forwards only (its backward is a placeholder)
    from neural_arch.nn.differential_attention import DifferentialAttention, DifferentialTransformer, DifferentialTransformerBlock
    for tag, kw, B, S, mask_kind in (("diff", dict(d_model=64, n_heads=4), 2, 9, None),
                                     ("diff_bias_keymask", dict(d_model=48, n_heads=6, bias=True, lambda_init=0.8), 2, 7, "key"),
                                     ("diff_full_mask", dict(d_model=32, n_heads=2, lambda_init=0.3), 1, 6, "causal")):
        np.random.seed(23)
        m = DifferentialAttention(**kw)
        m.lambda_param.data = (m.lambda_param.data + rng.standard_normal(m.lambda_param.shape).astype(np.float32) * 0.1).astype(np.float32)
        names = ["lambda_param", "W_q1", "W_k1", "W_v1", "W_q2", "W_k2", "W_v2", "W_o"]
        if kw.get("bias"):
            names += ["b_q1", "b_k1", "b_v1", "b_q2", "b_k2", "b_v2", "b_o"]
            for n in names[8:]:
                getattr(m, n).data = (rng.standard_normal(getattr(m, n).shape) * 0.1).astype(np.float32)
        x = rng.standard_normal((B, S, kw["d_model"])).astype(np.float32)
        mask = None
        if mask_kind == "key":
            mask = np.where(rng.random((B, 1, 1, S)) < 0.3, -1e4, 0.0).astype(np.float32)
            mask[..., 0] = 0.0
        elif mask_kind == "causal":
            mask = np.triu(np.full((S, S), -1e9, dtype=np.float32), k=1)
        out, (a1, a2) = m(Tensor(x), None if mask is None else Tensor(mask), return_attention_weights=True)
        g[f"{tag}_x"], g[f"{tag}_out"], g[f"{tag}_a1"], g[f"{tag}_a2"] = x, np.asarray(out.data), a1, a2
        if mask is not None:
            g[f"{tag}_mask"] = mask
        for n in names:
            g[f"{tag}_{n}"] = np.asarray(getattr(m, n).data)
    np.random.seed(24)
    blk = DifferentialTransformerBlock(d_model=32, n_heads=4, d_ff=48)
    x = rng.standard_normal((2, 6, 32)).astype(np.float32)
    g["diffblock_x"], g["diffblock_out"] = x, np.asarray(blk(Tensor(x)).data)
    np.random.seed(25)
    net = DifferentialTransformer(n_layers=2, d_model=32, n_heads=4, d_ff=48, vocab_size=50, max_seq_len=16)
    ids = rng.integers(0, 50, (2, 7))
    g["diffnet_ids"], g["diffnet_logits"] = ids, np.asarray(net(Tensor(ids)).data)


ROBERTA_TINY = dict(vocab_size=97, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64, max_position_embeddings=20)


def gen_roberta(g):
    """Seeded forwards of models/language/roberta.py (tiny config): masked-LM logits + loss with a padding mask, the sequence
    classifier's logits + loss; pins the position-id offset, the one-row token-type table and the init quirk (LayerNorm gains
    redrawn from N(0, 0.02^2))."""
    from neural_arch.core import Tensor
    from neural_arch.models.language.roberta import RoBERTaForMaskedLM, RoBERTaForSequenceClassification
    from neural_arch.optimization_config import configure
    configure(enable_jit=False, enable_fusion=False)
    rng = np.random.default_rng(600)
    ids = rng.integers(3, 97, (3, 11))
    mask = np.ones((3, 11), dtype=np.float32)
    mask[0, 8:] = 0
    mask[2, 5:] = 0
    ids[mask == 0] = 1
    labels = rng.integers(0, 97, (3, 11))
    np.random.seed(31)
    m = RoBERTaForMaskedLM(**ROBERTA_TINY)
    out = m(Tensor(ids), attention_mask=Tensor(mask), labels=Tensor(labels))
    g["ids"], g["mask"], g["labels"] = ids, mask, labels
    g["mlm_logits"], g["mlm_loss"] = np.asarray(out["logits"].data), np.asarray(out["loss"].data)
    g["ln_gain_first"] = np.asarray(m.roberta.embeddings.LayerNorm.weight.data)
    np.random.seed(32)
    c = RoBERTaForSequenceClassification(num_labels=3, **ROBERTA_TINY)
    cl = rng.integers(0, 3, (3,))
    out = c(Tensor(ids), attention_mask=Tensor(mask), labels=Tensor(cl))
    g["cls_labels"], g["cls_logits"], g["cls_loss"] = cl, np.asarray(out["logits"].data), np.asarray(out["loss"].data)


CLIP_TINY = dict(embed_dim=24, image_resolution=16, vision_layers=2, vision_width=64, vision_patch_size=8, context_length=9, vocab_size=50,
                 transformer_width=32, transformer_heads=4, transformer_layers=2)


def gen_clip(g):
    """Seeded forward of models/multimodal/clip.py (tiny config): normalised image / text embeddings, logits and the symmetric
    contrastive loss."""
    from neural_arch.core import Tensor
    from neural_arch.models.multimodal.clip import CLIP
    from neural_arch.optimization_config import configure
    configure(enable_jit=False, enable_fusion=False)
    rng = np.random.default_rng(700)
    img = rng.standard_normal((3, 3, 16, 16)).astype(np.float32)
    txt = rng.integers(0, 50, (3, 9))
    np.random.seed(41)
    m = CLIP(**CLIP_TINY)
    out = m(Tensor(img), Tensor(txt))
    g["img"], g["txt"] = img, txt
    g["image_embeds"], g["text_embeds"] = np.asarray(out["image_embeds"].data), np.asarray(out["text_embeds"].data)
    g["logits_per_image"], g["loss"] = np.asarray(out["logits_per_image"].data), np.asarray(out["loss"].data)
    g["pos_emb_sample"] = np.asarray(m.transformer.positional_embedding.data)[:2]


PRENORM_CASES = {
    "gelu_ln": dict(d_model=32, num_layers=2, num_heads=4, d_ff=64, max_seq_len=16, vocab_size=60),
    "swiglu_rms_untied": dict(d_model=32, num_layers=2, num_heads=2, d_ff=48, max_seq_len=16, vocab_size=60, activation="swiglu",
                              normalization="rmsnorm", tie_embeddings=False, scale_embeddings=False, rope_base=500.0),
    "relu_norope": dict(d_model=24, num_layers=1, num_heads=3, d_ff=40, max_seq_len=16, vocab_size=60, activation="relu", use_rope=False),
}


def gen_prenorm(g):
    """Seeded forwards of models/language/modern_transformer.py (tiny configs covering gelu/swiglu/relu, LayerNorm/RMSNorm, tied and
    untied heads, RoPE on/off), with and without a padding mask."""
    from neural_arch.core import Tensor
    from neural_arch.models.language.modern_transformer import PreNormTransformer
    from neural_arch.optimization_config import configure
    configure(enable_jit=False, enable_fusion=False)
    rng = np.random.default_rng(800)
    ids = rng.integers(0, 60, (3, 10))
    mask = np.ones((3, 10), dtype=np.int64)
    mask[0, 7:] = 0
    mask[2, 4:] = 0
    g["ids"], g["mask"] = ids, mask
    for i, (tag, kw) in enumerate(PRENORM_CASES.items()):
        np.random.seed(50 + i)
        m = PreNormTransformer(**kw)
        g[f"{tag}_logits"] = np.asarray(m(Tensor(ids))["logits"].data)
        g[f"{tag}_logits_masked"] = np.asarray(m(Tensor(ids), attention_mask=Tensor(mask))["logits"].data)
        g[f"{tag}_logits_pos3"] = np.asarray(m(Tensor(ids[:, :5]), start_pos=3)["logits"].data)


def gen_lion(g):
    """optim/lion.py trajectories: plain, with weight decay + maximize, and a gradient sequence containing exact zeros and a NaN."""
    from neural_arch.core import Parameter
    from neural_arch.optim.lion import Lion
    rng = np.random.default_rng(900)
    p0 = rng.standard_normal((6, 9)).astype(np.float32)
    grads = [(rng.standard_normal((6, 9)) * (10.0 ** rng.integers(-3, 1))).astype(np.float32) for _ in range(7)]
    grads[2][0, :3] = 0.0
    grads[0][1, 1] = 0.0            # c = 0 on the very first step -> sign 0 -> no update of that element
    grads[4][2, 2] = np.nan
    g["p0"], g["grads"] = p0, np.stack(grads)
    for name, kw in (("plain", dict(lr=1e-2)), ("wd_max", dict(lr=3e-3, weight_decay=0.1, maximize=True, betas=(0.8, 0.95))), ("biglr", dict(lr=2.5, beta1=0.5))):
        p = Parameter(p0.copy())
        opt = Lion({"w": p}, **kw)
        traj = []
        for gr in grads:
            p.grad = gr.copy()
            opt.step()
            traj.append(np.asarray(p.data).copy())
        g[f"{name}_traj"], g[f"{name}_m"] = np.stack(traj), np.asarray(opt.state["w"]["momentum_buffer"])


def gen_translation(g):
    """BASELINE configs[0]: the encoder-decoder Transformer of examples/translation/model_v2.py (the reference's own
    CPU-runnable case), seeded, at the config's width (d_model 128, 4 heads, 3 + 3 layers, d_ff 256) with a small vocabulary:
    forward logits for a padded synthetic batch (ids uniform in [4, V), trailing pad 0 of random length), with and
    without the look-ahead / padding masks the training script passes."""
    import importlib.util
    from neural_arch.core import Tensor
    spec = importlib.util.spec_from_file_location("ref_translation_model_v2", "/root/reference/examples/translation/model_v2.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(900)
    V, B, S = 120, 4, 12
    np.random.seed(0)
    model = mod.TranslationTransformer(src_vocab_size=V, tgt_vocab_size=V + 7, d_model=128, n_heads=4, n_layers=3, d_ff=256, max_seq_len=32)
    src = rng.integers(4, V, (B, S))
    tgt = rng.integers(4, V + 7, (B, S))
    for r in range(B):
        pad = int(rng.integers(0, 6))
        if pad:
            src[r, S - pad:] = 0
            tgt[r, S - pad:] = 0
    g["src"], g["tgt"] = src, tgt
    g["logits"] = model(Tensor(src), Tensor(tgt)).data
    src_mask = model.create_padding_mask(src)
    tgt_mask = model.create_look_ahead_mask(S)
    g["src_mask"], g["tgt_mask"] = (src_mask if src_mask is not None else np.zeros((B, S), np.float32)), tgt_mask
    g["logits_masked"] = model(Tensor(src), Tensor(tgt), src_mask=src_mask, tgt_mask=tgt_mask).data
    g["src_emb_sample"] = model.src_embedding.weight.data[:3]
    g["out_w_sample"] = model.output_projection.weight.data[:2]
    g["n_params"] = np.array(sum(int(np.prod(p.data.shape)) for p in model.parameters()))


def main():
    _ref()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, fn in (("functional", gen_functional), ("layers", gen_layers), ("optim", gen_optim), ("gpt2", gen_gpt2_pieces), ("bert_vit", gen_bert_vit),
                     ("sampler", gen_sampler), ("train_glue", gen_train_glue), ("modern", gen_modern), ("roberta", gen_roberta), ("clip", gen_clip), ("prenorm", gen_prenorm), ("lion", gen_lion), ("translation", gen_translation)):
        if only and name not in only:
            continue
        g = {}
        fn(g)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **{k: np.asarray(v) for k, v in g.items()})
        print(f"{name}: {len(g)} arrays")


if __name__ == "__main__":
    main()
