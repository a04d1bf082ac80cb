"""CPU ORACLE (test infrastructure, NOT product code): the reference's GPT-2 / training step restated in NumPy
with a CONNECTED graph (decision 2): every per-op forward/backward formula is the reference's own
(oracle/np_ops.py cites the lines), composed with the chain rule restored where the reference re-wraps
intermediates in fresh leaf tensors (models/language/gpt2.py:232-234,260,321-335,474-489).

Weights are drawn with the reference's exact NumPy RNG calls in the reference's construction order
(models/language/gpt2.py:143-373, nn/optimized.py:104-134, nn/embedding.py:19-24), so
``np.random.seed(s); GPT2Oracle(cfg)`` reproduces ``np.random.seed(s); neural_arch...GPT2(cfg)`` bit for bit
(checked by oracle/validate_against_reference.py when /root/reference is present).
"""
from __future__ import annotations

import numpy as np

from . import np_ops as O


def _he_uniform(out_f, in_f):
    lim = np.sqrt(6.0 / in_f)
    return np.random.uniform(-lim, lim, (out_f, in_f)).astype(np.float32)


class GPT2Oracle:
    def __init__(self, config, grad_clip=True, dtype=np.float32):
        """dtype=np.float64 runs the same formulas in double precision on the same (fp32-drawn) weights: the yardstick that
        tells how far TWO fp32 evaluation orders of this model may legitimately be apart (tests/test_gpu_train_parity.py)."""
        self.dtype = np.dtype(dtype)
        self.cfg = dict(config)
        d, V, L = config["n_embd"], config["vocab_size"], config["n_layer"]
        self.H = config["n_head"]
        self.hd = d // self.H
        self.eps = config.get("layer_norm_epsilon", 1e-5)
        self.grad_clip = grad_clip
        P = {}
        # construction order of GPT2LMHead.__init__ -> GPT2Model.__init__ (wte, then blocks, ln_f, _init_weights), then lm_head
        scale = 1.0 / np.sqrt(d)
        P["transformer.wte.weight"] = np.random.uniform(-scale, scale, (V, d)).astype(np.float32)
        for i in range(L):
            p = f"transformer.h.{i}."
            P[p + "ln_1.weight"] = np.ones(d, np.float32)
            P[p + "attn.c_attn.weight"] = _he_uniform(3 * d, d)
            P[p + "attn.c_attn.bias"] = np.zeros(3 * d, np.float32)
            P[p + "attn.c_proj.weight"] = _he_uniform(d, d)
            P[p + "attn.c_proj.bias"] = np.zeros(d, np.float32)
            P[p + "ln_2.weight"] = np.ones(d, np.float32)
            P[p + "mlp.swiglu.w1.weight"] = _he_uniform(2 * d, d)
            P[p + "mlp.swiglu.w2.weight"] = _he_uniform(2 * d, d)
            P[p + "mlp.swiglu.w3.weight"] = _he_uniform(d, 2 * d)
        P["transformer.ln_f.weight"] = np.ones(d, np.float32)
        P["transformer.wte.weight"] = np.random.randn(V, d).astype(np.float32) * 0.02      # _init_weights
        _ = _he_uniform(V, d)  # lm_head = Linear(d, V, bias=False) draws its own weight before being tied away
        self.params = P if self.dtype == np.float32 else {k: v.astype(self.dtype) for k, v in P.items()}
        self.grads = {}

    # tied: lm_head.weight IS transformer.wte.weight
    def forward(self, input_ids, labels=None):
        P, H, hd = self.params, self.H, self.hd
        B, S = input_ids.shape
        d = P["transformer.ln_f.weight"].shape[0]
        caches = []
        x = O.embedding_fwd(P["transformer.wte.weight"], input_ids)
        cos, sin = O.rope_tables(hd, S)
        scale = self.dtype.type(1.0 / np.sqrt(hd))
        cos, sin = cos.astype(self.dtype), sin.astype(self.dtype)
        for i in range(self.cfg["n_layer"]):
            p = f"transformer.h.{i}."
            c = {}
            h1, c["ln1"] = O.rmsnorm_fwd(x, P[p + "ln_1.weight"], self.eps)
            qkv, c["c_attn"] = O.linear_fwd(h1, P[p + "attn.c_attn.weight"], P[p + "attn.c_attn.bias"])
            q5 = qkv.reshape(B, S, 3, H, hd).transpose(2, 0, 3, 1, 4)
            q, k, v = q5[0], q5[1], q5[2]
            qr, kr = O.rope_fwd(q, cos, sin).astype(self.dtype), O.rope_fwd(k, cos, sin).astype(self.dtype)
            a, c["attn"] = O.attention_fwd(qr, kr, v, scale, "causal")
            a2 = a.transpose(0, 2, 1, 3).reshape(B, S, d)
            pr, c["c_proj"] = O.linear_fwd(a2, P[p + "attn.c_proj.weight"], P[p + "attn.c_proj.bias"])
            x2 = pr + x
            h2, c["ln2"] = O.rmsnorm_fwd(x2, P[p + "ln_2.weight"], self.eps)
            g1, c["w1"] = O.linear_fwd(h2, P[p + "mlp.swiglu.w1.weight"], None)
            u, c["w2"] = O.linear_fwd(h2, P[p + "mlp.swiglu.w2.weight"], None)
            gate = O.silu_fwd(g1)
            c["g1"], c["gate"], c["u"] = g1, gate, u
            m, c["w3"] = O.linear_fwd(gate * u, P[p + "mlp.swiglu.w3.weight"], None)
            x = m + x2
            caches.append(c)
        hf, cf = O.rmsnorm_fwd(x, P["transformer.ln_f.weight"], self.eps)
        logits, cl = O.linear_fwd(hf.reshape(B * S, d), P["transformer.wte.weight"], None)
        self._cache = (input_ids, caches, cf, cl, cos, sin, (B, S, d))
        out = {"logits": logits.reshape(B, S, -1)}
        if labels is not None:
            loss, cce = O.cross_entropy_fwd(logits, labels.reshape(-1), "mean")
            out["loss"] = loss
            self._cache_ce = cce
        return out

    def backward(self):
        """d(loss)/d(params) into self.grads (accumulating, like Tensor.backward)."""
        P, H, hd = self.params, self.H, self.hd
        input_ids, caches, cf, cl, cos, sin, (B, S, d) = self._cache
        cg = (lambda g: O.clip_grad(g, True)) if self.grad_clip else (lambda g: g)
        G = {}

        def acc(name, g):
            G[name] = g if name not in G else G[name] + g

        dlogits = cg(O.cross_entropy_bwd(np.float32(1.0), self._cache_ce))
        dhf, dW, _ = O.linear_bwd(dlogits, cl)
        acc("transformer.wte.weight", cg(dW))
        dx, dw = O.rmsnorm_bwd(cg(dhf.reshape(B, S, d)), cf)
        acc("transformer.ln_f.weight", cg(dw))
        dx = cg(dx)
        for i in reversed(range(self.cfg["n_layer"])):
            p = f"transformer.h.{i}."
            c = caches[i]
            # x = m + x2
            dm = dx
            dgu, dW3, _ = O.linear_bwd(dm, c["w3"])
            acc(p + "mlp.swiglu.w3.weight", cg(dW3))
            dgu = cg(dgu)
            dgate, du = cg(dgu * c["u"]), cg(dgu * c["gate"])
            dg1 = cg(O.silu_bwd(dgate, c["g1"]))
            dh2a, dW1, _ = O.linear_bwd(dg1, c["w1"])
            dh2b, dW2, _ = O.linear_bwd(du, c["w2"])
            acc(p + "mlp.swiglu.w1.weight", cg(dW1))
            acc(p + "mlp.swiglu.w2.weight", cg(dW2))
            dx2n, dw2 = O.rmsnorm_bwd(cg(dh2a + dh2b), c["ln2"])
            acc(p + "ln_2.weight", cg(dw2))
            dx2 = cg(dx + dx2n)
            # x2 = pr + x
            da2, dWp, dbp = O.linear_bwd(dx2, c["c_proj"])
            acc(p + "attn.c_proj.weight", cg(dWp))
            acc(p + "attn.c_proj.bias", cg(dbp))
            da = cg(da2).reshape(B, S, H, hd).transpose(0, 2, 1, 3)
            dq, dk, dv = O.attention_bwd(da, c["attn"])
            dq, dk = O.rope_bwd(cg(dq), cos, sin), O.rope_bwd(cg(dk), cos, sin)
            dqkv = np.stack([cg(dq), cg(dk), cg(dv)], axis=0).transpose(1, 3, 0, 2, 4).reshape(B, S, 3 * d).astype(self.dtype)
            dh1, dWa, dba = O.linear_bwd(dqkv, c["c_attn"])
            acc(p + "attn.c_attn.weight", cg(dWa))
            acc(p + "attn.c_attn.bias", cg(dba))
            dxn, dw1 = O.rmsnorm_bwd(cg(dh1), c["ln1"])
            acc(p + "ln_1.weight", cg(dw1))
            dx = cg(dx2 + dxn)
        acc("transformer.wte.weight", O.embedding_bwd_fast(dx.astype(self.dtype), input_ids, P["transformer.wte.weight"].shape[0]))
        for k, g in G.items():
            self.grads[k] = g.astype(self.dtype) if k not in self.grads else self.grads[k] + g.astype(self.dtype)
        return self.grads

    def zero_grad(self):
        self.grads = {}


class OracleAdamW:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, variant="adamw"):
        self.params = params
        self.h = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.state = {k: O.AdamState(v.shape) for k, v in params.items()}
        self.fn = O.adamw_step if variant == "adamw" else O.adam_step

    def step(self, grads):
        for k, g in grads.items():
            self.params[k] = self.fn(self.params[k], g, self.state[k], **self.h)


def train_steps(model: GPT2Oracle, opt: OracleAdamW, batches):
    """forward -> CE -> backward -> AdamW for each (input_ids, labels); returns the loss trajectory."""
    losses = []
    for ids, labels in batches:
        model.zero_grad()
        out = model.forward(ids, labels)
        losses.append(float(out["loss"]))
        opt.step(model.backward())
    return losses


class GPT2OracleBF16(GPT2Oracle):
    """The same model with the B200 backend's bf16 compute mode restated (np_ops.bf16_round at every point where the CUDA
    path stores or feeds bfloat16; fp32 everywhere else, like the kernels): the yardstick that separates "inherent bf16
    rounding noise" from kernel error.  Rounding points, with the code that makes them:
      * GEMM operands: norm outputs, attention output, SwiGLU output, dlogits and the bf16 twins of the residual-stream
        gradient are bf16; weights are read through their bf16 shadow (nfb200/functional/fused.py:_bf16_shadow);
      * GEMM outputs that feed another tensor-core op are stored bf16 (qkv, gate|up, logits in a training step, every dX);
        outputs that join the residual stream stay fp32 (kernels.linear_forward);
      * RoPE: forward, q and k are rotated in the qkv GEMM's epilogue BEFORE the rounding to bf16 (csrc/gemm_tc.cu, rope_cols); backward,
        the packed bf16 dqkv buffer is un-rotated in place (csrc/attention.cu k_rope);
      * flash attention: np_ops.attention_fwd_bf16 / attention_bwd_bf16;
      * SwiGLU and its backward read / write bf16 (csrc/elementwise.cu k_swiglu_fwd / k_swiglu_bwd);
      * cross entropy reads bf16 logits, does its math in fp32 and writes bf16 dlogits (csrc/softmax_ce.cu);
      * the reference's +-10 gradient clip sits in the dX GEMM epilogues and the norm backward (before the rounding).
    Weight / bias / norm-parameter gradients accumulate in fp32 from those bf16 operands."""

    def shadow_weights(self):
        """bf16 shadow of every weight matrix (what the tensor-core GEMMs read)."""
        return {k: O.bf16_round(v) for k, v in self.params.items() if v.ndim == 2}

    def block_forward(self, i, x, W=None):
        """One Transformer block on an fp32 residual-stream input x (B, S, d) -> (fp32 output, cache)."""
        r = O.bf16_round
        P, H, hd = self.params, self.H, self.hd
        W = W or self.shadow_weights()
        B, S, d = x.shape
        cos, sin = O.rope_tables(hd, S)
        scale = np.float32(1.0 / np.sqrt(hd))
        p = f"transformer.h.{i}."
        c = {"x": x, "cos": cos, "sin": sin}
        h1f, c["ln1"] = O.rmsnorm_fwd(x, P[p + "ln_1.weight"], self.eps)
        h1 = r(h1f)
        # the rotary embedding of q and k rides in the projection's epilogue: rotate the fp32 accumulator, THEN round to bf16
        qkv_f = np.matmul(h1, W[p + "attn.c_attn.weight"].T) + P[p + "attn.c_attn.bias"]
        q5 = qkv_f.reshape(B, S, 3, H, hd).transpose(2, 0, 3, 1, 4)
        qr, kr, v = r(O.rope_fwd(q5[0], cos, sin)), r(O.rope_fwd(q5[1], cos, sin)), r(q5[2])
        qkv = r(qkv_f)
        a, c["attn"] = O.attention_fwd_bf16(qr, kr, v, scale, "causal")
        a2 = a.transpose(0, 2, 1, 3).reshape(B, S, d)
        x2 = (np.matmul(a2, W[p + "attn.c_proj.weight"].T) + P[p + "attn.c_proj.bias"]) + x
        h2f, c["ln2"] = O.rmsnorm_fwd(x2, P[p + "ln_2.weight"], self.eps)
        h2 = r(h2f)
        g1 = r(np.matmul(h2, W[p + "mlp.swiglu.w1.weight"].T))
        u = r(np.matmul(h2, W[p + "mlp.swiglu.w2.weight"].T))
        act = r(O.silu_fwd(g1) * u)
        out = (np.matmul(act, W[p + "mlp.swiglu.w3.weight"].T) + x2).astype(np.float32)
        c.update(h1=h1, qkv=qkv, qr=qr, kr=kr, a2=a2, x2=x2, h2=h2, g1=g1, u=u, act=act, out=out)
        return out, c

    def block_backward(self, i, dx, c, acc, W=None):
        """Backward of block i for the fp32 residual-stream gradient dx; parameter gradients go through acc(name, g).
        Returns the gradient of the block input."""
        r = O.bf16_round
        H, hd = self.H, self.hd
        W = W or self.shadow_weights()
        cg = (lambda g: O.clip_grad(g, True)) if self.grad_clip else (lambda g: g)
        B, S, d = dx.shape
        cos, sin = c["cos"], c["sin"]
        p = f"transformer.h.{i}."

        def f2(t):
            return t.reshape(-1, t.shape[-1])

        dz = r(dx)                                                      # bf16 twin of the residual-stream gradient
        dact = r(cg(np.matmul(dz, W[p + "mlp.swiglu.w3.weight"])))
        acc(p + "mlp.swiglu.w3.weight", np.matmul(f2(dz).T, f2(c["act"])))
        sg = 1.0 / (1.0 + np.exp(-c["g1"]))
        dg1 = r(dact * c["u"] * (sg * (1.0 + c["g1"] * (1.0 - sg))))
        du = r(dact * c["g1"] * sg)
        dh2 = r(cg(np.matmul(dg1, W[p + "mlp.swiglu.w1.weight"]) + np.matmul(du, W[p + "mlp.swiglu.w2.weight"])))   # one GEMM over [w1; w2]
        acc(p + "mlp.swiglu.w1.weight", np.matmul(f2(dg1).T, f2(c["h2"])))
        acc(p + "mlp.swiglu.w2.weight", np.matmul(f2(du).T, f2(c["h2"])))
        dx2n, dw2 = O.rmsnorm_bwd(dh2, c["ln2"])
        acc(p + "ln_2.weight", dw2)
        dx2 = cg(dx + dx2n).astype(np.float32)
        dz2 = r(dx2)
        da2 = r(cg(np.matmul(dz2, W[p + "attn.c_proj.weight"])))
        acc(p + "attn.c_proj.weight", np.matmul(f2(dz2).T, f2(c["a2"])))
        acc(p + "attn.c_proj.bias", f2(dz2).sum(axis=0))
        da = da2.reshape(B, S, H, hd).transpose(0, 2, 1, 3)
        dq, dk, dv = O.attention_bwd_bf16(da, c["attn"])
        dq, dk = r(O.rope_bwd(dq, cos, sin)), r(O.rope_bwd(dk, cos, sin))
        dqkv = np.stack([dq, dk, dv], axis=0).transpose(1, 3, 0, 2, 4).reshape(B, S, 3 * d).astype(np.float32)
        dh1 = r(cg(np.matmul(dqkv, W[p + "attn.c_attn.weight"])))
        acc(p + "attn.c_attn.weight", np.matmul(f2(dqkv).T, f2(c["h1"])))
        acc(p + "attn.c_attn.bias", f2(dqkv).sum(axis=0))
        dxn, dw1 = O.rmsnorm_bwd(dh1, c["ln1"])
        acc(p + "ln_1.weight", dw1)
        return cg(dx2 + dxn).astype(np.float32)

    def forward(self, input_ids, labels=None):
        r = O.bf16_round
        P = self.params
        B, S = input_ids.shape
        d = P["transformer.ln_f.weight"].shape[0]
        caches = []
        x = O.embedding_fwd(P["transformer.wte.weight"], input_ids)
        W = self.shadow_weights()
        for i in range(self.cfg["n_layer"]):
            x, c = self.block_forward(i, x, W)
            caches.append(c)
        hff, cf = O.rmsnorm_fwd(x, P["transformer.ln_f.weight"], self.eps)
        hf = r(hff).reshape(B * S, d)
        logits = np.matmul(hf, W["transformer.wte.weight"].T)
        train = labels is not None
        if train:
            logits = r(logits)   # a training step keeps the (T, V) logits in bf16 (models/gpt2.py GPT2LMHead.forward)
        self._cache = (input_ids, caches, cf, hf, (B, S, d), W)
        self._dbg = {"hf": hf, "logits": logits}
        out = {"logits": logits.reshape(B, S, -1)}
        if train:
            loss, cce = O.cross_entropy_fwd(logits.astype(np.float32), labels.reshape(-1), "mean")
            out["loss"] = loss
            self._cache_ce = cce
        return out

    def backward(self):
        r = O.bf16_round
        P = self.params
        input_ids, caches, cf, hf, (B, S, d), W = self._cache
        cg = (lambda g: O.clip_grad(g, True)) if self.grad_clip else (lambda g: g)
        G = {}

        def acc(name, g):
            G[name] = g if name not in G else G[name] + g

        dlogits = r(O.cross_entropy_bwd(np.float32(1.0), self._cache_ce))
        dhf = r(cg(np.matmul(dlogits, W["transformer.wte.weight"])))
        acc("transformer.wte.weight", np.matmul(dlogits.T, hf))
        dx, dw = O.rmsnorm_bwd(dhf.reshape(B, S, d), cf)
        acc("transformer.ln_f.weight", dw)
        dx = cg(dx).astype(np.float32)
        self._dbg_dx = {}
        for i in reversed(range(self.cfg["n_layer"])):
            self._dbg_dx[i] = dx     # the residual-stream gradient entering block i (tests feed it to the CUDA block)
            dx = self.block_backward(i, dx, caches[i], acc, W)
        acc("transformer.wte.weight", O.embedding_bwd_fast(dx.astype(np.float32), input_ids, P["transformer.wte.weight"].shape[0]))
        for k, g in G.items():
            g = cg(g).astype(np.float32)
            self.grads[k] = g if k not in self.grads else self.grads[k] + g
        return self.grads
