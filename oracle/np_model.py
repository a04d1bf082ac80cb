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
    def __init__(self, config, grad_clip=True):
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
        self.params = P
        self.grads = {}

    # tied: lm_head.weight IS transformer.wte.weight
    def forward(self, input_ids, labels=None):
        P, H, hd = self.params, self.H, self.hd
        B, S = input_ids.shape
        d = P["transformer.ln_f.weight"].shape[0]
        caches = []
        x = O.embedding_fwd(P["transformer.wte.weight"], input_ids)
        cos, sin = O.rope_tables(hd, S)
        scale = np.float32(1.0 / np.sqrt(hd))
        for i in range(self.cfg["n_layer"]):
            p = f"transformer.h.{i}."
            c = {}
            h1, c["ln1"] = O.rmsnorm_fwd(x, P[p + "ln_1.weight"], self.eps)
            qkv, c["c_attn"] = O.linear_fwd(h1, P[p + "attn.c_attn.weight"], P[p + "attn.c_attn.bias"])
            q5 = qkv.reshape(B, S, 3, H, hd).transpose(2, 0, 3, 1, 4)
            q, k, v = q5[0], q5[1], q5[2]
            qr, kr = O.rope_fwd(q, cos, sin).astype(np.float32), O.rope_fwd(k, cos, sin).astype(np.float32)
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
            dqkv = np.stack([cg(dq), cg(dk), cg(dv)], axis=0).transpose(1, 3, 0, 2, 4).reshape(B, S, 3 * d).astype(np.float32)
            dh1, dWa, dba = O.linear_bwd(dqkv, c["c_attn"])
            acc(p + "attn.c_attn.weight", cg(dWa))
            acc(p + "attn.c_attn.bias", cg(dba))
            dxn, dw1 = O.rmsnorm_bwd(cg(dh1), c["ln1"])
            acc(p + "ln_1.weight", cg(dw1))
            dx = cg(dx2 + dxn)
        acc("transformer.wte.weight", O.embedding_bwd_fast(dx.astype(np.float32), input_ids, P["transformer.wte.weight"].shape[0]))
        for k, g in G.items():
            self.grads[k] = g.astype(np.float32) if k not in self.grads else self.grads[k] + g.astype(np.float32)
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
