"""Pre-norm RoPE Transformer (models/language/modern_transformer.py) against seeded forwards of the unmodified
reference (tests/golden/prenorm.npz, oracle/gen_golden.py::gen_prenorm): host path here, device paths under -m gpu."""
import os

import numpy as np
import pytest

import nfb200  # noqa: F401
from nfb200.core import Tensor
from nfb200.models import ModelError, PreNormTransformer, PreNormTransformerConfig, prenorm_transformer_small

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prenorm.npz"))
CASES = {
    "gelu_ln": dict(d_model=32, num_layers=2, num_heads=4, d_ff=64, max_seq_len=16, vocab_size=60),
    "swiglu_rms_untied": dict(d_model=32, num_layers=2, num_heads=2, d_ff=48, max_seq_len=16, vocab_size=60, activation="swiglu",
                              normalization="rmsnorm", tie_embeddings=False, scale_embeddings=False, rope_base=500.0),
    "relu_norope": dict(d_model=24, num_layers=1, num_heads=3, d_ff=40, max_seq_len=16, vocab_size=60, activation="relu", use_rope=False),
}


def _model(tag, device=None):
    np.random.seed(50 + list(CASES).index(tag))
    m = PreNormTransformer(**CASES[tag])
    return m.to(device) if device else m


def _run_all(m, dev, get, tol):
    t = (lambda a: Tensor(a, device=dev)) if dev else Tensor
    tag = m._tag
    np.testing.assert_allclose(get(m(t(G["ids"]))["logits"]), G[f"{tag}_logits"], rtol=tol, atol=tol)
    np.testing.assert_allclose(get(m(t(G["ids"]), attention_mask=t(G["mask"]))["logits"]), G[f"{tag}_logits_masked"], rtol=tol, atol=tol)
    np.testing.assert_allclose(get(m(t(G["ids"][:, :5]), start_pos=3)["logits"]), G[f"{tag}_logits_pos3"], rtol=tol, atol=tol)


@pytest.mark.parametrize("tag", list(CASES))
def test_prenorm_forward_matches_reference_host(tag):
    assert np.abs(G[f"{tag}_logits"] - G[f"{tag}_logits_masked"]).max() > 1e-4          # the padding mask does change the reference output
    m = _model(tag)
    m._tag = tag
    _run_all(m, None, lambda x: np.asarray(x.data), 2e-6)


def test_prenorm_config_validation_and_graph():
    for kw in (dict(d_model=0), dict(d_model=30, num_heads=4), dict(num_layers=0), dict(d_ff=0), dict(max_seq_len=0), dict(vocab_size=0)):
        with pytest.raises(ModelError):
            PreNormTransformerConfig(**kw)
    m = _model("swiglu_rms_untied")
    out = m(Tensor(G["ids"]), attention_mask=Tensor(G["mask"]), labels=Tensor(G["ids"]), output_hidden_states=True)
    assert len(out["hidden_states"]) == 3
    out["loss"].backward()
    assert all(p.grad is not None for _, p in m.named_parameters())       # one autograd graph (the reference's is severed)
    names = {n for n, _ in m.named_parameters()}
    assert {"token_embedding", "layer_0.self_attention.q_proj.weight", "layer_1.feed_forward.up_proj.weight", "final_norm.weight",
            "output_projection.weight"} <= names
    assert m._count_parameters() == sum(int(p.size) for p in m.parameters())
    assert prenorm_transformer_small.__name__ == "prenorm_transformer_small"


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_prenorm_device_fp32_matches_reference(be, tag):
    m = _model(tag, "cuda")
    m._tag = tag
    _run_all(m, "cuda", lambda x: x.data.get().astype(np.float32), 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("activation,norm", [("gelu", "layernorm"), ("swiglu", "rmsnorm")])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_prenorm_real_head_dim_train_steps_device_vs_host(be, precision, activation, norm):
    """head_dim 64: concatenated q/k/v GEMM -> interleaved RoPE in place on the packed buffer -> flash attention with the padding
    mask; three AdamW steps with the LM loss, device against host."""
    from nfb200 import kernels as K
    from nfb200.core import set_default_device
    from nfb200.optim import AdamW
    rng = np.random.default_rng(12)
    ids = rng.integers(0, 150, (2, 40))
    mask = np.ones((2, 40), dtype=np.int64)
    mask[1, 31:] = 0
    res = {}
    for dev in ("cpu", "cuda"):
        set_default_device(dev)
        K.set_precision(precision if dev == "cuda" else "fp32")
        try:
            np.random.seed(3)
            m = PreNormTransformer(d_model=128, num_layers=2, num_heads=2, d_ff=256, max_seq_len=64, vocab_size=150, activation=activation,
                                   normalization=norm)
            opt = AdamW(m.parameters(), lr=1e-4, moment_dtype="float64")
            losses = []
            for _ in range(3):
                opt.zero_grad()
                l = m(Tensor(ids, device=dev), attention_mask=Tensor(mask, device=dev), labels=Tensor(ids, device=dev))["loss"]
                l.backward()
                opt.step()
                losses.append(float(l.item()))
            res[dev] = losses
        finally:
            K.set_precision("fp32")
            set_default_device("cpu")
    np.testing.assert_allclose(res["cuda"], res["cpu"], rtol=5e-5 if precision == "fp32" else 1e-2)
    assert res["cpu"][-1] < res["cpu"][0]
