"""RoBERTa against seeded forwards of the unmodified reference (tests/golden/roberta.npz, made by
oracle/gen_golden.py::gen_roberta): host path here (always), device fp32 / bf16 paths under -m gpu."""
import os

import numpy as np
import pytest

import nfb200  # noqa: F401
from nfb200.core import Tensor
from nfb200.models import RoBERTaConfig, RoBERTaForMaskedLM, RoBERTaForSequenceClassification, roberta_base

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "roberta.npz"))
TINY = dict(vocab_size=97, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64, max_position_embeddings=20)


def _mlm(device=None):
    np.random.seed(31)
    m = RoBERTaForMaskedLM(**TINY)
    return m.to(device) if device else m


def test_roberta_forward_matches_reference_host():
    m = _mlm()
    np.testing.assert_array_equal(np.asarray(m.roberta.embeddings.LayerNorm.weight.data), G["ln_gain_first"])   # the init quirk, same RNG order
    out = m(Tensor(G["ids"]), attention_mask=Tensor(G["mask"]), labels=Tensor(G["labels"]))
    np.testing.assert_allclose(np.asarray(out["logits"].data), G["mlm_logits"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(float(out["loss"].item()), float(G["mlm_loss"]), rtol=1e-6)
    np.random.seed(32)
    c = RoBERTaForSequenceClassification(num_labels=3, **TINY)
    out = c(Tensor(G["ids"]), attention_mask=Tensor(G["mask"]), labels=Tensor(G["cls_labels"]))
    np.testing.assert_allclose(np.asarray(out["logits"].data), G["cls_logits"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(float(out["loss"].item()), float(G["cls_loss"]), rtol=1e-6)


def test_roberta_config_and_errors():
    cfg = RoBERTaConfig()
    assert (cfg.vocab_size, cfg.max_position_embeddings, cfg.type_vocab_size, cfg.layer_norm_eps, cfg.pad_token_id) == (50265, 514, 1, 1e-5, 1)
    m = _mlm()
    names = [n for n, _ in m.named_parameters()]
    assert "roberta.embeddings.word_embeddings.weight" in names and "roberta.encoder.layer.1.output.dense.weight" in names and "lm_head.bias" in names
    assert m.roberta.embeddings.token_type_embeddings.weight.shape == (1, 32)
    with pytest.raises(ValueError):
        m.roberta()
    with pytest.raises(ValueError):
        m.roberta(Tensor(G["ids"]), inputs_embeds=Tensor(np.zeros((3, 11, 32), dtype=np.float32)))
    assert roberta_base.__name__ == "roberta_base"


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_roberta_device_matches_reference(be, precision):
    from nfb200 import kernels as K
    K.set_precision(precision)
    try:
        m = _mlm("cuda")
        out = m(Tensor(G["ids"], device="cuda"), attention_mask=Tensor(G["mask"], device="cuda"), labels=Tensor(G["labels"], device="cuda"))
        got = out["logits"].data.get().astype(np.float32)
        if precision == "fp32":
            np.testing.assert_allclose(got, G["mlm_logits"], rtol=2e-5, atol=5e-6)
            np.testing.assert_allclose(float(out["loss"].item()), float(G["mlm_loss"]), rtol=1e-5)
        else:
            scale = float(np.abs(G["mlm_logits"]).max())
            assert float(np.abs(got - G["mlm_logits"]).max()) <= 2e-2 * scale
            np.testing.assert_allclose(float(out["loss"].item()), float(G["mlm_loss"]), rtol=1e-2)
        out["loss"].backward()
        assert m.roberta.embeddings.position_embeddings.weight.grad is not None
        pg = m.roberta.embeddings.position_embeddings.weight.grad.get()
        assert np.all(pg[:2] == 0) and np.any(pg[2:13] != 0) and np.all(pg[13:] == 0)   # positions pad+1 .. pad+S only
    finally:
        K.set_precision("fp32")


@pytest.mark.gpu
def test_roberta_real_head_dim_train_steps_device_equals_host(be):
    """hd = 64: the flash kernels with the per-key padding mask run inside RoBERTa; three AdamW steps on device == on host."""
    from nfb200.core import set_default_device
    from nfb200.optim import AdamW
    rng = np.random.default_rng(9)
    ids = rng.integers(3, 211, (2, 24))
    mask = np.ones((2, 24), np.float32)
    mask[1, 17:] = 0
    res = {}
    for dev in ("cpu", "cuda"):
        set_default_device(dev)
        try:
            np.random.seed(4)
            m = RoBERTaForMaskedLM(vocab_size=211, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                                   max_position_embeddings=40)
            opt = AdamW(m.parameters(), lr=1e-3, moment_dtype="float64")
            losses = []
            for _ in range(3):
                opt.zero_grad()
                l = m(Tensor(ids, device=dev), attention_mask=Tensor(mask, device=dev), labels=Tensor(ids, device=dev))["loss"]
                l.backward()
                opt.step()
                losses.append(float(l.item()))
            res[dev] = losses
        finally:
            set_default_device("cpu")
    np.testing.assert_allclose(res["cuda"], res["cpu"], rtol=2e-5)
    assert res["cpu"][-1] < res["cpu"][0]
