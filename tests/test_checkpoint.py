"""Checkpoint files (nfb200/checkpoint.py) on the host path: a run that is saved after two steps, restored
into freshly constructed objects and continued must be bit-identical to the uninterrupted run; the JSON record keeps the
layout of the reference's own checkpoint files (examples/training/gpt2_training.py:528-547)."""
import json

import numpy as np
import pytest

CFG = {"vocab_size": 67, "n_positions": 16, "n_embd": 32, "n_layer": 2, "n_head": 2}


def _make():
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    from nfb200.optim import lr_scheduler as S
    np.random.seed(0)
    model = GPT2(CFG)
    opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.1)
    return model, opt, S.StepLR(opt, step_size=2, gamma=0.5)


def _step(model, opt, sch, ids):
    from nfb200.core import Tensor
    labels = np.concatenate([ids[:, 1:], np.zeros((ids.shape[0], 1), ids.dtype)], axis=1)
    opt.zero_grad()
    loss = model(Tensor(ids), labels=Tensor(labels))["loss"]
    loss.backward()
    opt.step()
    sch.step()
    return float(loss.item())


def test_save_restore_continue_is_bit_identical(tmp_path):
    from nfb200.checkpoint import load_checkpoint, save_checkpoint
    rng = np.random.default_rng(3)
    batches = [rng.integers(0, CFG["vocab_size"], (2, 16)) for _ in range(4)]
    model, opt, sch = _make()
    ref_losses = [_step(model, opt, sch, b) for b in batches]
    ref_sd, ref_lr = model.state_dict(), opt.lr

    model, opt, sch = _make()
    first = [_step(model, opt, sch, b) for b in batches[:2]]
    path = save_checkpoint(str(tmp_path / "ckpt" / "gpt2_epoch_1"), model, opt, sch, epoch=0, step=2,
                           config={"vocab_size": 67, "learning_rate": 1e-3}, metrics={"train": {"loss": first[-1]}})
    assert path.endswith("gpt2_epoch_1.json")
    rec = json.load(open(path))
    assert {"epoch", "step", "config", "metrics"} <= set(rec) and rec["epoch"] == 0 and rec["step"] == 2      # the reference's record
    assert rec["metrics"]["train"]["loss"] == first[-1] and rec["config"]["vocab_size"] == 67
    with np.load(tmp_path / "ckpt" / "gpt2_epoch_1.npz") as z:
        assert set(k[6:] for k in z.files if k.startswith("model/")) == set(model.state_dict())              # state_dict()'s keys
        moments = [k for k in z.files if k.endswith("/exp_avg")]
        assert moments and all(z[k].dtype == np.float64 for k in moments)                                        # fp64, as upstream
    assert not list((tmp_path / "ckpt").glob("*.tmp"))

    np.random.seed(123)                                      # a differently initialised model: everything must come from the file
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    from nfb200.optim import lr_scheduler as S
    model2 = GPT2(CFG)
    opt2 = AdamW(model2.parameters(), lr=0.5, weight_decay=0.1)
    sch2 = S.StepLR(opt2, step_size=2, gamma=0.5)
    back = load_checkpoint(path, model2, opt2, sch2)
    assert back["step"] == 2 and opt2.step_count == 2
    rest = [_step(model2, opt2, sch2, b) for b in batches[2:]]
    assert first + rest == ref_losses
    assert opt2.lr == ref_lr
    sd2 = model2.state_dict()
    assert all(np.array_equal(ref_sd[k], sd2[k]) for k in ref_sd)


def test_load_errors(tmp_path):
    from nfb200.checkpoint import load_checkpoint, save_checkpoint
    model, opt, sch = _make()
    path = save_checkpoint(str(tmp_path / "c"), model)
    with pytest.raises(KeyError):
        load_checkpoint(path, model, opt)                    # saved without optimizer state
    (tmp_path / "ref.json").write_text(json.dumps({"epoch": 2, "step": 150, "config": {}, "metrics": {}}))
    with pytest.raises(ValueError):
        load_checkpoint(str(tmp_path / "ref.json"), model)   # a reference metrics record holds no weights
