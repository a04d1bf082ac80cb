"""examples/train_gpt2.py (the workload of the reference's examples/training/gpt2_training.py on this package's API) at a
size the host path finishes in seconds: the loss falls, every epoch leaves a reference-layout checkpoint record plus
weights, and the weights load back."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpt2_training_example_on_the_host_path(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    try:
        import train_gpt2
    finally:
        sys.path.pop(0)
    from nfb200.core import set_default_device
    try:
        res = train_gpt2.main(["--device", "cpu", "--vocab-size", "64", "--n-embd", "32", "--n-layer", "2", "--n-head", "2", "--n-ctx", "16",
                               "--batch-size", "8", "--learning-rate", "0.003", "--num-epochs", "3", "--train-size", "96", "--val-size", "16",
                               "--checkpoint-dir", str(tmp_path / "ck"), "--out", str(tmp_path / "metrics.json")])
    finally:
        set_default_device("cpu")
    assert len(res["train_losses"]) == 3 and res["train_losses"][-1] < res["train_losses"][0] - 0.05
    assert res["val_losses"][-1] < res["val_losses"][0]
    assert json.load(open(tmp_path / "metrics.json"))["best_val_perplexity"] == res["best_val_perplexity"]
    rec = json.load(open(tmp_path / "ck" / "gpt2_epoch_3.json"))
    assert rec["epoch"] == 2 and rec["step"] == 3 * (96 // 8) and rec["config"]["n_embd"] == 32
    assert set(rec["metrics"]) == {"train", "val"} and {"loss", "perplexity", "accuracy", "time"} <= set(rec["metrics"]["val"])
    from nfb200.checkpoint import load_checkpoint
    from nfb200.models import GPT2
    np.random.seed(5)
    model = GPT2({"vocab_size": 64, "n_embd": 32, "n_layer": 2, "n_head": 2, "n_ctx": 16, "n_positions": 16})
    load_checkpoint(str(tmp_path / "ck" / "gpt2_epoch_3"), model)
    with np.load(tmp_path / "ck" / "gpt2_epoch_3.npz") as z:
        assert np.array_equal(model.state_dict()["transformer.wte.weight"], z["model/transformer.wte.weight"])
