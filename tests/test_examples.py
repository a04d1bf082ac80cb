"""examples/train_{gpt2,bert_mlm,vit}.py (the workloads of the reference's examples/training/*_training.py on this
package's API) at sizes the host path finishes in seconds: the loss falls, every epoch leaves a reference-layout checkpoint
record plus weights, and the weights load back; the GPT-2 script also runs as two ranks under the launcher."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _example(name):
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    try:
        return importlib.import_module(name)
    finally:
        sys.path.pop(0)


def test_gpt2_training_example_on_the_host_path(tmp_path):
    train_gpt2 = _example("train_gpt2")
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


def test_gpt2_training_example_two_ranks_under_the_launcher(tmp_path):
    """The same script under `nfb200.distributed.launcher` with two host ranks (gloo): DistributedSampler shards,
    DistributedDataParallel averages the gradients, rank 0 alone writes the metrics and the checkpoints."""
    import socket
    from nfb200.distributed import launch_distributed_training
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    args = ["--device", "cpu", "--vocab-size", "64", "--n-embd", "32", "--n-layer", "2", "--n-head", "2", "--n-ctx", "16", "--batch-size", "4",
            "--learning-rate", "0.003", "--num-epochs", "2", "--train-size", "96", "--val-size", "16", "--checkpoint-dir", str(tmp_path / "ck"),
            "--out", str(tmp_path / "metrics.json")]
    env_before = os.environ.get("NFB200_PORT_OFFSET")
    os.environ["NFB200_PORT_OFFSET"] = "0"
    try:
        rc = launch_distributed_training(os.path.join(ROOT, "examples", "train_gpt2.py"), 2, 1, 0, "127.0.0.1", port, "auto", str(tmp_path / "logs"), *args)
    finally:
        if env_before is None:
            os.environ.pop("NFB200_PORT_OFFSET", None)
        else:
            os.environ["NFB200_PORT_OFFSET"] = env_before
    assert rc == 0, (tmp_path / "logs" / "rank_0_stderr.log").read_text()[-2000:] + (tmp_path / "logs" / "rank_1_stderr.log").read_text()[-2000:]
    res = json.load(open(tmp_path / "metrics.json"))
    assert len(res["train_losses"]) == 2 and np.isfinite(res["train_losses"]).all() and res["train_losses"][1] < res["train_losses"][0]
    rec = json.load(open(tmp_path / "ck" / "gpt2_epoch_2.json"))
    assert rec["step"] == 2 * (48 // 4)                       # each rank sees half of the 96 sequences per epoch
    assert "Epoch 2/2" in (tmp_path / "logs" / "rank_0_stdout.log").read_text()
    assert "Epoch" not in (tmp_path / "logs" / "rank_1_stdout.log").read_text()


def test_bert_mlm_and_vit_training_examples_on_the_host_path(tmp_path):
    from nfb200.core import set_default_device
    try:
        bert = _example("train_bert_mlm").main(
            ["--device", "cpu", "--vocab-size", "50", "--hidden-size", "32", "--num-hidden-layers", "2", "--num-attention-heads", "2",
             "--intermediate-size", "64", "--max-seq-length", "16", "--batch-size", "8", "--learning-rate", "0.003", "--num-epochs", "3",
             "--train-size", "96", "--val-size", "16", "--checkpoint-dir", str(tmp_path / "bert")])
        vit = _example("train_vit").main(
            ["--device", "cpu", "--image-size", "16", "--patch-size", "4", "--num-classes", "4", "--d-model", "32", "--depth", "2", "--num-heads", "2",
             "--batch-size", "8", "--learning-rate", "0.003", "--num-epochs", "3", "--train-size", "96", "--val-size", "16",
             "--checkpoint-dir", str(tmp_path / "vit")])
    finally:
        set_default_device("cpu")
    for res in (bert, vit):
        assert np.isfinite(res["train_losses"]).all() and res["train_losses"][-1] < res["train_losses"][0]
    rec = json.load(open(tmp_path / "bert" / "bert_epoch_3.json"))
    assert rec["config"]["hidden_size"] == 32 and rec["step"] == 36
    rec = json.load(open(tmp_path / "vit" / "vit_epoch_3.json"))
    assert rec["config"]["patch_size"] == 4 and (tmp_path / "vit" / "vit_epoch_3.npz").exists()
