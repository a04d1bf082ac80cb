"""Round-2 device tests for what round 1 only exercised on the host path (VERDICT r1 "missing" 4 / 5):
  * BASELINE configs[0], the translation encoder-decoder (examples/translation/model_v2.py), on the device: logits against
    the golden vectors of the UNMODIFIED reference and a 5-step Adam trajectory against this package's host path;
  * checkpoint save / restore with device tensors: bit-identical resume, in fp32 and in the bf16 mode (whose bf16 shadow
    weights must follow a load_state_dict -- ADVICE r1 high);
  * two optimizers over decay / no-decay subsets of one arena;
  * class ids outside the vocabulary in all three cross-entropy kernels."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture
def cuda_default():
    from nfb200 import kernels as K
    from nfb200.core import set_default_device
    set_default_device("cuda")
    yield
    K.set_wgrad_overlap(False)
    K.set_precision("fp32")
    set_default_device("cpu")


def test_translation_model_logits_on_device_match_the_reference(be, cuda_default):
    """Same seeded weights (drawn on the host, as the reference draws them), same padded batch: device fp32 logits within the
    host test's tolerance of the golden logits produced by the unmodified reference (oracle/gen_golden.py)."""
    from nfb200.core import Tensor
    from nfb200.models import TranslationTransformer
    g = np.load(os.path.join(GOLD, "translation.npz"))
    V = 120
    np.random.seed(0)
    model = TranslationTransformer(src_vocab_size=V, tgt_vocab_size=V + 7, d_model=128, n_heads=4, n_layers=3, d_ff=256, max_seq_len=32)
    assert np.array_equal(model.src_embedding.weight.numpy()[:3], g["src_emb_sample"])
    tol = dict(rtol=1e-5, atol=1e-5 * float(np.abs(g["logits"]).max()))
    out = model(Tensor(g["src"], device="cuda"), Tensor(g["tgt"], device="cuda"))
    assert out.device.is_cuda
    np.testing.assert_allclose(out.numpy(), g["logits"], **tol)
    masked = model(Tensor(g["src"], device="cuda"), Tensor(g["tgt"], device="cuda"), src_mask=g["src_mask"], tgt_mask=g["tgt_mask"])
    np.testing.assert_allclose(masked.numpy(), g["logits_masked"], **tol)


def _translation_run(device, steps=5):
    from nfb200.core import Tensor, set_default_device
    from nfb200.functional import cross_entropy_loss, reshape
    from nfb200.models import TranslationTransformer
    from nfb200.optim import Adam
    set_default_device(device)
    rng = np.random.default_rng(1234)
    V, B, S = 1000, 32, 20
    src, tgt = rng.integers(4, V, (B, S)), rng.integers(4, V, (B, S + 1))
    for r in range(B):
        pad = int(rng.integers(0, 6))
        if pad:
            src[r, S - pad:] = 0
            tgt[r, S + 1 - pad:] = 0
    np.random.seed(0)
    model = TranslationTransformer(V, V, d_model=128, n_heads=4, n_layers=3, d_ff=256, max_seq_len=32)
    opt = Adam(model.parameters(), lr=1e-3, **({"moment_dtype": "float64"} if device == "cuda" else {}))
    losses, grads = [], None
    for i in range(steps):
        opt.zero_grad()
        logits = model(Tensor(src, device=device), Tensor(tgt[:, :-1], device=device))
        loss = cross_entropy_loss(reshape(logits, (B * S, V)), Tensor(tgt[:, 1:].reshape(-1), device=device))
        loss.backward()
        if i == 0:
            grads = {n: (p.grad.get() if hasattr(p.grad, "get") else np.asarray(p.grad)).copy() for n, p in model.named_parameters() if p.grad is not None}
        opt.step()
        losses.append(float(loss.item()))
    return losses, grads


def test_translation_config_full_size_step_on_device_matches_host_path(be, cuda_default):
    """BASELINE configs[0] at its full size (vocab 1000, d 128, 4 heads, 3 + 3 layers, d_ff 256, batch 32 x 20, Adam 1e-3):
    first-step gradients of every parameter and the 5-step loss trajectory of the device fp32 path against this package's
    host path (pinned to the reference's logits by tests/test_oracle_golden.py) at the fp32 contract."""
    want_l, want_g = _translation_run("cpu")
    got_l, got_g = _translation_run("cuda")
    assert set(got_g) == set(want_g)
    for n, w in want_g.items():
        scale = float(np.abs(w).max()) + 1e-30
        err = float(np.abs(got_g[n] - w).max()) / scale
        assert err <= 2e-5, f"translation grad {n}: {err:.3e} of the largest entry"
    for a, b in zip(got_l, want_l):
        assert abs(a - b) <= 2e-5 * abs(b), (got_l, want_l)


# ---------------------------------------------------------------------------------------------- checkpoints on the device
CKPT_CFG = {"vocab_size": 211, "n_positions": 64, "n_embd": 128, "n_layer": 2, "n_head": 2}


def _ckpt_objects(seed, lr):
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    from nfb200.optim import lr_scheduler as S
    np.random.seed(seed)
    model = GPT2(CKPT_CFG)
    opt = AdamW(model.parameters(), lr=lr, weight_decay=0.1)
    return model, opt, S.StepLR(opt, step_size=2, gamma=0.5)


def _ckpt_step(model, opt, sch, ids):
    from nfb200.core import Tensor
    labels = np.concatenate([ids[:, 1:], np.zeros((ids.shape[0], 1), ids.dtype)], axis=1)
    opt.zero_grad()
    loss = model(Tensor(ids, device="cuda"), labels=Tensor(labels, device="cuda"))["loss"]
    loss.backward()
    opt.step()
    sch.step()
    return float(loss.item())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_checkpoint_save_restore_continue_on_device(be, cuda_default, tmp_path, precision):
    """nfb200/checkpoint.py with cuda tensors (nn/module.py:112-123 state_dict / load_state_dict, optim/adam.py:279-311 state):
    save after two steps, restore into DIFFERENTLY initialised objects whose optimizer (and with it the flat arena and, in
    bf16 mode, the bf16 shadow weights the tensor-core GEMMs read) already exists, continue: losses and final weights are
    bit-identical to the uninterrupted run.  In bf16 mode the first resumed loss would differ if load_state_dict left the
    shadow at the old weights."""
    from nfb200 import kernels as K
    from nfb200.checkpoint import load_checkpoint, save_checkpoint
    K.set_precision(precision)
    rng = np.random.default_rng(3)
    batches = [rng.integers(0, CKPT_CFG["vocab_size"], (2, 64)) for _ in range(4)]
    model, opt, sch = _ckpt_objects(0, 1e-3)
    ref_losses = [_ckpt_step(model, opt, sch, b) for b in batches]
    ref_sd = model.state_dict()
    model, opt, sch = _ckpt_objects(0, 1e-3)
    first = [_ckpt_step(model, opt, sch, b) for b in batches[:2]]
    path = save_checkpoint(str(tmp_path / "ck" / "gpt2_epoch_1"), model, opt, sch, epoch=0, step=2)
    model2, opt2, sch2 = _ckpt_objects(123, 0.5)
    _ckpt_step(model2, opt2, sch2, batches[3])        # the other run has trained: arena, moments and bf16 shadow are live and different
    back = load_checkpoint(path, model2, opt2, sch2)
    assert back["step"] == 2
    rest = [_ckpt_step(model2, opt2, sch2, b) for b in batches[2:]]
    assert first + rest == ref_losses, (first + rest, ref_losses)
    sd2 = model2.state_dict()
    assert all(np.array_equal(ref_sd[k], sd2[k]) for k in ref_sd)


def test_load_state_dict_refreshes_bf16_shadow(be, cuda_default):
    """Build model + optimizer (bf16 shadow live), THEN load weights: the next forward must see them."""
    from nfb200 import kernels as K
    from nfb200.core import Tensor
    K.set_precision("bf16")
    ids = np.random.default_rng(0).integers(0, CKPT_CFG["vocab_size"], (2, 64))
    a, _, _ = _ckpt_objects(0, 1e-3)
    want = a(Tensor(ids, device="cuda"))["logits"].numpy()
    b, _, _ = _ckpt_objects(99, 1e-3)
    before = b(Tensor(ids, device="cuda"))["logits"].numpy()
    assert not np.allclose(before, want, atol=1e-3)
    b.load_state_dict(a.state_dict())
    np.testing.assert_array_equal(b(Tensor(ids, device="cuda"))["logits"].numpy(), want)
    # a float64 state dict must not cut the parameters off from the arena either
    b.load_state_dict({k: v.astype(np.float64) for k, v in a.state_dict().items()})
    assert all(getattr(p, "_arena_bound", False) for p in b.parameters())
    np.testing.assert_array_equal(b(Tensor(ids, device="cuda"))["logits"].numpy(), want)


def test_two_optimizers_over_subsets_of_one_arena(be, cuda_default):
    """Decay / no-decay groups handed to two AdamW instances after a first optimizer (or a DDP wrapper) has formed the flat
    arena: both work on the SAME flat parameter / gradient buffers (no re-homing that would cut the other one off), each
    touching only its own runs.  Equals one AdamW per group built on a fresh model."""
    from nfb200.core import Tensor
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    from nfb200.optim.arena import ParamArena
    ids = np.random.default_rng(1).integers(0, CKPT_CFG["vocab_size"], (2, 64))
    labels = np.concatenate([ids[:, 1:], np.zeros((2, 1), ids.dtype)], axis=1)

    def groups(model):
        named = dict(model.named_parameters())
        decay = {n: p for n, p in named.items() if p.ndim == 2}
        rest = {n: p for n, p in named.items() if p.ndim != 2}
        return decay, rest

    def run(shared_arena):
        np.random.seed(0)
        model = GPT2(CKPT_CFG)
        if shared_arena:
            home = ParamArena(list(model.parameters()), np.float32, 0)   # what DistributedDataParallel does
        decay, rest = groups(model)
        o1 = AdamW(decay, lr=1e-3, weight_decay=0.1, moment_dtype="float64")
        o2 = AdamW(rest, lr=1e-3, weight_decay=0.0, moment_dtype="float64")
        if shared_arena:
            assert o1.arena.flat_g is home.flat_g and o2.arena.flat_g is home.flat_g and o1.arena.subset_of is home
            assert all(p._arena is home for p in model.parameters())
        losses = []
        for _ in range(3):
            o1.zero_grad()
            o2.zero_grad()
            loss = model(Tensor(ids, device="cuda"), labels=Tensor(labels, device="cuda"))["loss"]
            loss.backward()
            o1.step()
            o2.step()
            losses.append(float(loss.item()))
        return losses, model.state_dict()

    l_a, sd_a = run(True)
    l_b, sd_b = run(False)
    assert l_a == l_b
    assert all(np.array_equal(sd_a[k], sd_b[k]) for k in sd_a)
    # parameters spread over two different arenas cannot be re-homed silently
    np.random.seed(0)
    m = GPT2(CKPT_CFG)
    decay, rest = groups(m)
    AdamW(decay, lr=1e-3)
    with pytest.raises(RuntimeError, match="different flat arenas"):
        AdamW(dict(m.named_parameters()), lr=1e-3)


@pytest.mark.parametrize("rows,V,logits_bf16", [(8, 100, False), (6, 50257, False), (6, 50257, True), (5, 30522, True)])
def test_cross_entropy_class_id_outside_the_vocabulary(be, rows, V, logits_bf16):
    """NumPy raises IndexError for p[i, t] with t outside [-V, V) (functional/loss.py:55-60); the kernels used to read out of
    bounds.  Now: NaN loss and a zero gradient row for that sample, the device index-error flag is raised, and
    every other row is untouched.  Covers the generic, register-resident and bulk (bf16, wide-row) kernels."""
    import ctypes
    from nfb200 import _lib
    from nfb200.array import B200Array, bfloat16, dtype_code
    from oracle import np_ops as O
    rng = np.random.default_rng(V + rows)
    ld = (V + 7) // 8 * 8
    x = rng.standard_normal((rows, ld)).astype(np.float32)
    t = rng.integers(0, V, rows).astype(np.int64)
    t[1], t[3] = -V - 3, V + 5        # below -V (an ignore_index such as -100 on a small vocabulary) and beyond the last class
    t[2] = -1                         # NumPy-style negative index: valid (last class)
    xd = be.from_numpy(x)
    if logits_bf16:
        xd = xd.astype(bfloat16)
        x = xd.get().astype(np.float32)
    loss = B200Array.empty((rows,), np.float32)
    dl = B200Array.empty((rows, ld), np.float32)
    flag = ctypes.c_int(0)
    _lib.call("nfb_index_error_check", ctypes.byref(flag))     # clear
    _lib.call("nfb_cross_entropy_ex", dtype_code(xd.dtype), xd.ptr, be.from_numpy(t).ptr, 1, loss.ptr, dl.ptr, 0, rows, V, ld, ld, 1.0 / rows)
    got_l, got_d = loss.get(), dl.get()[:, :V]
    _lib.call("nfb_index_error_check", ctypes.byref(flag))
    assert flag.value == 1
    ok = np.array([i not in (1, 3) for i in range(rows)])
    tt = np.where(t < 0, t + V, t)
    want_l, cache = O.cross_entropy_fwd(x[ok][:, :V], tt[ok], "none")
    probs = cache[0].copy()
    probs[np.arange(ok.sum()), tt[ok]] -= 1.0
    assert np.isnan(got_l[1]) and np.isnan(got_l[3])
    assert not got_d[1].any() and not got_d[3].any()
    np.testing.assert_allclose(got_l[ok], want_l, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(got_d[ok], probs / rows, rtol=1e-4, atol=1e-7)
