"""Lion optimizer (optim/lion.py) against trajectories of the unmodified reference (tests/golden/lion.npz): host path here,
the fused device kernel (fp64 and fp32 momentum, arena + bf16 shadow) under -m gpu."""
import os

import numpy as np
import pytest

import nfb200  # noqa: F401
from nfb200.core import Parameter
from nfb200.optim import Lion, OptimizerError

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "lion.npz"))
CASES = {"plain": dict(lr=1e-2), "wd_max": dict(lr=3e-3, weight_decay=0.1, maximize=True, betas=(0.8, 0.95)), "biglr": dict(lr=2.5, beta1=0.5)}


def _run(tag, device=None, **extra):
    p = Parameter(G["p0"].copy(), device=device) if device else Parameter(G["p0"].copy())
    opt = Lion({"w": p}, **CASES[tag], **extra)
    traj = []
    for gr in G["grads"]:
        p.grad = gr.copy()
        opt.step()
        traj.append(np.asarray(p.data.get() if device else p.data).copy())
    return np.stack(traj), opt


@pytest.mark.parametrize("tag", list(CASES))
def test_lion_host_matches_reference(tag):
    traj, opt = _run(tag)
    np.testing.assert_array_equal(traj, G[f"{tag}_traj"])                    # same fp64 -> fp32 operation order: bit-exact
    np.testing.assert_array_equal(np.asarray(opt.state["w"]["momentum_buffer"]), G[f"{tag}_m"])
    assert opt.state["w"]["step"] == 7 and opt.get_statistics()["total_steps"] == 7
    sd = opt.get_state_dict()
    other = Lion({"w": Parameter(G["p0"].copy())})
    other.load_state_dict(sd)
    assert other.lr == opt.lr and other.beta2 == opt.beta2 and other.maximize == opt.maximize
    np.testing.assert_array_equal(other.state["w"]["momentum_buffer"], sd["state"]["w"]["momentum_buffer"])


def test_lion_validation():
    p = {"w": Parameter(np.zeros(3, dtype=np.float32))}
    for kw in (dict(lr=-1.0), dict(beta1=1.0), dict(beta2=-0.1), dict(weight_decay=-1.0), dict(betas=(0.9,))):
        with pytest.raises(OptimizerError):
            Lion(p, **kw)
    opt = Lion(p)
    opt.set_lr(0.5)
    assert opt.get_lr() == 0.5 and "Lion(lr=0.5" in repr(opt)
    with pytest.raises(OptimizerError):
        opt.set_lr(-1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("moments", ["float64", "float32"])
@pytest.mark.parametrize("tag", list(CASES))
def test_lion_device_matches_reference(be, tag, moments):
    traj, opt = _run(tag, device="cuda", moment_dtype=moments)
    want = G[f"{tag}_traj"]
    if moments == "float64":
        np.testing.assert_array_equal(traj, want)                             # the reference's arithmetic, element for element
        fin = np.isfinite(G[f"{tag}_m"])                                      # fused multiply-adds in the kernel: last-bit fp64 differences
        np.testing.assert_allclose(opt.state["w"]["momentum_buffer"].get()[fin], G[f"{tag}_m"][fin], rtol=1e-13, atol=1e-300)
    else:
        # fp32 momentum: sign(c) can differ only where |c| is at rounding level; everything else is identical
        diff = np.abs(traj - want)
        assert (diff > 1e-6).mean() < 0.02 and diff.max() <= 2.0 * CASES[tag].get("lr", 1e-4) * 7 + 1e-6
        finite = np.isfinite(G[f"{tag}_m"])
        np.testing.assert_allclose(opt.state["w"]["momentum_buffer"].get()[finite], G[f"{tag}_m"][finite], rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_lion_large_arena_bf16_shadow_and_graph_lr(be):
    """Several parameters in one arena (vector kernel + scalar tail), bf16 shadow refreshed by the same launch, and the learning
    rate read on the device in capturable mode."""
    from nfb200 import kernels as K
    from nfb200.array import B200Array
    rng = np.random.default_rng(5)
    K.set_precision("bf16")
    try:
        shapes = [(257, 129), (1031,), (64, 64, 3)]
        host = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        params = {f"p{i}": Parameter(h.copy(), device="cuda") for i, h in enumerate(host)}
        opt = Lion(params, lr=1e-2, weight_decay=0.05, moment_dtype="float32")
        opt.capturable = True
        ref = [h.copy() for h in host]
        mom = [np.zeros(s, dtype=np.float64) for s in shapes]
        for step in range(3):
            lr = 1e-2 * (0.5 ** step)
            opt.lr = lr
            for i, p in enumerate(params.values()):
                g = rng.standard_normal(shapes[i]).astype(np.float32)
                p.grad = g
                c = 0.9 * mom[i] + 0.1 * g.astype(np.float64)
                upd = np.clip((lr * np.sign(c) + lr * 0.05 * ref[i].astype(np.float64)).astype(np.float32), -1, 1)
                ref[i] = ref[i] - upd
                mom[i] = 0.99 * mom[i] + 0.01 * g.astype(np.float64)
            opt.step()
        for i, p in enumerate(params.values()):
            got = p.data.get()
            assert (np.abs(got - ref[i]) > 1e-6).mean() < 0.01
            bf = p._bf16.get().astype(np.float32)
            np.testing.assert_allclose(bf, got, rtol=2 ** -8, atol=1e-30)
    finally:
        K.set_precision("fp32")
