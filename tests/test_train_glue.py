"""Training-loop glue either side of the optimizer: learning-rate schedulers and the loss/gradient scaler,
pinned to traces of the UNMODIFIED reference (tests/golden/train_glue.npz, made by oracle/gen_golden.py::gen_train_glue).
CPU part: host logic + the NumPy path.  The device path of the same traces is in tests/test_gpu_train_glue.py."""
import json
import os

import numpy as np
import pytest

import nfb200  # noqa: F401
from nfb200.core import Parameter, Tensor
from nfb200.optim import AdamW, OptimizerError
from nfb200.optim import lr_scheduler as S
from nfb200.optimization import (AdvancedGradScaler, NumericalError, ScalerConfig, ScalingStrategy, check_gradients_finite,
                                 clip_gradients_by_norm, create_scaler)

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_glue.npz"))
CASES = json.loads(bytes(G["sched_cases"]).decode())
SHAPES = [tuple(s) for s in json.loads(bytes(G["scaler_shapes"]).decode())]


def _opt(lr):
    return AdamW({"w": Parameter(np.ones(3, dtype=np.float32))}, lr=lr)


@pytest.mark.parametrize("case", CASES["sched"], ids=[c[0] for c in CASES["sched"]])
def test_scheduler_traces_match_reference(case):
    tag, name, kw, lr0, epochs = case
    opt = _opt(lr0)
    sch = getattr(S, name)(opt, **kw)
    trace = [opt.lr]
    for _ in range(epochs):
        sch.step()
        trace.append(opt.lr)
    np.testing.assert_array_equal(np.asarray(trace, dtype=np.float64), G[f"sched_{tag}"])   # bit-exact: same operation order
    assert sch.get_last_lr() == [opt.lr]
    sd = sch.state_dict()
    assert sd["last_epoch"] == sch.last_epoch


@pytest.mark.parametrize("case", CASES["plateau"], ids=[c[0] for c in CASES["plateau"]])
def test_reduce_on_plateau_matches_reference(case):
    tag, kw, lr0, metrics = case
    opt = _opt(lr0)
    sch = S.ReduceLROnPlateau(opt, **kw)
    trace = []
    for m in metrics:
        sch.step(m)
        trace.append(opt.lr)
    np.testing.assert_array_equal(np.asarray(trace, dtype=np.float64), G[f"sched_{tag}"])


def test_scheduler_validation_errors():
    opt = _opt(1e-3)
    for make in (lambda: S.StepLR(opt, 0), lambda: S.StepLR(opt, 2, gamma=1.5), lambda: S.ExponentialLR(opt, 0.0),
                 lambda: S.CosineAnnealingLR(opt, 0), lambda: S.CosineAnnealingLR(opt, 5, eta_min=-1.0),
                 lambda: S.LinearLR(opt, start_factor=0.0), lambda: S.LinearLR(opt, total_iters=0), lambda: S.WarmupLR(opt, 0),
                 lambda: S.WarmupLR(opt, 3, warmup_factor=1.5), lambda: S.PolynomialLR(opt, 0), lambda: S.PolynomialLR(opt, 3, power=0.0),
                 lambda: S.PolynomialLR(opt, 3, end_lr=-1.0), lambda: S.ReduceLROnPlateau(opt, mode="avg"),
                 lambda: S.ReduceLROnPlateau(opt, factor=1.0), lambda: S.ReduceLROnPlateau(opt, patience=-1),
                 lambda: S.ReduceLROnPlateau(opt, threshold_mode="x"), lambda: S.StepLR(object(), 3),
                 lambda: S.ChainedScheduler([S.StepLR(opt, 2)], [3]), lambda: S.ChainedScheduler([S.StepLR(opt, 2)] * 3, [5, 4])):
        with pytest.raises(OptimizerError):
            make()
    with pytest.raises(NotImplementedError):
        S.LRScheduler(opt)


class _Opt:   # the scaler only touches .parameters and .step()
    def __init__(self, params):
        self.parameters, self.steps = params, 0

    def step(self):
        self.steps += 1


SCALER_CFG = {
    "dyn": dict(init_scale=1024.0, growth_interval=3, gradient_clip_threshold=1.0),
    "noclip": dict(init_scale=96.0, growth_interval=2, growth_factor=1.5, gradient_clip_threshold=0.0, max_loss_scale=200.0),
    "fixed": dict(init_scale=8.0, strategy=ScalingStrategy.FIXED, gradient_clip_threshold=0.5),
    "conservative": dict(init_scale=4.0, strategy=ScalingStrategy.CONSERVATIVE, consecutive_success_threshold=2, overflow_cooldown=3,
                         gradient_clip_threshold=1.0, min_loss_scale=2.0),
}


def run_scaler_trace(tag, make_params, read_grad, tensor_grads=False):
    """Replay the golden gradient sequence through AdvancedGradScaler.step; returns nothing, asserts against the trace."""
    sc = AdvancedGradScaler(ScalerConfig(**SCALER_CFG[tag]))
    params = make_params()
    opt = _Opt(params) if not hasattr(params, "step") else params
    plist = list(opt.parameters.values())
    sizes = [int(np.prod(s)) for s in SHAPES]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for t in range(G[f"scaler_{tag}_in"].shape[0]):
        row = G[f"scaler_{tag}_in"][t]
        for i, p in enumerate(plist):
            gr = row[offs[i]:offs[i + 1]].reshape(SHAPES[i]).copy()
            p.grad = Tensor(gr) if tensor_grads else gr
        ok = sc.step(opt)
        assert ok == bool(G[f"scaler_{tag}_taken"][t]), f"step {t}"
        assert sc.get_scale() == G[f"scaler_{tag}_scale"][t], f"step {t}"
        got = np.concatenate([read_grad(p).ravel() for p in plist])
        want = G[f"scaler_{tag}_out"][t]
        fin = np.isfinite(want)
        np.testing.assert_array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], want[fin], rtol=2e-6, atol=0, err_msg=f"step {t}")
    st = sc.get_statistics()
    np.testing.assert_array_equal([st["total_overflows"], st["successful_steps"], st["growth_tracker"]], G[f"scaler_{tag}_stats"])
    np.testing.assert_allclose(np.asarray(sc._gradient_norms_history, dtype=np.float64), G[f"scaler_{tag}_norm_hist"], rtol=2e-6)
    return sc, opt


@pytest.mark.parametrize("tensor_grads", [False, True], ids=["array_grads", "tensor_grads"])
@pytest.mark.parametrize("tag", list(SCALER_CFG))
def test_grad_scaler_trace_host(tag, tensor_grads):
    class P:   # reference-style holder when grads are Tensors (p.grad.data is rebound in place)
        grad = None
    if tensor_grads:
        make = lambda: {f"p{i}": P() for i in range(len(SHAPES))}                      # noqa: E731
        read = lambda p: np.asarray(p.grad.data, dtype=np.float32)                      # noqa: E731
    else:
        make = lambda: {f"p{i}": Parameter(np.zeros(s, dtype=np.float32)) for i, s in enumerate(SHAPES)}   # noqa: E731
        read = lambda p: np.asarray(p.grad, dtype=np.float32)                           # noqa: E731
    sc, opt = run_scaler_trace(tag, make, read, tensor_grads)
    assert opt.steps == int(G[f"scaler_{tag}_optsteps"])
    # state round trip
    other = AdvancedGradScaler()
    other.load_state_dict(sc.state_dict())
    assert other.get_scale() == sc.get_scale() and other.config.strategy == sc.config.strategy
    assert other.get_statistics()["total_overflows"] == sc.get_statistics()["total_overflows"]


def test_scale_loss_and_errors():
    sc = AdvancedGradScaler(ScalerConfig(init_scale=512.0))
    w = Parameter(np.asarray([0.37], dtype=np.float32))
    loss = (w * 1.0).sum() if hasattr(w * 1.0, "sum") else w
    out = sc.scale(loss)
    np.testing.assert_allclose(np.asarray(out.data).reshape(()), G["scaler_scaled_loss"], rtol=0, atol=0)
    out.backward()
    np.testing.assert_allclose(np.asarray(w.grad), [10.0])      # d(512 w)/dw = 512, clipped to the engine's +-10 (Tensor.backward contract)
    with pytest.raises(TypeError):
        sc.scale(3.0)
    with pytest.raises(NumericalError):
        sc.scale(Tensor(np.asarray(np.inf, dtype=np.float32)))
    sc.disable()
    assert sc.scale(loss) is loss and sc.unscale_(_Opt({})) is True
    assert create_scaler("nonsense").config.strategy == ScalingStrategy.DYNAMIC
    assert create_scaler("fixed", init_scale=4.0).get_scale() == 4.0
    with pytest.raises(ValueError):
        AdvancedGradScaler().unscale_(object())


def test_gradient_helpers_host():
    sizes = [int(np.prod(s)) for s in SHAPES]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    params = {f"p{i}": Parameter(np.zeros(s, dtype=np.float32)) for i, s in enumerate(SHAPES)}
    for i, p in enumerate(params.values()):
        p.grad = G["helper_in"][offs[i]:offs[i + 1]].reshape(SHAPES[i]).copy()
    opt = _Opt(params)
    fin, st = check_gradients_finite(opt)
    np.testing.assert_allclose([float(fin), st["params_with_grad"], st["finite_grads"], st["max_grad_norm"], st["avg_grad_norm"]],
                               G["helper_stats"], rtol=1e-6)
    total = clip_gradients_by_norm(opt, 0.75)
    np.testing.assert_allclose(total, G["helper_total_norm"], rtol=1e-6)
    got = np.concatenate([np.asarray(p.grad).ravel() for p in params.values()])
    np.testing.assert_allclose(got, G["helper_clipped"], rtol=2e-6)
