"""Device path of the training-loop glue: the loss/gradient scaler's fused unscale + per-parameter clip
kernels against the reference traces (tests/golden/train_glue.npz) and against NumPy at sizes with many chunks, and
learning-rate schedulers acting on a training step that is replayed from a CUDA graph (device-side learning rate)."""
import numpy as np
import pytest

from test_train_glue import G, SCALER_CFG, SHAPES, _Opt, run_scaler_trace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arena", [True, False], ids=["arena_one_call", "per_parameter"])
@pytest.mark.parametrize("tag", list(SCALER_CFG))
def test_grad_scaler_trace_device(be, tag, arena):
    from nfb200.core import Parameter
    from nfb200.optim import SGD

    def make():
        params = {f"p{i}": Parameter(np.zeros(s, dtype=np.float32), device="cuda") for i, s in enumerate(SHAPES)}
        return SGD(params, lr=0.0) if arena else params     # SGD re-homes the gradients into one flat arena
    sc, opt = run_scaler_trace(tag, make, lambda p: p.grad.get())
    if arena:
        assert opt.step_count == int(G[f"scaler_{tag}_optsteps"])
    else:
        assert opt.steps == int(G[f"scaler_{tag}_optsteps"])


@pytest.mark.parametrize("bad_at", [None, 0, 2, 4])
def test_grad_unscale_large_matches_numpy(be, bad_at):
    """Many chunks per parameter, sizes that are not multiples of 4 or of the chunk, a parameter smaller than a vector;
    one parameter made non-finite: everything in front of it is unscaled + clipped, it and everything behind is untouched."""
    from nfb200.core import Parameter
    from nfb200.optim import AdamW
    from nfb200.optimization import AdvancedGradScaler, ScalerConfig
    rng = np.random.default_rng(3)
    shapes = [(1031, 517), (3,), (50000,), (129, 64, 7), (1 << 16,)]
    params = {f"w{i}": Parameter(np.zeros(s, dtype=np.float32), device="cuda") for i, s in enumerate(shapes)}
    opt = AdamW(params, lr=0.0, moment_dtype="float32")
    scale = 24.0                                                    # not a power of two: true division, not a reciprocal multiply
    grads = [(rng.standard_normal(s) * 10.0 ** rng.integers(-4, -1) * scale).astype(np.float32) for s in shapes]
    if bad_at is not None:
        grads[bad_at].flat[grads[bad_at].size // 2] = np.nan
    for p, g in zip(params.values(), grads):
        p.grad = g
    sc = AdvancedGradScaler(ScalerConfig(init_scale=scale, gradient_clip_threshold=0.05))
    ok = sc.unscale_(opt)
    assert ok == (bad_at is None)
    for i, (p, g) in enumerate(zip(params.values(), grads)):
        got = p.grad.get()
        if bad_at is not None and i >= bad_at:
            np.testing.assert_array_equal(got, g)                   # untouched (NaN compares via array_equal's equal_nan default False ->
            continue                                                # handled: only the bad parameter holds a NaN)
        u = g / np.float32(scale)
        n = np.float32(np.sqrt(np.sum(u.astype(np.float64) ** 2)))
        if n > 0.05:
            u = u * (np.float32(0.05) / (n + np.float32(1e-6)))
        np.testing.assert_allclose(got, u, rtol=2e-6, atol=0)


def test_gradient_helpers_device(be):
    from nfb200.core import Parameter
    from nfb200.optimization import check_gradients_finite, clip_gradients_by_norm
    sizes = [int(np.prod(s)) for s in SHAPES]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    params = {f"p{i}": Parameter(np.zeros(s, dtype=np.float32), device="cuda") for i, s in enumerate(SHAPES)}
    for i, p in enumerate(params.values()):
        p.grad = G["helper_in"][offs[i]:offs[i + 1]].reshape(SHAPES[i]).copy()
    opt = _Opt(params)
    fin, st = check_gradients_finite(opt)
    np.testing.assert_allclose([float(fin), st["params_with_grad"], st["finite_grads"], st["max_grad_norm"], st["avg_grad_norm"]],
                               G["helper_stats"], rtol=1e-6)
    np.testing.assert_allclose(clip_gradients_by_norm(opt, 0.75), G["helper_total_norm"], rtol=1e-6)
    got = np.concatenate([p.grad.get().ravel() for p in params.values()])
    np.testing.assert_allclose(got, G["helper_clipped"], rtol=2e-6)
    list(params.values())[1].grad = np.full(SHAPES[1], np.inf, dtype=np.float32)
    fin, st = check_gradients_finite(opt)
    assert not fin and st["infinite_grads"] == 1 and st["finite_grads"] == len(SHAPES) - 1


CFG = {"vocab_size": 50257, "n_positions": 128, "n_embd": 256, "n_layer": 2, "n_head": 4}


@pytest.mark.parametrize("schedule", ["cosine_warmup_scheduler", "adamw_builtin_warmup_cosine", "plateau"])
def test_lr_schedule_acts_on_graph_replays(be, schedule):
    """The same 8 training steps run (a) eagerly and (b) with steps 3.. replayed from ONE captured CUDA graph; the scheduler
    is stepped between replays exactly as in the eager loop.  The fused AdamW reads lr from device memory, so both runs see
    the same learning rates and end at the same loss; (c) a graphed run whose scheduler is never stepped must differ."""
    from nfb200 import kernels as K
    from nfb200.core import Tensor, set_default_device
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    from nfb200.optim import lr_scheduler as S
    from nfb200.training import GraphedStep
    set_default_device("cuda")
    K.set_precision("bf16")
    rng = np.random.default_rng(11)
    ids = rng.integers(0, CFG["vocab_size"], (2, 128))
    labels = np.concatenate([ids[:, 1:], np.zeros((2, 1), dtype=ids.dtype)], axis=1)
    nsteps = 8

    def run(graphed, use_schedule=True):
        np.random.seed(0)
        model = GPT2(CFG)
        if schedule == "adamw_builtin_warmup_cosine":
            kw = dict(lr_scheduler="warmup_cosine", lr_scheduler_params={"warmup_steps": 3, "T_max": 8, "eta_min": 1e-5}) if use_schedule else {}
            opt = AdamW(model.parameters(), lr=2e-3, weight_decay=0.1, moment_dtype="float32", **kw)
            sched_step = lambda loss: None                                                    # noqa: E731
        else:
            opt = AdamW(model.parameters(), lr=2e-3, weight_decay=0.1, moment_dtype="float32")
            if schedule == "plateau":
                sch = S.ReduceLROnPlateau(opt, factor=0.25, patience=0, threshold=0.5, threshold_mode="abs")   # every step counts as "bad"
                sched_step = (lambda loss: sch.step(1.0)) if use_schedule else (lambda loss: None)            # noqa: E731
            else:
                sch = S.get_cosine_schedule_with_warmup(opt, 3, nsteps) if use_schedule else None
                sched_step = (lambda loss: sch.step()) if use_schedule else (lambda loss: None)               # noqa: E731
        ids_t, lab_t = Tensor(ids, device="cuda"), Tensor(labels, device="cuda")

        def step():
            opt.zero_grad()
            loss = model(ids_t, labels=lab_t)["loss"]
            loss.backward()
            opt.step()
            return loss
        lrs, losses, gs = [], [], None
        for i in range(nsteps):
            lrs.append(opt.lr)
            if graphed and i == 2:
                gs = GraphedStep(step, warmup=1, optimizers=[opt])      # runs ONE real step (this one), then captures
                loss = None                                              # (the captured call itself did not execute)
            elif gs is not None:
                loss = gs()
            else:
                loss = step()
            if schedule == "adamw_builtin_warmup_cosine":
                lrs[-1] = opt.lr                                         # the built-in schedule sets lr inside step()
            losses.append(float(loss.item()) if loss is not None else None)
            sched_step(losses[-1])
        sd = opt.get_state_dict()
        assert sd["step_count"] == nsteps and all(s["step"] == nsteps for s in sd["state"].values())
        return lrs, losses

    try:
        e_lr, e_loss = run(False)
        g_lr, g_loss = run(True)
        c_lr, c_loss = run(True, use_schedule=False)
        assert e_lr == g_lr and len(set(e_lr)) > 2, (e_lr, g_lr)
        keep = [i for i, x in enumerate(g_loss) if x is not None]
        assert keep == [0, 1, 3, 4, 5, 6, 7]
        np.testing.assert_allclose([g_loss[i] for i in keep], [e_loss[i] for i in keep], rtol=2e-4)
        assert abs(c_loss[-1] - e_loss[-1]) > 1e-3 * abs(e_loss[-1]), (c_loss, e_loss)   # frozen lr is a different trajectory
    finally:
        K.set_precision("fp32")
        set_default_device("cpu")


def test_sgd_lr_schedule_acts_on_graph_replays(be):
    """SGD (optim/sgd.py:87-115) with momentum under a StepLR schedule: steps 3.. replayed from one captured graph must
    follow the eager trajectory (the kernel reads the learning rate from device memory, nfb_sgd_step_ex), and a graphed
    run whose scheduler is never stepped must not."""
    from nfb200 import kernels as K
    from nfb200.core import Tensor, set_default_device
    from nfb200.models import GPT2
    from nfb200.optim import SGD
    from nfb200.optim import lr_scheduler as S
    from nfb200.training import GraphedStep
    set_default_device("cuda")
    K.set_precision("bf16")
    rng = np.random.default_rng(12)
    ids = rng.integers(0, CFG["vocab_size"], (2, 128))
    labels = np.concatenate([ids[:, 1:], np.zeros((2, 1), dtype=ids.dtype)], axis=1)
    nsteps = 8

    def run(graphed, use_schedule=True):
        np.random.seed(0)
        model = GPT2(CFG)
        opt = SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True)
        sch = S.StepLR(opt, step_size=2, gamma=0.3) if use_schedule else None
        ids_t, lab_t = Tensor(ids, device="cuda"), Tensor(labels, device="cuda")

        def step():
            opt.zero_grad()
            loss = model(ids_t, labels=lab_t)["loss"]
            loss.backward()
            opt.step()
            return loss
        lrs, losses, gs = [], [], None
        for i in range(nsteps):
            lrs.append(opt.lr)
            if graphed and i == 2:
                gs = GraphedStep(step, warmup=1, optimizers=[opt])
                loss = None
            else:
                loss = gs() if gs is not None else step()
            losses.append(float(loss.item()) if loss is not None else None)
            if sch is not None:
                sch.step()
        assert opt.step_count == nsteps
        return lrs, losses

    try:
        e_lr, e_loss = run(False)
        g_lr, g_loss = run(True)
        c_lr, c_loss = run(True, use_schedule=False)
        assert e_lr == g_lr and len(set(e_lr)) > 2, (e_lr, g_lr)
        keep = [i for i, x in enumerate(g_loss) if x is not None]
        np.testing.assert_allclose([g_loss[i] for i in keep], [e_loss[i] for i in keep], rtol=2e-4)
        assert abs(c_loss[-1] - e_loss[-1]) > 1e-3 * abs(e_loss[-1]), (c_loss, e_loss)
    finally:
        K.set_precision("fp32")
        set_default_device("cpu")


def test_clip_grad_norm_wrapper_on_device(be):
    """optim/clip.py on cuda tensors: the norm, the directly scaled gradients, and the form deferred into the fused AdamW
    kernel (device coefficient) against the host path of the same function."""
    from nfb200.core import Parameter
    from nfb200.optim import AdamW, clip_grad_norm
    rng = np.random.default_rng(21)
    shapes = [(40, 33), (517,), (64, 64)]
    vals = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    grads = [rng.standard_normal(s).astype(np.float32) * 3 for s in shapes]

    def run(device, deferred):
        ps = {f"p{i}": Parameter(v.copy(), device=device) for i, v in enumerate(vals)}
        opt = AdamW(ps, lr=1e-2, weight_decay=0.0, moment_dtype="float32")
        for p, g in zip(ps.values(), grads):
            p.grad = g.copy()
        norm = clip_grad_norm(ps, 1.0, optimizer=opt if deferred else None)
        norm = float(np.asarray(norm.get() if hasattr(norm, "get") else norm).reshape(-1)[0])
        g_after = [np.asarray(p.grad.get() if hasattr(p.grad, "get") else p.grad) for p in ps.values()]
        opt.step()
        return norm, g_after, [p.numpy() for p in ps.values()]

    n_host, g_host, p_host = run("cpu", False)
    for deferred in (False, True):
        n_dev, g_dev, p_dev = run("cuda", deferred)
        np.testing.assert_allclose(n_dev, n_host, rtol=1e-6)
        if not deferred:
            for a, b in zip(g_dev, g_host):
                np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9)
        for a, b in zip(p_dev, p_host):
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
