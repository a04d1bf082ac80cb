#!/usr/bin/env python
"""bench.py -- training tokens/s of the reference-architecture GPT-2 small (BASELINE.json configs[1]:
"GPT-2 small (model zoo, random-init) causal LM training, batch 8 x seq 512, on 1 B200") through the nfb200
public API (Tensor / nn / optim on the b200 backend), one step = forward + cross-entropy + backward + AdamW
over one synthetic (8, 512) token batch.

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (N > 1 under torchrun: data parallel)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm: the UNMODIFIED reference (oracle/_ref) on the host cores

One JSON line on stdout (rank 0).
  value             whole-job tokens/s with the batch resident in HBM, device-timed with CUDA events (max over ranks)
  e2e               the same step fed from pinned HOST buffers (H2D of ids/labels and D2H of the loss inside the timed region)
  roofline          the tcgen05 GEMM (dominant kernel): CUDA-event pairs around every launch of a single-stream replay of the
                    step (`frac`), next to the kernel's own %globaltimer span (`frac_device_clock`), vs the measured bf16 peak
  roofline_classes  every kernel class of the step (tensor: GEMM, attention; HBM: norm, cross entropy, optimizer, elementwise):
                    algorithmic work / event-timed duration vs the measured peaks
  other_configs     BASELINE.json configs[2..4]: BERT-base MLM B32xS512 bf16, ViT-B/16 B64 (data parallel at N GPUs) and, at
                    8 GPUs, GPT-2 medium B64xS1024 -- step time and throughput of each on the same N GPUs
  ddp_parity        (N > 1) a 3-step check that N ranks on their shards == one rank on the global batch
  cpu_baseline      (N = 1) the reference's own CPU step on a bounded sample, and the oracle port's
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "neural-network-from-scratch_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

GPT2_SMALL = {"vocab_size": 50257, "n_positions": 1024, "n_embd": 768, "n_layer": 12, "n_head": 12}
BATCH, SEQ = 8, 512
TRAIN_FLOP_PER_TOKEN = 712.88e6  # SURVEY App. C (GEMM + logits + attention, x3 for training)
# dram__bytes_read.sum + dram__bytes_write.sum per k_gemm_tc launch, launch-weighted over one step (profiles/r02_ncu_gemm_full.txt)
GEMM_DRAM_BYTES_PER_LAUNCH = 47.0e6
REF_DIR = os.path.join(ROOT, "oracle", "_ref", "neural_arch")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index=0):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(rng, batch=BATCH, seq=SEQ, vocab=GPT2_SMALL["vocab_size"]):
    ids = rng.integers(0, vocab, (batch, seq)).astype(np.int64)
    labels = np.concatenate([ids[:, 1:], np.zeros((batch, 1), dtype=np.int64)], axis=1)  # shifted by one, last = 0
    return ids, labels


# ------------------------------------------------------------------------------------------ reference / CPU arm
def host_threads():
    """Make the BLAS / OpenMP pools use every host core (torchrun exports OMP_NUM_THREADS=1 to its workers, which silently
    turned the CPU arm single-threaded at N > 1) and return the thread count the BLAS pool REALLY has."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        pools = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        if pools:
            return int(max(pools))
    except Exception:
        pass
    return int(os.environ.get("OMP_NUM_THREADS", n))


def cpu_oracle_run(steps, warmup, sample_batch):
    """The oracle PORT of the reference NumPy path with a CONNECTED graph (forward + CE + full backward + AdamW over all
    109 M parameters) at full model size on `sample_batch` of the 8 sequences per step."""
    from oracle.np_model import GPT2Oracle, OracleAdamW
    cores = host_threads()
    np.random.seed(0)
    model = GPT2Oracle(GPT2_SMALL, grad_clip=True)
    opt = OracleAdamW(model.params, lr=5e-4, weight_decay=0.1)
    rng = np.random.default_rng(1234)
    ids, labels = synthetic_batch(rng, sample_batch, SEQ)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad()
        model.forward(ids, labels)
        opt.step(model.backward())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    return {"value": sample_batch * SEQ / t, "unit": "tokens/s", "cores": cores, "kind": "port", "step_s": t,
            "sample": f"oracle port (NumPy/OpenBLAS on {cores} threads; connected graph: forward + CE + full backward + AdamW), full GPT-2 small, "
                      f"{sample_batch} of {BATCH} sequences x {SEQ} tokens per step, median of {steps}"}


def cpu_reference_run(steps, warmup, sample_batch, device="cpu"):
    """The UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py; its missing `core` from oracle/compat) through its
    own public API and stock configuration (Numba JIT + fusion defaults), as examples/training/gpt2_training.py:344-420 drives
    it: model(input_ids) -> logits re-wrapped in a fresh Tensor -> cross_entropy_loss -> loss.backward() -> AdamW.step().
    NOTE: that re-wrap severs the graph -- the reference's backward stops at the logits and its optimizer finds no
    gradients -- so a reference "step" is a forward pass + loss; the port above does the whole step."""
    os.environ["NEURAL_ARCH_REFERENCE"] = REF_DIR
    if os.path.join(ROOT, "oracle", "compat") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "compat"))
    import neural_arch  # noqa: F401
    from neural_arch.core import Tensor, get_default_device, set_default_device
    from neural_arch.functional import cross_entropy_loss
    from neural_arch.models.language.gpt2 import GPT2
    from neural_arch.optim import AdamW
    cores = host_threads()
    prev_device = get_default_device()
    set_default_device(device)      # "cpu": the reference's NumPy backend on the host cores (the CPU arm)
    try:
        return _reference_steps(Tensor, cross_entropy_loss, GPT2, AdamW, cores, steps, warmup, sample_batch, device)
    finally:
        set_default_device(prev_device)


def _reference_steps(Tensor, cross_entropy_loss, GPT2, AdamW, cores, steps, warmup, sample_batch, device):
    np.random.seed(0)
    model = GPT2(GPT2_SMALL)
    opt = AdamW(model.parameters(), lr=5e-4, weight_decay=0.1)
    rng = np.random.default_rng(1234)
    ids, labels = synthetic_batch(rng, sample_batch, SEQ)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = model(Tensor(ids))
        logits = out["logits"] if isinstance(out, dict) else out
        v = logits.shape[-1]
        loss = cross_entropy_loss(Tensor(logits.data.reshape(-1, v), requires_grad=True), Tensor(labels.reshape(-1)))
        loss.backward()
        opt.step()
        float(loss.data)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    where = f"{cores} BLAS threads" if device == "cpu" else f"Device.{device} tensors = the b200 backend through the reference's own call sites"
    return {"value": sample_batch * SEQ / t, "unit": "tokens/s", "cores": cores, "kind": "reference", "step_s": t,
            "sample": f"unmodified reference (oracle/_ref, stock Numba-JIT / fusion config, {where}), gpt2_small, the training step of "
                      f"examples/training/gpt2_training.py (forward + loss; its backward stops at the logits, SURVEY F6), {sample_batch} of {BATCH} sequences "
                      f"x {SEQ} tokens per step, median of {steps} after {warmup} warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # EXACTLY --steps timed steps after --warmup untimed ones, like the other arm; each step is a bounded sample of the workload
    # (whole sequences of the batch) sized from a one-sequence probe so that the run ends within a few minutes
    steps, warm = max(1, args.steps), max(0, args.warmup)
    run = cpu_reference_run if os.path.isdir(REF_DIR) else cpu_oracle_run
    probe = run(1, 1, 1)
    budget_s = float(os.environ.get("NFB200_REFERENCE_BUDGET_S", "120"))
    sample = int(budget_s / ((steps + warm) * max(probe["step_s"], 1e-3)))
    sample = max(1, min(BATCH, sample))
    r = run(steps, warm, sample)
    desc = r["sample"]
    line = {"impl": "reference", "metric": "training tokens/sec", "value": r["value"], "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": r["step_s"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "GPT-2 small (reference variant: RoPE+RMSNorm+SwiGLU, tied head) causal-LM train step, batch 8 x seq 512", "sample": desc},
            "cpu_baseline": {"value": r["value"], "unit": "tokens/s", "cores": r["cores"], "kind": r["kind"], "sample": desc},
            "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ per-class work accounting
DT_SIZE = {0: 4, 1: 8, 2: 4, 3: 8, 4: 1, 5: 2}   # NfbDType -> bytes


def _attn_flops(B, H, S, Skv, D, mode, backward):
    f = 4.0 * B * H * S * Skv * D * (2.0 if backward else 1.0)    # fwd 4 S^2 hd, bwd 2x (the QK^T recompute is not counted)
    return f * (0.5 if (mode & 1) else 1.0)                        # causal: halved (stated in the line)


def work_of(name, a):
    """(class, 'flop' | 'byte', ALGORITHMIC amount) of one C-ABI call from its arguments (include/nfb200.h signatures)."""
    if name == "nfb_gemm_bf16":
        return "gemm", "flop", 2.0 * a[2] * a[3] * a[4]
    if name == "nfb_gemm_bf16_grouped":
        return "gemm", "flop", sum(2.0 * a[3][g] * a[4][g] * a[5][g] for g in range(a[0]))
    if name == "nfb_flash_attn_fwd_gqa":
        return "attention", "flop", _attn_flops(a[6], a[7], a[9], a[20], a[10], a[22], False)
    if name == "nfb_flash_attn_bwd_gqa":
        return "attention", "flop", _attn_flops(a[10], a[11], a[13], a[24], a[14], a[26], True)
    if name == "nfb_norm_fwd":      # read x (+ residual, write the sum), write y
        n = float(a[11]) * a[12]
        return "norm", "byte", n * (DT_SIZE[a[1]] + DT_SIZE[a[2]] + (2 * DT_SIZE[a[1]] if a[4] else 0))
    if name == "nfb_norm_bwd":      # read dy, x (+ skip gradient), write dx (+ bf16 twin)
        n = float(a[13]) * a[14]
        return "norm", "byte", n * (DT_SIZE[a[2]] + 2 * DT_SIZE[a[1]] + (DT_SIZE[a[1]] if a[8] else 0) + (2 if a[16] else 0))
    if name == "nfb_cross_entropy_ex":
        return "cross_entropy", "byte", float(a[7]) * a[8] * (DT_SIZE[a[0]] + (DT_SIZE[a[6]] if a[5] else 0))
    if name in ("nfb_adam_step", "nfb_adam_step_ex"):   # p r+w, g r, m r+w, v r+w (+ bf16 shadow w)
        msz = 8 if a[1] == 1 else 4
        return "optimizer", "byte", float(a[7]) * (8 + 4 + 4 * msz + (2 if a[6] else 0))
    if name == "nfb_swiglu_fwd":
        return "elementwise", "byte", float(a[4]) * a[5] * 3 * DT_SIZE[a[0]]
    if name == "nfb_swiglu_bwd":
        return "elementwise", "byte", float(a[6]) * a[7] * 5 * DT_SIZE[a[0]]
    if name == "nfb_rope":
        return "elementwise", "byte", float(a[6]) * a[7] * a[8] * a[9] * max(1, a[4]) * 2 * DT_SIZE[a[0]]
    if name == "nfb_colsum":
        return "elementwise", "byte", float(a[2]) * a[3] * DT_SIZE[a[0]]
    if name == "nfb_gather_rows":
        return "elementwise", "byte", float(a[6]) * a[7] * (DT_SIZE[a[1]] + DT_SIZE[a[2]])
    if name == "nfb_scatter_add_rows":
        return "elementwise", "byte", float(a[5]) * a[6] * (DT_SIZE[a[1]] + 8)
    if name == "nfb_cast":
        return "elementwise", "byte", float(a[4]) * (DT_SIZE[a[0]] + DT_SIZE[a[1]])
    if name == "nfb_unary":
        return "elementwise", "byte", float(a[4]) * 2 * DT_SIZE[a[1]]
    if name == "nfb_unary_bwd":
        return "elementwise", "byte", float(a[5]) * 3 * DT_SIZE[a[1]]
    if name == "nfb_binary_scalar":
        return "elementwise", "byte", float(a[5]) * 2 * DT_SIZE[a[1]]
    if name == "nfb_add3":
        return "elementwise", "byte", float(a[4]) * 16
    return None


PROFILED = ("nfb_gemm_bf16", "nfb_gemm_bf16_grouped", "nfb_flash_attn_fwd_gqa", "nfb_flash_attn_bwd_gqa", "nfb_norm_fwd", "nfb_norm_bwd",
            "nfb_cross_entropy_ex", "nfb_adam_step", "nfb_adam_step_ex", "nfb_swiglu_fwd", "nfb_swiglu_bwd", "nfb_rope", "nfb_colsum", "nfb_gather_rows",
            "nfb_scatter_add_rows", "nfb_cast", "nfb_unary", "nfb_unary_bwd", "nfb_binary_scalar", "nfb_add3")


# ------------------------------------------------------------------------------------------ this repo's arm
class Job:
    """One model + optimizer (+ DDP wrapper) + a persistent synthetic batch + the step closure."""

    def __init__(self, model, opt, ddp, args_kw, loss_of=None):
        self.model, self.opt, self.ddp = model, opt, ddp
        self.fwd = ddp if ddp is not None else model
        self.args, self.kwargs = args_kw
        self.loss_of = loss_of or (lambda out: out["loss"])

    def step(self):
        self.opt.zero_grad()
        loss = self.loss_of(self.fwd(*self.args, **self.kwargs))
        loss.backward()
        if self.ddp is not None:
            self.ddp.step_optimizer(self.opt)   # bucket-pipelined: finish_gradient_sync() + opt.step()
        else:
            self.opt.step()
        return loss


def run_b200(args):
    import nfb200
    from nfb200 import _lib, kernels as K, runtime as R
    from nfb200.array import B200Array
    from nfb200.core import Tensor, set_default_device
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    from nfb200.training import GraphedStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.init(local)
    if not nfb200.get_backend("b200").is_available:
        raise RuntimeError("bench.py: the b200 backend is unavailable (no libnfb200.so or no GPU); there is no CPU fallback")
    dev = f"cuda:{local}"
    set_default_device(dev)
    K.set_precision(args.precision)
    K.set_wgrad_overlap(args.wgrad_overlap)
    K.set_wgrad_group(args.wgrad_group)
    if args.grad_dtype == "auto":
        args.grad_dtype = "bfloat16" if (world >= 4 and args.precision == "bf16") else "float32"
    if world > 1:
        from nfb200.distributed import DistributedDataParallel, all_reduce_max_host, init_process_group
        from nfb200.distributed import barrier as dist_barrier
        init_process_group("nccl")
        if args.gemm_sm_limit >= 0:
            # data parallel: the persistent GEMM grids leave the other SMs to NCCL's channel CTAs while buckets are on the wire
            _lib.call("nfb_gemm_bf16_set_sm_limit", args.gemm_sm_limit)

    def barrier():
        if world > 1:
            dist_barrier()
        R.synchronize()

    def wrap(model):
        if world <= 1:
            return None
        return DistributedDataParallel(model, bucket_size_mb=args.bucket_mb, grad_dtype=args.grad_dtype if args.precision == "bf16" else "float32",
                                       sparse_embedding_grads=args.sparse_embedding_grads)

    def timed(fn, n):
        e0, e1 = R.Event(), R.Event()
        barrier()
        e0.record()
        out = None
        for _ in range(n):
            out = fn()
        e1.record()
        e1.synchronize()
        barrier()
        ms = e0.elapsed_ms(e1)
        if world > 1:
            ms = all_reduce_max_host(ms)
        return ms, out

    def capture(job, warm):
        for _ in range(warm):
            job.step()
        barrier()
        if not args.graph:
            return job.step
        g = GraphedStep(job.step, warmup=2, optimizers=[job.opt])
        for _ in range(2):
            g()
        barrier()
        return g

    # ====================================================================== main config: GPT-2 small B8 x S512
    np.random.seed(0)  # identical replicas on every rank (the reference relies on seeding too)
    model = GPT2(GPT2_SMALL)
    ddp = wrap(model)
    opt = AdamW(model.parameters(), lr=5e-4, weight_decay=0.1, moment_dtype="float32")
    if ddp is None and args.step_in_backward:
        opt.enable_step_in_backward(min_run_elems=args.sib_min)   # AdamW launches ride the side stream behind their weight-gradient GEMMs
    rng = np.random.default_rng(1234 + rank)  # each rank gets its own shard: weak scaling, per-GPU batch fixed
    ids_h, labels_h = synthetic_batch(rng)
    job = Job(model, opt, ddp, ((Tensor(ids_h, device=dev),), {"labels": Tensor(labels_h, device=dev)}))
    run = capture(job, max(args.warmup, 3))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _l0 = _lib.launch_count()
    job.step()                                   # kernels per step (counted on an eager step; a replay launches the same kernels)
    launches_per_step = _lib.launch_count() - _l0
    # ---- timed region 1: inputs resident in HBM (the step streams >> L2 (126 MB) of weights / activations / optimizer state)
    ms, loss = timed(run, args.steps)
    final_loss = float(loss.item())

    # ---- timed region 1b: per-kernel measurement on a SINGLE-STREAM capture of the same step (with the weight-gradient branch
    # on its side stream two kernels share the SMs and a pair of events no longer brackets one kernel's own duration):
    # CUDA-event pairs (external event-record nodes of the graph) around every profiled C-ABI call + the GEMM kernel's own
    # %globaltimer span; durations are those of the last replay.  Kept apart from region 1 so the event nodes do not perturb `value`.
    prof, spans, ms_prof = None, None, None
    if args.profile and args.graph:
        try:
            K.set_wgrad_overlap(False)
            job.step()   # warm the allocator in this mode
            cap = 4096
            span_buf = B200Array.empty((2 * cap,), np.int64)
            records, spans_meta = [], []
            _lib.PROFILE_NAMES = frozenset(PROFILED)

            def profiled_step():
                _lib.PROFILE = records
                K.GEMM_SPAN = {"buf": span_buf, "cap": cap, "meta": spans_meta}
                try:
                    return job.step()
                finally:
                    _lib.PROFILE = None
                    K.GEMM_SPAN = None

            gprof = GraphedStep(profiled_step, warmup=1, optimizers=[opt])
            prof = records[len(records) // 2:]   # GraphedStep ran the step twice (eager warm-up, then the capture): the captured half
            gprof()
            _lib.call("nfb_memset", span_buf.ptr, 0xFF, 2 * cap * 8)   # arm: the next replay stamps min(start) / ~max(end)
            R.synchronize()
            ms_prof, _ = timed(gprof, 1)
            R.synchronize()
            raw = span_buf.get().view(np.uint64).reshape(cap, 2)
            armed = np.uint64(0xFFFFFFFFFFFFFFFF)
            spans = [float(np.uint64(~raw[i, 1]) - raw[i, 0]) * 1e-6 for i in range(cap) if raw[i, 0] != armed]   # ns -> ms, in launch order
            _ = sum(a.elapsed_ms(b) for _, _, a, b in prof)   # raises if captured events cannot be read
            K.set_wgrad_overlap(args.wgrad_overlap)
            if rank != 0:
                prof = None
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"bench.py: per-kernel profile unavailable ({type(exc).__name__}: {exc})\n")
            _lib.PROFILE = None
            K.GEMM_SPAN = None
            K.set_wgrad_overlap(args.wgrad_overlap)
            prof = None
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2 (e2e): ids/labels come from PINNED host memory every step, the loss goes back to the host
    import ctypes as C
    nbytes = ids_h.nbytes
    pin = C.c_void_p()
    _lib.call("nfb_host_alloc", 2 * nbytes + 64, C.byref(pin))
    pin_ids = np.frombuffer((C.c_char * nbytes).from_address(pin.value), dtype=np.int64).reshape(ids_h.shape)
    pin_lab = np.frombuffer((C.c_char * nbytes).from_address(pin.value + nbytes), dtype=np.int64).reshape(ids_h.shape)
    pin_ids[...] = ids_h
    pin_lab[...] = labels_h
    loss_host = np.zeros(1, np.float32)
    dev_ids, dev_lab = B200Array.empty(ids_h.shape, np.int64), B200Array.empty(ids_h.shape, np.int64)
    job_e2e = Job(model, opt, ddp, ((Tensor(dev_ids),), {"labels": Tensor(dev_lab)}))
    run_e2e = GraphedStep(job_e2e.step, warmup=1, optimizers=[opt]) if args.graph else job_e2e.step
    e2e_steps = max(2, min(args.steps, 10))

    def e2e_step():
        _lib.call("nfb_memcpy_h2d_async", dev_ids.ptr, pin_ids.ctypes.data, nbytes)
        _lib.call("nfb_memcpy_h2d_async", dev_lab.ptr, pin_lab.ctypes.data, nbytes)
        l = run_e2e()
        _lib.call("nfb_memcpy_d2h", loss_host.ctypes.data, l.data.ptr, 4)  # blocks until the step's loss is on the host

    ms_e2e, _ = timed(e2e_step, e2e_steps)
    _lib.call("nfb_host_free", pin)
    del run, run_e2e, job, job_e2e, model, opt, ddp
    R.empty_cache()

    # ====================================================================== other configs (BASELINE.json configs[2..4])
    other = {}
    if args.other_configs:
        def other_config(key, build, batch, units, unit, flop_per_unit, lr, wd, steps):
            try:
                np.random.seed(0)
                m = build()
                d = wrap(m)
                o = AdamW(m.parameters(), lr=lr, weight_decay=wd, moment_dtype="float32")
                j = Job(m, o, d, batch())
                r = capture(j, 3)
                t_ms, l = timed(r, steps)
                per = t_ms / steps
                other[key] = {"ms_per_step": per, "value": units * world * steps / (t_ms / 1e3), "unit": unit, "per_gpu_units": units, "n_gpus": world,
                              "model_tflops_per_gpu": units * flop_per_unit / (per / 1e3) / 1e12, "loss": float(l.item()), "steps": steps}
                del r, j, o, d, m
            except Exception as exc:  # noqa: BLE001 -- one config failing must not take the headline line down
                other[key] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            R.synchronize()
            R.empty_cache()

        from nfb200.models import BERTConfig, BERTForMaskedLM, gpt2_medium, vit_b_16
        orng = np.random.default_rng(4321 + rank)

        def bert_batch():
            ids = orng.integers(0, 30522, (32, 512))
            return (Tensor(ids, device=dev),), {"attention_mask": Tensor(np.ones((32, 512), np.float32), device=dev), "labels": Tensor(ids, device=dev)}

        def vit_batch():
            img = orng.standard_normal((64, 3, 224, 224)).astype(np.float32)
            return (Tensor(img, device=dev),), {"labels": Tensor(orng.integers(0, 1000, (64,)), device=dev)}

        def g2m_batch():
            ids, lab = synthetic_batch(orng, 8, 1024)
            return (Tensor(ids, device=dev),), {"labels": Tensor(lab, device=dev)}

        other_config("bert_base_mlm_b32_s512_bf16", lambda: BERTForMaskedLM(BERTConfig()), bert_batch, 32 * 512, "tokens/s", 706.88e6, 2e-5, 0.01, 8)
        other_config("vit_b16_b64_224", lambda: vit_b_16(num_classes=1000), vit_batch, 64, "images/s", 105.38e9, 1e-3, 0.1, 8)
        if world == 8 or args.gpt2_medium:
            other_config("gpt2_medium_b8_s1024_per_gpu", gpt2_medium, g2m_batch, 8 * 1024, "tokens/s", 2120.72e6, 5e-4, 0.1, 6)

    # ====================================================================== data-parallel parity (N > 1)
    ddp_parity = None
    if world > 1 and args.ddp_parity:
        try:
            ddp_parity = ddp_parity_check(rank, world, dev)
        except Exception as exc:  # noqa: BLE001
            ddp_parity = {"error": f"{type(exc).__name__}: {str(exc)[:300]}", "ok": False}

    if rank != 0:
        return
    tokens_per_step = BATCH * SEQ * world
    value = tokens_per_step * args.steps / (ms / 1e3)
    e2e_value = tokens_per_step * e2e_steps / (ms_e2e / 1e3)
    pk = peaks()
    roof, classes = None, None
    if prof:
        per_class = {}
        gemm_ev = []
        for name, a, e0, e1 in prof:
            w = work_of(name, a)
            if w is None:
                continue
            cls, kind, amount = w
            t = e0.elapsed_ms(e1)
            c = per_class.setdefault(cls, {"kind": kind, "work": 0.0, "ms": 0.0, "n": 0})
            c["work"] += amount
            c["ms"] += t
            c["n"] += 1
            if cls == "gemm":
                gemm_ev.append((amount, t))
        # GEMM: event-pair durations vs the kernels' own %globaltimer spans (same launches, same replay)
        n_g = len(gemm_ev)
        span_ms = spans[-n_g:] if spans and len(spans) >= n_g else None
        tot_flop = sum(f for f, _ in gemm_ev)
        tot_ms = sum(t for _, t in gemm_ev)
        achieved = tot_flop / (tot_ms / 1e3) / 1e12
        inflation_us = None
        roof = {"bound": "tensor", "kernel": "k_gemm_tc (tcgen05/TMEM bf16 GEMM; the four weight gradients of a block are one grouped stream-K launch)",
                "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
                "achieved_event_pairs": achieved, "frac_event_pairs": achieved / pk["bf16_tflops_sustained"], "avg_launch_ms_event_pairs": tot_ms / n_g,
                "traffic": GEMM_DRAM_BYTES_PER_LAUNCH,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per k_gemm_tc launch, ncu --set full capture of this command, "
                                  "launch-weighted mean over the launches of one step (profiles/r02_ncu_gemm_full.txt)",
                "peak_source": pk["source"] + " (sustained: kernel timed inside a long step)",
                "launches_timed": n_g, "avg_launch_ms": tot_ms / n_g, "avg_launch_gflop": tot_flop / n_g / 1e9, "share_of_step": tot_ms / ms_prof,
                "timing": "CUDA event pairs (external event-record nodes) captured into a single-stream graph of the same step around every GEMM launch; durations of the last replay",
                "algorithmic_flops": "2*M*N*K per contraction (fwd, dgrad and wgrad GEMMs of every Linear + logits), summed over the launches of one step"}
        if span_ms:
            sp_tot = sum(span_ms)
            inflation_us = (tot_ms - sp_tot) / n_g * 1e3
            dc = tot_flop / (sp_tot / 1e3) / 1e12
            # headline = the kernel's own duration on the device clock (what ncu's gpu__time_duration measures, taken live and warm);
            # the event-pair figure is kept next to it: an event-record NODE on each side of a kernel node costs several us of
            # graph-dependency latency, which is not time the kernel runs
            roof.update({"achieved": dc, "frac": dc / pk["bf16_tflops_sustained"], "avg_launch_ms": sp_tot / n_g, "share_of_step": sp_tot / ms_prof,
                         "event_pair_inflation_us_per_launch": inflation_us,
                         "timing": "per launch, live inside the timed single-stream replay of the step: the kernel's own %globaltimer span (first CTA ENTRY -> last CTA "
                                   "exit, prologue included: the quantity ncu reports as gpu__time_duration), CUDA events bracketing the replay; "
                                   "*_event_pairs = CUDA event pairs (external event-record graph nodes) around the same launches of the same replay, "
                                   "which add event_pair_inflation_us_per_launch of graph-node latency to every launch"})
        classes = {}
        for cls, c in per_class.items():
            tensor = c["kind"] == "flop"
            peak = pk["bf16_tflops_sustained"] if tensor else pk["hbm_gbs"]
            ach = c["work"] / (c["ms"] / 1e3) / (1e12 if tensor else 1e9)
            entry = {"bound": "tensor" if tensor else "hbm", "achieved": ach, "peak": peak, "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak,
                     "launches": c["n"], "ms_per_step": c["ms"], "share_of_step": c["ms"] / ms_prof}
            if inflation_us is not None and inflation_us > 0:   # the same event-node latency sits on every event pair: net of it
                net = c["ms"] - c["n"] * inflation_us * 1e-3
                if net > 0:
                    entry["frac_net_of_event_latency"] = c["work"] / (net / 1e3) / (1e12 if tensor else 1e9) / peak
            classes[cls] = entry
        classes["_notes"] = ("algorithmic work per call from its arguments (SURVEY 8d): GEMM 2MNK; attention 4*S*Skv*hd forward, 2x backward, HALVED for causal; "
                             "norm / cross-entropy / AdamW / elementwise: bytes read + written once at their storage types (bf16 activations, fp32 residual stream / "
                             "master weights / moments); durations = CUDA event pairs in the single-stream replay (each pair carries the event-node latency "
                             "reported as roofline.event_pair_inflation_us_per_launch; frac_net_of_event_latency subtracts it)")
    line = {"metric": "training tokens/sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": "GPT-2 small (reference variant: RoPE+RMSNorm+SwiGLU, tied head) causal-LM train step (fwd + CE + bwd + AdamW)",
                       "per_gpu_batch": BATCH, "seq_len": SEQ, "global_batch": BATCH * world, "parallelism": f"dp{world}",
                       **({"gradient_buckets": f"{args.bucket_mb:g} MB NCCL all-reduce (ncclAvg, {args.grad_dtype if args.precision == 'bf16' else 'float32'} on the wire, fp32 arena) "
                                               "overlapped with backward, AdamW pipelined behind them",
                           "gemm_sm_limit": args.gemm_sm_limit} if world > 1 else {}),
                       "precision": "bf16 tensor-core GEMM/attention operands, fp32 accumulate / residual stream / master weights / Adam moments" if args.precision == "bf16" else "fp32",
                       "l2": "no explicit flush: each step streams ~3 GB of weights/optimizer state and ~5 GB of activations, far above the 126 MB L2",
                       "grad_clip": "reference +-10 clip on fp32 gradients and in the dgrad GEMM epilogues"},
            "model_tflops": value * TRAIN_FLOP_PER_TOKEN / 1e12, "mfu_vs_sustained_peak": value * TRAIN_FLOP_PER_TOKEN / 1e12 / (pk["bf16_tflops_sustained"] * world),
            "final_loss": final_loss,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": int(2 * nbytes), "d2h_bytes_per_step": 4, "steps": e2e_steps},
            "gpu_launches": int(launches_per_step * args.steps), "cuda_graph": bool(args.graph), "clocks": clocks}
    if roof is not None:
        line["roofline"] = roof
        line["roofline_classes"] = classes
    if other:
        line["other_configs"] = other
    if ddp_parity is not None:
        line["ddp_parity"] = ddp_parity
    if world == 1 and not args.no_cpu_baseline:
        port = cpu_oracle_run(1, 0, 2)
        if os.path.isdir(REF_DIR):
            ref = cpu_reference_run(1, 1, 1)
            line["cpu_baseline"] = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["port"] = {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")}
            try:   # the same UNMODIFIED reference step with Device.cuda tensors: its own code on the b200 backend (the drop-in)
                from nfb200.array import set_implicit_host_conversion
                set_implicit_host_conversion(True)
                drop = cpu_reference_run(2, 1, BATCH, device=dev)
                line["reference_on_b200"] = {"value": drop["value"], "unit": "tokens/s", "sample": drop["sample"]}
            except Exception as exc:  # noqa: BLE001
                line["reference_on_b200"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
        else:
            line["cpu_baseline"] = {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)


def ddp_parity_check(rank, world, dev):
    """N ranks, each on its own shard, against ONE process on the global batch (distributed/data_parallel.py:271-287: the
    all-reduce(AVERAGE) of per-shard mean gradients is the global-batch mean gradient): first-step gradients of every
    parameter and a 3-step AdamW loss trajectory, fp32 contraction path so that the comparison is at the 1e-5 level.
    Every rank runs the data-parallel job; rank 0 also runs the single-process one (it can draw every rank's shard)."""
    from nfb200 import kernels as K
    from nfb200.core import Tensor
    from nfb200.distributed import DistributedDataParallel, all_reduce_max_host, get_backend
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    cfg = {"vocab_size": 4099, "n_positions": 128, "n_embd": 256, "n_layer": 2, "n_head": 4}
    B, S, steps = 2, 128, 3
    prev = K.get_precision()
    K.set_precision("fp32")
    try:
        def shard(r, step):
            g = np.random.default_rng(9000 + 31 * step + r)
            ids = g.integers(0, cfg["vocab_size"], (B, S))
            return ids, np.concatenate([ids[:, 1:], np.zeros((B, 1), dtype=ids.dtype)], axis=1)

        def run(parallel):
            np.random.seed(0)
            m = GPT2(cfg)
            d = DistributedDataParallel(m, bucket_size_mb=1.0) if parallel else None
            o = AdamW(m.parameters(), lr=5e-4, weight_decay=0.1, moment_dtype="float64")
            losses, grads = [], None
            for s in range(steps):
                if parallel:
                    ids, lab = shard(rank, s)
                else:
                    parts = [shard(r, s) for r in range(world)]
                    ids, lab = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
                o.zero_grad()
                loss = (d if parallel else m)(Tensor(ids, device=dev), labels=Tensor(lab, device=dev))["loss"]
                loss.backward()
                if parallel:
                    d.finish_gradient_sync()
                if s == 0:
                    grads = {n: p.grad.get().copy() for n, p in m.named_parameters()}
                o.step()
                losses.append(float(loss.item()))
            return losses, grads

        par_l, par_g = run(True)
        # the data-parallel loss of a rank is its SHARD's loss: the global-batch loss is the mean over ranks
        g = get_backend()
        mean_l = [float(np.mean(g.tcp.allgather(float(x)))) for x in par_l]
        res = None
        if rank == 0:
            one_l, one_g = run(False)
            gmax = max(float(np.abs(par_g[n] - one_g[n]).max() / (np.abs(one_g[n]).max() + 1e-30)) for n in one_g)
            lmax = max(abs(a - b) / abs(b) for a, b in zip(mean_l, one_l))
            res = {"grad_max_rel": gmax, "loss_max_rel": lmax, "ok": bool(gmax <= 1e-5 and lmax <= 1e-5), "steps": steps,
                   "what": f"{world} ranks x ({B} x {S}) shards vs one process on the global ({B * world} x {S}) batch, GPT-2 d256 L2 H4 V4099, fp32 path; "
                           "first-step gradients relative to each tensor's largest entry, 3-step AdamW loss trajectory (mean over ranks)"}
        all_reduce_max_host(0.0)   # keep the ranks together
        return res
    finally:
        K.set_precision(prev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bucket-mb", type=float, default=100.0,
                    help="gradient all-reduce bucket size (the reference's default is 25; 100 measured 2.5 %% faster at 4 GPUs and 3.5 %% at 8, profiles/r01_ddp_explore.txt)")
    ap.add_argument("--grad-dtype", default="auto", choices=["auto", "float32", "bfloat16"],
                    help="data parallel, bf16 precision: dtype of the gradient buckets on the wire (fp32 arena either way); auto = bfloat16 from 4 GPUs on "
                         "(measured: 6.30 vs 6.46 ms at 8 GPUs, 6.14 vs 6.11 ms at 2, profiles/r02_ddp_sweep.txt)")
    ap.add_argument("--sparse-embedding-grads", action="store_true", help="data parallel: exchange embedding-table row gradients as (ids, rows)")
    ap.add_argument("--gemm-sm-limit", type=int, default=-1,
                    help="data parallel: SMs the persistent GEMM grids may occupy (leaves the rest to NCCL's channel CTAs); -1 = all")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every kernel from Python instead of replaying a captured CUDA graph")
    ap.add_argument("--no-profile", "--no-profile-gemm", dest="profile", action="store_false", help="skip the per-kernel roofline pass")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false", help="skip BERT-base / ViT-B/16 (/ GPT-2 medium at 8 GPUs)")
    ap.add_argument("--gpt2-medium", action="store_true", help="run the GPT-2 medium per-GPU shard (B8 x S1024) at any N (default: only at 8 GPUs)")
    ap.add_argument("--no-ddp-parity", dest="ddp_parity", action="store_false")
    ap.add_argument("--step-in-backward", dest="step_in_backward", action="store_true",
                    help="AdamW launches per layer from inside backward on the side stream (measured: within 0.6 %% of the default, one launch after backward)")
    ap.add_argument("--sib-min", type=int, default=1 << 21, help="step-in-backward: smallest contiguous run of finished parameters that gets its own AdamW launch")
    ap.add_argument("--wgrad-group", type=int, default=4, help="weight-gradient GEMMs per grouped stream-K launch (1 = one launch each)")
    ap.add_argument("--no-wgrad-overlap", dest="wgrad_overlap", action="store_false", help="keep the weight-gradient GEMMs on the compute stream")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
