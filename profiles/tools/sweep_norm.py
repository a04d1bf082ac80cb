"""Grid-size sweep of the RMSNorm / LayerNorm kernels at the GPT-2 small shape (T = 4096 rows of 768), same cold-L2
graph-replay method as bench_bw.py.  NFB_NORM_{FWD,BWD}_GRID are read by csrc/norm.cu on every launch."""
import os, sys, json, numpy as np
sys.path.insert(0, "neural-network-from-scratch_b200"); sys.path.insert(0, ".")
import nfb200
from nfb200 import kernels as Kn, runtime as R
from nfb200.array import B200Array, bfloat16
be = nfb200.get_backend("b200")
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6532.2
REP = 10
rng = np.random.default_rng(0)
def f32(*shape): return B200Array.from_numpy(rng.standard_normal(shape).astype(np.float32))

def timed(fn):
    def run(with_fn):
        for _ in range(REP):
            R.flush_l2()
            if with_fn: fn()
    fn(); R.synchronize()
    ga = R.Graph()
    with ga: run(True)
    gb = R.Graph()
    with gb: run(False)
    R.synchronize()
    ta = min(R.time_ms(ga.replay, iters=3, warmup=1) for _ in range(3))
    tb = min(R.time_ms(gb.replay, iters=3, warmup=1) for _ in range(3))
    return (ta - tb) / REP

for T in (4096, 16384):
    d = 768
    x = f32(T, d); w = f32(d); dres = f32(T, d); dyb = Kn.to_bf16(f32(T, d)); dw = f32(d)
    y, mean, rstd = be.norm_forward(x, w, None, 1e-5, rms=True, out_dtype=bfloat16)
    timed(lambda: be.norm_forward(x, w, None, 1e-5, rms=True, out_dtype=bfloat16))
    for g in (0, 148, 296, 444, 592, 740, 888):
        os.environ["NFB_NORM_FWD_GRID"] = str(g)
        ms = timed(lambda: be.norm_forward(x, w, None, 1e-5, rms=True, out_dtype=bfloat16))
        print(f"T={T} rmsnorm fwd grid={g or 'default':>7} {ms*1e3:7.2f} us  {T*d*6/ms/1e6/PEAK*100:5.1f}%", flush=True)
    os.environ["NFB_NORM_FWD_GRID"] = "0"
    for g in (0, 128, 148, 171, 205, 256, 296, 342, 512):
        os.environ["NFB_NORM_BWD_GRID"] = str(g)
        ms = timed(lambda: be.norm_backward(dyb, x, w, mean, rstd, rms=True, dresidual=dres, dweight_out=dw, accumulate=True, clip=10.0, bf16_twin=True))
        print(f"T={T} rmsnorm bwd grid={g or 'default':>7} {ms*1e3:7.2f} us  {T*d*16/ms/1e6/PEAK*100:5.1f}%", flush=True)
    os.environ["NFB_NORM_BWD_GRID"] = "0"
