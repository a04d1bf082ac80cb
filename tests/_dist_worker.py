"""Worker for the multi-process distributed tests (spawned by tests/test_distributed_cpu.py and the GPU DDP test).
Usage: python _dist_worker.py <mode> <out_dir>   with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the env."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neural-network-from-scratch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import nfb200  # noqa: E402
from nfb200 import distributed as dist  # noqa: E402
from nfb200.core import Tensor, set_default_device  # noqa: E402


def collectives(out):
    rank, world = dist.get_rank(), dist.get_world_size()
    x = np.arange(6, dtype=np.float32).reshape(2, 3) + 10 * rank
    res = {}
    res["sum"] = dist.all_reduce(Tensor(x.copy()), dist.ReduceOp.SUM).numpy().tolist()
    res["avg"] = dist.all_reduce(Tensor(x.copy()), dist.ReduceOp.AVERAGE).numpy().tolist()
    res["max"] = dist.all_reduce(Tensor(x.copy()), dist.ReduceOp.MAX).numpy().tolist()
    res["gather"] = [t.numpy().tolist() for t in dist.all_gather(Tensor(x.copy()))]
    res["bcast"] = dist.broadcast(Tensor(x.copy()), src=1 if world > 1 else 0).numpy().tolist()
    y = np.arange(4 * world, dtype=np.float32).reshape(2 * world, 2) * (rank + 1)
    res["rs"] = dist.reduce_scatter(Tensor(y), dist.ReduceOp.SUM).numpy().tolist()
    dist.barrier()
    res["sampler"] = list(dist.DistributedSampler(10, shuffle=True, seed=3))
    json.dump(res, open(os.path.join(out, f"coll_{rank}.json"), "w"))


def helpers(out):
    """gather / scatter / send / recv / no_sync / info helpers of distributed/data_parallel.py:391-476."""
    rank, world = dist.get_rank(), dist.get_world_size()
    res = {"info": dist.get_distributed_info(), "avail": dist.is_distributed_training_available()}
    x = np.arange(4, dtype=np.float32) + 100 * rank
    g = dist.gather(Tensor(x), dst=1)
    res["gather"] = None if g is None else [t.numpy().tolist() for t in g]
    pieces = [Tensor(np.full((2, 2), 7.0 + r, np.float32)) for r in range(world)] if rank == 1 else None
    res["scatter"] = dist.scatter(pieces, src=1).numpy().tolist()
    if rank == 1:
        try:
            dist.scatter([Tensor(np.zeros(1, np.float32))] * (world + 1), src=1)
            res["scatter_err"] = False
        except ValueError:
            res["scatter_err"] = True
    buf = Tensor(np.zeros(3, np.float32))
    if rank == 0:
        dist.send(Tensor(np.array([1.0, 2.0, 3.0], np.float32)), dst=1, tag=5)
        res["recv"] = dist.recv(buf, src=1, tag=6).numpy().tolist()
    else:
        got = dist.recv(buf, src=0, tag=5).numpy()
        dist.send(Tensor(got * 2), dst=0, tag=6)
        res["recv"] = got.tolist()
    # no_sync: two local backward passes accumulate, the first synchronised one averages the accumulated sums
    from nfb200 import nn
    np.random.seed(0)
    lin = nn.Linear(3, 2)
    ddp = dist.DistributedDataParallel(lin)
    xin = Tensor(np.full((1, 3), 1.0 + rank, np.float32))
    lin.zero_grad()
    with dist.no_sync():
        ddp(xin).sum().backward()
        res["local"] = np.asarray(lin.weight.grad).tolist()
    ddp(xin).sum().backward()
    ddp.sync_gradients()
    res["synced"] = np.asarray(lin.weight.grad).tolist()
    dist.barrier()
    json.dump(res, open(os.path.join(out, f"help_{rank}.json"), "w"))


def ddp_train(out, device):
    """2 steps of data-parallel training of a tiny GPT-2; every rank sees its own half of the global batch."""
    from nfb200.models import GPT2
    from nfb200.optim import AdamW
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = {"vocab_size": 101, "n_positions": 64, "n_embd": 128, "n_layer": 2, "n_head": 2}
    set_default_device(device)
    np.random.seed(0)
    model = GPT2(cfg)
    ddp = dist.DistributedDataParallel(model, bucket_size_mb=0.25, sparse_embedding_grads=bool(int(os.environ.get("NFB200_TEST_SPARSE", "0"))),
                                       grad_dtype=os.environ.get("NFB200_TEST_GRAD_DTYPE", "float32"))
    which = os.environ.get("NFB200_TEST_OPT", "adamw")
    if which == "lion":
        from nfb200.optim import Lion
        opt = Lion(model.parameters(), lr=1e-4, weight_decay=0.01)
    elif which == "sgd":
        from nfb200.optim import SGD
        opt = SGD(model.parameters(), lr=1e-2, momentum=0.9)
    else:
        opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, moment_dtype="float64")
    rng = np.random.default_rng(7)
    losses = []
    for step_i in range(2):
        GB = 4  # global batch, split evenly over the ranks
        ids = rng.integers(0, 101, (GB, 32))
        labels = np.concatenate([ids[:, 1:], np.zeros((GB, 1), dtype=ids.dtype)], axis=1)
        per = GB // world
        sl = slice(rank * per, (rank + 1) * per)
        opt.zero_grad()
        loss = ddp(Tensor(ids[sl], device=device), labels=Tensor(labels[sl], device=device))["loss"]
        loss.backward()
        if step_i == 0:
            ddp.finish_gradient_sync()   # the reference's two-call form (data_parallel.py:271-287)
            opt.step()
        else:
            ddp.step_optimizer(opt)      # the optimizer kernel pipelined behind the bucket all-reduces
        losses.append(float(loss.item()))
    sd = {k: v.tolist() for k, v in list(model.state_dict().items())[:3]}
    json.dump({"losses": losses, "params": sd, "n_buckets": len(ddp.buckets)}, open(os.path.join(out, f"ddp_{rank}.json"), "w"))


if __name__ == "__main__":
    mode, out = sys.argv[1], sys.argv[2]
    backend = "nccl" if mode.endswith("gpu") else "gloo"
    dist.init_process_group(backend)
    if mode == "collectives":
        collectives(out)
    elif mode == "helpers":
        helpers(out)
    elif mode == "ddp_cpu":
        ddp_train(out, "cpu")
    elif mode == "ddp_gpu":
        if os.environ.get("NFB200_TEST_PRECISION"):
            from nfb200 import kernels as K
            K.set_precision(os.environ["NFB200_TEST_PRECISION"])
        ddp_train(out, f"cuda:{os.environ.get('LOCAL_RANK', '0')}")
    dist.destroy_process_group()
