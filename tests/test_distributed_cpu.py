"""Multi-process (world_size 2) tests of the distributed host logic on CPU: real TCP collectives ("gloo" backend of this
repo), DistributedSampler (bit-identical to distributed/data_parallel.py:304-388), DistributedDataParallel gradient
averaging == single-process training on the global batch (SURVEY 8e equivalence check)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "_dist_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(mode, out_dir, world=2, extra_env=None):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   NFB200_PORT_OFFSET="0")
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, WORKER, mode, str(out_dir)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = []
    for p in procs:
        o, _ = p.communicate(timeout=240)
        outs.append(o.decode(errors="replace"))
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return outs


def test_collectives_world2(tmp_path):
    _spawn("collectives", tmp_path)
    r0 = json.load(open(tmp_path / "coll_0.json"))
    r1 = json.load(open(tmp_path / "coll_1.json"))
    x0 = np.arange(6, dtype=np.float32).reshape(2, 3)
    x1 = x0 + 10
    for r in (r0, r1):
        assert np.array_equal(r["sum"], x0 + x1)
        assert np.array_equal(r["avg"], (x0 + x1) / 2)
        assert np.array_equal(r["max"], x1)
        assert np.array_equal(r["gather"], [x0, x1])
        assert np.array_equal(r["bcast"], x1)
    y = np.arange(8, dtype=np.float32).reshape(4, 2)
    assert np.array_equal(r0["rs"], (y * 1 + y * 2)[:2]) and np.array_equal(r1["rs"], (y * 1 + y * 2)[2:])
    # sampler: disjoint rank-strided shards of the same permutation
    perm = np.random.RandomState(3).permutation(10).tolist()
    assert r0["sampler"] == perm[0::2] and r1["sampler"] == perm[1::2]


def test_gather_scatter_send_recv_no_sync_world2(tmp_path):
    """The module-level helpers of distributed/data_parallel.py:391-476 and communication.py:585-592 between two processes."""
    _spawn("helpers", tmp_path)
    r0 = json.load(open(tmp_path / "help_0.json"))
    r1 = json.load(open(tmp_path / "help_1.json"))
    for r, h in enumerate((r0, r1)):
        assert h["info"] == {"available": True, "world_size": 2, "rank": r, "backend": "gloo"} and h["avail"] is True
        assert np.array_equal(h["scatter"], np.full((2, 2), 7.0 + r))          # rank r receives piece r of rank 1's list
    x0 = np.arange(4, dtype=np.float32)
    assert r0["gather"] is None and np.array_equal(r1["gather"], [x0, x0 + 100])   # only dst = 1 keeps the list
    assert r1["scatter_err"] is True                                             # wrong list length -> ValueError (upstream)
    assert r1["recv"] == [1.0, 2.0, 3.0] and r0["recv"] == [2.0, 4.0, 6.0]        # 0 -> 1, doubled, 1 -> 0
    # no_sync: inside the block each rank holds its own gradient of sum(x W): every row of dW equals x_r (or its transpose)
    g0, g1 = np.asarray(r0["local"]), np.asarray(r1["local"])
    assert np.all(g0 == 1.0) and np.all(g1 == 2.0)
    # the next synchronised backward accumulates once more locally (2 x_r) and averages over ranks: (2*1 + 2*2) / 2 = 3
    assert np.all(np.asarray(r0["synced"]) == 3.0) and np.all(np.asarray(r1["synced"]) == 3.0)


def test_helpers_without_process_group():
    from nfb200 import distributed as dist
    from nfb200.core import Tensor
    assert not dist.is_initialized()
    t = Tensor(np.ones(3, np.float32))
    assert dist.gather(t) == [t] and dist.scatter([t]) is t
    assert dist.get_distributed_info() == {"available": False, "world_size": 1, "rank": 0, "backend": None}
    assert dist.is_distributed_training_available() is False
    with dist.no_sync():
        pass


def test_sampler_padding_and_epoch():
    from nfb200.distributed import DistributedSampler
    s0 = DistributedSampler(10, num_replicas=4, rank=0, shuffle=False)
    s3 = DistributedSampler(10, num_replicas=4, rank=3, shuffle=False)
    assert len(s0) == 3 and list(s0) == [0, 4, 8] and list(s3) == [3, 7, 1]  # padded with the head of the index list
    s = DistributedSampler(7, num_replicas=2, rank=1, shuffle=True, seed=5, drop_last=True)
    a = list(s)
    s.set_epoch(1)
    b = list(s)
    assert len(a) == 3 and a != b
    assert a == np.random.RandomState(5).permutation(7).tolist()[:6][1::2]


def test_ddp_equals_single_process_on_global_batch(tmp_path):
    """mean-of-shard-means == global mean for equal shards: 2-rank DDP must reproduce 1-process training."""
    _spawn("ddp_cpu", tmp_path, world=2)
    d0 = json.load(open(tmp_path / "ddp_0.json"))
    d1 = json.load(open(tmp_path / "ddp_1.json"))
    one = tmp_path / "one"
    one.mkdir()
    _spawn("ddp_cpu", one, world=1)
    ref = json.load(open(one / "ddp_0.json"))
    for k in d0["params"]:
        np.testing.assert_allclose(d0["params"][k], d1["params"][k], rtol=0, atol=0)      # replicas stay identical
        np.testing.assert_allclose(d0["params"][k], ref["params"][k], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose((d0["losses"][0] + d1["losses"][0]) / 2, ref["losses"][0], rtol=1e-5)
