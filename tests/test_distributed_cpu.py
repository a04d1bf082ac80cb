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


def test_launcher_config_validation():
    """distributed/launcher.py:267-286: the same ValueErrors."""
    from nfb200.distributed import DistributedLauncher, LaunchConfig, find_free_port, get_local_ip
    for kw in (dict(nproc_per_node=0), dict(nnodes=0), dict(nnodes=2, node_rank=2), dict(master_addr="not-an-ip"), dict(master_port=80)):
        with pytest.raises(ValueError):
            DistributedLauncher(LaunchConfig(**kw))
    DistributedLauncher(LaunchConfig(nproc_per_node=8, nnodes=2, node_rank=1, master_addr="10.0.0.1", master_port=29500))
    assert 1024 <= find_free_port() <= 65535 and socket.inet_aton(get_local_ip())


def test_launcher_env_contract_logs_and_fail_fast(tmp_path):
    """Ranks get RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (launcher.py:145-162), per-rank log files (:164-183); one failing
    rank ends the job at once with that rank's exit code."""
    import time
    from nfb200.distributed import launch_distributed_training
    script = tmp_path / "train.py"
    script.write_text(
        "import json, os, sys, time\n"
        "keys = ['RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT', 'NCCL_ASYNC_ERROR_HANDLING', 'OMP_NUM_THREADS']\n"
        "json.dump({k: os.environ.get(k) for k in keys} | {'argv': sys.argv[1:]}, open(os.path.join(sys.argv[1], 'env_' + os.environ['RANK'] + '.json'), 'w'))\n"
        "print('hello from', os.environ['RANK'])\n"
        "if sys.argv[2] == 'fail':\n"
        "    if os.environ['RANK'] == '3': sys.exit(7)\n"
        "    time.sleep(120)\n")
    logs = tmp_path / "logs"
    rc = launch_distributed_training(str(script), 2, 2, 1, "127.0.0.1", 29611, "auto", str(logs), str(tmp_path), "ok")
    assert rc == 0
    for local, rank in ((0, 2), (1, 3)):                      # node 1 of 2, two ranks per node
        e = json.load(open(tmp_path / f"env_{rank}.json"))
        assert e == {"RANK": str(rank), "WORLD_SIZE": "4", "LOCAL_RANK": str(local), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29611",
                     "NCCL_ASYNC_ERROR_HANDLING": "1", "OMP_NUM_THREADS": "1", "argv": [str(tmp_path), "ok"]}
        assert (logs / f"rank_{rank}_stdout.log").read_text().strip() == f"hello from {rank}"
        assert (logs / f"rank_{rank}_stderr.log").exists()
    t0 = time.time()
    rc = launch_distributed_training(str(script), 2, 2, 1, "127.0.0.1", 29611, "auto", None, str(tmp_path), "fail")
    assert rc == 7 and time.time() - t0 < 30


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


def test_wire_format_is_a_closed_set_of_value_types():
    """The rendezvous / host collectives never unpickle: values travel in a fixed binary framing (any host that can
    reach the port could otherwise execute code in the training job)."""
    from nfb200.distributed import communication as C
    vals = [None, True, False, 7, -3, 2.5, "héllo", b"\x00\x01", np.arange(6, dtype=np.float32).reshape(2, 3), np.array([1, 2], np.int64),
            [1, "a", None], (2.0, [3]), {"k": np.zeros((0, 4), np.float64), "n": {"x": 1}}]
    for v in vals:
        buf = bytearray()
        C._encode(v, buf)
        got, pos = C._decode(memoryview(bytes(buf)), 0)
        assert pos == len(buf)
        if isinstance(v, np.ndarray):
            assert got.dtype == v.dtype and np.array_equal(got, v)
        elif isinstance(v, dict):
            assert got.keys() == v.keys() and np.array_equal(got["k"], v["k"]) and got["n"] == v["n"]
        else:
            assert got == v and type(got) is type(v)
    with pytest.raises(TypeError):
        C._encode(object(), bytearray())            # arbitrary objects cannot be sent
    with pytest.raises(TypeError):
        C._encode(np.array(["a"], dtype=object), bytearray())
    import pickle
    evil = pickle.dumps({"x": 1})
    with pytest.raises(ValueError):
        C._decode(memoryview(evil), 0)              # a pickle stream is not a message
    bad = bytearray()
    C._encode(np.zeros(4, np.float32), bad)
    bad[-17] ^= 0xFF                                 # corrupt the payload length
    with pytest.raises((ValueError, Exception)):
        C._decode(memoryview(bytes(bad)), 0)


def test_rendezvous_rejects_unauthenticated_and_out_of_range_peers():
    """Rank 0 admits a peer only with the magic, a rank in 1..world-1 that has not checked in yet and the job's HMAC; an
    intruder (wrong token, out-of-range or duplicate rank) is dropped and the real rank still gets its slot."""
    import struct
    import threading
    import time
    from nfb200.distributed import communication as C
    port = _free_port()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["NFB200_RDZV_TOKEN"] = "job-secret"
    result = {}

    def server():
        result["g"] = C._TcpGroup(0, 2, "127.0.0.1", port, 30)

    th = threading.Thread(target=server)
    th.start()
    time.sleep(0.3)

    def raw(payload):
        s = socket.create_connection(("127.0.0.1", port), timeout=5)
        s.sendall(payload)
        return s

    good_mac = C._hello_mac(1, 2)
    intruders = [raw(C._MAGIC + struct.pack("!I", 1) + b"\x00" * 32),            # wrong MAC
                 raw(C._MAGIC + struct.pack("!I", 7) + C._hello_mac(7, 2)),      # rank out of range (was an IndexError)
                 raw(C._MAGIC + struct.pack("!I", 0) + C._hello_mac(0, 2)),      # rank 0 is the server itself
                 raw(b"GARBAGE!" + struct.pack("!I", 1) + good_mac)]             # wrong magic
    time.sleep(0.3)
    assert th.is_alive(), "an intruder was admitted"
    peer = C._TcpGroup(1, 2, "127.0.0.1", port, 30)
    th.join(timeout=30)
    assert not th.is_alive()
    g0 = result["g"]
    t2 = threading.Thread(target=lambda: result.__setitem__("r", g0.allgather(np.arange(3))))
    t2.start()
    mine = peer.allgather(np.arange(3) + 10)
    t2.join(timeout=30)
    assert np.array_equal(mine[0], np.arange(3)) and np.array_equal(mine[1], np.arange(3) + 10)
    for s in intruders:
        s.close()
    peer.close()
    g0.close()
    del os.environ["NFB200_RDZV_TOKEN"]
