"""Process launcher with the reference's environment contract (distributed/launcher.py:24-287, env at :145-162):
spawns one process per rank with RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT set.

    python -m nfb200.distributed.launcher --nproc-per-node 8 train.py --args ...
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys


def launch(script: str, script_args, nproc_per_node: int, master_addr: str = "127.0.0.1", master_port: int = 29500, nnodes: int = 1,
           node_rank: int = 0) -> int:
    world = nproc_per_node * nnodes
    procs = []
    for local in range(nproc_per_node):
        env = dict(os.environ)
        env.update({"RANK": str(node_rank * nproc_per_node + local), "WORLD_SIZE": str(world), "LOCAL_RANK": str(local),
                    "MASTER_ADDR": master_addr, "MASTER_PORT": str(master_port), "NCCL_ASYNC_ERROR_HANDLING": "1"})
        procs.append(subprocess.Popen([sys.executable, script] + list(script_args), env=env))
    rc = 0
    try:
        for p in procs:
            r = p.wait()
            rc = rc or r
            if r != 0:  # one rank died: take the others down (exact PIDs we started)
                for q in procs:
                    if q.poll() is None:
                        q.terminate()
    except KeyboardInterrupt:
        for q in procs:
            if q.poll() is None:
                q.terminate()
        rc = 130
    return rc


def main(argv=None):
    ap = argparse.ArgumentParser(description="nfb200 distributed launcher")
    ap.add_argument("--nproc-per-node", "--nproc_per_node", type=int, default=1)
    ap.add_argument("--nnodes", type=int, default=1)
    ap.add_argument("--node-rank", "--node_rank", type=int, default=0)
    ap.add_argument("--master-addr", "--master_addr", default="127.0.0.1")
    ap.add_argument("--master-port", "--master_port", type=int, default=29500)
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    sys.exit(launch(a.script, a.script_args, a.nproc_per_node, a.master_addr, a.master_port, a.nnodes, a.node_rank))


if __name__ == "__main__":
    main()
