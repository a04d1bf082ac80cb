"""Process launcher with the reference's API and environment contract (distributed/launcher.py:24-287; env at :145-162).

    python -m nfb200.distributed.launcher --nproc-per-node 8 train.py --args ...

One process per GPU: rank r of this node gets RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT (plus the two
settings the reference adds, NCCL_ASYNC_ERROR_HANDLING=1 and OMP_NUM_THREADS=1) and selects `cuda:LOCAL_RANK`;
`nfb200.distributed.init_process_group()` in the script does the NCCL unique-id exchange over MASTER_ADDR:MASTER_PORT.

Same names as upstream -- `LaunchConfig`, `DistributedLauncher`, `launch_distributed_training`, `find_free_port`,
`get_local_ip` -- with these differences, all on the side of doing what the upstream docstrings promise:
* a rank that exits non-zero takes the whole job down at once (the survivors would only hang in their next collective);
  upstream counts "restarts" in a monitor thread but never restarts anything (launcher.py:192-211), so `max_restarts`
  and `monitor_interval` are accepted and only the latter is used (as the polling period);
* every rank runs in its own session and is terminated through the exact PIDs started here (SIGTERM, then SIGKILL after
  5 s), also on SIGINT / SIGTERM of the launcher;
* multi-node: run the launcher once per node with `nnodes` / `node_rank` (upstream's `MultiNodeLauncher` shells out to
  ssh; not mirrored).
"""
from __future__ import annotations

import argparse
import os
import signal
import socket
import subprocess
import sys
import time
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional


@dataclass
class LaunchConfig:
    """Field names and defaults of distributed/launcher.py:24-48."""
    nproc_per_node: int = 1
    nnodes: int = 1
    node_rank: int = 0
    master_addr: str = "127.0.0.1"
    master_port: int = 29500
    backend: str = "auto"
    use_env: bool = True
    max_restarts: int = 3
    monitor_interval: float = 1.0
    start_timeout: int = 600
    log_dir: Optional[str] = None
    redirect_stdout: bool = True
    redirect_stderr: bool = True


class _Workers:
    """The ranks of this node: spawn, watch, tear down."""

    def __init__(self, config: LaunchConfig):
        self.config = config
        self.procs: List[subprocess.Popen] = []
        self._files = []
        # shared secret of the job's TCP rendezvous (communication._hello_mac): drawn once per launch; a multi-node job
        # exports the same NFB200_RDZV_TOKEN on every node
        import secrets
        self._token = os.environ.get("NFB200_RDZV_TOKEN") or secrets.token_hex(16)

    def env_for(self, rank: int, world_size: int, local_rank: int) -> dict:
        env = dict(os.environ)
        if self.config.use_env:
            env.update({"RANK": str(rank), "WORLD_SIZE": str(world_size), "LOCAL_RANK": str(local_rank),
                        "MASTER_ADDR": self.config.master_addr, "MASTER_PORT": str(self.config.master_port),
                        "NCCL_ASYNC_ERROR_HANDLING": "1", "OMP_NUM_THREADS": "1", "NFB200_RDZV_TOKEN": self._token})
        return env

    def _log_files(self, rank: int):
        c = self.config
        if not c.log_dir:
            return None, None
        d = Path(c.log_dir)
        d.mkdir(parents=True, exist_ok=True)
        out = open(d / f"rank_{rank}_stdout.log", "w") if c.redirect_stdout else None
        err = open(d / f"rank_{rank}_stderr.log", "w") if c.redirect_stderr else None
        self._files += [f for f in (out, err) if f is not None]
        return out, err

    def start(self, script: str, script_args: List[str]) -> bool:
        c = self.config
        world = c.nnodes * c.nproc_per_node
        for local in range(c.nproc_per_node):
            rank = c.node_rank * c.nproc_per_node + local
            out, err = self._log_files(rank)
            try:
                self.procs.append(subprocess.Popen([sys.executable, script] + list(script_args), env=self.env_for(rank, world, local),
                                                   stdout=out, stderr=err, start_new_session=True))
            except OSError as e:
                sys.stderr.write(f"launcher: cannot start rank {rank}: {e}\n")
                self.stop()
                return False
        return True

    def wait(self) -> int:
        """Exit code of the job: 0 when every rank returned 0, else the largest code among the ranks that failed on their own
        (a signal death counts as 128 + n); the ranks taken down in response do not count."""
        codes = [None] * len(self.procs)

        def code(r):
            return r if r >= 0 else 128 - r
        failed: List[int] = []
        try:
            while any(c is None for c in codes):
                for i, p in enumerate(self.procs):
                    if codes[i] is None and p.poll() is not None:
                        codes[i] = code(p.returncode)
                failed = [c for c in codes if c]
                if failed:
                    self.stop()                                   # fail fast: the rest would hang in their next collective
                    break
                if any(c is None for c in codes):
                    time.sleep(min(0.2, max(0.01, self.config.monitor_interval)))
        finally:
            self._close_files()
        return max(failed) if failed else 0

    @staticmethod
    def _signal_group(p: subprocess.Popen, sig: int):
        """Signal the session this launcher created for rank process p (its own children included), never anything else."""
        try:
            os.killpg(p.pid, sig)          # start_new_session=True made p the leader of a group with the same id
        except (OSError, AttributeError):
            try:
                p.send_signal(sig)
            except OSError:
                pass

    def stop(self):
        """SIGTERM to every live rank started here, SIGKILL after 5 s."""
        live = [p for p in self.procs if p.poll() is None]
        for p in live:
            self._signal_group(p, signal.SIGTERM)
        deadline = time.time() + 5.0
        for p in live:
            try:
                p.wait(timeout=max(0.0, deadline - time.time()))
            except subprocess.TimeoutExpired:
                self._signal_group(p, signal.SIGKILL)
                p.wait()

    def _close_files(self):
        for f in self._files:
            try:
                f.close()
            except OSError:
                pass
        self._files = []


class DistributedLauncher:
    """distributed/launcher.py:237-287: validates the configuration (same `ValueError`s) and runs the ranks of this node."""

    def __init__(self, config: LaunchConfig):
        self.config = config
        self._validate_config()

    def _validate_config(self):
        c = self.config
        if c.nproc_per_node <= 0:
            raise ValueError("nproc_per_node must be positive")
        if c.nnodes <= 0:
            raise ValueError("nnodes must be positive")
        if not 0 <= c.node_rank < c.nnodes:
            raise ValueError("node_rank must be in range [0, nnodes)")
        try:
            socket.inet_aton(c.master_addr)
        except OSError:
            raise ValueError(f"Invalid master_addr: {c.master_addr}") from None
        if not 1024 <= c.master_port <= 65535:
            raise ValueError("master_port must be in range [1024, 65535]")

    def launch(self, training_script: str, *script_args) -> int:
        workers = _Workers(self.config)

        def on_signal(signum, frame):
            workers.stop()
            sys.exit(128 + signum)
        previous = {}
        for sig in (signal.SIGINT, signal.SIGTERM):
            try:
                previous[sig] = signal.signal(sig, on_signal)
            except ValueError:          # not the main thread: the caller keeps its own handlers
                pass
        try:
            if not workers.start(training_script, [str(a) for a in script_args]):
                return 1
            return workers.wait()
        finally:
            for sig, h in previous.items():
                signal.signal(sig, h)


def launch_distributed_training(training_script: str, nproc_per_node: int = 1, nnodes: int = 1, node_rank: int = 0,
                                master_addr: str = "127.0.0.1", master_port: int = 29500, backend: str = "auto",
                                log_dir: Optional[str] = None, *script_args) -> int:
    """distributed/launcher.py:363-401 (same positional order)."""
    config = LaunchConfig(nproc_per_node=nproc_per_node, nnodes=nnodes, node_rank=node_rank, master_addr=master_addr,
                          master_port=master_port, backend=backend, log_dir=log_dir)
    return DistributedLauncher(config).launch(training_script, *script_args)


def launch(script: str, script_args, nproc_per_node: int, master_addr: str = "127.0.0.1", master_port: int = 29500, nnodes: int = 1,
           node_rank: int = 0, log_dir: Optional[str] = None) -> int:
    """Shorthand used by the command line below."""
    return launch_distributed_training(script, nproc_per_node, nnodes, node_rank, master_addr, master_port, "auto", log_dir, *script_args)


def find_free_port() -> int:
    """A port the kernel just handed out on this host (distributed/launcher.py:404-410)."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("", 0))
        return s.getsockname()[1]


def get_local_ip() -> str:
    """The address of the interface a default route would use; 127.0.0.1 without one (distributed/launcher.py:413-422).
    A UDP connect sends no packet, so this works without network access."""
    try:
        with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s:
            s.connect(("8.8.8.8", 80))
            return s.getsockname()[0]
    except OSError:
        return "127.0.0.1"


def main(argv=None):
    ap = argparse.ArgumentParser(description="nfb200 distributed launcher")
    ap.add_argument("--nproc-per-node", "--nproc_per_node", type=int, default=1)
    ap.add_argument("--nnodes", type=int, default=1)
    ap.add_argument("--node-rank", "--node_rank", type=int, default=0)
    ap.add_argument("--master-addr", "--master_addr", default="127.0.0.1")
    ap.add_argument("--master-port", "--master_port", type=int, default=29500)
    ap.add_argument("--log-dir", "--log_dir", default=None)
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    sys.exit(launch(a.script, a.script_args, a.nproc_per_node, a.master_addr, a.master_port, a.nnodes, a.node_rank, a.log_dir))


if __name__ == "__main__":
    main()
