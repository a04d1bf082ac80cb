"""Checkpoints either side of the training step.

The reference has no weight file format: `Module.state_dict()` returns `{dotted_name: ndarray}` (nn/module.py:112-123),
the optimizers return nested dicts with fp64 moment arrays (optim/adam.py:279-311), and its training scripts only write a
JSON record `{"epoch", "step", "config", "metrics"}` per epoch (examples/training/gpt2_training.py:528-547,
checkpoints/gpt2/gpt2_epoch_3.json).  This module keeps all three interchangeable:

* `<path>.json` -- that same JSON record (so whatever reads the reference's checkpoints reads these), plus a `"state"` entry
  holding the scalar part of the optimizer / scheduler / scaler state and, for every array, the key it has in
* `<path>.npz`  -- one NumPy archive: `model/<dotted_name>` for the parameters (exactly `state_dict()`'s keys and fp32
  values) and `optimizer/...`, `scheduler/...`, `scaler/...` for array-valued state (Adam moments as fp64, the reference's
  format).

Device parameters are snapshotted through `state_dict()` / `get_state_dict()` (one device->host copy per tensor); nothing
here touches the GPU otherwise.  Files are written to a temporary name and renamed, so a crash never leaves a torn
checkpoint.  Under data parallelism every rank holds the same state: call `save_checkpoint` from every rank and only
rank 0 writes (the others return None)."""
from __future__ import annotations

import json
import os
import tempfile
from typing import Any, Dict, Optional

import numpy as np

_ARRAY = "__array__"


def _split(obj: Any, prefix: str, arrays: Dict[str, np.ndarray]):
    """Nested state -> JSON-able skeleton; every ndarray leaf moves to `arrays[prefix/...]` and leaves its key behind."""
    if isinstance(obj, dict):
        return {str(k): _split(v, f"{prefix}/{k}", arrays) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_split(v, f"{prefix}/{i}", arrays) for i, v in enumerate(obj)]
    if hasattr(obj, "get") and hasattr(obj, "shape") and not isinstance(obj, np.ndarray):   # a device array
        obj = obj.get()
    if isinstance(obj, np.ndarray):
        arrays[prefix] = obj
        return {_ARRAY: prefix}
    if isinstance(obj, np.generic):
        return obj.item()
    return obj


def _join(skel: Any, arrays) -> Any:
    if isinstance(skel, dict):
        if set(skel) == {_ARRAY}:
            return np.asarray(arrays[skel[_ARRAY]])
        return {k: _join(v, arrays) for k, v in skel.items()}
    if isinstance(skel, list):
        return [_join(v, arrays) for v in skel]
    return skel


def _state_of(obj):
    for name in ("get_state_dict", "state_dict"):
        fn = getattr(obj, name, None)
        if callable(fn):
            return fn()
    raise TypeError(f"{type(obj).__name__} has no state_dict()")


def _base_lrs(sch):
    """The reference's scheduler state is only `last_epoch` (optim/lr_scheduler.py:50-56): the base learning rates are
    whatever `optimizer.lr` was when the scheduler was constructed.  Kept next to it so that a resumed run does not depend
    on re-creating the optimizer with the original learning rate first."""
    out = {"base_lrs": [float(x) for x in getattr(sch, "base_lrs", [])]}
    kids = getattr(sch, "schedulers", None)
    if kids:
        out["schedulers"] = [_base_lrs(k) for k in kids]
    return out


def _set_base_lrs(sch, saved):
    if saved.get("base_lrs") and hasattr(sch, "base_lrs"):
        sch.base_lrs = list(saved["base_lrs"])
    for k, sv in zip(getattr(sch, "schedulers", None) or [], saved.get("schedulers", [])):
        _set_base_lrs(k, sv)


def _atomic_write(path: str, write):
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=d, prefix=os.path.basename(path) + ".", suffix=".tmp")
    try:
        with os.fdopen(fd, "wb") as f:
            write(f)
        os.replace(tmp, path)
    except BaseException:
        try:
            os.unlink(tmp)
        except OSError:
            pass
        raise


def save_checkpoint(path: str, model, optimizer=None, scheduler=None, scaler=None, epoch: int = 0, step: int = 0,
                    config: Optional[dict] = None, metrics: Optional[dict] = None) -> Optional[str]:
    """Write `<path>.npz` + `<path>.json`; returns the JSON path (None on the ranks that do not write)."""
    from . import distributed as dist
    if dist.is_initialized() and dist.get_rank() != 0:
        return None
    base = path[:-5] if path.endswith(".json") else (path[:-4] if path.endswith(".npz") else path)
    arrays: Dict[str, np.ndarray] = {}
    for name, value in model.state_dict().items():
        arrays[f"model/{name}"] = np.asarray(value)
    state = {}
    for key, obj in (("optimizer", optimizer), ("scheduler", scheduler), ("scaler", scaler)):
        if obj is not None:
            state[key] = _split(_state_of(obj), key, arrays)
    if scheduler is not None:
        state["scheduler_base_lrs"] = _base_lrs(scheduler)
    record = {"epoch": int(epoch), "step": int(step), "config": dict(config or {}), "metrics": dict(metrics or {}),
              "state": state, "arrays": os.path.basename(base) + ".npz", "format": "nfb200-checkpoint-1"}
    _atomic_write(base + ".npz", lambda f: np.savez(f, **arrays))
    _atomic_write(base + ".json", lambda f: f.write(json.dumps(record, indent=2, default=str).encode()))
    return base + ".json"


def load_checkpoint(path: str, model=None, optimizer=None, scheduler=None, scaler=None, strict: bool = True) -> dict:
    """Restore whatever is passed in from `<path>.json` / `<path>.npz`; returns the JSON record (epoch, step, config,
    metrics, ...).  `model.load_state_dict` copies the parameters back to the device they live on."""
    base = path[:-5] if path.endswith(".json") else (path[:-4] if path.endswith(".npz") else path)
    with open(base + ".json") as f:
        record = json.load(f)
    if record.get("format") != "nfb200-checkpoint-1":
        raise ValueError(f"{base}.json is not an nfb200 checkpoint (a reference metrics record has no weights to load)")
    with np.load(os.path.join(os.path.dirname(os.path.abspath(base)), record["arrays"])) as z:
        if model is not None:
            sd = {k[len("model/"):]: z[k] for k in z.files if k.startswith("model/")}
            model.load_state_dict(sd, strict=strict)
        for key, obj in (("optimizer", optimizer), ("scheduler", scheduler), ("scaler", scaler)):
            if obj is None:
                continue
            if key not in record["state"]:
                raise KeyError(f"{base}.json holds no {key} state")
            obj.load_state_dict(_join(record["state"][key], z))
        if scheduler is not None and "scheduler_base_lrs" in record["state"]:
            _set_base_lrs(scheduler, record["state"]["scheduler_base_lrs"])
    return record
