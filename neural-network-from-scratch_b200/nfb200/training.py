"""Training-loop helpers on the device: CUDA-graph capture of a whole training step.

The reference's step is thousands of tiny NumPy calls; ours is ~500 kernel launches whose host-side cost (Python +
ctypes, ~20-40 us each) exceeds their device time on a B200.  ``GraphedStep`` runs the step a few times eagerly
(warming the caching allocator so no cudaMalloc happens later), captures one step into a CUDA graph on the compute
stream -- including the data-parallel NCCL bucket all-reduces on the communication stream -- and then replays it
with one launch per step.  Requirements on ``step_fn``: static shapes, inputs read from persistent device tensors
(copy new batches INTO them), no device->host reads inside (fetch the loss after the replay), optimizer in
``capturable`` mode (device-side step counter)."""
from __future__ import annotations

from . import runtime as R


class GraphedStep:
    def __init__(self, step_fn, warmup: int = 3, optimizers=()):
        for opt in optimizers:
            opt.capturable = True
        self.step_fn = step_fn
        self.out = None
        for _ in range(max(1, warmup)):
            self.out = step_fn()
        R.synchronize()
        self.graph = R.Graph()
        with self.graph:
            self.out = step_fn()
        R.synchronize()
        self.replays = 0
        self._opts = list(optimizers)
        # the captured step advanced the host-side counters (and a built-in schedule) once without running on the device
        for opt in self._opts:
            opt.step_count -= 1
            if hasattr(opt, "state"):
                for st in {id(s): s for s in opt.state.values()}.values():
                    st["step"] -= 1
            if hasattr(opt, "_update_learning_rate"):
                opt._update_learning_rate()       # a built-in schedule was advanced by the capture call: back to this step's lr

    def __call__(self):
        """One training step = one graph launch.  Learning rates are read on the device: whatever a scheduler (or the
        optimizer's built-in schedule) wrote to ``opt.lr`` before this call is what the replayed AdamW kernels use."""
        for opt in self._opts:
            if hasattr(opt, "_graph_replayed"):
                opt._graph_replayed(before=True)
            else:
                opt.step_count += 1
        self.graph.replay()
        self.replays += 1
        return self.out
