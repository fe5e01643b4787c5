"""A whole training step - loss, backward, gradient all-reduce, optimizer update - as ONE CUDA graph.

The PEAGNN step is ~270 kernel launches (aggregations, projections, fusion, scoring, Adam); on one
GPU the host keeps ahead of the device, but once the propagation is row-sharded over several GPUs each
rank's kernels shrink and the step becomes launch-bound.  Capturing the step removes the per-launch host
cost: a replay is a single submission.  Validated on one GPU (tests/test_gpu_model.py, bench.py).  With an
``allreduce`` hook the NCCL collectives are captured too (capture mode 'thread_local', see __init__).

Contract (the usual CUDA-graph one):
  * run at least one eager step first, so every lazily built structure (CSR views, kernel attributes,
    NCCL buffers, optimizer state) exists before the capture - and drop every reference to that step's
    loss / outputs: tensors that still carry its autograd graph keep AccumulateGrad nodes tied to the
    default stream alive, and the capture fails with a stream-capture-invalidated error;
  * batches must keep the captured shape - other shapes (the last, short batch of an epoch) go through
    the eager path;
  * do not call ``optimizer.zero_grad()`` between replays: gradients are static tensors owned by the
    graph, and every replay overwrites them;
  * the optimizer must be capturable (``torch.optim.Adam(..., fused=True, capturable=True)``).
"""
import torch

from . import _lib
from . import functional as F_


class GraphedTrainStep(object):
    def __init__(self, model, optimizer, example_batch, allreduce=None, profile=False):
        """Trains ONE eager step on ``example_batch`` (its loss is ``first_loss``), then captures the step.
        ``allreduce``: optional callable run between backward and the optimizer step (multi-GPU:
        ``lambda: distributed.allreduce_gradients(params)``).
        ``profile``: capture an external CUDA event node on both sides of every C-ABI launch; after each
        replay (and a synchronize) ``profile_events`` = [(tag, algorithmic bytes, start, end)] holds that
        replay's per-launch times - measurement builds only, the event nodes serialise the graph."""
        self.model, self.optimizer, self.allreduce = model, optimizer, allreduce
        self.static_batch = example_batch.clone()
        self.shape = tuple(example_batch.shape)
        self.profile_events = []
        saved_lib_profile, _lib.profile = _lib.profile, None
        try:
            # autograd's AccumulateGrad nodes remember the stream they were created on; nodes born on the
            # default stream in earlier eager steps would drag the legacy stream into the capture.  One step
            # on a side stream (with every reference to the old autograd graph dropped) re-creates them there.
            model.cached_repr = None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                optimizer.zero_grad(set_to_none=True)
                self.first_loss = self._eager(self.static_batch).detach().clone()
            torch.cuda.current_stream().wait_stream(side)
            model.cached_repr = None
            optimizer.zero_grad(set_to_none=True)             # backward allocates the static .grad tensors in the pool
            torch.cuda.synchronize()
            count0 = _lib.load().peagnn_launch_count()
            self.graph = torch.cuda.CUDAGraph()
            if profile:
                _lib.profile = self.profile_events
            # with collectives in the step, NCCL's watchdog thread keeps polling CUDA events of earlier work while this
            # thread captures; under the default 'global' capture mode that poll is an illegal call and takes the
            # process group down (the 2-GPU hang of round 1) - 'thread_local' confines the check to this thread
            mode = 'thread_local' if allreduce is not None or torch.distributed.is_initialized() else 'global'
            with torch.cuda.graph(self.graph, capture_error_mode=mode):
                self.static_loss = self._eager(self.static_batch)
            self.launches_per_replay = int(_lib.load().peagnn_launch_count() - count0)
        finally:
            _lib.profile = saved_lib_profile

    def _eager(self, batch):
        loss = self.model.loss(batch)
        loss.backward()
        if self.allreduce is not None:
            self.allreduce()
        self.optimizer.step()
        return loss

    def __call__(self, batch):
        """Runs one step on ``batch`` (device or pinned host tensor); returns the loss tensor (static:
        read it before the next call)."""
        if tuple(batch.shape) != self.shape:                  # e.g. the short last batch of an epoch
            batch = batch.to(self.static_batch.device, non_blocking=True)
            self.optimizer.zero_grad(set_to_none=False)
            return self._eager(batch).detach()
        self.static_batch.copy_(batch, non_blocking=True)
        self.graph.replay()
        return self.static_loss.detach()
