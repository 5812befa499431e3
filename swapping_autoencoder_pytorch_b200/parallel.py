"""Data-parallel execution: one process per GPU, one NCCL all-reduce of a flat gradient bucket per backward.

Replaces reference ``models/__init__.py:75-93`` (``MultiGPUModelWrapper`` over single-process
``nn.DataParallel``: per-call parameter broadcast, scatter/gather, reduce onto GPU 0) and keeps its call
surface: ``wrapper(*args, command=...)``, ``.singlegpu_model``, ``.get_parameters_for_mode``, ``.save``,
``.opt``.  Design:

* the batch is partitioned by image across ranks (each rank is handed its own shard; per-rank batch even);
* parameters are replicated once (broadcast from rank 0 at construction), optimizer state is replicated;
* because ``requires_grad`` is toggled between the D and G half-steps (reference optimizer :44-49), stock DDP's
  fixed reducer does not fit; instead every parameter carries a post-accumulate hook that queues ONE
  end-of-backward callback; the callback packs the gradients that exist (the active group) into a persistent
  flat fp32 bucket with a single kernel, all-reduces it (sum) over NCCL / NVLink, and unpacks scaled by
  1/world — so ``loss.backward(); optimizer.step()`` in the unchanged optimizer sees averaged gradients;
* per-sample losses are means over the local shard; the average of equal-sized shard means equals the
  reference's mean over the gathered global batch (reference optimizer :75, :93).
"""
import os

import torch
import torch.distributed as dist

from . import backend


def init_distributed(backend_name=None):
    """Join the torchrun-provided process group (no-op for a single process).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend_name is None:
            backend_name = "nccl" if torch.cuda.is_available() else "gloo"
        if backend_name == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend_name)
    return rank, world, local


class GradientBucket:
    """Persistent flat buffer + cached device pointer tables for one set of gradient tensors."""

    def __init__(self, device):
        self.device = device
        self.flat = None
        self.tables = {}     # key (tuple of data_ptrs) -> (ptrs, offsets, sizes, total)

    def _table(self, grads):
        key = tuple(g.data_ptr() for g in grads)
        hit = self.tables.get(key)
        if hit is not None:
            return hit
        sizes = [g.numel() for g in grads]
        offsets, total = [], 0
        for s in sizes:
            offsets.append(total)
            total += (s + 3) // 4 * 4            # keep every segment 16-byte aligned
        dev = self.device
        entry = (torch.tensor(key, dtype=torch.int64, device=dev), torch.tensor(offsets, dtype=torch.int64, device=dev),
                 torch.tensor(sizes, dtype=torch.int64, device=dev), total, offsets, sizes)
        if len(self.tables) > 16:
            self.tables.clear()
        self.tables[key] = entry
        return entry

    def all_reduce_mean(self, grads, world):
        if not grads:
            return
        if self.device.type != "cuda":
            # host path used by the gloo unit tests: same flatten / reduce / scatter-back arithmetic in torch
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            flat.div_(world)
            o = 0
            for g in grads:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
            return
        ptrs, offsets, sizes, total, _, _ = self._table(grads)
        if self.flat is None or self.flat.numel() < total:
            self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        flat = self.flat[:total]
        k = backend.kernels()
        k.bucket_pack(ptrs, offsets, sizes, len(grads), flat)
        dist.all_reduce(flat)
        k.bucket_unpack(ptrs, offsets, sizes, len(grads), flat, 1.0 / world)


class MultiGPUModelWrapper:
    def __init__(self, opt, model):
        self.opt = opt
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if opt.num_gpus > 0:
            model = model.to(torch.device("cuda", torch.cuda.current_device()))
        self.singlegpu_model = model
        self.parallelized_model = model          # attribute kept for callers of the reference wrapper
        self.device = next(model.parameters()).device
        self._pending = False
        self.suspend_reduce = False           # set while a half-step is being captured into a CUDA graph (graphs.py)
        self._bucket = GradientBucket(self.device)
        model(command="per_gpu_initialize")
        if self.world > 1:
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t, src=0)
            for p in model.parameters():
                p.register_post_accumulate_grad_hook(self._on_grad)

    # the hook fires once per parameter per backward; only the first one queues the callback
    def _on_grad(self, param):
        if self.suspend_reduce:
            return
        if not self._pending:
            self._pending = True
            torch.autograd.Variable._execution_engine.queue_callback(self._reduce_gradients)

    def _reduce_gradients(self):
        self._pending = False
        grads = [p.grad for p in self.singlegpu_model.parameters()
                 if p.requires_grad and p.grad is not None]
        self._bucket.all_reduce_mean(grads, self.world)

    def reduce_gradients_now(self):
        """explicit form of the end-of-backward callback (used after a CUDA-graph replay, where autograd does not run)"""
        if self.world > 1:
            self._reduce_gradients()

    def get_parameters_for_mode(self, mode):
        return self.singlegpu_model.get_parameters_for_mode(mode)

    def save(self, total_steps_so_far):
        if self.rank == 0:
            self.singlegpu_model.save(total_steps_so_far)

    def shard(self, batch):
        """This rank's slice of a global batch (dim 0), mirroring DataParallel's scatter."""
        if self.world == 1:
            return batch
        n = batch.shape[0]
        assert n % self.world == 0, "global batch must divide evenly across ranks"
        per = n // self.world
        return batch[self.rank * per:(self.rank + 1) * per]

    def __call__(self, *args, **kwargs):
        return self.singlegpu_model(*args, **kwargs)
