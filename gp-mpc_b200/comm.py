"""Host-side rendezvous for the multi-GPU path (one process per GPU).

Only small host objects travel here (the 128-byte NCCL id, fitted hyper rows,
ragged 'points'-mode results); the data path's all-gather of predictions runs
inside libgpmpc on NCCL.  torch.distributed is plumbing: any initialised
backend (nccl on the GPU box, gloo in CPU tests) works.
"""
from __future__ import annotations

import os


class Comm:
    """rank/world + object collectives.  Comm() with no process group is the
    single-process identity."""

    def __init__(self, group=None):
        self._dist = None
        self.group = group
        self.rank, self.world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank = dist.get_rank(group)
                self.world = dist.get_world_size(group)
        except ImportError:
            pass

    @classmethod
    def from_env(cls):
        """Initialise torch.distributed from torchrun's environment when WORLD_SIZE > 1."""
        world = int(os.environ.get('WORLD_SIZE', '1'))
        if world > 1:
            import torch
            import torch.distributed as dist
            if not dist.is_initialized():
                backend = 'nccl' if torch.cuda.is_available() else 'gloo'
                if backend == 'nccl':
                    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
                dist.init_process_group(backend=backend)
        return cls()

    def allgather_object(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self._dist.all_gather_object(out, obj, group=self.group)
        return out

    def broadcast_object(self, obj, src=0):
        if self.world == 1:
            return obj
        box = [obj if self.rank == src else None]
        self._dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def barrier(self):
        if self.world > 1:
            self._dist.barrier(group=self.group)
