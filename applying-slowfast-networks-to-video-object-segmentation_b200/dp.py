"""Data-parallel plumbing for the hot path: one process per GPU, clips / sequences sharded by rank,
and exactly one collective per optimizer step -- a sum all-reduce of the trainable gradients in a single flat
f32 bucket (NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests).  BatchNorm statistics stay
per rank / per call like the reference (no SyncBN: the reference normalises over a single clip batch).
The reference has no distributed code on this path; this is new functionality."""
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Contiguous, balanced split of clip indices: the first n_clips % world ranks get one extra clip."""
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def shard_sequences(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Inference sharding by whole sequence: greedy longest-first onto the least-loaded rank (sequences have
    unequal length; no frame halo is needed because a sequence never crosses ranks)."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        out[r].append(i)
        loads[r] += lengths[i]
    return out


class GradArena:
    """Every trainable gradient as a slice of ONE flat f32 buffer, grouped so that each group is a contiguous range.

    ``groups``: [(name, [parameters])] in the order the backward pass FINISHES them (for the hot path: the roi_heads
    parameters -- 87 % of the bytes, fc6 alone 70 % -- are complete before the SlowFast module's backward starts).  While
    ``ops.GRAD_ARENA`` points here the libsfvos backward passes accumulate straight into the slices and hand them to autograd,
    which adopts them as ``p.grad`` (no copy), so ``all_reduce(range(name))`` can start the moment a group is done and overlap
    the rest of the backward pass; there is no pack / torch.cat pass and ``unpack`` is a no-op.  ``zero()`` once per optimizer
    step; between micro-batches set ``p.grad = None`` (the slices keep accumulating and are adopted again)."""

    def __init__(self, groups, device=None):
        self.groups = [(name, [p for p in ps if p.requires_grad]) for name, ps in groups]
        params = [p for _, ps in self.groups for p in ps]
        device = device if device is not None else params[0].device
        self.slices, self.ranges = {}, {}
        off = 0
        for name, ps in self.groups:
            start = off
            for p in ps:
                assert p.dtype == torch.float32, "the arena holds f32 master gradients"
                self.slices[p.data_ptr()] = (off, p.numel(), tuple(p.shape))
                off += (p.numel() + 3) // 4 * 4                      # 16-byte aligned slices (vector atomics)
            self.ranges[name] = (start, off)
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.params = params

    def view(self, param, shape=None):
        hit = self.slices.get(param.data_ptr())
        if hit is None:
            return None
        off, n, pshape = hit
        return self.flat[off:off + n].view(tuple(shape) if shape is not None else pshape)     # a NEW tensor object every time

    def range(self, name):
        a, b = self.ranges[name]
        return self.flat[a:b]

    def zero(self):
        self.flat.zero_()

    def adopted(self):
        """True iff every parameter's .grad IS its arena slice (checked once after a warm-up step)."""
        return all(p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * self.slices[p.data_ptr()][0] for p in self.params)


class GradBucket:
    """All trainable gradients of ``params`` viewed as one contiguous f32 buffer.

    ``pack()`` copies the .grad tensors in, ``all_reduce()`` sums over ranks (async on CUDA when
    ``async_op=True``) and divides by the world size, ``unpack()`` points every p.grad at its slice (no copy)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += p.numel()

    def pack(self):
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()]
            if p.grad is None:
                view.zero_()
            else:
                view.copy_(p.grad.reshape(-1))
        return self.flat

    def all_reduce(self, world: int = None, group=None, average: bool = True):
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = world or dist.get_world_size(group)
        if world == 1:
            return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            if self.flat.is_cuda:
                from . import ops
                ops.axpby(self.flat, self.flat, 1.0 / world, 0.0)
            else:
                self.flat.mul_(1.0 / world)
        return self.flat

    def unpack(self):
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)


def dp_train_step(step, features, optimizer, bucket: GradBucket, average: bool = True):
    """forward + backward of ``step`` (a workload.HotPathStep or anything with .forward(features) -> (loss, _))
    on this rank's clips, one gradient all-reduce, optimizer step.  Returns the local loss tensor."""
    optimizer.zero_grad(set_to_none=True)
    loss, _ = step.forward(features)
    loss.backward()
    bucket.pack()
    bucket.all_reduce(average=average)
    bucket.unpack()
    optimizer.step()
    return loss.detach()
