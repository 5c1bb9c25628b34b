"""world_size-2 gloo tests (CPU) of the data-parallel host logic: clip / sequence sharding and the single flat
gradient all-reduce.  'all-reduced grads on n ranks == grads of one process over all the clips' (SURVEY 8(e))."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfvos_b200 import dp


def test_shard_clips_balanced_and_complete():
    for n, w in [(64, 8), (64, 4), (7, 2), (3, 4), (0, 2)]:
        shards = [dp.shard_clips(n, r, w) for r in range(w)]
        assert sorted(i for s in shards for i in s) == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_shard_sequences_longest_first():
    lengths = [50, 100, 75, 60, 90, 55, 80, 70, 65, 95, 51, 99, 52, 98, 53, 97]
    shards = dp.shard_sequences(lengths, 8)
    assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(lengths) // 2


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
        mine = dp.shard_clips(8, rank, world)
        loss = ((model(x[mine]) - y[mine]) ** 2).sum()
        loss.backward()
        bucket = dp.GradBucket(model.parameters())
        bucket.pack()
        bucket.all_reduce(average=False)
        bucket.unpack()
        if rank == 0:
            torch.save([p.grad.clone() for p in model.parameters()], out)
        # every rank must hold identical gradients after the collective
        ref = [p.grad.clone() for p in model.parameters()]
        for t in ref:
            t2 = t.clone()
            dist.broadcast(t2, src=0)
            assert torch.equal(t, t2)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_equals_single_process(tmp_path):
    out = str(tmp_path / "grads.pt")
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    model = _model()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
    ((model(x) - y) ** 2).sum().backward()
    for a, p in zip(got, model.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-6)


# ---- gradient arena: the backward pass writes into the flat buffer, ranges are all-reduced as they complete ------------
class _ArenaLinear(torch.autograd.Function):
    """Stands in for a libsfvos backward: the weight gradient is accumulated into ``ops.new_grad(w)`` and that accumulator is
    what autograd receives (adopted as w.grad without a copy when w.grad is None)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return x @ w.t()

    @staticmethod
    def backward(ctx, g):
        from sfvos_b200 import ops
        x, w = ctx.saved_tensors
        gw = ops.new_grad(w)
        gw.add_(g.t() @ x)
        return g @ w, gw


def _arena_worker(rank, world, port, out):
    from sfvos_b200 import ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        w1, w2 = torch.nn.Parameter(torch.randn(5, 6)), torch.nn.Parameter(torch.randn(1, 5))
        arena = dp.GradArena([("head", [w2]), ("trunk", [w1])])
        ops.GRAD_ARENA = arena
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
        mine = dp.shard_clips(8, rank, world)
        arena.zero()
        for half in (mine[:len(mine) // 2], mine[len(mine) // 2:]):          # two micro-batches accumulate into the arena
            w1.grad = w2.grad = None
            h = torch.relu(_ArenaLinear.apply(x[half], w1))
            ((_ArenaLinear.apply(h, w2) - y[half]) ** 2).sum().backward()
        assert arena.adopted()                                               # .grad IS the arena slice: no pack pass
        dist.all_reduce(arena.range("head"))                                 # ranges are contiguous views of the flat buffer
        dist.all_reduce(arena.range("trunk"))
        if rank == 0:
            torch.save([w1.grad.clone(), w2.grad.clone()], out)
    finally:
        ops.GRAD_ARENA = None
        dist.destroy_process_group()


def test_gradient_arena_micro_batches_and_ranges(tmp_path):
    out = str(tmp_path / "arena.pt")
    port = 29950 + os.getpid() % 40
    mp.spawn(_arena_worker, args=(2, port, out), nprocs=2, join=True)
    g1, g2 = torch.load(out)
    torch.manual_seed(0)
    w1, w2 = torch.nn.Parameter(torch.randn(5, 6)), torch.nn.Parameter(torch.randn(1, 5))
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
    ((torch.relu(x @ w1.t()) @ w2.t() - y) ** 2).sum().backward()
    assert torch.allclose(g1, w1.grad, rtol=1e-5, atol=1e-5) and torch.allclose(g2, w2.grad, rtol=1e-5, atol=1e-5)
