"""Data-parallel plumbing shared by the trainer and bench.py: one process per GPU (torchrun), every rank
owns full replicas and a shard of the batch; the only collective on the data path is the SUM/N all-reduce
of each network's flat gradient arena (NCCL over NVLink on GPUs; gloo on CPU in the tests).

Replaces the reference's nn.DataParallel (train.py:168-169): 7 parameter broadcasts + scatter/gather +
reduce-add per iteration become zero broadcasts and two bucketed all-reduces."""
import torch
import torch.distributed as dist


def shard_batch(t, rank, world):
    """Contiguous equal shard of the leading (batch) dimension."""
    n = t.shape[0]
    if n % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (n, world))
    per = n // world
    return t[rank * per:(rank + 1) * per]


class GradExchange:
    def __init__(self, world, comm_stream=None):
        self.world = world
        self.comm_stream = comm_stream
        self._works = []

    def broadcast_params(self, arenas, src=0):
        for a in arenas:
            dist.broadcast(a, src=src)

    def allreduce_async(self, grad_arena):
        """Average `grad_arena` over the ranks; on CUDA the collective runs on the side stream after the
        work already enqueued on the current stream."""
        if self.world == 1:
            return
        if grad_arena.is_cuda and self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                grad_arena.div_(self.world)
                self._works.append(dist.all_reduce(grad_arena, op=dist.ReduceOp.SUM, async_op=True))
        else:
            grad_arena.div_(self.world)
            self._works.append(dist.all_reduce(grad_arena, op=dist.ReduceOp.SUM, async_op=True))

    def allreduce_sums(self, t):
        """In-place SUM over the ranks of a small tensor on the current stream (the masked sums of the losses in the
        exact-global mode: a handful of doubles; stream-ordered, no host synchronisation)."""
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def pending(self):
        return len(self._works)

    def wait(self):
        for w in self._works:
            w.wait()
        self._works = []
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
