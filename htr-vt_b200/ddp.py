"""Data-parallel training: one process per GPU, batch sharding, NCCL all-reduce (AVG) of the gradients.

The reference is single-GPU; this is the B200-native scale-out of its training step.
The encoder's backward writes every parameter gradient into ONE flat fp32 buffer laid out in
named_parameters() order, so the exchange is three or four NCCL calls: the transformer segment is reduced as
soon as the transformer backward finishes (overlapping the stem backward, which is ~80 % of the
step), the last stem layer when its blocks are done, the small remainder of the stem at the end.  BatchNorm statistics stay per rank (the reference has no SyncBN).
"""
import torch
import torch.distributed as dist


class GradAllReduce(object):
    def __init__(self, group=None):
        self.group = group
        self.pending = []

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def reduce_async(self, flat_segment: torch.Tensor) -> None:
        if self.world == 1 or flat_segment.numel() == 0:
            return
        op = dist.ReduceOp.AVG if flat_segment.is_cuda else dist.ReduceOp.SUM
        work = dist.all_reduce(flat_segment, op=op, group=self.group, async_op=True)
        self.pending.append((work, flat_segment, op))

    def finish(self) -> None:
        for work, seg, op in self.pending:
            work.wait()
            if op == dist.ReduceOp.SUM:           # gloo (CPU tests) has no AVG
                seg.div_(self.world)
        self.pending = []


def broadcast_tensors(tensors, group=None, src=0):
    """Make every rank start from rank `src`'s values (what torch DDP does at construction): parameters AND buffers
    (BatchNorm running statistics, counters).  Tensors are packed per dtype so the exchange is a handful of collectives."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    with torch.no_grad():
        for (_dtype, _dev), ts in by_dtype.items():
            flat = torch.cat([t.detach().reshape(-1) for t in ts])
            dist.broadcast(flat, src=src, group=group)
            off = 0
            for t in ts:
                t.copy_(flat[off:off + t.numel()].view(t.shape))
                off += t.numel()


def segment_bounds(names, numels, first_transformer_prefix="blocks."):
    """Offsets of the [stem | transformer] split of the flat gradient buffer.  Every tensor starts on a 256-byte
    boundary (64 floats): weight-gradient GEMMs reduce-add into these views through TMA tensor maps, which need
    16-byte aligned bases.  Tensor i occupies [offs[i], offs[i] + numels[i]); the padding stays zero."""
    offs = [0]
    for n in numels:
        offs.append(offs[-1] + (n + 63) // 64 * 64)
    split = len(names)
    for i, n in enumerate(names):
        if n.startswith(first_transformer_prefix):
            split = i
            break
    return offs, offs[split]


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a batch for this rank (inference / evaluation sharding)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
