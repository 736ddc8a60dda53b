"""CPU, world_size 2 over gloo: the gradient exchange used inside backward averages flat segments across ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from importlib import import_module
    import htrvt_b200  # noqa: F401
    ddp = import_module("htr-vt_b200.ddp")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = ["mask_token", "patch_embed.conv1.weight", "blocks.0.attn.qkv.weight", "head.bias"]
    numels = [8, 18, 24, 5]
    offs, split = ddp.segment_bounds(names, numels)
    flat = torch.arange(offs[-1], dtype=torch.float32) * (rank + 1)
    sync = ddp.GradAllReduce()
    sync.reduce_async(flat[split:])        # transformer segment first (overlaps the stem backward on GPU)
    sync.reduce_async(flat[:split])
    sync.finish()
    want = torch.arange(offs[-1], dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, want)
    lo, hi = ddp.shard_batch(11, rank, world)
    # enable_data_parallel() broadcasts parameters AND buffers from rank 0 (torch DDP semantics): ranks seeded
    # differently must end up with rank 0's replica, BatchNorm statistics and counters included
    from functools import partial
    H = import_module("htr-vt_b200.model.HTR_VT")
    torch.manual_seed(100 + rank)
    m = H.MaskedAutoencoderViT(12, img_size=[64, 128], patch_size=(4, 64), embed_dim=256, depth=1, num_heads=2,
                               mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    with torch.no_grad():
        m.patch_embed.bn1.running_mean.add_(float(rank + 1))
        m.patch_embed.bn1.num_batches_tracked.add_(3 * (rank + 1))
    m.enable_data_parallel()
    digest = torch.stack([v.double().sum() for v in m.state_dict().values()])
    both = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(both, digest)
    same = all(torch.equal(both[0], b) for b in both)
    out[rank] = (bool(ok), lo, hi, bool(same), int(m.patch_embed.bn1.num_batches_tracked), m.dp_rank)
    dist.destroy_process_group()


def test_grad_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 6, 6, 11)
    assert out[0][3] and out[1][3]                       # identical replicas after the broadcast
    assert out[0][4] == out[1][4] == 3                   # rank 0's BatchNorm counter everywhere
    assert (out[0][5], out[1][5]) == (0, 1)
