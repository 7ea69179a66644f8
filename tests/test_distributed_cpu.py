"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding and the flat-gradient
all-reduce used by data-parallel training."""
import os
import socket

import torch
import torch.multiprocessing as mp

from nerf_simple_b200.engine import shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 640000, 2560000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.trainer import allreduce_mean_, attach_flat_grad, flatten_parameters
    torch.manual_seed(0)
    net = Nerf()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = flatten_parameters(net)
    grad = attach_flat_grad(net)
    # views alias the flat buffers, state_dict is unchanged
    ok = all(torch.equal(before[k], v) for k, v in net.state_dict().items())
    ok &= 595844 <= flat.numel() <= 595844 + 3 * 24 and all(p.grad.data_ptr() >= grad.data_ptr() for p in net.parameters())
    for p in net.parameters():
        p.grad.fill_(float(rank + 1))
    allreduce_mean_(grad, world)
    ok &= all(bool(torch.allclose(p.grad, torch.full_like(p.grad, (1 + world) / 2))) for p in net.parameters())
    ok &= float(grad.sum()) == 595844 * (1 + world) / 2          # padding floats stay zero
    # ray sharding + gather reassembles the frame in order
    n = 1001
    b, e = shard_range(n, rank, world)
    local = torch.arange(b, e, dtype=torch.float32)[:, None].repeat(1, 4)
    from nerf_simple_b200.engine import gather_shards
    full = gather_shards(local, n, rank, world)
    ok &= bool(full.shape == (n, 4) and torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32)))
    n = 1000                                                      # equal bands: the single-buffer collective
    b, e = shard_range(n, rank, world)
    full = gather_shards(torch.arange(b, e, dtype=torch.float32)[:, None].repeat(1, 4), n, rank, world)
    ok &= bool(full.shape == (n, 4) and torch.equal(full[:, 3], torch.arange(n, dtype=torch.float32)))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
