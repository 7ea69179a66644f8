"""Multi-GPU check (torchrun, one rank per GPU): a ray-sharded frame assembled by all_gather equals the bands
rendered on one GPU, and a data-parallel Trainer keeps its replicas bit-identical.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded_render.py"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops
from nerf_simple_b200.engine import FrameRenderer, render_sharded, shard_range
from nerf_simple_b200.nets import Nerf
from nerf_simple_b200.trainer import Trainer
from nerf_simple_b200.xyz import poses_to_render

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = Nerf().to(dev)
poses = torch.stack(poses_to_render(4, -30, 3)).to(dev)
H = W = 100
n = H * W
with torch.no_grad():
    rgb, disp = render_sharded(FrameRenderer(net, H, W, 138.9, N=64, seed=3), poses, 1, rank, world)
    parts = []
    for r in range(world):                       # every rank re-renders every band locally (same seed, same offset)
        b, e = shard_range(n, r, world)
        parts.append(FrameRenderer(net, H, W, 138.9, N=64, seed=3).render_rays(poses, n + b, e - b)[0])
    ok = torch.equal(rgb.reshape(n, 3), torch.cat(parts))
# data-parallel training: different batches per rank, one all-reduce, replicas stay identical
rays = ops.generate_rays(poses, H, W, 138.9)
gt = torch.sigmoid(rays[:, 3:6] * 3)
tr = Trainer(net, rays, gt, N=32, batch_size=1024, seed=1 + rank, world_size=world)
losses = [tr.step(sync_loss=True) for _ in range(30)]
flat = tr.flat_param.clone()
ref = flat.clone()
dist.broadcast(ref, 0)
same = torch.equal(flat, ref)
res = torch.tensor([float(ok), float(same), losses[0], losses[-1]], device=dev)
allres = [torch.empty_like(res) for _ in range(world)]
dist.all_gather(allres, res)
if rank == 0:
    for r, t in enumerate(allres):
        print(f"rank {r}: sharded frame == bands {bool(t[0])}, replicas identical {bool(t[1])}, loss {t[2]:.4f} -> {t[3]:.4f}")
    assert all(bool(t[0]) and bool(t[1]) for t in allres)
    assert all(float(t[3]) < float(t[2]) for t in allres)
    print("multi-GPU check ok")
dist.destroy_process_group()
