"""Stress of the peer-memory all-reduce (torchrun, one rank per GPU): many short data-parallel steps (small batch,
so the flag protocol of adam_allreduce_p2p_kernel runs every ~0.2 ms), deliberately skewed ranks, then the replicas
must still be bit-identical and the loss finite.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 scripts/stress_p2p.py [steps]"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops
from nerf_simple_b200.nets import Nerf
from nerf_simple_b200.trainer import Trainer
from nerf_simple_b200.xyz import poses_to_render

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = Nerf().to(dev)
poses = torch.stack(poses_to_render(4, -30, 3)).to(dev)
rays = ops.generate_rays(poses, 64, 64, 88.8)
gt = torch.sigmoid(rays[:, 3:6] * 3)
tr = Trainer(net, rays, gt, N=32, batch_size=512, seed=1, world_size=world)
t0 = time.perf_counter()
for i in range(steps):
    tr.step()
    if i % 997 == rank * 131 % 997:      # skew: one rank at a time falls behind by a host-side stall
        torch.cuda.synchronize()
        time.sleep(0.002)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ref = tr.flat_param.clone()
dist.broadcast(ref, 0)
same = torch.tensor([float(torch.equal(tr.flat_param, ref)), float(torch.isfinite(tr.last_loss))], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"stress_p2p: {world} ranks, {steps} steps in {dt:.1f} s ({dt / steps * 1e3:.3f} ms/step), mode: {tr.launch_mode}; "
          f"replicas identical: {bool(same[0])}, loss finite: {bool(same[1])}, loss {float(tr.last_loss):.5f}", flush=True)
tr.close()
dist.destroy_process_group()
sys.exit(0 if bool(same.min() > 0) else 1)
