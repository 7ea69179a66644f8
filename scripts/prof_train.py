import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops
from nerf_simple_b200.nets import Nerf
from nerf_simple_b200.trainer import Trainer
from nerf_simple_b200.xyz import poses_to_render
torch.manual_seed(0)
net = Nerf().cuda()
poses = torch.stack(poses_to_render(4, -30, 25)).cuda()
rays = ops.generate_rays(poses, 400, 400, 555.5); gt = torch.rand(rays.shape[0], 3, device="cuda")
tr = Trainer(net, rays, gt, N=64, batch_size=4096)
for _ in range(5): tr.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): tr.step()
e1.record(); torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1) / 30)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): tr.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=50))
