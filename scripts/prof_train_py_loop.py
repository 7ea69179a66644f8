"""Developer probe: where the time of the reference's own training-loop body (train.py:47-57) goes on this engine."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops, config
from nerf_simple_b200.nets import Nerf
from nerf_simple_b200.dataload import RayGenerator
from nerf_simple_b200.rendering import render_nerf
from nerf_simple_b200.xyz import poses_to_render
B, N = 4096, 64
poses = torch.stack(poses_to_render(4, -30, 25)).cuda()
rays_table = ops.generate_rays(poses, 400, 400, 555.5)
class _RG:
    rays_dataset = {"train": rays_table.cpu()}
    select = RayGenerator.select
    _select_device = RayGenerator._select_device
rg = _RG()
train_imgs = torch.rand(rays_table.shape[0], 3).double()
config.set_sampler("philox"); config.set_select("device")
torch.manual_seed(0)
net = Nerf().cuda()
opt = torch.optim.Adam(net.parameters(), lr=5e-4)
crit = torch.nn.MSELoss()
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0; return time.perf_counter()
def step(timed):
    t = time.perf_counter()
    rays, ray_ids = rg.select(mode="train", N=B)
    if timed: t = tick("select", t)
    gt = train_imgs[ray_ids, :].float().cuda()
    if timed: t = tick("gt gather+h2d", t)
    opt.zero_grad()
    if timed: t = tick("zero_grad", t)
    rgb, depth, alpha, acc, w = render_nerf(rays.cuda(), net, N)
    if timed: t = tick("render_nerf", t)
    loss = crit(rgb, gt)
    if timed: t = tick("loss", t)
    loss.backward()
    if timed: t = tick("backward", t)
    opt.step()
    if timed: t = tick("adam", t)
for _ in range(5): step(False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step(False)
torch.cuda.synchronize()
print("untimed loop ms/step", (time.perf_counter() - t0) / 20 * 1e3)
for _ in range(20): step(True)
for k, v in T.items(): print(f"{k:16s} {v / 20 * 1e3:8.3f} ms")
