"""Soak test of the chain kernels (GPU): N renders of the same 800x800x64 frame must be bit-identical (a missing barrier
or an overlapping shared-memory write shows up as run-to-run differences), in the separate-kernel and the fused path and
in the bf16x3 mode; then S training steps must stay finite with a falling loss.
usage: python scripts/soak.py [renders=100] [train_steps=5000]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops
from nerf_simple_b200.engine import FrameRenderer
from nerf_simple_b200.nets import Nerf
from nerf_simple_b200.trainer import Trainer
from nerf_simple_b200.xyz import poses_to_render

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
torch.manual_seed(0)
net = Nerf().cuda()
poses = torch.stack(poses_to_render(4, -30, 30)).cuda()
f = 800 / (2 * np.tan(0.6911112070083618 / 2))
with torch.no_grad():
    for name, kw, n in (("bf16 separate kernels", dict(precision="bf16", fused=False), R), ("bf16 fused render", dict(precision="bf16", fused=True), R),
                        ("bf16x3", dict(precision="bf16x3", fused=False), max(4, R // 10))):
        first, bad = None, 0
        t0 = time.time()
        for i in range(n):
            r = FrameRenderer(net, 800, 800, f, N=64, seed=7, **kw)
            rgb, disp = r.render_frame(poses, 3)
            if first is None:
                first = (rgb.clone(), disp.clone())
            elif not (torch.equal(rgb, first[0]) and torch.equal(disp, first[1])):
                bad += 1
        torch.cuda.synchronize()
        print(f"{name:24s}: {n} renders of 640,000 rays x 64, {bad} differ from the first, finite={bool(torch.isfinite(first[0]).all())}, {time.time() - t0:.1f} s", flush=True)
rays = ops.generate_rays(poses[:25], 400, 400, 555.5)
gt = (0.5 + 0.5 * torch.sin(rays[:, :3] * 3.0)).contiguous()      # a smooth target the net can fit
tr = Trainer(net, rays, gt, N=64, batch_size=4096)
losses = []
t0 = time.time()
for i in range(S):
    loss = tr.step()
    if i % max(1, S // 10) == 0 or i == S - 1:
        losses.append(float(loss))
torch.cuda.synchronize()
print(f"training: {S} steps of 4096 rays x 64 in {time.time() - t0:.1f} s, loss samples {['%.4f' % l for l in losses]}, "
      f"params finite={all(bool(torch.isfinite(p).all()) for p in net.parameters())}")
tr.close()
