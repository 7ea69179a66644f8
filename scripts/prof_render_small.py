"""Developer probe: three 400x400x64 frames through FrameRenderer (for `ncu -k regex:FwdEpi -s 1 -c 1 ...`)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200.engine import FrameRenderer
from nerf_simple_b200.nets import Nerf
from nerf_simple_b200.xyz import poses_to_render
torch.manual_seed(0)
net = Nerf().cuda()
poses = torch.stack(poses_to_render(4, -30, 30)).cuda()
rend = FrameRenderer(net, 400, 400, 400 / (2 * np.tan(0.6911112070083618 / 2)), N=64, seed=1, precision="bf16")
for i in range(3):
    rend.render_frame(poses, i)
torch.cuda.synchronize()
print("ok")
