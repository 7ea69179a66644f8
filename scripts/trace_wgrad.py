"""Developer probe: per-CTA end times of mlp_wgrad_tc_kernel.  Build a library with -DNB_WG_TRACE
(scripts/build_variants.sh mlp_tc.cu trace "-DNB_WG_TRACE"), run this under NERF_B200_LIB=...lib_trace.so and
grep WGTRACE: block, first item, first tile, last item, last tile, global timer (ns) at exit and at entry.
Last measured: CTA durations 390-416 us (mean 404) inside a 420 us launch, 2.0-3.1 us per tile by item kind."""
import sys, os, torch
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
for _ in range(4): tr.step()
torch.cuda.synchronize()
