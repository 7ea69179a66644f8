"""Small-shape run of the tcgen05 / TMA kernels (forward, forward with saved tiles, delta chain, wgrad, fused render)
for   compute-sanitizer --tool memcheck python scripts/sanitize_chain_kernels.py
Ragged sizes on purpose: 3 rays x 37 samples (one partial tile), 5 x 64 (3 tiles -> 4 training tiles), 129 x 64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import _lib, config, ops          # noqa: E402
from nerf_simple_b200.nets import Nerf                  # noqa: E402

_lib.load()
config.set_precision("bf16")
torch.manual_seed(0)
net = Nerf().cuda()
for B, N in ((3, 37), (5, 64), (129, 64)):
    rays = torch.randn(B, 6, device="cuda")
    ts = ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=0)
    with torch.no_grad():
        out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N)
    net.zero_grad()
    out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N).view(B, N, 4)
    rgb = ops.composite_apply(out, ts, rays, dirs_mode=1)[0]
    rgb.square().mean().backward()
    if N in (32, 64, 128):
        with torch.no_grad():
            ops.render_fused(net, N, rays=rays, seed=1, offset=0)
torch.cuda.synchronize()
print("chain kernels ok", float(sum(p.grad.abs().sum() for p in net.parameters())))
