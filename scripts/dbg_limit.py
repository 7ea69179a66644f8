import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import config, ops, _lib
from nerf_simple_b200.nets import Nerf
net = Nerf().cuda(); config.set_precision("bf16")
B, N = 640000, 64
rays = torch.randn(B, 6, device="cuda"); rays[:, :3] *= 0.1
ts = ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=0)
with torch.no_grad():
    for _ in range(2): out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"NB200_DBG={os.environ.get('NB200_DBG','0')}: {ms:.3f} ms {B*N*1186816/ms/1e9:.1f} TFLOP/s")
