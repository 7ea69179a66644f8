"""Developer probe (GPU): bf16 tcgen05 backward vs golden gradients, per tensor; timing of a train step."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import config, ops, _lib
from nerf_simple_b200.nets import Nerf
from nerf_simple_b200.rendering import render_nerf

g = dict(np.load("tests/golden/case_train_b64_n64.npz"))
W = dict(np.load("tests/golden/weights_seed0.npz"))
net = Nerf().cuda()
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
rays = torch.from_numpy(g["rays"]).cuda(); gt = torch.from_numpy(g["gt"]).cuda()
for prec in ("fp32", "bf16"):
    config.set_precision(prec); config.set_sampler("reference")
    net.zero_grad(); torch.manual_seed(1)
    rgb, *_ = render_nerf(rays, net, 64)
    loss = torch.nn.functional.mse_loss(rgb, gt); loss.backward(); torch.cuda.synchronize()
    print(prec, "loss", loss.item(), "golden", float(g["loss"]))
    for k, p in net.named_parameters():
        ref = g["grad." + k]; sc = np.abs(ref).max()
        err = np.abs(p.grad.cpu().numpy() - ref).max()
        print(f"   {k:28s} scale {sc:.3e} relerr {err/sc:.3e}")
# timing of fwd+bwd at 4096 x 64
config.set_precision("bf16"); config.set_sampler("philox")
B, N = 4096, 64
rays = torch.randn(B, 6, device="cuda"); rays[:, :3] *= 0.1
gt = torch.rand(B, 3, device="cuda")
for it in range(3):
    net.zero_grad(); rgb, *_ = render_nerf(rays, net, N); torch.nn.functional.mse_loss(rgb, gt).backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    net.zero_grad(); rgb, *_ = render_nerf(rays, net, N); torch.nn.functional.mse_loss(rgb, gt).backward()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"bf16 train step (render_nerf+mse+backward, no optimizer) {B}x{N}: {ms:.3f} ms  {B/ms*1e3/1e6:.2f} Mrays/s  {B*N*3489024/ms/1e9:.1f} TFLOP/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(3):
        net.zero_grad(); rgb, *_ = render_nerf(rays, net, N); torch.nn.functional.mse_loss(rgb, gt).backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
