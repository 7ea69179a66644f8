import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops, _lib
lib = _lib.load()
B, N = 640000 * 2, 64
outs = torch.randn(B, N, 4, device="cuda"); ts = ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=0)
rays = torch.randn(B, 6, device="cuda"); d_rgb = torch.randn(B, 3, device="cuda"); d_outs = torch.empty_like(outs)
for _ in range(3):
    ops.composite_apply(outs, ts, rays, dirs_mode=1, want_alpha_weights=False)
    lib.nb200_composite_backward(_lib.ptr(outs), _lib.ptr(ts), _lib.ptr(rays), 1, _lib.ptr(d_rgb), None, None, None, None, B, N, _lib.ptr(d_outs), _lib.stream_ptr())
torch.cuda.synchronize(); print("ok")
