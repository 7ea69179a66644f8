"""Developer probe: exercise the plain-CUDA streaming kernels (no tcgen05/TMA) on ragged shapes, for
  compute-sanitizer --tool memcheck python scripts/sanitize_stream_kernels.py
(one tool per gpurun call, smallest case)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
for N in (2, 7, 31, 32, 37, 64, 96, 100, 128, 192, 200, 256, 300):
    for B in (1, 2, 3, 5, 77):
        outs = torch.randn(B, N, 4, device="cuda", requires_grad=True)
        ts = ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=3)
        rays = torch.randn(B, 6, device="cuda")
        for aw in (False, True):
            res = ops.composite_apply(outs, ts, rays, dirs_mode=1, want_alpha_weights=aw)
            sum(r.sum() for r in res).backward()
        res = ops.composite_apply(outs, ts, rays[:, 3:].contiguous(), dirs_mode=0)
        res[0].sum().backward()
        u = torch.rand(B, N, device="cuda")
        ops.stratified_ts(B, N, 2, 6, u=u)
poses = torch.eye(4, device="cuda")[None].repeat(3, 1, 1)
ops.generate_rays(poses, 17, 13, 20.0, ray_begin=5, n_rays=3 * 17 * 13 - 5)
table = torch.randn(1001, 6, device="cuda"); gtt = torch.rand(1001, 3, device="cuda")
r = torch.empty(333, 6, device="cuda"); g = torch.empty(333, 3, device="cuda"); ids = torch.empty(333, dtype=torch.int64, device="cuda")
assert lib.nb200_select_rays(_lib.ptr(table), _lib.ptr(gtt), 1001, 1, 0, 333, _lib.ptr(r), _lib.ptr(g), _lib.ptr(ids), _lib.stream_ptr()) == 0
d = torch.empty(333, 3, device="cuda"); loss = torch.zeros((), device="cuda")
assert lib.nb200_mse_loss_grad(_lib.ptr(g), _lib.ptr(gtt[:333].contiguous()), 333, _lib.ptr(d), _lib.ptr(loss), _lib.stream_ptr()) == 0
n = 10007
p, gr, m, v = (torch.randn(n, device="cuda") for _ in range(4))
v.abs_()
assert lib.nb200_adam_step(_lib.ptr(p), _lib.ptr(gr), _lib.ptr(m), _lib.ptr(v), n, 1, 5e-4, 0.9, 0.999, 1e-8, _lib.stream_ptr()) == 0
from nerf_simple_b200.hierarchical import sample_pdf_merge
w = torch.rand(77, 64, device="cuda"); tsc = ops.stratified_ts(77, 64, 2, 6, device="cuda", seed=1, offset=0)
sample_pdf_merge(tsc, w, 128, seed=3, offset=0)
ops.positional_encoding(torch.randn(1001, 6, device="cuda"))
torch.cuda.synchronize()
print("stream kernels ok")
