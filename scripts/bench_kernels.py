"""Developer probe (GPU): achieved HBM bandwidth of the streaming kernels (algorithmic bytes)."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import ops, _lib
lib = _lib.load()
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for N in (64, 128):
    B = 640000 * 4 if N == 64 else 640000 * 2          # >> L2: 1.3 GB of (r,g,b,sigma)
    outs = torch.randn(B, N, 4, device="cuda")
    ts = ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=0)
    rays = torch.randn(B, 6, device="cuda")
    d_rgb = torch.randn(B, 3, device="cuda")
    ms = timeit(lambda: ops.composite_apply(outs, ts, rays, dirs_mode=1, want_alpha_weights=False))
    by = B * (20 * N + 20 + 24)
    print(f"composite fwd  N={N}: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s  ({by/ms/1e6/PEAK*100:.1f}% of {PEAK:.0f})  {B/ms/1e3:.1f} Mrays/s")
    ms = timeit(lambda: ops.composite_apply(outs, ts, rays, dirs_mode=1, want_alpha_weights=True))
    by = B * (28 * N + 20 + 24)
    print(f"composite fwd+aw N={N}: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s  ({by/ms/1e6/PEAK*100:.1f}%)")
    d_outs = torch.empty_like(outs)
    def bwd():
        rc = lib.nb200_composite_backward(_lib.ptr(outs), _lib.ptr(ts), _lib.ptr(rays), 1, _lib.ptr(d_rgb), None, None, None, None,
                                          B, N, _lib.ptr(d_outs), _lib.stream_ptr())
        assert rc == 0
    ms = timeit(bwd)
    by = B * (36 * N + 12 + 24)
    print(f"composite bwd  N={N}: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s  ({by/ms/1e6/PEAK*100:.1f}%)")
    ms = timeit(lambda: ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=0))
    by = B * N * 4
    print(f"sampler philox N={N}: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s  ({by/ms/1e6/PEAK*100:.1f}%)")
    del outs, ts, d_outs
poses = torch.eye(4, device="cuda")[None].repeat(30, 1, 1)
ms = timeit(lambda: ops.generate_rays(poses, 1600, 1600, 2222.2, 0, 30 * 1600 * 1600 // 2))
by = 30 * 1600 * 1600 // 2 * 24
print(f"raygen: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s ({by/ms/1e6/PEAK*100:.1f}%)")
