"""A/B of library builds on the three MLP kernels of one training step (4096 rays x 64):
    python scripts/ab_train_kernels.py lib_a.so lib_b.so ...
Each library in its own process (NERF_B200_LIB); prints CUDA-event times of forward-with-save and backward."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import ctypes as C
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    from nerf_simple_b200 import _lib
    probe = C.CDLL(_lib.LIB_PATH)
    _lib.SYMBOLS = {k: v for k, v in _lib.SYMBOLS.items() if hasattr(probe, k)}
    from nerf_simple_b200 import ops
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.trainer import Trainer
    from nerf_simple_b200.xyz import poses_to_render
    torch.manual_seed(0)
    net = Nerf().cuda()
    poses = torch.stack(poses_to_render(4, -30, 25)).cuda()
    rays = ops.generate_rays(poses, 400, 400, 555.5); gt = torch.rand(rays.shape[0], 3, device="cuda")
    tr = Trainer(net, rays, gt, N=64, batch_size=4096, precision=os.environ.get("AB_PRECISION", "bf16"))
    for _ in range(30): tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): tr.step()
    e1.record(); torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 200
    tr.part_events.clear()
    for _ in range(20): tr.step(time_parts=True)
    torch.cuda.synchronize()
    f = np.mean([e[0].elapsed_time(e[1]) for e in tr.part_events]); b = np.mean([e[2].elapsed_time(e[3]) for e in tr.part_events])
    print(f"{os.path.basename(_lib.LIB_PATH):20s} {os.environ.get('AB_PRECISION', 'bf16'):15s} step {step:7.4f} ms   fwd+save {f:7.4f}   bwd {b:7.4f}", flush=True)
else:
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, NERF_B200_LIB=os.path.abspath(lib)))
