"""A/B of library builds on the render hot loop:  python scripts/ab_render.py lib_a.so lib_b.so ...
Each library runs in its own process (NERF_B200_LIB); prints ms per 800x800x64 frame (device time, 8 frames).
AB_PRECISION=bf16x3 times the error-compensated mode."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import ctypes as C
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    from nerf_simple_b200 import _lib
    probe = C.CDLL(_lib.LIB_PATH)
    _lib.SYMBOLS = {k: v for k, v in _lib.SYMBOLS.items() if hasattr(probe, k)}      # older builds lack newer symbols
    from nerf_simple_b200.engine import FrameRenderer
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.xyz import poses_to_render
    torch.manual_seed(0)
    net = Nerf().cuda()
    poses = torch.stack(poses_to_render(4, -30, 30)).cuda()
    rend = FrameRenderer(net, 800, 800, 800 / (2 * np.tan(0.6911112070083618 / 2)), N=64, seed=1, precision=os.environ.get("AB_PRECISION", "bf16"))
    for i in range(3):
        rend.render_frame(poses, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        rend.render_frame(poses, i)
    e1.record()
    torch.cuda.synchronize()
    print(f"{os.path.basename(_lib.LIB_PATH):28s} {e0.elapsed_time(e1) / 8:8.3f} ms/frame", flush=True)
else:
    for lib in sys.argv[1:]:
        env = dict(os.environ, NERF_B200_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env)
