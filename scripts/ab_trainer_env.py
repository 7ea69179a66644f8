"""A/B of Trainer settings given as environment assignments, same box, each in its own process:
    python scripts/ab_trainer_env.py NB200_TRAINER_FORK=0 NB200_TRAINER_FORK=1 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import torch
    sys.path.insert(0, ROOT)
    from nerf_simple_b200 import ops
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.trainer import Trainer
    from nerf_simple_b200.xyz import poses_to_render
    torch.manual_seed(0)
    net = Nerf().cuda()
    poses = torch.stack(poses_to_render(4, -30, 25)).cuda()
    rays = ops.generate_rays(poses, 400, 400, 555.5); gt = torch.rand(rays.shape[0], 3, device="cuda")
    tr = Trainer(net, rays, gt, N=64, batch_size=4096)
    for _ in range(50): tr.step()
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300): tr.step()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 300)
    print(f"{sys.argv[2]:28s} step ms {res}  loss {float(tr.last_loss):.4f}  {tr.launch_mode}", flush=True)
else:
    for setting in sys.argv[1:] * 2:
        k, v = setting.split("=")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", setting], env=dict(os.environ, **{k: v}))
