"""Developer probe (GPU): bf16 tcgen05 forward vs fp32 SIMT forward vs golden; quick timings."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import config, ops, _lib
from nerf_simple_b200.nets import Nerf

g = dict(np.load("tests/golden/case_train_b64_n64.npz"))
W = dict(np.load("tests/golden/weights_seed0.npz"))
net = Nerf().cuda()
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
q = torch.from_numpy(g["query"]).cuda()
ref = torch.from_numpy(g["out"]).cuda()
with torch.no_grad():
    config.set_precision("fp32"); o32 = net(q)
    print("fp32 max err", float((o32 - ref).abs().max()))
    config.set_precision("bf16"); o16 = net(q)
    torch.cuda.synchronize()
    err = (o16 - ref).abs()
    print("bf16 max err per channel", err.max(0).values.tolist(), "mean", err.mean(0).tolist())
    print("ref absmax per channel", ref.abs().max(0).values.tolist())
    bad = (err.max(1).values > 1e-2).nonzero().flatten()
    print("rows > 1e-2:", bad.numel(), bad[:20].tolist())
    if bad.numel():
        print(o16[bad[:4]], ref[bad[:4]])
    # ragged
    o = net(q[:1037]); print("ragged err", float((o - ref[:1037]).abs().max()))
    # timing: training-size and render-size forward, rays mode
    for B, N in ((4096, 64), (65536, 64), (640000, 64)):
        rays = torch.randn(B, 6, device="cuda"); rays[:, :3] *= 0.1; 
        ts = ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=0)
        for prec in ("bf16",) + (("fp32",) if B <= 65536 else ()):
            config.set_precision(prec)
            for _ in range(2): out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 5
            for _ in range(reps): out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            fl = B * N * 1186816 / (ms * 1e-3) / 1e12
            print(f"{prec} fwd B={B} N={N}: {ms:.3f} ms  {B/(ms*1e-3)/1e6:.2f} Mrays/s  {fl:.1f} TFLOP/s  finite={bool(torch.isfinite(out).all())}")
