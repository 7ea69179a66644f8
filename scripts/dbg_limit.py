import sys, os, torch, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import config, ops, _lib
from nerf_simple_b200.nets import Nerf
net = Nerf().cuda(); config.set_precision("bf16")
B, N = 640000, 64
rays = torch.randn(B, 6, device="cuda"); rays[:, :3] *= 0.1
ts = ops.stratified_ts(B, N, 2, 6, device="cuda", seed=1, offset=0)
lines = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
reps = int(os.environ.get("REPS", "40"))
with torch.no_grad():
    for _ in range(3): out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
proc.terminate()
ms = e0.elapsed_time(e1) / reps
sel = [l for t, l in lines if t0 + 0.3 <= t <= t1]
print(f"NB200_DBG={os.environ.get('NB200_DBG','0')}: {ms:.3f} ms {B*N*1186816/ms/1e9:.1f} TFLOP/s | smi samples {len(sel)}: first {sel[:2]} last {sel[-2:]}")
