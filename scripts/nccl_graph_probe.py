"""Which way of putting the gradient all-reduce into the training step's CUDA graph works on this stack?
Run under torchrun with 2 ranks; each variant in its own process with a timeout (a hang must not take the box):
    python scripts/nccl_graph_probe.py            # driver: launches the variants
Variants: plain torch.cuda.graph capture of dist.all_reduce after eager warm-up collectives, with the
capture_error_mode values, and NCCL_GRAPH_REGISTER=0.  Prints one line per variant."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    import torch
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    mode = os.environ["PROBE_MODE"]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    x = torch.full((595856,), float(rank + 1), device=dev)
    y = torch.zeros_like(x)
    for _ in range(3):                                  # eager warm-up: communicator + channels exist before capture
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                          # warm-up on the side stream too
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    x.fill_(float(rank + 1))
    torch.cuda.synchronize()
    with torch.cuda.graph(g, capture_error_mode=mode):
        y.copy_(x)
        dist.all_reduce(y, op=dist.ReduceOp.AVG)
        y.mul_(2.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        g.replay()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    want = 2.0 * (1 + dist.get_world_size()) / 2
    ok = bool((y == want).all())
    if rank == 0:
        print(f"PROBE mode={mode} NCCL_GRAPH_REGISTER={os.environ.get('NCCL_GRAPH_REGISTER', '-')} ok={ok} replay_us={dt * 1e6:.1f}", flush=True)
    dist.destroy_process_group()


def main():
    if "PROBE_MODE" in os.environ and "RANK" in os.environ:
        return worker()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    port = 29600
    for mode in ("thread_local", "global", "relaxed"):
        for reg in (None, "0"):
            env = dict(os.environ, PROBE_MODE=mode)
            if reg is not None:
                env["NCCL_GRAPH_REGISTER"] = reg
            port += 1
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                   "--master-port", str(port), os.path.abspath(__file__)]
            proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, start_new_session=True)
            try:
                so, se = proc.communicate(timeout=120)
                lines = [ln for ln in so.splitlines() if ln.startswith("PROBE")]
                print(lines[-1] if lines else f"PROBE mode={mode} reg={reg} rc={proc.returncode} no result: {se[-400:]!r}", flush=True)
            except subprocess.TimeoutExpired:
                import signal
                os.killpg(proc.pid, signal.SIGKILL)          # the whole session: torchrun and its workers, nothing else
                proc.communicate()
                print(f"PROBE mode={mode} NCCL_GRAPH_REGISTER={reg} HUNG (120 s timeout)", flush=True)


if __name__ == "__main__":
    main()
