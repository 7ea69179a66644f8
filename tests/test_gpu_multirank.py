"""Two-rank NCCL test of the north star's multi-GPU split: ray bands of one frame + one all_gather, and
data-parallel training with one all-reduce per step.  Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box,
where tests/test_distributed_cpu.py (gloo, world size 2) covers the host logic."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("p2p", ["0", "1"])
def test_two_rank_sharded_render_and_dp_training(p2p):
    env = dict(os.environ, NB200_P2P_ALLREDUCE=p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "check_multirank.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=150, env=env, cwd=ROOT)
    logdir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(logdir):          # keep the ranks' full output next to the other GPU-session logs
        with open(os.path.join(logdir, f"multirank_p2p{p2p}.log"), "w") as fh:
            fh.write(out.stdout + "\n---- stderr ----\n" + out.stderr)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["all_ranks_ok"] and res["sharded_frame_bit_identical"] and res["replicas_identical_after_30_steps"]
    assert ("fused into the Adam kernel" in res["launch_mode"]) == (p2p == "1"), (res["launch_mode"], res.get("p2p_error"))
