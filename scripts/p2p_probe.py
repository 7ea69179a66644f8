"""Debug probe for the peer-memory all-reduce (2 ranks):  torchrun ... scripts/p2p_probe.py"""
import os, sys, ctypes
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
n = 4096
from nerf_simple_b200.trainer import PeerBuffer
flag_off = n * 4
buf = PeerBuffer(flag_off + _lib.P2P_FLAG_WORDS * 4, dev)
grad = buf.tensor(torch.float32, n)
flags = buf.tensor(torch.int32, _lib.P2P_FLAG_WORDS, flag_off)
grad.fill_(float(rank + 1))
torch.cuda.synchronize()
print(rank, "ptrs", [hex(q) for q in buf.ptrs], flush=True)
dist.barrier()
state = torch.zeros(_lib.TRAIN_STATE_BYTES, dtype=torch.uint8, device=dev)
_lib.check(lib.nb200_train_state_init(_lib.ptr(state), 0, 0, 0, 0.1, _lib.stream_ptr(dev)))
p = torch.zeros(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
gp, fp = buf.ptr_array(0), buf.ptr_array(flag_off)
for it in range(3):
    rc = lib.nb200_adam_allreduce_p2p(_lib.ptr(p), gp, fp, rank, world, _lib.ptr(m), _lib.ptr(v), n, _lib.ptr(state), 0.9, 0.999, 1e-8, _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    print(rank, "iter", it, "rc", rc, "p[0]", float(p[0]), "flags", flags.tolist()[:20], flush=True)
    _lib.check(lib.nb200_train_state_advance(_lib.ptr(state), 1, 1, 1.0, _lib.stream_ptr(dev)))
dist.barrier()
dist.destroy_process_group()
