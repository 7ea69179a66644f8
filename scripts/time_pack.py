"""CUDA-event time of nb200_pack_weights (runs once per optimizer step)."""
import ctypes as C
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_simple_b200 import _lib
probe = C.CDLL(_lib.LIB_PATH)
_lib.SYMBOLS = {k: v for k, v in _lib.SYMBOLS.items() if hasattr(probe, k)}
from nerf_simple_b200.nets import Nerf
lib = _lib.load()
net = Nerf().cuda()
params = net.kernel_params()
for prec in (_lib.BF16, _lib.BF16X3):
    buf = torch.empty(lib.nb200_packed_weights_bytes(prec), dtype=torch.uint8, device="cuda")
    pa = _lib.ptr_array(params)
    for _ in range(20):
        lib.nb200_pack_weights(prec, pa, _lib.ptr(buf), _lib.stream_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500):
        lib.nb200_pack_weights(prec, pa, _lib.ptr(buf), _lib.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print(f"{os.path.basename(_lib.LIB_PATH)} pack_weights precision {prec}: {e0.elapsed_time(e1) / 500 * 1e3:.2f} us per call (2 kernels)")
