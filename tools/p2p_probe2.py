#!/usr/bin/env python
"""Single process, two GPUs: kernel on cuda:0 storing into memory of cuda:1 after cudaDeviceEnablePeerAccess."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymn_b200 import _lib

torch.cuda.set_device(0)
n = 1000
src = torch.arange(n, dtype=torch.float32, device="cuda:0")
off = torch.arange(n, dtype=torch.int64, device="cuda:0")
dst = torch.zeros(n, dtype=torch.float32, device="cuda:1")
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
mode = sys.argv[1] if len(sys.argv) > 1 else "cudart"
if mode == "cudart":
    rt = ctypes.CDLL("libcudart.so.12")
    print("setdevice", rt.cudaSetDevice(ctypes.c_int(0)), "enable", rt.cudaDeviceEnablePeerAccess(ctypes.c_int(1), ctypes.c_uint(0)),
          "last", rt.cudaGetLastError())
elif mode == "torchcopy":
    tmp = torch.zeros(16, device="cuda:0")
    dst[:16].copy_(tmp)          # 0 -> 1
    tmp.copy_(dst[:16])          # 1 -> 0
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print("copies done")
_lib.call("mn_votes_unpack", src, off, n, 1, n, dst, n)
torch.cuda.synchronize(0)
print(mode, "store ok:", bool(torch.equal(dst.cpu(), src.cpu())))
