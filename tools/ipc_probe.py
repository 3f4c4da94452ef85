"""Probe: do CUDA IPC memory handles work between the torchrun ranks of one box?  (multi-GPU design check)"""
import ctypes
import os

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = ctypes.CDLL("libcudart.so.12")


class Handle(ctypes.Structure):
    _fields_ = [("reserved", ctypes.c_ubyte * 64)]


rt.cudaIpcGetMemHandle.argtypes = [ctypes.POINTER(Handle), ctypes.c_void_p]
rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), Handle, ctypes.c_uint]
ptr = ctypes.c_void_p()
assert rt.cudaMalloc(ctypes.byref(ptr), ctypes.c_size_t(1 << 20)) == 0
handle = Handle()
rc = rt.cudaIpcGetMemHandle(ctypes.byref(handle), ptr)
print(rank, "get handle rc", rc, flush=True)
mine = torch.tensor(list(handle.reserved), dtype=torch.uint8, device="cuda")
allh = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allh, mine)
peers = []
for r in range(world):
    if r == rank:
        peers.append(ptr)
        continue
    hb = Handle()
    hb.reserved[:] = allh[r].cpu().tolist()
    p = ctypes.c_void_p()
    rc = rt.cudaIpcOpenMemHandle(ctypes.byref(p), hb, ctypes.c_uint(1))
    print(rank, "open handle of", r, "rc", rc, hex(p.value or 0), flush=True)
    peers.append(p)
# every rank writes its rank+1 into slot `rank` of every peer's buffer
src = torch.full((16,), float(rank + 1), dtype=torch.float64, device="cuda")
for r in range(world):
    dst = ctypes.c_void_p(peers[r].value + rank * 128)
    rc = rt.cudaMemcpy(dst, ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(128), ctypes.c_int(3))
    assert rc == 0, rc
torch.cuda.synchronize()
dist.barrier()
out = torch.empty(16 * world, dtype=torch.float64, device="cuda")
rt.cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ptr, ctypes.c_size_t(128 * world), ctypes.c_int(3))
torch.cuda.synchronize()
print(rank, "buffer", out.view(world, 16)[:, 0].tolist(), flush=True)
can = ctypes.c_int()
for r in range(world):
    if r != local:
        rt.cudaDeviceCanAccessPeer(ctypes.byref(can), local, r)
        print(rank, "can access peer", r, can.value, flush=True)
dist.destroy_process_group()
