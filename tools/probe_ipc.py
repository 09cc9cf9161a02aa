"""Probe (2 GPUs, torchrun): libcldet peer buffers (cudaMalloc + IPC opened with MY device current) + a kernel on MY device
writing the PEER's memory."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cl_object_detection_b200 import _lib  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
lib = _lib.load()


def say(*a):
    torch.cuda.synchronize()
    print(rank, *a, flush=True)


ptr = ctypes.c_void_p()
hbuf = ctypes.create_string_buffer(64)
_lib.check(lib.cldet_peer_alloc(1024, ctypes.byref(ptr), hbuf))
handles = [None] * world
dist.all_gather_object(handles, hbuf.raw)
say('alloc', hex(ptr.value))
peers = []
for r in range(world):
    if r == rank:
        peers.append(ptr.value)
        continue
    q = ctypes.c_void_p()
    _lib.check(lib.cldet_peer_open(handles[r], ctypes.byref(q)))
    peers.append(q.value)
    say('opened peer', r, hex(q.value))
dist.barrier()
st_ = torch.cuda.current_stream().cuda_stream


class Holder:
    def __init__(self, p, n):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f4', 'data': (p, False), 'version': 2}


mine = torch.as_tensor(Holder(ptr.value, 256), device=dev)
mine[:] = -3.0
torch.cuda.synchronize()
dist.barrier()
for r in range(world):
    rc = lib.cldet_clip_boxes(peers[r] + 4 * 8 * rank, 1, 7 + rank, 9 + rank, st_)      # kernel on MY device, memory of rank r
    say('kernel on peer', r, 'rc', rc)
dist.barrier()
say('my buffer', mine[:20].tolist())
dist.destroy_process_group()
