"""Probe: does this box expose NVLS multicast through torch's symmetric memory (plumbing only)?"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device="cuda")
    h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print("rank", rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr),
          "signal_pad_ptrs", len(h.signal_pad_ptrs), flush=True)
    import ctypes
    drv = ctypes.CDLL("libcuda.so.1")
    v = ctypes.c_int()
    drv.cuDeviceGetAttribute(ctypes.byref(v), 132, local)   # CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
    print("rank", rank, "MULTICAST_SUPPORTED", v.value, flush=True)
except Exception as e:
    print("rank", rank, "symm_mem failed:", repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
