"""Device-memory helpers.  torch is used for plumbing only (allocation, streams, torch.distributed):
the kernels run through the C ABI on raw pointers."""
import ctypes as C

import numpy as np
import torch

_DT = {np.dtype(np.uint64): torch.int64, np.dtype(np.int64): torch.int64, np.dtype(np.uint8): torch.uint8,
       np.dtype(np.int8): torch.int8, np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32}


def to_device(a, device="cuda:0"):
    """numpy array -> torch tensor on the device (uint64 travels as int64 bit pattern)."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a)
    return t.to(device, non_blocking=False)


def empty(n, dtype, device="cuda:0"):
    return torch.empty(n, dtype=_DT[np.dtype(dtype)], device=device)


def ptr(t):
    return C.c_void_p(t.data_ptr())


def to_numpy_u64(t):
    return t.cpu().numpy().view(np.uint64)


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
