"""Device plumbing: PyTorch owns device memory and streams, nothing else.

Every helper here fails loudly without a CUDA device -- the package has no CPU path.
"""
import os

import numpy as np
import torch

from . import _lib

_DTYPES = {"float32": (torch.float32, _lib.F32), "float64": (torch.float64, _lib.F64),
           "f32": (torch.float32, _lib.F32), "f64": (torch.float64, _lib.F64)}


def resolve_dtype(dtype):
    """'float32' / 'float64' (or numpy / torch dtypes) -> (torch dtype, C-ABI code)."""
    if isinstance(dtype, torch.dtype):
        key = {torch.float32: "float32", torch.float64: "float64"}.get(dtype)
    else:
        key = str(np.dtype(dtype)) if not isinstance(dtype, str) else dtype
    if key not in _DTYPES:
        raise ValueError(f"dtype must be float32 or float64, got {dtype!r}")
    return _DTYPES[key]


def device(dev=None):
    if not torch.cuda.is_available():
        raise _lib.BmcError("pybmc_b200 needs a CUDA device (sm_100a); there is no CPU path")
    _lib.load()
    if dev is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(dev)


def stream_ptr(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def to_device(a, dev, dtype=torch.float64):
    """Host array -> contiguous device tensor through pinned staging memory."""
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    host = torch.from_numpy(np.array(a, dtype=np.float64, order="C", copy=True))
    if host.numel() > 4096:
        host = host.pin_memory()
    t = host.to(dev, non_blocking=True)
    return t if dtype == torch.float64 else t.to(dtype)


def to_host(t):
    """Device tensor -> NumPy array.  Large results land in page-locked memory taken from PyTorch's
    caching host allocator (a fresh cudaHostAlloc costs ~1 ms per MB, a cached block nothing), ~10x
    faster than a pageable copy; the returned array owns the block until it is garbage-collected."""
    t = t.detach()
    if t.device.type != "cuda" or t.numel() * t.element_size() < (1 << 20):
        return t.cpu().numpy()
    t = t.contiguous()
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return buf.numpy()


def fresh_seed():
    """The reference's generators are seeded from OS entropy on every call
    (pybmc/inference_utils.py:52, pybmc/sampling_utils.py:55); so is ours unless a seed is given."""
    return int.from_bytes(os.urandom(8), "little")
