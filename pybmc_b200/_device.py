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


def on(dev):
    """Context manager making ``dev`` the current CUDA device: the C ABI launches on the current device, so
    every public entry point wraps its launches in this (``device=`` may name any visible GPU)."""
    return torch.cuda.device(dev)


def on_own_device(method):
    """Decorator for methods of objects that carry their device in ``self.dev``."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.dev):
            return method(self, *args, **kwargs)
    return wrapper


def stream_ptr(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def to_device(a, dev, dtype=torch.float64):
    """Host array -> contiguous device tensor.  Large arrays are staged through a page-locked block taken
    from PyTorch's caching host allocator (a cached block costs nothing; ``Tensor.pin_memory()`` would
    page-lock a fresh allocation on every call, ~1 ms per MB) and copied asynchronously."""
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    arr = np.asarray(a, dtype=np.float64)
    if arr.size <= 4096:
        t = torch.from_numpy(np.array(arr, order="C", copy=True)).to(dev)
    else:
        stage = torch.empty(arr.shape, dtype=torch.float64, pin_memory=True)
        np.copyto(stage.numpy(), arr)              # one pass: gathers non-contiguous input as it goes
        t = stage.to(dev, non_blocking=True)
        # the caching host allocator keeps `stage` alive until the copy has run (stream-ordered reuse)
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
