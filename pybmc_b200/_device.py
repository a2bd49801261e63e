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


_STAGE_POOL = None
_STAGE_CHUNK = 2 << 20          # bytes per staging task


def _stage_pool():
    """A few host threads for filling page-locked staging blocks (NumPy copies release the GIL).  One thread moves
    ~6 GB/s, which made the staging pass -- not PCIe -- the largest item between the device-timed and the
    end-to-end prediction figures (33 MB per step: 5 ms).  BMC_STAGE_THREADS overrides the count."""
    global _STAGE_POOL
    if _STAGE_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        n = int(os.environ.get("BMC_STAGE_THREADS", "0")) or max(1, min(8, (os.cpu_count() or 1) // max(
            1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))))
        _STAGE_POOL = (ThreadPoolExecutor(max_workers=n), n)
    return _STAGE_POOL


def to_device(a, dev, dtype=torch.float64):
    """Host array -> contiguous device tensor.  Large arrays are staged through a page-locked block taken
    from PyTorch's caching host allocator (a cached block costs nothing; ``Tensor.pin_memory()`` would
    page-lock a fresh allocation on every call, ~1 ms per MB), filled in row chunks by a few threads, each
    chunk copied to the device asynchronously as soon as it is staged (the DMA of one chunk runs under the
    staging of the next)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    arr = np.asarray(a, dtype=np.float64)
    if arr.size <= 4096:
        t = torch.from_numpy(np.array(arr, order="C", copy=True)).to(dev)
        return t if dtype == torch.float64 else t.to(dtype)
    stage = torch.empty(arr.shape, dtype=torch.float64, pin_memory=True)
    view = stage.numpy()
    rows = arr.shape[0]
    row_bytes = max(1, arr.nbytes // max(rows, 1))
    step = max(1, _STAGE_CHUNK // row_bytes)
    pool, n_threads = _stage_pool()
    if arr.nbytes < 2 * _STAGE_CHUNK or rows < 2 or n_threads < 2:
        np.copyto(view, arr)              # one pass: gathers non-contiguous input as it goes
        t = stage.to(dev, non_blocking=True)
    else:
        t = torch.empty(arr.shape, dtype=torch.float64, device=dev)
        cuts = list(range(0, rows, step)) + [rows]
        jobs = [pool.submit(np.copyto, view[lo:hi], arr[lo:hi]) for lo, hi in zip(cuts[:-1], cuts[1:])]
        with torch.cuda.device(dev):
            for job, lo, hi in zip(jobs, cuts[:-1], cuts[1:]):
                job.result()
                t[lo:hi].copy_(stage[lo:hi], non_blocking=True)
    # the caching host allocator keeps `stage` alive until the copies have run (stream-ordered reuse)
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
