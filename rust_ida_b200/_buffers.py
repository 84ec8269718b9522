"""Buffer plumbing: numpy arrays are HOST buffers (the end-to-end path copies them), torch CUDA tensors
are DEVICE buffers (no copies; work is enqueued on torch's current stream).  torch is used for device
memory and streams only -- no torch op is on the compute path."""
import numpy as np

from . import _lib

try:  # torch is optional for host-buffer use
    import torch
except Exception:  # pragma: no cover
    torch = None


def is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def describe(x, dtype, what):
    """-> (pointer, mem).  Checks dtype / contiguity; never copies."""
    if x is None:
        return None, None
    if is_torch(x):
        want = {np.float64: torch.float64, np.int64: torch.int64, np.int32: torch.int32}[dtype]
        if x.dtype != want or not x.is_contiguous():
            raise TypeError("%s must be a contiguous %s tensor" % (what, want))
        if not x.is_cuda:
            raise TypeError("%s: torch tensors must live on a CUDA device (use numpy for host buffers)" % what)
        return x.data_ptr(), _lib.MEM_DEVICE
    if isinstance(x, np.ndarray):
        if x.dtype != dtype or not x.flags["C_CONTIGUOUS"]:
            raise TypeError("%s must be a C-contiguous %s array" % (what, np.dtype(dtype)))
        return x.ctypes.data, _lib.MEM_HOST
    raise TypeError("%s: expected numpy array or torch CUDA tensor, got %r" % (what, type(x)))


def common_mem(*mems):
    ms = {m for m in mems if m is not None}
    if len(ms) != 1:
        raise TypeError("all buffers of one call must be in the same memory space (all numpy or all CUDA tensors)")
    return ms.pop()


def current_stream(mem, device=None):
    """torch's current stream OF `device` (not of torch's current device) for device buffers"""
    if mem == _lib.MEM_DEVICE and torch is not None and torch.cuda.is_available():
        return torch.cuda.current_stream(device).cuda_stream
    return None


def device_of(*xs):
    """CUDA device index shared by the torch tensors among xs (None if there are none); raises if they disagree"""
    devs = {x.device.index for x in xs if is_torch(x) and x.is_cuda}
    if len(devs) > 1:
        raise ValueError("buffers live on different CUDA devices: %s" % sorted(devs))
    return devs.pop() if devs else None


def tensors_device(*xs):
    """context manager making the CUDA device of the tensors among xs current for a Level-1 call (whose kernels run on
    the current device), restoring the previous one afterwards; a no-op for host buffers"""
    import contextlib
    dev = device_of(*xs)
    if dev is None or torch is None:
        return contextlib.nullcontext(), None
    return torch.cuda.device(dev), dev


class on_device:
    """Run C-ABI calls with `device` current and put the caller's current device back afterwards (every compute entry
    point of the library calls cudaSetDevice on the handle's device)."""

    def __init__(self, device):
        self.device, self.prev = device, None

    def __enter__(self):
        if torch is not None and torch.cuda.is_available() and self.device is not None:
            self.prev = torch.cuda.current_device()
        return self

    def __exit__(self, *a):
        if self.prev is not None and self.prev != self.device:
            torch.cuda.set_device(self.prev)
        return False


def empty_like_space(ref, shape, dtype):
    """allocate an output in the same memory space as `ref`"""
    if is_torch(ref):
        want = {np.float64: torch.float64, np.int64: torch.int64, np.int32: torch.int32}[dtype]
        return torch.empty(shape, dtype=want, device=ref.device)
    return np.empty(shape, dtype=dtype)
