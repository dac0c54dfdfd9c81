"""ctypes binding of lib/libgdmae_b200.so (the C ABI declared in include/gdmae_b200.h).

There is NO fallback: if the shared object is missing or a call fails, an exception is raised.
PyTorch is used only for device memory and streams (``tensor.data_ptr()``,
``torch.cuda.current_stream()``).
"""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgdmae_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "gdmae_b200.h")

_lib = None


class GdmaeError(RuntimeError):
    pass


def exported_symbols_from_header():
    """Names of every function declared in include/gdmae_b200.h (used by the CPU export test)."""
    with open(HEADER) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gdmae_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GdmaeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(gd-mae_b200 has no CPU / PyTorch fallback path)")
        L = ctypes.CDLL(LIB_PATH)
        L.gdmae_last_error.restype = ctypes.c_char_p
        for name in exported_symbols_from_header():
            fn = getattr(L, name)
            if name.endswith("_bytes"):
                fn.restype = ctypes.c_size_t
            elif name == "gdmae_launch_count" or name.endswith("_offset"):
                fn.restype = ctypes.c_int64
            elif name != "gdmae_last_error":
                fn.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise GdmaeError(f"{what} failed (rc={rc}): {lib().gdmae_last_error().decode()}")


def P(t):
    """Device pointer of a tensor (NULL for None)."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "gd-mae_b200 ops need contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def i64(v):
    return ctypes.c_int64(int(v))


def f32(v):
    return ctypes.c_float(float(v))


def farr(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def iarr(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def parr(tensors):
    return (ctypes.c_void_p * len(tensors))(*[0 if t is None else t.data_ptr() for t in tensors])


# optional per-kernel timing (bench.py): name -> list of (start_event, end_event, algorithmic_bytes)
KERNEL_TIMERS = None


class timed:
    """with timed("sra_fwd", nbytes): <C-ABI call>  - CUDA events on the launching stream."""

    def __init__(self, name, nbytes):
        self.name, self.nbytes = name, nbytes

    def __enter__(self):
        if KERNEL_TIMERS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if KERNEL_TIMERS is not None:
            self.e1.record()
            KERNEL_TIMERS.setdefault(self.name, []).append((self.e0, self.e1, self.nbytes))
        return False


_workspaces = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer per (device, stream) - the C ABI never allocates; work issued on a side stream
    (GDMAE.prefetch_index) must not share scratch with the main stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf
