"""Zero-copy views of a GpuStore's device planes as torch tensors (plumbing only:
torch.distributed / NCCL collectives and synthetic-data fills operate on these views;
no product arithmetic runs through torch)."""
from __future__ import annotations

from . import _native as N


class _Plane:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def values_tensor(store):
    """float32 tensor aliasing the store's cells (length = store.size)."""
    import torch

    if store.size == 0:
        return torch.empty(0, dtype=torch.float32, device="cuda")
    return torch.as_tensor(_Plane(N.lib().olap_store_values_ptr(store._h), store.size, "<f4"), device="cuda")


def status_tensor(store):
    """uint8 tensor aliasing the store's status plane, or None when it has none."""
    import torch

    ptr = N.lib().olap_store_status_ptr(store._h)
    if not ptr:
        return None
    if store.size == 0:
        return torch.empty(0, dtype=torch.uint8, device="cuda")
    return torch.as_tensor(_Plane(ptr, store.size, "|u1"), device="cuda")


def use_torch_stream(stream=None):
    """Make the native library enqueue on a torch stream (default: the current one)."""
    import torch

    stream = stream or torch.cuda.current_stream()
    N.check(N.lib().olap_set_stream(stream.cuda_stream))
    return stream
