"""Device-memory plumbing shared by the host modules (torch is used for allocation and streams only)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_workspaces: dict = {}


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"expected a torch.Tensor, got {type(t).__name__}")
        if not t.is_cuda:
            raise RuntimeError(
                "optimaltextures_b200 runs on a B200 only: got a CPU tensor and there is no CPU fallback "
                "(move the tensor to cuda, or use the reference implementation)"
            )
        if dev is not None and t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
        dev = t.device
    return dev


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous (copies only when needed)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """A scratch buffer per device AND current stream that only ever grows (so steady-state calls never allocate;
    work enqueued on two streams never shares scratch)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


def stream_ptr(dev: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def call(name: str, *args) -> None:
    _lib.check(getattr(_lib.lib(), name)(*args))
