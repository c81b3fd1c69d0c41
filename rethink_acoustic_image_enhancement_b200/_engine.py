"""Host-side plumbing shared by the drop-in modules: packed-weight cache, workspace, stream.

PyTorch is used only for device memory (torch.empty), the current CUDA stream and parameters;
every arithmetic op of the forward runs in libkdlae_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib


def default_precision() -> str:
    p = os.environ.get("KDLAE_B200_PRECISION", "bf16").lower()
    if p not in _lib.PRECISIONS:
        raise ValueError(f"KDLAE_B200_PRECISION must be one of {list(_lib.PRECISIONS)}, got {p!r}")
    return p


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """Per-module runtime state: packed weights (invalidated when parameters change) and workspace."""

    def __init__(self, kind: str):
        self.kind = kind
        self._packed = {}      # (device, prec) -> (signature, uint8 tensor)
        self._ws = {}          # (device, stream) -> uint8 tensor
        self._params = None    # cached parameter / buffer list in state_dict() order
        self._checked = set()

    # -- device / library -------------------------------------------------------------------
    def require_cuda(self, t: torch.Tensor, name: str) -> None:
        if not t.is_cuda:
            raise RuntimeError(
                f"{name}: the B200-native forward needs CUDA tensors (got device {t.device}); "
                "there is no CPU fallback - move the module and inputs to a B200 GPU")
        lib = _lib.load()
        idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
        if idx not in self._checked:
            _lib.check(lib.kdlae_device_check(idx), "kdlae_device_check")
            self._checked.add(idx)

    @staticmethod
    def stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    # -- packed weights ---------------------------------------------------------------------
    def packed(self, tensors: Sequence[Optional[torch.Tensor]], device: torch.device, prec: int, nbytes: int, pack_fn):
        """Return the packed-weight blob, re-packing when any source tensor changed (version / storage)."""
        sig = tuple((0, 0) if t is None else (t.data_ptr(), t._version) for t in tensors)
        key = (device, prec)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        srcs: List[Optional[torch.Tensor]] = []
        for t in tensors:
            if t is None:
                srcs.append(None)
                continue
            if t.device != device:
                raise RuntimeError(f"parameter on {t.device} but input on {device}: call module.to(device) first")
            srcs.append(t.detach().to(torch.float32).contiguous())
        arr = (C.c_void_p * len(srcs))(*[_ptr(t) for t in srcs])
        blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        pack_fn(arr, len(srcs), blob)
        # packing is enqueued on the current stream; keep srcs alive until it has run
        torch.cuda.current_stream(device).synchronize()
        self._packed[key] = (sig, blob)
        return blob

    def invalidate(self) -> None:
        self._packed.clear()
        self._params = None

    def tensors(self, module) -> List[Optional[torch.Tensor]]:
        """The module's parameters and buffers in state_dict() order, cached: walking state_dict() (483 entries for the
        teacher) on every forward costs more host time than the launch of a small forward.  The cache holds the tensor objects
        themselves, so in-place updates (optimizer steps) are seen through their version counters; anything that rebinds
        them (load_state_dict, .to(), _apply) goes through invalidate()."""
        if self._params is None:
            self._params = [v for v in module.state_dict(keep_vars=True).values()]
        return self._params

    # -- workspace --------------------------------------------------------------------------
    def workspace(self, device: torch.device, nbytes: int) -> torch.Tensor:
        """Scratch buffer of the forward, one per (device, stream): two forwards of one module on two streams never share
        activations (nothing would order the second call's writes against the first call's reads), and a buffer is only ever
        freed / regrown from the stream it was allocated on, which is the order the caching allocator guarantees."""
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            self._ws[key] = None
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    @staticmethod
    def pick_micro_batch(batch: int, bytes_for_one: int, device: torch.device, cap: int = 16) -> int:
        """Largest micro-batch whose workspace fits a conservative share of free HBM (B200: 180 GB)."""
        free, _total = torch.cuda.mem_get_info(device)
        budget = min(int(free * 0.6), 48 << 30)
        mb = max(1, min(batch, cap, budget // max(1, bytes_for_one)))
        return int(mb)
