"""Host/device mirrored typed array -- the slice of ssrlcv::Unity<T> the matching path uses.

ref: include/Unity.cuh:144-391 (class), :646-665 setMemoryState, :819-854 transferMemoryTo,
:923-971 checkpoint, :466-531 checkpoint constructor.

Host memory is a numpy structured array (dtypes in ssrlcv_b200.types), device memory is a
flat torch.uint8 CUDA tensor (torch is only the allocator / copy engine here).
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np

from . import checkpoint as ck


class MemoryState(IntEnum):
    null = 0
    cpu = 1
    gpu = 2
    both = 3


null, cpu, gpu, both = MemoryState.null, MemoryState.cpu, MemoryState.gpu, MemoryState.both


class IllegalUnityTransition(RuntimeError):
    pass


class NullUnityException(RuntimeError):
    pass


def _torch():
    import torch

    return torch


class Unity:
    def __init__(self, data, dtype: np.dtype, state: MemoryState = cpu, device: int = 0, n: int | None = None):
        """data: numpy array (state cpu/both), torch uint8 CUDA tensor (state gpu/both) or None (zero-filled, needs n)."""
        self.dtype = np.dtype(dtype)
        self.device_index = device
        self.host: np.ndarray | None = None
        self.device = None          # torch.uint8 tensor, flat
        self.state = MemoryState(state)
        self.fore = MemoryState(state)
        self._packed = None         # cache slot used by MatchFactory (packed descriptor set)
        if data is None:
            if n is None:
                raise ValueError("Unity(None, ...) needs n")
            self.numElements = int(n)
            if self.state in (cpu, both):
                self.host = np.zeros(self.numElements, self.dtype)
            if self.state in (gpu, both):
                self.device = _torch().zeros(self.numElements * self.dtype.itemsize, dtype=_torch().uint8,
                                             device=f"cuda:{device}")
        elif isinstance(data, np.ndarray):
            if self.state == gpu:
                raise IllegalUnityTransition("host data given with state gpu")
            self.host = np.ascontiguousarray(data, self.dtype)
            self.numElements = len(self.host)
            if self.state == both:
                self.fore = cpu
                self._upload()
                self.fore = both
        else:  # torch tensor
            if self.state == cpu:
                raise IllegalUnityTransition("device data given with state cpu")
            t = data.contiguous().view(_torch().uint8).reshape(-1)
            self.device = t
            self.numElements = t.numel() // self.dtype.itemsize if n is None else int(n)
            if self.state == both:
                self._download()

    # -- basic accessors ------------------------------------------------------------------------
    def size(self) -> int:
        return self.numElements

    def getMemoryState(self) -> MemoryState:
        return self.state

    def getFore(self) -> MemoryState:
        return self.fore

    def setFore(self, state: MemoryState) -> None:
        self.fore = MemoryState(state)

    def device_ptr(self) -> int:
        if self.device is None:
            raise NullUnityException("no device memory")
        return self.device.data_ptr()

    # -- transfers ------------------------------------------------------------------------------
    def _upload(self) -> None:
        torch = _torch()
        src = torch.from_numpy(self.host.view(np.uint8).reshape(-1))
        if self.device is None or self.device.numel() != src.numel():
            self.device = torch.empty(src.numel(), dtype=torch.uint8, device=f"cuda:{self.device_index}")
        self.device.copy_(src)
        self._packed = None

    def _download(self) -> None:
        raw = self.device[: self.numElements * self.dtype.itemsize].cpu().numpy()
        self.host = raw.view(self.dtype).copy()

    def transferMemoryTo(self, state: MemoryState) -> None:
        state = MemoryState(state)
        if self.state == null:
            raise NullUnityException("cannot transfer a null Unity")
        if state == cpu:
            if self.device is None:
                raise IllegalUnityTransition("no device memory to transfer from")
            self._download()
            if self.state == gpu:
                self.state = both
        elif state == gpu:
            if self.host is None:
                raise IllegalUnityTransition("no host memory to transfer from")
            self._upload()
            if self.state == cpu:
                self.state = both
        else:
            raise IllegalUnityTransition("transferMemoryTo takes cpu or gpu")
        self.fore = both

    def setMemoryState(self, state: MemoryState) -> None:
        state = MemoryState(state)
        if state == self.state:
            return
        if state == null:
            self.clear()
            return
        if state in (gpu, both) and self.device is None:
            self._upload()
        elif state in (gpu, both) and self.fore == cpu:
            self._upload()
        if state in (cpu, both) and (self.host is None or self.fore == gpu):
            self._download()
        if state == gpu:
            self.host = None
        if state == cpu:
            self.device = None
            self._packed = None
        self.state = state
        self.fore = state

    def setData(self, data, n: int, state: MemoryState) -> None:
        self.__init__(data, self.dtype, state, self.device_index, n)

    def clear(self) -> None:
        self.host = None
        self.device = None
        self._packed = None
        self.numElements = 0
        self.state = null
        self.fore = null

    # -- checkpoints ----------------------------------------------------------------------------
    def checkpoint(self, id: int, dirPath: str = "./") -> str:
        if self.state == null:
            raise NullUnityException("cannot write a checkpoint with a null Unity<T>")
        if self.host is None or self.fore == gpu:
            self._download()
        kind = next(k for k, dt in ck.TYPE_DTYPES.items() if dt == self.dtype)
        return ck.write_uty(dirPath, kind, self.host, int(self.state), id)

    @classmethod
    def from_checkpoint(cls, path: str, device: int = 0) -> "Unity":
        f = ck.read_uty(path)
        u = cls(f.data, f.data.dtype, cpu, device)
        if f.state != cpu:
            u.setMemoryState(MemoryState(f.state))
        return u

    def to_numpy(self) -> np.ndarray:
        if self.host is None or self.fore == gpu:
            self._download()
        return self.host
