"""``LfmComm``: the C-ABI's own collective handle (``lfm_comm_*``, include/lfm_b200.h) from Python.

The sharded batched path needs two collectives -- an integer MIN all-reduce of best-objective keys and an all-gather
of the per-rank winners.  ``multi_start_fit`` runs them through ``torch.distributed`` by default; pass ``comm=LfmComm``
to run them through the library's own NCCL communicator instead (what a host without PyTorch binds).  The 128-byte
NCCL id travels from rank 0 to the other ranks by whatever the launcher offers: ``from_torch_distributed`` uses a
broadcast over an already initialised process group (gloo or nccl), ``from_file`` a shared file.
"""
from __future__ import annotations

import ctypes as C
import os
import time

from . import _lib

ID_BYTES = 128


class LfmComm:
    def __init__(self, world: int, rank: int, id_bytes: bytes):
        if len(id_bytes) != ID_BYTES:
            raise ValueError(f"the NCCL id is {ID_BYTES} bytes")
        _lib.require_device()
        self._h = C.c_void_p()
        buf = C.create_string_buffer(id_bytes, ID_BYTES)
        _lib.check(_lib.lib().lfm_comm_create(C.byref(self._h), int(world), int(rank), buf), "lfm_comm_create")
        self.world, self.rank = int(world), int(rank)

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(ID_BYTES)
        _lib.check(_lib.lib().lfm_comm_unique_id(buf), "lfm_comm_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls) -> "LfmComm":
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.zeros(ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return cls(world, rank, bytes(t.cpu().numpy().tobytes()))

    @classmethod
    def from_file(cls, path: str, world: int, rank: int, timeout: float = 60.0) -> "LfmComm":
        if rank == 0:
            tmp = path + ".tmp"
            with open(tmp, "wb") as fh:
                fh.write(cls.unique_id())
            os.replace(tmp, path)
        t0 = time.time()
        while not (os.path.exists(path) and os.path.getsize(path) == ID_BYTES):
            if time.time() - t0 > timeout:
                raise TimeoutError(f"no NCCL id at {path}")
            time.sleep(0.01)
        with open(path, "rb") as fh:
            return cls(world, rank, fh.read())

    def allreduce_min_i64(self, t, stream=None) -> None:
        """In-place MIN all-reduce of an int64 CUDA tensor on `stream` (default: torch's current stream)."""
        import torch

        assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()
        s = stream if stream is not None else torch.cuda.current_stream()
        _lib.check(_lib.lib().lfm_comm_allreduce_min_i64(self._h, t.data_ptr(), t.numel(), s.cuda_stream),
                   "lfm_comm_allreduce_min_i64")

    def allgather_f64(self, send, recv, stream=None) -> None:
        """recv (world x send.numel()) <- send of every rank, fp64 CUDA tensors."""
        import torch

        assert send.is_cuda and recv.is_cuda and send.dtype == torch.float64 and recv.dtype == torch.float64
        assert recv.numel() == self.world * send.numel() and send.is_contiguous() and recv.is_contiguous()
        s = stream if stream is not None else torch.cuda.current_stream()
        _lib.check(_lib.lib().lfm_comm_allgather_f64(self._h, send.data_ptr(), recv.data_ptr(), send.numel(), s.cuda_stream),
                   "lfm_comm_allgather_f64")

    def close(self) -> None:
        if self._h:
            _lib.lib().lfm_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
