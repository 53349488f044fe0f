"""Multi-GPU exchange of per-shard verdicts through the C ABI (include/kgv.h "multi-GPU", csrc/kgv_comm.cu).

One context and one communicator per GPU.  The bootstrap (who is rank 0, how the 128-byte NCCL id and the 64-byte peer handles
reach the other ranks) belongs to the host: `ShardComm.from_torch_distributed` uses an initialised torch.distributed group for
exactly that and nothing else - every byte of the data path goes through libkgv (NCCL called from C, or peer stores over NVLink).
"""
import ctypes

import numpy as np

ID_BYTES, HANDLE_BYTES = 128, 64


class ShardComm:
    def __init__(self, ctx, n_ranks, rank, nccl_id=None, slice_capacity=0):
        self.ctx, self.n_ranks, self.rank = ctx, int(n_ranks), int(rank)
        h = ctypes.c_void_p()
        idp = (ctypes.c_uint8 * ID_BYTES).from_buffer_copy(bytes(nccl_id)) if nccl_id is not None else None
        ctx._check(ctx._lib.kgv_comm_create(ctx._h, self.n_ranks, self.rank, ctypes.addressof(idp) if idp is not None else None, int(slice_capacity), ctypes.byref(h)))
        self._h = h

    @staticmethod
    def unique_id(lib):
        buf = (ctypes.c_uint8 * ID_BYTES)()
        rc = lib.kgv_comm_unique_id(ctypes.addressof(buf))
        if rc != 0:
            raise RuntimeError(f"kgv_comm_unique_id failed ({rc}): NCCL unavailable")
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, ctx, slice_capacity=1 << 22, nccl=True, peer=True, group=None):
        """Bootstrap over an initialised torch.distributed group (control plane only)."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", ctx.device)
        nccl_id = None
        if nccl:
            t = torch.zeros(ID_BYTES, dtype=torch.uint8, device=dev)
            if rank == 0:
                t.copy_(torch.frombuffer(bytearray(cls.unique_id(ctx._lib)), dtype=torch.uint8))
            dist.broadcast(t, src=0, group=group)
            nccl_id = bytes(t.cpu().numpy().tobytes())
        c = cls(ctx, world, rank, nccl_id=nccl_id, slice_capacity=slice_capacity if peer else 0)
        if peer and world > 1:
            mine = torch.frombuffer(bytearray(c.export_handle()), dtype=torch.uint8).to(dev)
            allh = torch.zeros(world * HANDLE_BYTES, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            c.import_handles(allh.cpu().numpy().tobytes())
        return c

    def export_handle(self):
        buf = (ctypes.c_uint8 * HANDLE_BYTES)()
        self.ctx._check(self.ctx._lib.kgv_comm_export(self._h, ctypes.addressof(buf)))
        return bytes(buf)

    def import_handles(self, handles):
        assert len(handles) == self.n_ranks * HANDLE_BYTES
        buf = (ctypes.c_uint8 * len(handles)).from_buffer_copy(handles)
        self.ctx._check(self.ctx._lib.kgv_comm_import(self._h, ctypes.addressof(buf)))

    @staticmethod
    def connect_local(comms):
        """several contexts of ONE process: direct peer access instead of IPC handles"""
        arr = (ctypes.c_void_p * len(comms))(*[c._h for c in comms])
        rc = comms[0].ctx._lib.kgv_comm_connect_local(arr, len(comms))
        if rc != 0:
            raise RuntimeError(f"kgv_comm_connect_local failed ({rc})")

    # ---- data path (device pointers) ----
    def allgather(self, local_ptr, nbytes_per_rank, all_ptr):
        self.ctx._check(self.ctx._lib.kgv_shard_allgather(self.ctx._h, self._h, local_ptr, int(nbytes_per_rank), all_ptr))

    def publish_bitmap(self, status_ptr, n):
        e = ctypes.c_uint64()
        self.ctx._check(self.ctx._lib.kgv_shard_publish_bitmap(self.ctx._h, self._h, status_ptr, int(n), ctypes.byref(e)))
        return int(e.value)

    def publish_bytes(self, src_ptr, nbytes):
        e = ctypes.c_uint64()
        self.ctx._check(self.ctx._lib.kgv_shard_publish_bytes(self.ctx._h, self._h, src_ptr, int(nbytes), ctypes.byref(e)))
        return int(e.value)

    def wait(self, epoch, nbytes_per_rank=0, all_ptr=None):
        self.ctx._check(self.ctx._lib.kgv_shard_wait(self.ctx._h, self._h, int(epoch), int(nbytes_per_rank), all_ptr))

    def shard_validation(self, on=True):
        """kgv_set_sharding: split the signature checks of this context's validation calls over the ranks"""
        self.ctx._check(self.ctx._lib.kgv_set_sharding(self.ctx._h, self._h if on else None))

    def close(self):
        if self._h:
            self.ctx._lib.kgv_set_sharding(self.ctx._h, None)
            self.ctx._lib.kgv_comm_destroy(self._h)
            self._h = None
