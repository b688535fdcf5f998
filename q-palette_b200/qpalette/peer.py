"""Exchange region for row-sharded decode (one process per GPU): a device allocation every rank maps from every peer over
CUDA IPC, so that kernels can push their slice of a gathered buffer straight into the peers' memory over NVLink
(`qp_fused_norm_had_xchg`, include/qpalette.h).  torch.distributed is used only to pass the 64-byte IPC handles around."""
import ctypes

import torch
import torch.distributed as dist

from ._cabi import Xchg, check, lib


class _Span:
    """lets torch wrap a raw device pointer (no copy): __cuda_array_interface__ v2"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerRegion:
    FLAG_ALIGN = 256

    def __init__(self, layout, nsites, rank, world, group, device):
        """layout: list of (name, nbytes) buffers placed in the region; flags for `nsites` exchange sites follow them."""
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.offsets, off = {}, 0
        for name, nbytes in layout:
            self.offsets[name] = (off, nbytes)
            off += (nbytes + 255) & ~255
        self.flags_offset = off
        self.nbytes = off + ((nsites * world * 4 + 255) & ~255)
        # every rank runs the same sequence of collectives whatever fails locally, and raises only at the end
        err, raw = None, None
        self.base, self.peer_bases = None, []
        try:
            base = ctypes.c_void_p()
            check(lib().qp_peer_alloc(ctypes.byref(base), self.nbytes))
            self.base = base.value
            handle = ctypes.create_string_buffer(64)
            check(lib().qp_peer_export(self.base, handle))
            raw = handle.raw
        except Exception as ex:
            err = ex
        handles = [None] * world
        dist.all_gather_object(handles, raw, group=group)
        if err is None and any(h is None for h in handles):
            err = RuntimeError("a peer could not export its exchange region")
        if err is None:
            try:
                for r, hb in enumerate(handles):
                    if r == rank:
                        self.peer_bases.append(self.base)
                    else:
                        ptr = ctypes.c_void_p()
                        check(lib().qp_peer_import(ctypes.create_string_buffer(hb, 64), ctypes.byref(ptr)))
                        self.peer_bases.append(ptr.value)
            except Exception as ex:
                err = ex
        dist.barrier(group=group)  # every region is mapped everywhere before anyone pushes
        if err is not None:
            raise err
        i64 = dict(dtype=torch.int64, device=device)
        self.d_bases = torch.tensor(self.peer_bases, **i64)
        self.d_flags = torch.tensor([b + self.flags_offset for b in self.peer_bases], **i64)
        self.epoch = torch.zeros(nsites, dtype=torch.int32, device=device)
        self._keep = []

    def tensor(self, name, dtype):
        """torch view of a buffer of the local region"""
        off, nbytes = self.offsets[name]
        t = torch.as_tensor(_Span(self.base + off, nbytes), device=self.device).view(dtype)
        self._keep.append(t)
        return t

    def xchg(self, name, site):
        """descriptor for completing buffer `name` (each rank owns an equal contiguous slice) at exchange site `site`"""
        off, nbytes = self.offsets[name]
        assert nbytes % (16 * self.world) == 0, f"{name}: {nbytes} bytes do not split into 16-byte aligned slices"
        return Xchg(self.d_bases.data_ptr(), self.d_flags.data_ptr(), self.epoch.data_ptr(), off, nbytes // self.world,
                    self.rank, self.world, site)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for r, b in enumerate(self.peer_bases):
            if r != self.rank:
                lib().qp_peer_close(b)
        dist.barrier(group=self.group)
        lib().qp_peer_free(self.base)
        self.peer_bases = []
