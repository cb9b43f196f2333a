"""Multi-GPU plumbing for the ctypes mirror: one process per GPU (torchrun), torch.distributed only for the rendezvous
(the NCCL unique id travels over it); the boundary all-gather / cost all-reduce of the hot path are issued by
libgvib200.so itself on the problem's stream, through the NCCL library torch already loaded."""
from __future__ import annotations

import ctypes as C
import glob
import os

from . import capi


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def find_libnccl() -> str:
    import torch
    roots = [os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib"),
             os.path.join(os.path.dirname(torch.__file__), "lib"), "/usr/lib/x86_64-linux-gnu"]
    for r in roots:
        for cand in sorted(glob.glob(os.path.join(r, "libnccl.so*"))):
            return cand
    raise ImportError("libnccl not found")


def attach_nccl(ctx: "capi.Context", rank: int, world: int):
    """Create an NCCL communicator over the torch.distributed world and attach it to the gvib200 context."""
    import torch
    import torch.distributed as dist
    path = find_libnccl()
    nccl = C.CDLL(path, mode=C.RTLD_GLOBAL)
    uid = _UniqueId()
    if rank == 0:
        rc = nccl.ncclGetUniqueId(C.byref(uid))
        assert rc == 0, f"ncclGetUniqueId -> {rc}"
    box = [bytes(uid.internal) if rank == 0 else None]
    # c_char arrays stop at NUL when read through .value; go through the raw buffer instead
    if rank == 0:
        box = [C.string_at(C.addressof(uid), 128)]
    dist.broadcast_object_list(box, src=0)
    C.memmove(C.addressof(uid), box[0], 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    torch.cuda.synchronize()
    rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    assert rc == 0, f"ncclCommInitRank -> {rc}"
    lib = capi.load_library()
    capi._check(lib.gvib200_ctx_set_comm(ctx.h, comm, rank, world, path.encode()))
    ctx._nccl = (nccl, comm)
    return comm
