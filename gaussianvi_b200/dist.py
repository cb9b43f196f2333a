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


def attach_nccl(ctx: "capi.Context", rank: int, world: int, mailbox: bool = True):
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
    if mailbox and world > 1:
        attach_mailbox(ctx, rank, world)
    return comm


def attach_mailbox(ctx: "capi.Context", rank: int, world: int) -> bool:
    """Peer-memory exchange: every rank's mailbox handle (CUDA IPC, 64 bytes) travels over torch.distributed, then each
    rank maps its peers' mailboxes (gvib200_ctx_mailbox_create / _connect).  Returns False (and leaves the NCCL path in
    place) when some rank cannot export or map a mailbox, e.g. GPUs without peer access."""
    import torch.distributed as dist
    lib = capi.load_library()
    buf = C.create_string_buffer(64)
    ok = lib.gvib200_ctx_mailbox_create(ctx.h, buf, C.c_size_t(64)) >= 0
    handles = [None] * world
    dist.all_gather_object(handles, bytes(buf.raw) if ok else None)
    if any(h is None for h in handles):
        return False
    blob = C.create_string_buffer(b"".join(handles), 64 * world)
    ok = lib.gvib200_ctx_mailbox_connect(ctx.h, world, rank, blob, C.c_size_t(64)) >= 0
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if not any(flags):
        return False
    if not all(flags):
        raise RuntimeError("gvib200 mailbox: some ranks connected and some did not: " + lib.gvib200_last_error().decode())
    return True


def shard_problems(n_problems: int, rank: int, world: int):
    """Independent problems (BASELINE cfg5, SURVEY 8(e) first bullet) shard over the ranks as contiguous ranges of the
    problem index; returns (first, count) of `rank`.  No collective touches the iteration path: every rank runs its range
    as one block-diagonal batch, the results are gathered once at the end (gather_problem_results)."""
    if not (0 <= rank < world) or n_problems < 0:
        raise ValueError("shard_problems: bad arguments")
    base, extra = divmod(n_problems, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_problem_results(local, group=None):
    """One gather of per-problem results (e.g. the final means [count, S*d]) at the end of a sharded cfg5 run:
    returns the list of every rank's array on every rank (torch.distributed.all_gather_object; any backend)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out
