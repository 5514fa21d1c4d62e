"""Multi-GPU plumbing: one process per GPU (time_tuning.py:516-521,717), clips sharded by rank,
and the library's own NCCL communicator for the Sinkhorn marginal all-reduce
(my_utils.py:259-272).  torch.distributed is used only to hand the 128-byte NCCL unique id
from rank 0 to the other ranks."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _cabi, ops


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous equal shards; like the reference's equal-B assumption (my_utils.py:257) the
    item count must divide evenly."""
    if n_items % world_size:
        raise ValueError(f"{n_items} clips do not shard evenly over {world_size} ranks")
    per = n_items // world_size
    return range(rank * per, (rank + 1) * per)


def broadcast_bytes(payload, nbytes: int, src: int = 0, device="cpu") -> bytes:
    """Broadcast a fixed-size byte string from `src` over the default process group."""
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm(p2p: bool = True):
    """Create the library communicator for the current default process group (no-op for 1 rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    if ops._comm["handle"] is not None:
        return ops._comm["handle"]
    lib = _cabi.lib()
    rank, ws = dist.get_rank(), dist.get_world_size()
    uid = None
    if rank == 0:
        raw = C.create_string_buffer(_cabi.UNIQUE_ID_BYTES)
        _cabi.check(lib.timet_comm_unique_id(raw), "comm_unique_id")
        uid = raw.raw
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    uid = broadcast_bytes(uid, _cabi.UNIQUE_ID_BYTES, 0, dev)
    handle = C.c_void_p()
    _cabi.check(lib.timet_comm_init(uid, rank, ws, C.byref(handle)), "comm_init")
    ops._comm.update(handle=handle, world_size=ws, rank=rank)
    if p2p and dist.get_backend() == "nccl" and os.environ.get("TIMET_SK_P2P", "1") != "0":
        # NVLink peer-memory exchange for the resident Sinkhorn kernel: all-gather the 64-byte CUDA IPC handles.
        # Every rank must end up on the same path, so success is agreed with a MIN all-reduce.
        ok = 1
        blob = b""
        try:
            raw = C.create_string_buffer(_cabi.IPC_HANDLE_BYTES)
            _cabi.check(lib.timet_comm_p2p_handle(handle, raw), "comm_p2p_handle")
            mine = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).to(dev)
        except RuntimeError:
            ok, mine = 0, torch.zeros(_cabi.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(ws)]
        dist.all_gather(allh, mine)
        if ok:
            try:
                blob = b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh)
                _cabi.check(lib.timet_comm_p2p_connect(handle, blob), "comm_p2p_connect")
            except RuntimeError:
                ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            _cabi.check(lib.timet_comm_p2p_disable(handle), "comm_p2p_disable")
        ops._comm["p2p"] = bool(int(flag.item()))
    else:
        ops._comm["p2p"] = False
    # Every rank has loaded (or built) the library and mapped its peers by now; make that a synchronisation point so
    # that the first in-kernel peer exchange does not start minutes apart on different ranks.  (The in-kernel waits are
    # wall-clock bounded anyway: env TIMET_P2P_TIMEOUT_S, default 600 s.)
    dist.barrier()
    return handle


def destroy_comm():
    if ops._comm["handle"] is not None:
        _cabi.check(_cabi.lib().timet_comm_destroy(ops._comm["handle"]), "comm_destroy")
        ops._comm.update(handle=None, world_size=1, rank=0, p2p=False)
