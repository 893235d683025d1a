"""Row sharding used by the multi-GPU path (mirror of sb::comm_slice in csrc/gpu/comm.cu).

A block of `total` rows is cut into `world` slices of `chunk = ceil(total / world)` rows (the last ones may be short
or empty); rank r owns rows [r*chunk, min(total, (r+1)*chunk)).  The slices are exchanged with one all-gather of
`chunk` padded rows per rank, so the gathered buffer has world*chunk rows of which the first `total` are meaningful.
"""
from __future__ import annotations


def chunk_rows(total: int, world: int) -> int:
    return (total + world - 1) // world


def slice_rows(total: int, world: int, rank: int) -> tuple[int, int]:
    c = chunk_rows(total, world)
    begin = min(total, rank * c)
    return begin, min(total, begin + c)


def init_comm(lib, dist, device=None) -> None:
    """Create the library's NCCL communicator from an initialised torch.distributed process group: rank 0 makes the
    128-byte unique id, it is broadcast through the process group, every rank calls spasm_b200_comm_init."""
    import ctypes as C

    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        lib.spasm_b200_comm_unique_id(buf)
    t = torch.tensor(list(buf), dtype=torch.uint8, device=device if device is not None else "cpu")
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    lib.spasm_b200_comm_init(rank, world, raw)
