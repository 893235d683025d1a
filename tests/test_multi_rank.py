"""Host-side logic of the multi-GPU path on CPU: two gloo ranks cut a block of rows with the library's slicing rule,
each fills its slice, one padded all-gather rebuilds the block on every rank (what comm.cu does with ncclAllGather on
the dense blocks)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, width, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from spasm_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunk = sharding.chunk_rows(total, world)
    begin, end = sharding.slice_rows(total, world, rank)
    full = np.arange(total * width, dtype=np.int32).reshape(total, width) * 7 % 42013      # what one GPU would compute
    mine = torch.zeros((chunk, width), dtype=torch.int32)
    mine[: end - begin] = torch.from_numpy(full[begin:end].copy())
    gathered = [torch.zeros((chunk, width), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, mine)
    block = torch.cat(gathered)[:total].numpy()
    ok = np.array_equal(block, full)
    # the timing reduction of bench.py: max over ranks
    v = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    out[rank] = bool(ok and v.item() == world)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1000, 371, 1, 7])
def test_two_rank_row_sharding_roundtrip(total):
    import torch.multiprocessing as mp
    world, width = 2, 33
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    procs = [mp.Process(target=_worker, args=(r, world, port, total, width, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(0) is True and out.get(1) is True


def test_slices_cover_every_row_once():
    from spasm_b200 import sharding
    for total in (0, 1, 5, 8, 1000, 1001, 2371):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                b, e = sharding.slice_rows(total, world, r)
                assert 0 <= b <= e <= total and e - b <= sharding.chunk_rows(total, world)
                seen += list(range(b, e))
            assert seen == list(range(total))


def _worker_pieces(rank, world, port, nrows, cap, out):
    """the exchange of spasm_rref / spasm_kernel (csrc/gpu/api.cu: exchange_pieces) with gloo: every rank reduces a
    contiguous slice of the rows in batches of `cap` rows -> variable-length CSR pieces; one all-gather of the piece
    sizes, then one broadcast per piece from its owner; every rank ends with all the pieces in row order."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from spasm_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    lens = rng.integers(1, 9, size=nrows)                      # what one GPU would compute: row i has lens[i] entries
    ptr = np.concatenate([[0], np.cumsum(lens)])
    vals = (np.arange(ptr[-1], dtype=np.int64) * 13 % 42013).astype(np.int32)
    begin, end = sharding.slice_rows(nrows, world, rank)
    mine = [(b, min(end, b + cap)) for b in range(begin, end, cap)]
    MAXP = 8
    meta = torch.zeros(2 * MAXP + 1, dtype=torch.int64)
    meta[0] = len(mine)
    for k, (b, e) in enumerate(mine):
        meta[1 + 2 * k], meta[2 + 2 * k] = e - b, int(ptr[e] - ptr[b])
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    rows_seen, chunks = 0, []
    for r in range(world):
        for k in range(int(metas[r][0])):
            nr, nz = int(metas[r][1 + 2 * k]), int(metas[r][2 + 2 * k])
            if r == rank:
                b, e = mine[k]
                buf = torch.from_numpy(vals[ptr[b]:ptr[e]].copy())
            else:
                buf = torch.zeros(nz, dtype=torch.int32)
            if nz:
                dist.broadcast(buf, src=r)
            chunks.append(buf.numpy())
            rows_seen += nr
    got = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
    out[rank] = bool(rows_seen == nrows and np.array_equal(got, vals))
    dist.destroy_process_group()


@pytest.mark.parametrize("nrows,cap", [(1000, 300), (5, 300), (1, 1), (0, 4), (999, 1000)])
def test_two_rank_variable_length_piece_exchange(nrows, cap):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    procs = [mp.Process(target=_worker_pieces, args=(r, world, port, nrows, cap, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(0) is True and out.get(1) is True
