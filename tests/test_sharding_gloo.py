"""CPU tests of the N>1 host logic: block-range sharding + the single scatter / gather(v),
run with world_size 2 and 3 on the gloo backend.  The per-rank "engine" here is the oracle
(this is tests/: allowed) -- what is under test is snappier_b200/sharding.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snappier_b200 import sharding
from tests import helpers as H


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 2, 7, 8, 9, 1000, 1 << 20):
        for w in (1, 2, 3, 4, 8):
            r = sharding.shard_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, n_blocks: int):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as O
    from snappier_b200.batch import pack
    try:
        blocks = H.synthetic_blocks(31337, n_blocks, size=8192)
        blocks = [b[: 8192 - 37 * (i % 5)] for i, b in enumerate(blocks)]  # ragged
        if rank == 0:
            base, off, ln = pack(blocks)
            t = (torch.from_numpy(base), torch.from_numpy(off.astype(np.int64)), torch.from_numpy(ln.astype(np.int32)))
        else:
            t = (None, None, None)
        my_base, my_off, my_len, first, n_total = sharding.scatter_batch(*t, src=0, device=torch.device("cpu"))
        lo, hi = sharding.shard_range(n_blocks, world, rank)
        assert (first, n_total) == (lo, n_blocks) and my_off.numel() == hi - lo
        # this rank's shard is byte-identical to the corresponding source blocks
        mb = my_base.numpy()
        mine = [mb[int(o): int(o) + int(l)].tobytes() for o, l in zip(my_off, my_len)]
        assert mine == blocks[lo:hi]
        # compress the shard into slack-y slots (like the GPU compressor does), then gather(v)
        pitch = O.get_max_compressed_length(8192)
        slots = np.zeros(max(len(mine), 1) * pitch, np.uint8)
        s_off = np.arange(len(mine), dtype=np.int64) * pitch
        s_len = np.zeros(len(mine), np.int32)
        for i, b in enumerate(mine):
            c = O.compress(b)[1]
            slots[i * pitch: i * pitch + len(c)] = np.frombuffer(c, np.uint8)
            s_len[i] = len(c)
        g_base, g_off, g_len = sharding.gather_batch(torch.from_numpy(slots), torch.from_numpy(s_off),
                                                     torch.from_numpy(s_len), dst=0)
        if rank == 0:
            gb = g_base.numpy()
            got = [gb[int(o): int(o) + int(l)].tobytes() for o, l in zip(g_off, g_len)]
            assert got == [O.compress(b)[1] for b in blocks]  # order preserved, bit-exact
            assert [O.decompress(c)[1] for c in got] == blocks
        else:
            assert g_base is None
        # checksum-of-checksums all-reduce (the only collective of the "born sharded" bench path)
        local = torch.tensor([sum(O.crc32c(b) for b in mine)], dtype=torch.int64)
        dist.all_reduce(local)
        assert int(local) == sum(O.crc32c(b) for b in blocks)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_blocks", [(2, 23), (3, 10), (2, 1)])
def test_scatter_process_gather_round_trip(oracle, world, n_blocks):
    mp.spawn(_worker, args=(world, _free_port(), n_blocks), nprocs=world, join=True)


def _subgroup_worker(rank: int, world: int, port: int, n_blocks: int):
    """Scatter / gather inside a SUB-group whose ranks differ from the global ones (group {1, 2} of a 3-rank world: group
    rank r is global rank r + 1): sharding.py must translate group ranks to the global ranks torch.distributed's
    broadcast / point-to-point calls take."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snappier_b200.batch import pack
    try:
        grp = dist.new_group(ranks=[1, 2])  # every rank calls new_group; only 1 and 2 are members
        if rank in (1, 2):
            g_rank = dist.get_rank(grp)
            assert g_rank == rank - 1
            blocks = [bytes([i % 251]) * (100 + 13 * i) for i in range(n_blocks)]
            if g_rank == 0:
                base, off, ln = pack(blocks)
                t = (torch.from_numpy(base), torch.from_numpy(off.astype(np.int64)), torch.from_numpy(ln.astype(np.int32)))
            else:
                t = (None, None, None)
            my_base, my_off, my_len, first, n_total = sharding.scatter_batch(*t, src=0, device=torch.device("cpu"), group=grp)
            lo, hi = sharding.shard_range(n_blocks, 2, g_rank)
            mb = my_base.numpy()
            mine = [mb[int(o): int(o) + int(l)].tobytes() for o, l in zip(my_off, my_len)]
            assert (first, n_total) == (lo, n_blocks) and mine == blocks[lo:hi]
            # slots with slack, gathered to group rank 1 (= global rank 2)
            slots = np.zeros(max(len(mine), 1) * 1024, np.uint8)
            for i, b in enumerate(mine):
                slots[i * 1024: i * 1024 + len(b)] = np.frombuffer(b, np.uint8)
            g_base, g_off, g_len = sharding.gather_batch(torch.from_numpy(slots), torch.arange(len(mine), dtype=torch.int64) * 1024,
                                                         torch.tensor([len(b) for b in mine], dtype=torch.int32), dst=1, group=grp)
            if g_rank == 1:
                gb = g_base.numpy()
                assert [gb[int(o): int(o) + int(l)].tobytes() for o, l in zip(g_off, g_len)] == blocks
            else:
                assert g_base is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_scatter_gather_inside_a_subgroup():
    mp.spawn(_subgroup_worker, args=(3, _free_port(), 9), nprocs=3, join=True)
