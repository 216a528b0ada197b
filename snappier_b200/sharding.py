"""Block-range sharding of a batch across the GPUs of one box.

Snappy blocks are independent by construction (SnappyCompressor.cs:40-44: the input is
cut into 64 KiB fragments, the hash table is cleared per fragment, offsets are
fragment-relative), so the only multi-GPU strategy is data parallelism over blocks:

    rank r owns the contiguous block range  [r*N/W, (r+1)*N/W)

There is no exchange step inside the codec, hence no collective on the data path.  When
a batch is born on one rank, it is distributed with ONE scatter of the variable-size byte
ranges and collected with ONE gather(v) -- NCCL has no scatterv/gatherv, so both are a
single grouped send/recv (torch.distributed.batch_isend_irecv -> ncclGroupStart/End),
preceded by a broadcast / all-gather of the per-rank byte counts.

One process per GPU; works on NCCL (CUDA tensors) and on gloo (CPU tensors: that is how
tests/test_sharding_gloo.py covers the N>1 host logic without GPUs).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, order-preserving, balanced to within one item."""
    return n_items * rank // world, n_items * (rank + 1) // world


def shard_ranges(n_items: int, world: int) -> list[tuple[int, int]]:
    return [shard_range(n_items, world, r) for r in range(world)]


def _byte_span(off: torch.Tensor, length: torch.Tensor, lo: int, hi: int) -> tuple[int, int]:
    """Byte range [b0, b1) covered by items lo..hi-1 of a densely packed, ordered batch."""
    if hi <= lo:
        return 0, 0
    return int(off[lo]), int(off[hi - 1]) + int(length[hi - 1])


def _global_rank(group, r: int) -> int:
    """torch.distributed's broadcast / P2POp take GLOBAL ranks; `r` here is a rank inside `group`."""
    return r if group is None else dist.get_global_rank(group, r)


def scatter_batch(base: torch.Tensor | None, off: torch.Tensor | None, length: torch.Tensor | None,
                  src: int = 0, device: torch.device | None = None, group=None):
    """Scatter a densely packed batch (base u8, off i64[N], len i32[N]) from `src` (a rank of `group`) in
    contiguous block ranges.  Returns this rank's (base, off, len) with offsets rebased to
    0, plus (first_item, n_items_total).  Non-src ranks pass None for the three tensors."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gr = lambda r: _global_rank(group, r)
    if device is None:
        device = base.device if base is not None else torch.device("cpu")
    # 1) metadata: N, then per-rank (byte_lo, byte_hi)
    hdr = torch.zeros(1 + 2 * world, dtype=torch.int64, device=device)
    if rank == src:
        n = off.numel()
        hdr[0] = n
        for r, (lo, hi) in enumerate(shard_ranges(n, world)):
            b0, b1 = _byte_span(off, length, lo, hi)
            hdr[1 + 2 * r], hdr[2 + 2 * r] = b0, b1
    dist.broadcast(hdr, gr(src), group=group)
    h = hdr.cpu().tolist()
    n = h[0]
    lo, hi = shard_range(n, world, rank)
    my_b0, my_b1 = h[1 + 2 * rank], h[2 + 2 * rank]
    my_base = torch.empty(max(my_b1 - my_b0, 1), dtype=torch.uint8, device=device)
    my_off = torch.empty(hi - lo, dtype=torch.int64, device=device)
    my_len = torch.empty(hi - lo, dtype=torch.int32, device=device)
    # 2) ONE grouped send/recv for payload + per-item metadata
    ops = []
    if rank == src:
        for r, (rlo, rhi) in enumerate(shard_ranges(n, world)):
            b0, b1 = h[1 + 2 * r], h[2 + 2 * r]
            if r == src:
                my_base[: b1 - b0].copy_(base[b0:b1])
                my_off.copy_(off[rlo:rhi] - b0)
                my_len.copy_(length[rlo:rhi])
                continue
            if rhi > rlo:
                ops.append(dist.P2POp(dist.isend, base[b0:b1].contiguous(), gr(r), group))
                ops.append(dist.P2POp(dist.isend, (off[rlo:rhi] - b0).contiguous(), gr(r), group))
                ops.append(dist.P2POp(dist.isend, length[rlo:rhi].contiguous(), gr(r), group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, my_base[: my_b1 - my_b0], gr(src), group))
        ops.append(dist.P2POp(dist.irecv, my_off, gr(src), group))
        ops.append(dist.P2POp(dist.irecv, my_len, gr(src), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return my_base, my_off, my_len, lo, n


def gather_batch(base: torch.Tensor, off: torch.Tensor, length: torch.Tensor, dst: int = 0, group=None, engine=None):
    """gather(v): every rank contributes its items (base u8, off i64[n_r] (rank-local, any
    layout), len i32[n_r]); `dst` (a rank of `group`) receives them densely packed in rank order.  Returns
    (base, off, len) on dst, (None, None, None) elsewhere.  With `engine` (a snappier_b200.batch.Engine on this rank's
    GPU) the slots are packed by the CUDA pack kernels (snp_pack_batch); without, by torch indexing (CPU / gloo tests)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gr = lambda r: _global_rank(group, r)
    device = base.device
    # compact this rank's items (slots may have slack, e.g. compress output slots)
    n_r = off.numel()
    if engine is not None and base.is_cuda and n_r:
        stream = torch.cuda.current_stream(device).cuda_stream
        _, tot = engine.pack_batch_device(base, off, length, None, stream)      # size query
        total = int(tot)
        dense = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
        engine.pack_batch_device(base, off, length, dense, stream)
    else:
        lens64 = length.to(torch.int64)
        total = int(lens64.sum()) if n_r else 0
        dense_off = torch.cumsum(lens64, 0) - lens64 if n_r else lens64
        if n_r and not bool((dense_off == off).all()):
            idx = torch.repeat_interleave(off - dense_off, lens64) + torch.arange(total, device=device)
            dense = base[idx]
        else:
            dense = base[:total]
    counts = torch.tensor([n_r, total], dtype=torch.int64, device=device)
    allc = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(allc, counts, group=group)
    allc = [c.cpu().tolist() for c in allc]
    ops = []
    if rank == dst:
        n_tot = sum(c[0] for c in allc)
        b_tot = sum(c[1] for c in allc)
        out_base = torch.empty(max(b_tot, 1), dtype=torch.uint8, device=device)
        out_len = torch.empty(n_tot, dtype=torch.int32, device=device)
        i0 = b0 = 0
        for r, (nr, br) in enumerate(allc):
            if r == dst:
                out_base[b0:b0 + br].copy_(dense[:br])
                out_len[i0:i0 + nr].copy_(length)
            elif nr:
                ops.append(dist.P2POp(dist.irecv, out_base[b0:b0 + br], gr(r), group))
                ops.append(dist.P2POp(dist.irecv, out_len[i0:i0 + nr], gr(r), group))
            i0 += nr
            b0 += br
    elif n_r:
        ops.append(dist.P2POp(dist.isend, dense[:total].contiguous(), gr(dst), group))
        ops.append(dist.P2POp(dist.isend, length.contiguous(), gr(dst), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != dst:
        return None, None, None
    l64 = out_len.to(torch.int64)
    return out_base, torch.cumsum(l64, 0) - l64, out_len
