"""Multi-GPU plumbing for the block codec (BASELINE config 5).

Blocks are independent, so the codec itself needs no collective: every rank compresses / decompresses its own
contiguous range of blocks.  Communication exists only to (1) hand rank r its slice of the input when the input
lives on rank 0, (2) tell everybody every block's compressed size (so each rank knows where its bytes go in the
concatenated stream) and (3) collect the variable-length outputs on rank 0.  All three are written against
torch.distributed so the same code runs over NCCL/NVLink (GPU tensors) and over gloo (CPU tensors, used by the
CPU test-suite with world_size 2).

Every transfer is ONE grouped exchange (`batch_isend_irecv`: a single ncclGroupStart/End over NVLink): rank 0 posts
all of its sends -- or all of its receives, each straight into the final, prefix-summed place of the destination
buffer -- together; nothing is staged through temporaries.  The loops of the reference that these replace are the
per-block loops of lib/lizard_frame.c:544-549 (compress) and :1148-1169 (decode): what is scattered is their input
range, what is gathered is their concatenated output.
"""
import torch
import torch.distributed as dist


def block_range(n_blocks: int, rank: int, world: int):
    """Contiguous, balanced partition: the first (n_blocks % world) ranks own one extra block."""
    base, extra = divmod(n_blocks, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _run(ops):
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def scatter_blocks(src, total_bytes: int, block: int, device, group=None, out=None):
    """Rank 0 holds `src` (uint8 tensor of total_bytes on `device`); every rank returns (its slice, lo, hi).
    `out` (optional) = preallocated destination of at least the slice's size."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_blocks = (total_bytes + block - 1) // block
    lo, hi = block_range(n_blocks, rank, world)
    my_bytes = max(0, min(hi * block, total_bytes) - lo * block)
    mine = out[:my_bytes] if out is not None else torch.empty(my_bytes, dtype=torch.uint8, device=device)
    ops = []
    if rank == 0:
        for r in range(1, world):
            rlo, rhi = block_range(n_blocks, r, world)
            a, b = rlo * block, min(rhi * block, total_bytes)
            if b > a:
                ops.append(dist.P2POp(dist.isend, src[a:b], r, group))       # a slice of a 1-D tensor is contiguous
        _run(ops)
        mine.copy_(src[:my_bytes])
    elif my_bytes:
        _run([dist.P2POp(dist.irecv, mine, 0, group)])
    return mine, lo, hi


def exchange_sizes(my_sizes, n_blocks: int, group=None):
    """all_gather of per-block compressed sizes (int64 tensor, one entry per owned block).
    Returns (all sizes as one tensor of n_blocks entries, exclusive byte offset of this rank's first block)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per = max(block_range(n_blocks, r, world)[1] - block_range(n_blocks, r, world)[0] for r in range(world))
    padded = torch.zeros(per, dtype=torch.int64, device=my_sizes.device)
    padded[: my_sizes.numel()] = my_sizes.to(torch.int64)
    gathered = torch.empty(world * per, dtype=torch.int64, device=my_sizes.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = block_range(n_blocks, r, world)
        parts.append(gathered[r * per: r * per + hi - lo])
    sizes = torch.cat(parts)
    lo, _ = block_range(n_blocks, rank, world)
    return sizes, int(sizes[:lo].sum())


def _rank_bytes(all_sizes, n_blocks: int, world: int):
    """Bytes of the concatenated stream every rank owns (host list), from one device->host read of the prefix sums."""
    csum = torch.cumsum(all_sizes.to(torch.int64), 0)
    ends = [block_range(n_blocks, r, world)[1] for r in range(world)]
    idx = torch.tensor([e - 1 for e in ends if e > 0], dtype=torch.int64, device=csum.device)
    vals = csum[idx].tolist() if idx.numel() else []
    out, prev, k = [], 0, 0
    for e in ends:
        if e > 0:
            out.append(int(vals[k]) - prev)
            prev = int(vals[k])
            k += 1
        else:
            out.append(0)
    return out


def gather_stream(my_bytes, all_sizes, n_blocks: int, device, group=None, out=None):
    """Collect every rank's concatenated compressed bytes on rank 0, in block order: one grouped exchange, every
    receive lands at its final (prefix-summed) offset.  Returns the stream on rank 0 (a view of `out` if given)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per_rank = _rank_bytes(all_sizes, n_blocks, world)
    if rank != 0:
        if per_rank[rank]:
            _run([dist.P2POp(dist.isend, my_bytes[:per_rank[rank]], 0, group)])
        return None
    total = sum(per_rank)
    stream = out[:total] if out is not None else torch.empty(total, dtype=torch.uint8, device=device)
    ops, pos = [], per_rank[0]
    for r in range(1, world):
        if per_rank[r]:
            ops.append(dist.P2POp(dist.irecv, stream[pos:pos + per_rank[r]], r, group))
        pos += per_rank[r]
    _run(ops)
    stream[:per_rank[0]].copy_(my_bytes[:per_rank[0]])
    return stream


def scatter_stream(stream, all_sizes, n_blocks: int, device, group=None, out=None):
    """Inverse of gather_stream: rank 0 holds the concatenated compressed blocks (`stream`) and `all_sizes`; every rank
    returns (its slice of the stream, the sizes of its blocks, lo, hi).  The sizes travel by broadcast."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = all_sizes.to(device=device, dtype=torch.int64) if rank == 0 else torch.empty(n_blocks, dtype=torch.int64, device=device)
    dist.broadcast(sizes, src=0, group=group)
    lo, hi = block_range(n_blocks, rank, world)
    per_rank = _rank_bytes(sizes, n_blocks, world)
    my_bytes = per_rank[rank]
    mine = out[:my_bytes] if out is not None else torch.empty(my_bytes, dtype=torch.uint8, device=device)
    if rank == 0:
        ops, pos = [], per_rank[0]
        for r in range(1, world):
            if per_rank[r]:
                ops.append(dist.P2POp(dist.isend, stream[pos:pos + per_rank[r]], r, group))
            pos += per_rank[r]
        _run(ops)
        mine.copy_(stream[:my_bytes])
    elif my_bytes:
        _run([dist.P2POp(dist.irecv, mine, 0, group)])
    return mine, sizes[lo:hi], lo, hi


def gather_blocks(my_out, total_bytes: int, block: int, device, group=None, out=None):
    """Collect the decoded blocks (fixed size except the last one) on rank 0, in block order, each rank's range
    received in place."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_blocks = (total_bytes + block - 1) // block
    lo, hi = block_range(n_blocks, rank, world)
    if rank != 0:
        a, b = lo * block, min(hi * block, total_bytes)
        if b > a:
            _run([dist.P2POp(dist.isend, my_out[: b - a], 0, group)])
        return None
    whole = out[:total_bytes] if out is not None else torch.empty(total_bytes, dtype=torch.uint8, device=device)
    ops = []
    for r in range(1, world):
        rlo, rhi = block_range(n_blocks, r, world)
        a, b = rlo * block, min(rhi * block, total_bytes)
        if b > a:
            ops.append(dist.P2POp(dist.irecv, whole[a:b], r, group))
    _run(ops)
    b0 = min(hi * block, total_bytes)
    if b0 > 0:
        whole[:b0].copy_(my_out[:b0])
    return whole
