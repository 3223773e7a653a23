"""Multi-GPU plumbing for the block codec (BASELINE config 5).

Blocks are independent, so the codec itself needs no collective: every rank compresses / decompresses its own
contiguous range of blocks.  Communication exists only to (1) hand rank r its slice of the input when the input
lives on rank 0, (2) tell everybody every block's compressed size (so each rank knows where its bytes go in the
concatenated stream) and (3) collect the variable-length outputs on rank 0.  All three are written against
torch.distributed so the same code runs over NCCL/NVLink (GPU tensors) and over gloo (CPU tensors, used by the
CPU test-suite with world_size 2).
"""
import torch
import torch.distributed as dist


def block_range(n_blocks: int, rank: int, world: int):
    """Contiguous, balanced partition: the first (n_blocks % world) ranks own one extra block."""
    base, extra = divmod(n_blocks, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scatter_blocks(src, total_bytes: int, block: int, device, group=None):
    """Rank 0 holds `src` (uint8 tensor of total_bytes on `device`); every rank returns its own slice."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_blocks = (total_bytes + block - 1) // block
    lo, hi = block_range(n_blocks, rank, world)
    my_bytes = max(0, min(hi * block, total_bytes) - lo * block)
    mine = torch.empty(my_bytes, dtype=torch.uint8, device=device)
    if rank == 0:
        reqs = []
        for r in range(1, world):
            rlo, rhi = block_range(n_blocks, r, world)
            a, b = rlo * block, min(rhi * block, total_bytes)
            if b > a:
                reqs.append(dist.isend(src[a:b].contiguous(), dst=r, group=group))
        mine.copy_(src[: my_bytes])
        for q in reqs:
            q.wait()
    elif my_bytes:
        dist.recv(mine, src=0, group=group)
    return mine, lo, hi


def exchange_sizes(my_sizes, n_blocks: int, group=None):
    """all_gather of per-block compressed sizes (int64 tensor, one entry per owned block).
    Returns (all sizes as one tensor of n_blocks entries, exclusive byte offset of this rank's first block)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per = max(block_range(n_blocks, r, world)[1] - block_range(n_blocks, r, world)[0] for r in range(world))
    padded = torch.zeros(per, dtype=torch.int64, device=my_sizes.device)
    padded[: my_sizes.numel()] = my_sizes.to(torch.int64)
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = block_range(n_blocks, r, world)
        parts.append(gathered[r][: hi - lo])
    sizes = torch.cat(parts)
    lo, _ = block_range(n_blocks, rank, world)
    return sizes, int(sizes[:lo].sum())


def gather_stream(my_bytes, all_sizes, n_blocks: int, device, group=None):
    """Collect every rank's concatenated compressed bytes on rank 0, in block order.  Returns the stream on rank 0."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank != 0:
        if my_bytes.numel():
            dist.send(my_bytes.contiguous(), dst=0, group=group)
        return None
    out = torch.empty(int(all_sizes.sum()), dtype=torch.uint8, device=device)
    pos = 0
    for r in range(world):
        lo, hi = block_range(n_blocks, r, world)
        nbytes = int(all_sizes[lo:hi].sum())
        if r == 0:
            out[pos:pos + nbytes].copy_(my_bytes[:nbytes])
        elif nbytes:
            tmp = torch.empty(nbytes, dtype=torch.uint8, device=device)
            dist.recv(tmp, src=r, group=group)
            out[pos:pos + nbytes].copy_(tmp)
        pos += nbytes
    return out


def scatter_stream(stream, all_sizes, n_blocks: int, device, group=None):
    """Inverse of gather_stream: rank 0 holds the concatenated compressed blocks (`stream`) and `all_sizes`; every rank
    returns (its slice of the stream, the sizes of its blocks, lo, hi).  The sizes travel by broadcast."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = all_sizes.to(device=device, dtype=torch.int64) if rank == 0 else torch.empty(n_blocks, dtype=torch.int64, device=device)
    dist.broadcast(sizes, src=0, group=group)
    lo, hi = block_range(n_blocks, rank, world)
    my_bytes = int(sizes[lo:hi].sum())
    mine = torch.empty(my_bytes, dtype=torch.uint8, device=device)
    if rank == 0:
        reqs, pos = [], 0
        for r in range(world):
            rlo, rhi = block_range(n_blocks, r, world)
            nbytes = int(sizes[rlo:rhi].sum())
            if r == 0:
                mine.copy_(stream[pos:pos + nbytes])
            elif nbytes:
                reqs.append(dist.isend(stream[pos:pos + nbytes].contiguous(), dst=r, group=group))
            pos += nbytes
        for q in reqs:
            q.wait()
    elif my_bytes:
        dist.recv(mine, src=0, group=group)
    return mine, sizes[lo:hi], lo, hi


def gather_blocks(my_out, total_bytes: int, block: int, device, group=None):
    """Collect the decoded blocks (fixed size except the last one) on rank 0, in block order."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_blocks = (total_bytes + block - 1) // block
    if rank != 0:
        if my_out.numel():
            dist.send(my_out.contiguous(), dst=0, group=group)
        return None
    out = torch.empty(total_bytes, dtype=torch.uint8, device=device)
    for r in range(world):
        lo, hi = block_range(n_blocks, r, world)
        a, b = lo * block, min(hi * block, total_bytes)
        if b <= a:
            continue
        if r == 0:
            out[a:b].copy_(my_out[: b - a])
        else:
            tmp = torch.empty(b - a, dtype=torch.uint8, device=device)
            dist.recv(tmp, src=r, group=group)
            out[a:b].copy_(tmp)
    return out
