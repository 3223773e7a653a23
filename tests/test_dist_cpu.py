"""world_size-2 gloo test of the multi-GPU host logic (lizard_b200/dist.py): scatter input block ranges,
exchange compressed sizes, gather the concatenated stream, and the way back (scatter the stream, decode, gather blocks).  The per-rank codec is the CPU oracle here (this
test runs without a GPU); on the GPU box the same functions run over NCCL with the CUDA codec (bench.py)."""
import ctypes
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lizard_b200 as lz
from lizard_b200 import dist as lzdist
from tests import refs

BS = lz.BLOCK_SIZE


def _oracle_compress(block: bytes, level: int) -> bytes:
    L = ctypes.CDLL(os.path.join(refs.ROOT, "oracle", "liboracle.so"))
    L.oracle_Lizard_compress.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    cap = len(block) + 64
    dst = ctypes.create_string_buffer(cap)
    n = L.oracle_Lizard_compress(block, dst, len(block), cap, level)
    return dst.raw[:n]


def _oracle_decompress(comp: bytes, cap: int) -> bytes:
    L = ctypes.CDLL(os.path.join(refs.ROOT, "oracle", "liboracle.so"))
    L.oracle_Lizard_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    dst = ctypes.create_string_buffer(max(cap, 1))
    r = L.oracle_Lizard_decompress_safe(comp, dst, len(comp), cap)
    assert r >= 0, r
    return dst.raw[:r]


def _worker(rank, world, port, total, level, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_blocks = (total + BS - 1) // BS
        src = None
        if rank == 0:
            src = torch.frombuffer(bytearray(lz.datagen(total, 50, 4)), dtype=torch.uint8)
        mine, lo, hi = lzdist.scatter_blocks(src, total, BS, "cpu")
        raw = mine.numpy().tobytes()
        comp = [_oracle_compress(raw[i:i + BS], level) for i in range(0, len(raw), BS)]
        sizes = torch.tensor([len(c) for c in comp], dtype=torch.int64)
        all_sizes, my_off = lzdist.exchange_sizes(sizes, n_blocks)
        blob = torch.frombuffer(bytearray(b"".join(comp)), dtype=torch.uint8) if comp else torch.empty(0, dtype=torch.uint8)
        stream = lzdist.gather_stream(blob, all_sizes, n_blocks, "cpu")
        # and back: rank 0 owns the stream, every rank decodes its range, rank 0 collects the blocks
        part, my_sizes, lo2, hi2 = lzdist.scatter_stream(stream, all_sizes if rank == 0 else None, n_blocks, "cpu")
        assert (lo2, hi2) == (lo, hi) and my_sizes.tolist() == sizes.tolist()
        pb, pos, outs = part.numpy().tobytes(), 0, []
        for i, k in enumerate(my_sizes.tolist()):
            cap = min(BS, total - (lo + i) * BS)
            outs.append(_oracle_decompress(pb[pos:pos + k], cap))
            pos += k
        mine_out = torch.frombuffer(bytearray(b"".join(outs)), dtype=torch.uint8) if outs else torch.empty(0, dtype=torch.uint8)
        back = lzdist.gather_blocks(mine_out, total, BS, "cpu")
        if rank == 0:
            q.put((all_sizes.tolist(), stream.numpy().tobytes(), my_off, back.numpy().tobytes()))
        else:
            q.put((lo, hi, my_off))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5 * BS + 1234, 2 * BS, 1000])
def test_scatter_exchange_gather_world2(total):
    if not os.path.exists(os.path.join(refs.ROOT, "oracle", "liboracle.so")):
        pytest.skip("oracle not built")
    level, world = 10, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + total) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, level, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = lz.datagen(total, 50, 4)
    want = [_oracle_compress(data[i:i + BS], level) for i in range(0, len(data), BS)]
    r0 = [g for g in got if isinstance(g[0], list)][0]
    r1 = [g for g in got if not isinstance(g[0], list)][0]
    assert r0[0] == [len(w) for w in want]
    assert r0[1] == b"".join(want)
    assert r0[3] == data                                  # scatter_stream -> decode per rank -> gather_blocks
    lo, hi, off = r1
    assert (lo, hi) == lzdist.block_range(len(want), 1, 2)
    assert off == sum(len(w) for w in want[:lo])


def test_block_range_partition():
    for n in (0, 1, 7, 8, 8192, 65536 + 3):
        for w in (1, 2, 3, 8):
            ranges = [lzdist.block_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
