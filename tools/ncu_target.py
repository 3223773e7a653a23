#!/usr/bin/env python
"""Fixed launch sequence for ncu captures: 1 GiB datagen -P50, one level; WARM compress + decompress rounds, then ONE
compress and ONE decompress.  With `-k regex:lizard_` every kernel of the library matches; a decompress call is 3 launches
(plan, expand, token kernel) when the Huffman pre-pass is on, a compress call is 1.  So

  ncu --set full --import-source on --clock-control none -k regex:lizard_ -s $((WARM*4)) -c 4 -f -o out python tools/ncu_target.py --level L

captures exactly the last encode, plan, expand and decode launch.  Nothing printed under ncu is a bench value."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BS = 1 << 17


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=10)
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--variant", type=int, default=-1)
    a = ap.parse_args()
    import torch
    import lizard_b200 as lz
    dev = torch.device("cuda", 0)
    L = lz.lib()
    assert L.LizardB200_setDevice(0) == 0, L.LizardB200_lastError()
    if a.variant >= 0:
        L.LizardB200_setDecodeVariant.argtypes = [ctypes.c_int]
        assert L.LizardB200_setDecodeVariant(a.variant) == 0
    nbytes = a.size_mib << 20
    n = nbytes // BS
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    lz.datagen_into(h.data_ptr(), nbytes, 50.0, 0)
    d_src = h.to(dev)
    stride = (L.Lizard_compressBound(BS) + 15) // 16 * 16
    d_comp = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    d_back = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    so, co = idx * BS, idx * stride
    sl = torch.full((n,), BS, dtype=torch.int32, device=dev)
    cap = torch.full((n,), BS - 1, dtype=torch.int32, device=dev)
    bcap = torch.full((n,), BS, dtype=torch.int32, device=dev)
    cs = torch.zeros(n, dtype=torch.int32, device=dev)
    ds = torch.zeros(n, dtype=torch.int32, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(a.warm + 1):
        assert L.LizardB200_compress_device(d_src.data_ptr(), so.data_ptr(), sl.data_ptr(), d_comp.data_ptr(), co.data_ptr(),
                                            cap.data_ptr(), cs.data_ptr(), n, a.level, sp) == 0
        assert L.LizardB200_decompress_device(d_comp.data_ptr(), co.data_ptr(), cs.data_ptr(), d_back.data_ptr(), so.data_ptr(),
                                              bcap.data_ptr(), ds.data_ptr(), n, sp) == 0
    torch.cuda.synchronize()
    print("ok", bool(torch.equal(d_back, d_src)), int(cs.sum()))


if __name__ == "__main__":
    main()
