#!/usr/bin/env python
"""dec_bench.py -- kernel-only timing of the decode (and encode) kernel on device-resident data, several levels and both
decode schedules in one process.  A development tool: bench.py is the contract bench.  One JSON line per case.

  python tools/dec_bench.py [--size-mib 1024] [--levels 10,21,41] [--iters 5] [--variants 15,7,3] [--encode]
  variant bits: 1 pooled copy sweeps, 2 compact extension chain, 4 Huffman pre-pass, 8 token pre-pass (LizardB200_setDecodeVariant)
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BS = 1 << 17


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--levels", default="10,21,41")
    ap.add_argument("--variants", default="15,7,3")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--encode", action="store_true")
    ap.add_argument("--enc-shapes", default="", help="encode launch shapes to time, 'warps,tables,ctas' separated by ';' "
                    "(LIZARDB200_ENC_SHAPE is read at every launch); implies --encode")
    ap.add_argument("--no-decode", action="store_true")
    args = ap.parse_args()
    import torch
    import lizard_b200 as lz
    dev = torch.device("cuda", 0)
    L = lz.lib()
    L.LizardB200_setDecodeVariant.argtypes = [ctypes.c_int]
    assert L.LizardB200_setDevice(0) == 0, L.LizardB200_lastError()
    nbytes = args.size_mib << 20
    n = nbytes // BS
    h_src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    lz.datagen_into(h_src.data_ptr(), nbytes, 50.0, 0)
    d_src = h_src.to(dev)
    stride = (L.Lizard_compressBound(BS) + 15) // 16 * 16
    d_comp = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    d_back = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    d_src_off, d_comp_off = idx * BS, idx * stride
    d_src_len = torch.full((n,), BS, dtype=torch.int32, device=dev)
    d_cap = torch.full((n,), BS - 1, dtype=torch.int32, device=dev)
    d_back_cap = torch.full((n,), BS, dtype=torch.int32, device=dev)
    d_csize = torch.zeros(n, dtype=torch.int32, device=dev)
    d_dsize = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    def timed(fn, iters):
        fn(); fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        ev[0].record(stream)
        for k in range(iters):
            fn()
            ev[k + 1].record(stream)
        torch.cuda.synchronize()
        ts = [ev[k].elapsed_time(ev[k + 1]) for k in range(iters)]
        return min(ts), sum(ts) / len(ts)

    for level in [int(x) for x in args.levels.split(",")]:
        def compress():
            s = L.LizardB200_compress_device(d_src.data_ptr(), d_src_off.data_ptr(), d_src_len.data_ptr(), d_comp.data_ptr(),
                                             d_comp_off.data_ptr(), d_cap.data_ptr(), d_csize.data_ptr(), n, level, sp)
            assert s == 0, L.LizardB200_lastError()

        def decompress():
            s = L.LizardB200_decompress_device(d_comp.data_ptr(), d_comp_off.data_ptr(), d_csize.data_ptr(), d_back.data_ptr(),
                                               d_src_off.data_ptr(), d_back_cap.data_ptr(), d_dsize.data_ptr(), n, sp)
            assert s == 0, L.LizardB200_lastError()

        compress()
        torch.cuda.synchronize()
        ctot = int(d_csize.sum())
        algo = nbytes + ctot
        shapes = [x for x in args.enc_shapes.split(";") if x] or ([""] if args.encode else [])
        for shape in shapes:
            if shape:
                os.environ["LIZARDB200_ENC_SHAPE"] = shape
            best, avg = timed(compress, args.iters)
            print(json.dumps({"kernel": "encode", "level": level, "shape": shape or "default", "ms_best": round(best, 3),
                              "ms_avg": round(avg, 3), "MBps": round(nbytes / 1e6 / (avg / 1e3), 1),
                              "algo_GBps": round(algo / 1e9 / (avg / 1e3), 1), "compressed": int(d_csize.sum())}), flush=True)
        os.environ.pop("LIZARDB200_ENC_SHAPE", None)
        if shapes:
            compress()                                  # leave the default shape's output for the decode runs
            torch.cuda.synchronize()
        for v in ([] if args.no_decode else [int(x) for x in args.variants.split(",")]):
            assert L.LizardB200_setDecodeVariant(v) == 0
            d_back.zero_()
            best, avg = timed(decompress, args.iters)
            ok = bool(torch.equal(d_back, d_src)) and int((d_dsize != BS).sum()) == 0
            print(json.dumps({"kernel": "decode", "variant": v, "level": level, "ms_best": round(best, 3), "ms_avg": round(avg, 3),
                              "MBps": round(nbytes / 1e6 / (avg / 1e3), 1), "algo_GBps": round(algo / 1e9 / (avg / 1e3), 1),
                              "frac_of_6560": round(algo / 1e9 / (avg / 1e3) / 6560.6, 4), "round_trip_ok": ok}), flush=True)
        L.LizardB200_setDecodeVariant(7)


if __name__ == "__main__":
    main()
