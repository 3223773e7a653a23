#!/usr/bin/env python
"""Time the host-buffer frame round trip with the library's LIZARDB200_TRACE timeline, plus raw PCIe copy rates."""
import ctypes, os, sys, time
os.environ["LIZARDB200_TRACE"] = "1" if "--trace" in sys.argv else "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lizard_b200 as lz

n = 1 << 30
level = 10
L = lz.lib()
lz.bind_frame_api(L)
src = torch.frombuffer(bytearray(lz.datagen(n)), dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("H2D", lambda: d.copy_(src, non_blocking=True)), ("D2H", lambda: back.copy_(d, non_blocking=True))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("%s 1 GiB pinned: %.2f ms  %.1f GB/s" % (name, dt * 1e3, n / dt / 1e9))
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
t = time.perf_counter()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(s2):
    back.copy_(d, non_blocking=True)
d2.copy_(src, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("H2D + D2H concurrently, 1 GiB each: %.2f ms" % (dt * 1e3))
prefs = lz.make_prefs(level, 1, True, False, 0)
cap = L.LizardF_compressFrameBound(n, ctypes.byref(prefs))
frame = torch.empty(cap, dtype=torch.uint8).pin_memory()
dctx = ctypes.c_void_p(); L.LizardF_createDecompressionContext(ctypes.byref(dctx), 100)
for it in range(4):
    sys.stderr.flush()
    t = time.perf_counter()
    fs = L.LizardF_compressFrame(frame.data_ptr(), cap, src.data_ptr(), n, ctypes.byref(prefs))
    t1 = time.perf_counter()
    so, si = ctypes.c_size_t(n), ctypes.c_size_t(fs)
    r = L.LizardF_decompress(dctx, back.data_ptr(), ctypes.byref(so), frame.data_ptr(), ctypes.byref(si), None)
    t2 = time.perf_counter()
    print("iter %d: compressFrame %.2f ms, decompress %.2f ms, frame %d bytes, r=%d" % (it, (t1 - t) * 1e3, (t2 - t1) * 1e3, fs, r), file=sys.stderr)
assert torch.equal(back, src)
