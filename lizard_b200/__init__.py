"""lizard_b200 -- thin ctypes binding over liblizard_b200.so (the C-ABI in include/lizard_b200.h).

The product is the shared library: hand-written sm_100a CUDA kernels behind Lizard's own C API
(`Lizard_compress`, `Lizard_decompress_safe`, ...) plus batch entry points.  This module only exists so
the tests and bench.py (Python) can call that C-ABI; it contains no codec logic and there is NO fallback:
if the library is missing, or no B200 is present, calls raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIZARDB200_LIB") or os.path.join(_HERE, "liblizard_b200.so")     # the override is for A/B builds (tools/)
DATAGEN_PATH = os.path.join(os.path.dirname(_HERE), "tools", "libdatagen.so")   # bench / test input generator, not product code

BLOCK_SIZE = 1 << 17
_lib = None
_dg = None


class LizardB200Error(RuntimeError):
    pass


def lib():
    """Load liblizard_b200.so (built by __graft_entry__.build()); raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LizardB200Error(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        c_int_p = ctypes.POINTER(ctypes.c_int)
        vpp = ctypes.POINTER(ctypes.c_void_p)
        L.Lizard_versionNumber.restype = ctypes.c_int
        L.Lizard_compressBound.argtypes = [ctypes.c_int]
        L.Lizard_sizeofState.argtypes = [ctypes.c_int]
        L.Lizard_compress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.Lizard_compress_extState.argtypes = [ctypes.c_void_p] + L.Lizard_compress.argtypes
        L.Lizard_decompress_safe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.LizardB200_lastError.restype = ctypes.c_char_p
        L.LizardB200_launchCount.restype = ctypes.c_ulonglong
        L.LizardB200_compress_batch.argtypes = [vpp, c_int_p, vpp, c_int_p, c_int_p, ctypes.c_int, ctypes.c_int]
        L.LizardB200_decompress_batch.argtypes = [vpp, c_int_p, vpp, c_int_p, c_int_p, ctypes.c_int]
        dev_args = [ctypes.c_void_p] * 7 + [ctypes.c_uint]
        L.LizardB200_decompress_device.argtypes = dev_args + [ctypes.c_void_p]
        L.LizardB200_compress_device.argtypes = dev_args + [ctypes.c_int, ctypes.c_void_p]
        L.LizardB200_gather_device.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_uint, ctypes.c_void_p]
        L.LizardB200_compress_blocks.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.LizardB200_decompress_blocks.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def _check(status, what):
    if status != 0:
        raise LizardB200Error(f"{what} failed: status {status}: {lib().LizardB200_lastError().decode()}")


def compress_bound(n):
    return lib().Lizard_compressBound(n)


def compress(src: bytes, level: int, cap: int = None) -> bytes:
    """Lizard_compress on host bytes. Returns b'' when the library returns 0 (did not fit / failed)."""
    L = lib()
    cap = L.Lizard_compressBound(len(src)) if cap is None else cap
    dst = ctypes.create_string_buffer(max(cap, 1))
    n = L.Lizard_compress(src, dst, len(src), cap, level)
    return dst.raw[:n]


def decompress(src: bytes, max_size: int):
    """Lizard_decompress_safe on host bytes -> (return code, bytes)."""
    L = lib()
    dst = ctypes.create_string_buffer(max(max_size, 1))
    r = L.Lizard_decompress_safe(src, dst, len(src), max_size)
    return r, (dst.raw[:r] if r > 0 else b"")


def _batch(fn, units, caps, extra):
    n = len(units)
    srcs = (ctypes.c_void_p * n)()
    sizes = (ctypes.c_int * n)()
    dsts = (ctypes.c_void_p * n)()
    dcaps = (ctypes.c_int * n)()
    res = (ctypes.c_int * n)()
    keep, outs = [], []
    for i, u in enumerate(units):
        b = ctypes.create_string_buffer(bytes(u), max(len(u), 1))
        keep.append(b)
        srcs[i] = ctypes.cast(b, ctypes.c_void_p)
        sizes[i] = len(u)
        o = ctypes.create_string_buffer(max(caps[i], 1))
        outs.append(o)
        dsts[i] = ctypes.cast(o, ctypes.c_void_p)
        dcaps[i] = caps[i]
    st = fn(srcs, sizes, dsts, dcaps, res, n, *extra)
    _check(st, "batch call")
    return [(res[i], outs[i].raw[:res[i]] if res[i] > 0 else b"") for i in range(n)]


def compress_batch(units, level, caps=None):
    """LizardB200_compress_batch: list of bytes -> list of (result, compressed bytes)."""
    L = lib()
    caps = [L.Lizard_compressBound(len(u)) for u in units] if caps is None else caps
    return _batch(L.LizardB200_compress_batch, units, caps, (level,))


def decompress_batch(units, caps):
    """LizardB200_decompress_batch: list of compressed bytes -> list of (result, bytes)."""
    return _batch(lib().LizardB200_decompress_batch, units, caps, ())


def _load_dg():
    global _dg
    if _dg is None:
        if not os.path.exists(DATAGEN_PATH):
            raise LizardB200Error(f"{DATAGEN_PATH} not built")
        _dg = ctypes.CDLL(DATAGEN_PATH)
        _dg.lizb200_datagen.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_double, ctypes.c_double, ctypes.c_uint]
    return _dg


def datagen(size: int, match_pct: float = 50.0, seed: int = 0, lit_pct: float = 0.0) -> bytes:
    """Bytes identical to the reference's `datagen -g<size> -P<match_pct> -s<seed>` (tools/datagen.c)."""
    buf = ctypes.create_string_buffer(max(size, 1))
    if _load_dg().lizb200_datagen(buf, size, match_pct, lit_pct, seed) != 0:
        raise LizardB200Error("datagen failed")
    return buf.raw[:size]


def datagen_into(ptr: int, size: int, match_pct: float = 50.0, seed: int = 0, lit_pct: float = 0.0):
    """Same generator, writing into caller memory (e.g. a pinned torch tensor's data_ptr())."""
    if _load_dg().lizb200_datagen(ctypes.c_void_p(ptr), size, match_pct, lit_pct, seed) != 0:
        raise LizardB200Error("datagen failed")


# ---- frame layer (LizardF_*) -----------------------------------------------------------------------------------
class FrameInfo(ctypes.Structure):
    _fields_ = [("blockSizeID", ctypes.c_int), ("blockMode", ctypes.c_int), ("contentChecksumFlag", ctypes.c_int),
                ("frameType", ctypes.c_int), ("contentSize", ctypes.c_ulonglong), ("reserved", ctypes.c_uint * 2)]


class Preferences(ctypes.Structure):
    _fields_ = [("frameInfo", FrameInfo), ("compressionLevel", ctypes.c_int), ("autoFlush", ctypes.c_uint),
                ("reserved", ctypes.c_uint * 4)]


def make_prefs(level, block_id=1, independent=True, checksum=False, content_size=0):
    p = Preferences()
    p.frameInfo.blockSizeID = block_id
    p.frameInfo.blockMode = 1 if independent else 0
    p.frameInfo.contentChecksumFlag = 1 if checksum else 0
    p.frameInfo.contentSize = content_size
    p.compressionLevel = level
    return p


def bind_frame_api(L):
    """Set ctypes signatures of the LizardF_* symbols on a library handle (ours or the compiled reference)."""
    sz = ctypes.c_size_t
    L.LizardF_isError.argtypes = [sz]
    L.LizardF_getErrorName.argtypes = [sz]
    L.LizardF_getErrorName.restype = ctypes.c_char_p
    L.LizardF_compressFrameBound.restype = sz
    L.LizardF_compressFrameBound.argtypes = [sz, ctypes.c_void_p]
    L.LizardF_compressFrame.restype = sz
    L.LizardF_compressFrame.argtypes = [ctypes.c_void_p, sz, ctypes.c_void_p, sz, ctypes.c_void_p]
    L.LizardF_createCompressionContext.restype = sz
    L.LizardF_createCompressionContext.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint]
    L.LizardF_freeCompressionContext.argtypes = [ctypes.c_void_p]
    for name in ("LizardF_compressBegin",):
        getattr(L, name).restype = sz
        getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, sz, ctypes.c_void_p]
    L.LizardF_compressBound.restype = sz
    L.LizardF_compressBound.argtypes = [sz, ctypes.c_void_p]
    L.LizardF_compressUpdate.restype = sz
    L.LizardF_compressUpdate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, sz, ctypes.c_void_p, sz, ctypes.c_void_p]
    for name in ("LizardF_flush", "LizardF_compressEnd"):
        getattr(L, name).restype = sz
        getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, sz, ctypes.c_void_p]
    L.LizardF_createDecompressionContext.restype = sz
    L.LizardF_createDecompressionContext.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint]
    L.LizardF_freeDecompressionContext.restype = sz
    L.LizardF_freeDecompressionContext.argtypes = [ctypes.c_void_p]
    L.LizardF_decompress.restype = sz
    L.LizardF_decompress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(sz), ctypes.c_void_p,
                                     ctypes.POINTER(sz), ctypes.c_void_p]
    L.LizardF_getFrameInfo.restype = sz
    L.LizardF_getFrameInfo.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(sz)]
    return L


def frame_compress(L, data: bytes, prefs) -> bytes:
    """LizardF_compressFrame on a library handle; raises on a frame error."""
    cap = L.LizardF_compressFrameBound(len(data), ctypes.byref(prefs))
    dst = ctypes.create_string_buffer(cap)
    n = L.LizardF_compressFrame(dst, cap, data, len(data), ctypes.byref(prefs))
    if L.LizardF_isError(n):
        raise LizardB200Error("LizardF_compressFrame: " + L.LizardF_getErrorName(n).decode())
    return dst.raw[:n]


def frame_decompress(L, frame: bytes, out_cap: int, chunk: int = 0, dst_chunk: int = 0):
    """Feed a frame to LizardF_decompress (whole, or in `chunk`-byte pieces). Returns (last result, bytes)."""
    ctx = ctypes.c_void_p()
    L.LizardF_createDecompressionContext(ctypes.byref(ctx), 100)
    out = ctypes.create_string_buffer(max(out_cap, 1))
    src = ctypes.create_string_buffer(frame, max(len(frame), 1))
    ip = op = 0
    res = 1
    try:
        while ip < len(frame) or res != 0:
            n_in = len(frame) - ip if not chunk else min(chunk, len(frame) - ip)
            n_out = out_cap - op if not dst_chunk else min(dst_chunk, out_cap - op)
            si = ctypes.c_size_t(n_in)
            so = ctypes.c_size_t(n_out)
            res = L.LizardF_decompress(ctx, ctypes.byref(out, op), ctypes.byref(so), ctypes.byref(src, ip),
                                       ctypes.byref(si), None)
            if L.LizardF_isError(res):
                return res, out.raw[:op]
            ip += si.value
            op += so.value
            if si.value == 0 and so.value == 0 and (n_in == 0 or n_out == 0):
                break
            if res == 0 and ip >= len(frame):
                break
    finally:
        L.LizardF_freeDecompressionContext(ctx)
    return res, out.raw[:op]
