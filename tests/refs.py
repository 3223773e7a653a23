"""Test-only loaders for the checkers: the compiled reference (oracle/_ref) and our C restatement (oracle/liboracle.so)."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _bind(L):
    L.Lizard_compressBound.argtypes = [ctypes.c_int]
    L.Lizard_sizeofState.argtypes = [ctypes.c_int]
    L.Lizard_compress.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.Lizard_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    return L


def ref_parity():
    """Reference compiled with -DLIZARD_RESET_MEM: the bit-exact target for compression."""
    p = os.path.join(REF_DIR, "liblizard_ref_parity.so")
    return _bind(ctypes.CDLL(p)) if os.path.exists(p) else None


def ref_speed():
    p = os.path.join(REF_DIR, "liblizard_ref_speed.so")
    return _bind(ctypes.CDLL(p)) if os.path.exists(p) else None


def ref_compress(L, data: bytes, level: int, cap: int = None) -> bytes:
    cap = L.Lizard_compressBound(len(data)) if cap is None else cap
    dst = ctypes.create_string_buffer(max(cap, 1))
    n = L.Lizard_compress(data, dst, len(data), cap, level)
    return dst.raw[:n]


def ref_decompress(L, comp: bytes, cap: int):
    # twice the capacity: the reference does not charge a raw inner block against maxDecompressedSize
    # (lib/lizard_decompress.c:164-180 never reduces outputSize), so a compressed inner block behind a raw one may write up
    # to the raw block's size past the capacity; a result > cap tells the caller that this happened
    dst = ctypes.create_string_buffer(2 * max(cap, 1) + 64)
    r = L.Lizard_decompress_safe(comp, dst, len(comp), cap)
    return r, (dst.raw[:r] if r > 0 else b"")


_ORACLE = None


def oracle():
    """Our plain-C restatement (oracle/liboracle.so); None if not built."""
    global _ORACLE
    p = os.path.join(ROOT, "oracle", "liboracle.so")
    if _ORACLE is None and os.path.exists(p):
        L = ctypes.CDLL(p)
        L.oracle_Lizard_compress.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_Lizard_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        L.oracle_last_min_offset.restype = ctypes.c_uint
        _ORACLE = L
    return _ORACLE


def stream_obeys_min_offset(comp: bytes, cap: int) -> bool:
    """True when every match of a (successfully decoded) stream has offset >= 8, the rule all Lizard parsers
    enforce (LIZARD_*_MIN_OFFSET).  Below that the reference's 8-byte granule copies make ITS output depend on
    stale bytes beyond the write cursor, so byte parity is only defined for streams that obey the rule."""
    L = oracle()
    dst = ctypes.create_string_buffer(max(cap, 1) + 64)
    r = L.oracle_Lizard_decompress_safe(comp, dst, len(comp), cap)
    return r > 0 and L.oracle_last_min_offset() >= 8
