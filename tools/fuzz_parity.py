#!/usr/bin/env python
"""Randomised parity campaign on the CPU builds of the device code (TEST TOOL; needs oracle/_ref, i.e. the build container).

  python tools/fuzz_parity.py encode|decode <seed> <seconds>

encode: random inputs (datagen mixes, low-entropy noise, periodic data, far repeats, zeros; 0 .. 2 inner blocks), random
        level among the implemented ones, random capacity, packed or forced-plain (tagged) hash table; the 1-lane host
        build and the 32-lane emulated warp must return the reference's bytes (-DLIZARD_RESET_MEM build).
decode: valid streams of every level and damaged copies (bit flips, truncation, overwritten headers, appended bytes),
        random capacity; the 1-lane build, the 32-lane emulated warp and both with the Huffman / token pre-passes must
        return the reference's code, and its bytes when the stream obeys the min-offset rule.
The unit tests run fixed samples of the same generators; this is for long runs after touching a parser or the decoder.
"""
import ctypes
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import lizard_b200 as lz    # noqa: E402
from tests import refs      # noqa: E402

BS = 1 << 17
ENC_LEVELS = [10, 30, 11, 31, 21, 41, 22, 42, 13, 17]
DEC_LEVELS = [10, 30, 11, 21, 41, 17, 24, 45]


def gen(rnd, rng):
    k = rnd.randrange(6)
    size = rnd.choice([rnd.randrange(0, 300), rnd.randrange(300, 20000), rnd.randrange(20000, BS + 1), BS,
                       rnd.randrange(BS + 1, 2 * BS + 5000)])
    if k == 0:
        return lz.datagen(size, rnd.choice([10, 30, 50, 70, 90]), rnd.randrange(1000))
    if k == 1:
        return rng.integers(0, rnd.choice([2, 4, 16, 256]), size, dtype=np.uint8).tobytes()
    if k == 2:
        pat = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 40)))
        return (pat * (size // len(pat) + 1))[:size]
    if k == 3:
        a = bytearray(lz.datagen(size, 50, rnd.randrange(1000)))
        for _ in range(rnd.randrange(1, 6)):
            if size > 10:
                lo = rnd.randrange(size)
                hi = min(size, lo + rnd.randrange(1, 5000))
                a[lo:hi] = rng.integers(0, 256, hi - lo, dtype=np.uint8).tobytes()
        return bytes(a)
    if k == 4:
        a = bytearray(rng.integers(0, 256, size, dtype=np.uint8).tobytes())
        for _ in range(rnd.randrange(1, 30)):
            if size > 2000:
                n = rnd.randrange(8, 600)
                s, d = rnd.randrange(0, size - n), rnd.randrange(0, size - n)
                a[d:d + n] = a[s:s + n]
        return bytes(a)
    return bytes(size)


def raw_prefix(stream: bytes) -> int:
    """Bytes of the raw inner blocks a unit starts with, if anything follows them; else 0."""
    ip, total = 1, 0
    while ip + 4 <= len(stream) and stream[ip] == 0x80:
        n = int.from_bytes(stream[ip + 1:ip + 4], "little")
        total += n
        ip += 4 + n
    return total if ip < len(stream) else 0


def main():
    mode, seed, budget = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    ref = refs.ref_parity()
    assert ref is not None, "oracle/_ref not built"
    shim = ctypes.CDLL(os.path.join(ROOT, "lizard_b200", "libhostshim.so"))
    for f in (shim.lzb_host_compress, shim.lzb_emu_compress):
        f.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    shim.lzb_decompress_with_prepass.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_int)]
    rnd, rng = random.Random(seed), np.random.default_rng(seed)
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < budget:
        d = gen(rnd, rng)
        if mode == "encode":
            level = rnd.choice(ENC_LEVELS)
            if level in (13, 17) and len(d) > 40000:
                d = d[:40000]                               # the chain walk is slow under the coroutine emulator
            bound = len(d) + 2 + (len(d) // BS + 1) * 4
            cap = rnd.choice([bound, max(len(d) - 1, 1), bound])
            want = refs.ref_compress(ref, d, level, cap)
            plain = rnd.random() < 0.5
            shim.lzb_force_plain_table(1 if plain else 0)
            for name, f in (("host", shim.lzb_host_compress), ("emu", shim.lzb_emu_compress)):
                dst = ctypes.create_string_buffer(max(cap, 1) + 64)
                r = f(d, len(d), dst, cap, level)
                if dst.raw[:r] != want:
                    bad += 1
                    print("MISMATCH encode", name, level, len(d), cap, plain, r, len(want), flush=True)
        else:
            level = rnd.choice(DEC_LEVELS)
            comp = refs.ref_compress(ref, d, level)
            streams = [comp]
            for _ in range(3):
                b = bytearray(comp)
                m = rnd.randrange(5)
                if m == 0 and b:
                    b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
                elif m == 1:
                    b = b[: rnd.randrange(0, len(b) + 1)]
                elif m == 2 and b:
                    b[rnd.randrange(min(60, len(b)))] = rnd.randrange(256)
                elif m == 3:
                    b += bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 5)))
                elif b:
                    for _ in range(3):
                        b[rnd.randrange(len(b))] = rnd.randrange(256)
                streams.append(bytes(b))
            for s in streams:
                cap = rnd.choice([len(d), len(d), max(len(d) - 1, 0), len(d) + 50])
                rr, ro = refs.ref_decompress(ref, s, cap)
                # the reference does not charge raw inner blocks against the capacity.  With a raw block in
                # front of more data and less room than the unit needs, what it returns (success past the end of dst, or
                # an error met only because it went on) is not a target: we must refuse or agree, and stay inside dst.
                tainted = raw_prefix(s) > 0 and cap < len(d)
                defined = rr > 0 and rr <= cap and refs.stream_obeys_min_offset(s, cap)
                for dec_mode in (0, 1, 4, 5):
                    buf = ctypes.create_string_buffer(b"\xA5" * (cap + 64), cap + 64)
                    jd = ctypes.c_int(0)
                    r = shim.lzb_decompress_with_prepass(s, len(s), buf, cap, dec_mode, ctypes.byref(jd))
                    wrong = (r >= 0 and r != rr) if tainted else r != rr
                    if buf.raw[cap:] != b"\xA5" * 64:
                        wrong = True                        # wrote behind the capacity
                    if wrong or (defined and buf.raw[:r] != ro):
                        bad += 1
                        print("MISMATCH decode", dec_mode, level, len(d), len(s), cap, r, rr, flush=True)
        n += 1
    shim.lzb_force_plain_table(0)
    print("mode", mode, "seed", seed, "cases", n, "mismatches", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
