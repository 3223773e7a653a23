// decode.cuh -- device-side Lizard block decoder (one warp per independent compressed unit).
//
// A "unit" is exactly what one Lizard_decompress_safe() call receives: [level byte] followed by
// inner blocks of <= 128 KiB (reference: lib/lizard_decompress.c:115-264 Lizard_decompress_generic).
// Frame blocks of 128 KiB hold one inner block, so the frame layer hands us thousands of independent
// units per launch; one warp walks one unit, the grid is a persistent set of warps pulling unit
// indices from an atomic counter.
//
// Per inner block the warp
//   1. parses the 5 stream headers (all lanes redundantly; they are ~20 bytes),
//   2. expands Huffman'd streams into its private scratch (lane-parallel over the 4 bitstreams),
//   3. runs the token loop: lanes agree on the (uniform) cursor state and split every literal run and
//      every match copy between them.
// Accept/reject and the returned error codes follow lib/lizard_decompress_lz4.h:7-163 and
// lib/lizard_decompress_liz.h:14-220 check for check (see the comments at each test).
#pragma once
#include "common.cuh"
#include "entropy_dec.cuh"

namespace lzb {

#define LZB_FULL 0xffffffffu

struct DecodeBatch {
    const u8*  src_base;    // compressed bytes of all units
    const u64* src_off;     // [n] byte offset of unit i in src_base
    const u32* src_len;     // [n] compressed size of unit i
    u8*        dst_base;    // output arena
    const u64* dst_off;     // [n] where unit i decodes to
    const u32* dst_cap;     // [n] capacity available to unit i (maxDecompressedSize)
    int*       result;      // [n] Lizard_decompress_safe return value
    u32        n_units;
    u8*        scratch;     // n_warps * kDecScratchPerWarp bytes (Huffman-expanded streams)
    u32*       counter;     // work queue head
};

enum : u32 { kDecStreamScratch = kBlockSize + 64, kDecScratchPerWarp = 4 * kDecStreamScratch };

struct DecWarpShared {             // per-warp shared memory
    u16 table[1u << kHufTableLogMax];
    HufStatsScratch stats;
    u8  weights[256];
    u32 rank[kHufTableLogMax + 1];
    u32 pad[3];
};

// ---- warp-cooperative byte movers -------------------------------------------------------------
LZ_D void warp_copy(u8* dst, const u8* src, u32 n, u32 lane)
{
    for (u32 i = lane; i < n; i += 32) dst[i] = src[i];
}
LZ_D void warp_fill(u8* dst, u8 v, u32 n, u32 lane)
{
    for (u32 i = lane; i < n; i += 32) dst[i] = v;
}
// LZ77 match: dst[op+i] = dst[op-off+i] with byte-serial semantics.  An overlapping match is a
// periodic extension of the `off` bytes before op, so every source byte already exists.
LZ_D void warp_match(u8* dst, u32 op, u32 off, u32 len, u32 lane)
{
    const u8* s = dst + op - off;
    u8* d = dst + op;
    if (off >= len) { for (u32 i = lane; i < len; i += 32) d[i] = s[i]; }
    else if (off != 0) { for (u32 i = lane; i < len; i += 32) d[i] = s[i % off]; }
}

// ---- Huffman stream expansion -----------------------------------------------------------------
// single-symbol decode of one of the 4 segments by one lane; true when the bitstream ended exactly
LZ_D bool huf_lane_segment(u8* out, long count, const u8* src, u32 len, const u16* table, u32 tl, int* init_err)
{
    BitReader b;
    int e = bits_init(b, src, len);
    *init_err = e;
    if (e < 0) return false;
    long p = 0;
    for (;;) {
        if (bits_reload(b) != kBitsUnfinished) break;
        long k = count - p; if (k > 4) k = 4;
        if (k <= 0) break;
        for (long j = 0; j < k; ++j) out[p++] = (u8)hufx_sym(b, table, tl);
    }
    while (p < count) out[p++] = (u8)hufx_sym(b, table, tl);
    return bits_done(b);
}

// HUF_decompress for one stream; all lanes return the same value (n or negative)
LZ_D int warp_huf_decompress(u8* dst, u32 n, const u8* src, u32 c, DecWarpShared* sh, u32 lane)
{
    if (n == 0) return kErrDstSmall;
    if (c > n) return kErrCorrupt;
    if (c == n) { warp_copy(dst, src, n, lane); __syncwarp(); return (int)n; }
    if (c == 1) { warp_fill(dst, src[0], n, lane); __syncwarp(); return (int)n; }
    const u32 algo = huf_select_decoder(n, c);
    int h = 0; u32 tl = 0;
    if (lane == 0) {
        u32 nsym = 0;
        h = huf_read_stats(sh->weights, sh->rank, &nsym, &tl, src, c, &sh->stats);
        if (h >= 0) huf_fill_dtable(sh->table, sh->weights, sh->rank, nsym, tl);
    }
    h = __shfl_sync(LZB_FULL, h, 0);
    tl = __shfl_sync(LZB_FULL, tl, 0);
    if (h < 0) return h;
    if ((u32)h >= c) return kErrSrcSize;
    __syncwarp();
    const u8* pay = src + h;
    const u32 pc = c - (u32)h;
    if (pc < 10) return kErrCorrupt;
    const u32 l1 = rd_le16(pay), l2 = rd_le16(pay + 2), l3 = rd_le16(pay + 4);
    if (l1 + l2 + l3 + 6 > pc) return kErrCorrupt;
    const u32 l4 = pc - (l1 + l2 + l3 + 6);
    const long seg = (long)((n + 3) / 4);
    bool ok = true; int ierr = 0;
    if (lane < 4) {
        const u8* s = pay + 6 + (lane > 0 ? l1 : 0) + (lane > 1 ? l2 : 0) + (lane > 2 ? l3 : 0);
        u32 len = lane == 0 ? l1 : lane == 1 ? l2 : lane == 2 ? l3 : l4;
        long cnt = lane < 3 ? seg : (long)n - 3 * seg;
        if (cnt < 0) cnt = 0;
        ok = huf_lane_segment(dst + (long)lane * seg, cnt, s, len, sh->table, tl, &ierr);
    }
    // the reference initialises the four readers before decoding anything and returns the first failure
    for (int k = 0; k < 4; ++k) { int e = __shfl_sync(LZB_FULL, ierr, k); if (e < 0) return e; }
    const bool all_ok = __all_sync(LZB_FULL, ok);
    __syncwarp();
    if (all_ok) return (int)n;
    if (!algo) return kErrCorrupt;
    // the reference would have run its double-symbol decoder, which tolerates a few malformed tails
    int r = 0;
    if (lane == 0) r = huf_decode4_serial(dst, n, pay, pc, sh->table, tl, 1);
    r = __shfl_sync(LZB_FULL, r, 0);
    __syncwarp();
    return r;
}

// ---- one inner block's streams ------------------------------------------------------------------
struct Streams {
    const u8* flags;  u32 nflags;
    const u8* lits;   u32 nlits;
    const u8* off16;  u32 noff16;
    const u8* off24;  u32 noff24;
    const u8* src_end;                // end of the whole compressed unit (bound for stray reads)
};

// byte of a stream that the reference reads without an exact bound: real memory past a raw stream is
// the rest of the compressed unit; anything further reads as zero here
LZ_D u32 stray_byte(const u8* p, const Streams& s, const u8* unit_begin)
{
    // streams living in scratch never take this path past their end with a meaningful value
    return (p >= unit_begin && p < s.src_end) ? *p : 0u;
}

// length extension byte(s): b<254 -> b ; 254 -> LE16 ; 255 -> LE24  (lizard_decompress_lz4.h:50-61)
LZ_D u32 read_ext(const u8* lits, u32 nlits, u32& lp)
{
    u32 v = lits[lp];
    if (v >= 254) {
        u32 b1 = lp + 1 < nlits ? lits[lp + 1] : 0, b2 = lp + 2 < nlits ? lits[lp + 2] : 0;
        if (v == 254) { v = b1 | (b2 << 8); lp += 2; }
        else { u32 b3 = lp + 3 < nlits ? lits[lp + 3] : 0; v = b1 | (b2 << 8) | (b3 << 16); lp += 3; }
    }
    lp++;
    return v;
}

// fastLZ4 codewords (lib/lizard_decompress_lz4.h:7-163).  `op` is the offset inside the unit's output,
// `oend` the unit's capacity; matches may reach back to offset 0 of the unit.
LZ_D int decode_tokens_lz4(const Streams& s, u8* dst, u32 op0, u32 oend, u32 lane)
{
    const long nl = (long)s.nlits;
    long op = op0;
    u32 fp = 0, lp = 0;
    if (oend - op0 == 0) return (s.nflags == 1 && s.flags[0] == 0) ? 0 : -1;
    while (fp < s.nflags) {
        const u32 tok = s.flags[fp++];
        u32 len = tok & 15;
        if (len == 15) {
            if ((long)lp > nl - 5) return -(int)fp - 1;
            len = read_ext(s.lits, s.nlits, lp) + 15;
        }
        if (op + len > (long)oend - 16 || (long)lp + len > nl - 18) return -(int)fp - 1;
        warp_copy(dst + op, s.lits + lp, len, lane);
        op += len; lp += len;
        const u32 off = rd_le16(s.lits + lp); lp += 2;
        if ((long)off > op) return -(int)fp - 1;                 // match < lowLimit
        u32 ml = tok >> 4;
        if (ml == 15) {
            if ((long)lp > nl - 5) return -(int)fp - 1;
            ml = read_ext(s.lits, s.nlits, lp) + 15;
        }
        ml += kMinMatch;
        if (op + ml > (long)oend - 16) return -(int)fp - 1;
        __syncwarp();
        warp_match(dst, (u32)op, off, ml, lane);
        __syncwarp();
        op += ml;
    }
    const long rest = nl - (long)lp;
    if (rest < 0 || op + rest > (long)oend) return -(int)fp - 1;
    warp_copy(dst + op, s.lits + lp, (u32)rest, lane);
    __syncwarp();
    op += rest;
    return (int)(op - op0);
}

// LIZv1 codewords (lib/lizard_decompress_liz.h:14-220)
LZ_D int decode_tokens_lizv1(const Streams& s, u8* dst, u32 op0, u32 oend, u32 lane, const u8* unit_begin)
{
    const long nl = (long)s.nlits;
    long op = op0;
    u32 fp = 0, lp = 0, p16 = 0, p24 = 0;
    u32 last_off = 0;                                             // LIZARD_INIT_LAST_OFFSET per inner block
    if (oend - op0 == 0) return (s.nflags == 1 && s.flags[0] == 0) ? 0 : -1;
    while (fp < s.nflags) {
        const u32 tok = s.flags[fp++];
        u32 ml;
        if (tok >= 32) {
            u32 len = tok & 7;
            if (len == 7) {
                if ((long)lp > nl - 1) return -(int)fp - 1;
                len = read_ext(s.lits, s.nlits, lp) + 7;
            }
            if (op + len > (long)oend - 16 || (long)lp > nl - 16) return -(int)fp - 1;
            {   // the reference copies first and notices an over-long run later; never read past the stream
                u32 avail = (long)lp < nl ? (u32)(nl - lp) : 0;
                warp_copy(dst + op, s.lits + lp, len < avail ? len : avail, lane);
            }
            op += len; lp += len;
            if (p16 > s.noff16) return -(int)fp - 1;
            if ((tok >> 7) == 0) {                                // new 16-bit offset; bit 7 set = repeat last offset
                if (p16 + 2 <= s.noff16) last_off = rd_le16(s.off16 + p16);
                else last_off = stray_byte(s.off16 + p16, s, unit_begin) | (stray_byte(s.off16 + p16 + 1, s, unit_begin) << 8);
                p16 += 2;
            }
            ml = (tok >> 3) & 15;
            if (ml == 15) {
                if ((long)lp > nl - 1) return -(int)fp - 1;
                ml = read_ext(s.lits, s.nlits, lp) + 15;
            }
        } else if (tok < kLastLongOff) {
            if ((long)p24 > (long)s.noff24 - 3) return -(int)fp - 1;
            ml = tok + kMmLongOff;
            last_off = rd_le24(s.off24 + p24); p24 += 3;
        } else {
            if ((long)lp > nl - 1) return -(int)fp - 1;
            ml = read_ext(s.lits, s.nlits, lp) + kLastLongOff + kMmLongOff;
            if ((long)p24 > (long)s.noff24 - 3) return -(int)fp - 1;
            last_off = rd_le24(s.off24 + p24); p24 += 3;
        }
        if ((long)last_off > op) return -(int)fp - 1;             // match < lowLimit
        if (op + ml > (long)oend - 16) return -(int)fp - 1;
        __syncwarp();
        warp_match(dst, (u32)op, last_off, ml, lane);
        __syncwarp();
        op += ml;
    }
    const long rest = nl - (long)lp;
    if (rest < 0 || op + rest > (long)oend) return -(int)fp - 1;
    warp_copy(dst + op, s.lits + lp, (u32)rest, lane);
    __syncwarp();
    op += rest;
    return (int)(op - op0);
}

// One stream header.  Returns 1 on success, 0 on failure (Lizard_readStream, lizard_decompress.c:72-112).
// `ip` is an offset into the unit.
LZ_D int read_stream(bool huff, const u8* src, long csize, long& ip, u8* scratch, const u8** ptr, u32* len,
                     DecWarpShared* sh, u32 lane)
{
    if (!huff) {
        if (ip > csize - 3) return 0;
        *ptr = src + ip + 3;
        *len = rd_le24(src + ip);
        ip += 3 + (long)*len;
        return 1;
    }
    if (ip > csize - 6) return 0;
    const u32 n = rd_le24(src + ip), c = rd_le24(src + ip + 3);
    if (n > kBlockSize || ip + (long)c > csize - 6) return 0;
    const int r = warp_huf_decompress(scratch, n, src + ip + 6, c, sh, lane);
    if (r < 0 || (u32)r != n) return 0;
    ip += (long)c + 6;
    *ptr = scratch; *len = n;
    return 1;
}

// Lizard_decompress_safe for one unit; every lane returns the same value.
LZ_D int decode_unit(const u8* src, u32 csize_u, u8* dst, u32 cap, u8* scratch, DecWarpShared* sh, u32 lane)
{
    const long csize = (long)csize_u;
    if (csize < 1) return 0;
    const int level = src[0];
    if (level < (int)kMinLevel || level > (int)kMaxLevel) return -1;
    const bool lizv1 = level_is_lizv1(level);
    long ip = 1;
    long op = 0;
    while (ip < csize) {
        const u32 hdr = src[ip++];
        if (hdr == kFlagRaw) {
            if (ip > csize - 3) return -1;
            const u32 len = rd_le24(src + ip); ip += 3;
            if (ip + (long)len > csize || op + (long)len > (long)cap) return -1;
            warp_copy(dst + op, src + ip, len, lane);
            __syncwarp();
            op += len; ip += len;
            continue;
        }
        if (hdr & kFlagLen) return -1;
        if (ip > csize - 15) return -1;
        {   // lengths stream: always raw, always empty from this encoder, but honour its size field
            const long len_len = (long)rd_le24(src + ip);
            const long len_end = ip + 3 + len_len;
            if (len_end > csize - 3) return -1;
            ip = len_end;
        }
        Streams s;
        s.src_end = src + csize;
        if (!read_stream(hdr & kFlagOff16, src, csize, ip, scratch + 3 * kDecStreamScratch, &s.off16, &s.noff16, sh, lane)) return -1;
        if (!read_stream(hdr & kFlagOff24, src, csize, ip, scratch + 2 * kDecStreamScratch, &s.off24, &s.noff24, sh, lane)) return -1;
        if (!read_stream(hdr & kFlagFlags, src, csize, ip, scratch + 1 * kDecStreamScratch, &s.flags, &s.nflags, sh, lane)) return -1;
        if (!read_stream(hdr & kFlagLiterals, src, csize, ip, scratch, &s.lits, &s.nlits, sh, lane)) return -1;
        if (ip > csize) return -1;
        const int res = lizv1 ? decode_tokens_lizv1(s, dst, (u32)op, cap, lane, src)
                              : decode_tokens_lz4(s, dst, (u32)op, cap, lane);
        if (res <= 0) return res;
        op += res;
    }
    return (int)op;
}

}  // namespace lzb
