// entropy_dec.cuh -- serial (single-lane) pieces of the Huff0 decoder: backward bit reader,
// FSE-coded weight header, weight statistics, single-symbol decode table.
//
// These run on ONE lane per Huffman stream (they are tiny: <= 255 weights) and are compiled for
// both host and device so the CPU test-suite can check them against the reference library.
// Semantics follow (accept/reject included):
//   lib/entropy/bitstream.h:260-408      BIT_DStream_t: init / look / reload / endOfDStream
//   lib/entropy/entropy_common.c:71-160  FSE_readNCount
//   lib/entropy/fse_decompress.c:113-168 FSE_buildDTable, :220-294 FSE_decompress_usingDTable
//   lib/entropy/entropy_common.c:170-231 HUF_readStats
//   lib/entropy/huf_decompress.c:87-133  HUF_readDTableX2
#pragma once
#include "common.cuh"

namespace lzb {

enum : int { kErrGeneric = -1, kErrCorrupt = -20, kErrSrcSize = -72, kErrDstSmall = -70, kErrTableLog = -44,
             kErrMaxSymbol = -48 };

// ------------------------------------------------------------------------------------------
// Backward bit reader.  The encoder wrote LSB-first and closed with a single 1 bit; the decoder
// starts from the last byte, skips the padding above the end mark and reads toward the start.
// `used` counts consumed bits of the 64-bit window `win` (loaded little-endian at `ptr`).
// ------------------------------------------------------------------------------------------
struct BitReader {
    const u8* start;
    const u8* ptr;
    u64 win;
    u32 used;
};
enum : int { kBitsUnfinished = 0, kBitsEndOfBuffer = 1, kBitsCompleted = 2, kBitsOverflow = 3 };

LZ_HD int bits_init(BitReader& b, const u8* src, u32 len)
{
    b.start = src; b.ptr = src; b.win = 0; b.used = 0;
    if (len < 1) return kErrSrcSize;
    u8 last = src[len - 1];
    if (len >= 8) {
        b.ptr = src + len - 8;
        b.win = rd_le64(b.ptr);
        b.used = last ? 8 - highbit32(last) : 0;
    } else {
        u64 w = 0;
        for (u32 i = 0; i < len; ++i) w |= (u64)src[i] << (8 * i);
        b.win = w;
        b.used = (last ? 8 - highbit32(last) : 0) + (8 - len) * 8;
    }
    if (last == 0) return kErrGeneric;   // end mark missing
    return 0;
}

// up to 56 bits; n may be 0
LZ_HD u64 bits_look(const BitReader& b, u32 n) { return ((b.win << (b.used & 63)) >> 1) >> ((63 - n) & 63); }
LZ_HD u64 bits_read(BitReader& b, u32 n) { u64 v = bits_look(b, n); b.used += n; return v; }

LZ_HD int bits_reload(BitReader& b)
{
    if (b.used > 64) return kBitsOverflow;
    if (b.ptr >= b.start + 8) {
        b.ptr -= b.used >> 3;
        b.used &= 7;
        b.win = rd_le64(b.ptr);
        return kBitsUnfinished;
    }
    if (b.ptr == b.start) return b.used < 64 ? kBitsEndOfBuffer : kBitsCompleted;
    u32 nb = b.used >> 3;
    int st = kBitsUnfinished;
    if (b.ptr - nb < b.start) { nb = (u32)(b.ptr - b.start); st = kBitsEndOfBuffer; }
    b.ptr -= nb;
    b.used -= nb * 8;
    b.win = rd_le64(b.ptr);
    return st;
}
LZ_HD bool bits_done(const BitReader& b) { return b.ptr == b.start && b.used == 64; }

// ------------------------------------------------------------------------------------------
// FSE normalized-count header.  Returns bytes consumed (>0) or a negative error.
// norm[] needs 256 entries; *max_sv is in/out (in: alphabet limit, out: last symbol present).
// ------------------------------------------------------------------------------------------
LZ_HD int fse_read_ncount(short* norm, u32* max_sv, u32* table_log, const u8* src, u32 size)
{
    if (size < 4) return kErrSrcSize;
    const u8* const iend = src + size;
    const u8* ip = src;
    u32 stream = rd_le32(ip);
    int nb = (int)(stream & 15) + (int)kFseMinTableLog;
    if (nb > (int)kFseAbsMaxTableLog) return kErrTableLog;
    stream >>= 4;
    int bit_count = 4;
    *table_log = (u32)nb;
    int remaining = (1 << nb) + 1;
    int threshold = 1 << nb;
    nb++;
    u32 sym = 0;
    bool prev_zero = false;

    while (remaining > 1 && sym <= *max_sv) {
        if (prev_zero) {
            u32 n0 = sym;
            while ((stream & 0xFFFF) == 0xFFFF) {
                n0 += 24;
                if (ip < iend - 5) { ip += 2; stream = rd_le32(ip) >> bit_count; }
                else               { stream >>= 16; bit_count += 16; }
            }
            while ((stream & 3) == 3) { n0 += 3; stream >>= 2; bit_count += 2; }
            n0 += stream & 3;
            bit_count += 2;
            if (n0 > *max_sv) return kErrMaxSymbol;
            while (sym < n0) norm[sym++] = 0;
            if (ip <= iend - 7 || ip + (bit_count >> 3) <= iend - 4) {
                ip += bit_count >> 3;
                bit_count &= 7;
                stream = rd_le32(ip) >> bit_count;
            } else {
                stream >>= 2;
            }
        }
        {
            const int max = (2 * threshold - 1) - remaining;
            int count;
            if ((stream & (u32)(threshold - 1)) < (u32)max) {
                count = (int)(stream & (u32)(threshold - 1));
                bit_count += nb - 1;
            } else {
                count = (int)(stream & (u32)(2 * threshold - 1));
                if (count >= threshold) count -= max;
                bit_count += nb;
            }
            count--;                                   // stored with +1 so that -1 ("less than one") fits
            remaining -= count < 0 ? -count : count;
            norm[sym++] = (short)count;
            prev_zero = (count == 0);
            while (remaining < threshold) { nb--; threshold >>= 1; }

            if (ip <= iend - 7 || ip + (bit_count >> 3) <= iend - 4) {
                ip += bit_count >> 3;
                bit_count &= 7;
            } else {
                bit_count -= (int)(8 * (iend - 4 - ip));
                ip = iend - 4;
            }
            stream = rd_le32(ip) >> (bit_count & 31);
        }
    }
    if (remaining != 1) return kErrCorrupt;
    if (bit_count > 32) return kErrCorrupt;
    *max_sv = sym - 1;
    ip += (bit_count + 7) >> 3;
    return (int)(ip - src);
}

struct FseCell { u16 next; u8 sym; u8 nbits; };

// Spread symbols over the state table and derive (nbits, next-state base) per cell.
// cells[] needs 1 << table_log entries, scratch_next[] 256 entries.  Returns 0 or negative.
LZ_HD int fse_build_dtable(FseCell* cells, u16* scratch_next, const short* norm, u32 max_sv, u32 table_log)
{
    if (max_sv > 255) return kErrMaxSymbol;
    if (table_log > kFseMaxTableLog) return kErrTableLog;
    const u32 size = 1u << table_log;
    u32 high = size - 1;
    for (u32 s = 0; s <= max_sv; ++s) {
        if (norm[s] == -1) { cells[high--].sym = (u8)s; scratch_next[s] = 1; }
        else scratch_next[s] = (u16)norm[s];
    }
    const u32 mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 pos = 0;
    for (u32 s = 0; s <= max_sv; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            cells[pos].sym = (u8)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    if (pos != 0) return kErrGeneric;
    for (u32 u = 0; u < size; ++u) {
        u32 nx = scratch_next[cells[u].sym]++;
        u32 nbits = table_log - highbit32(nx);
        cells[u].nbits = (u8)nbits;
        cells[u].next = (u16)((nx << nbits) - size);
    }
    return 0;
}

// Two interleaved tANS states reading one backward stream (FSE_decompress_usingDTable_generic).
// Returns number of symbols written or negative.
LZ_HD int fse_decode_stream(u8* dst, u32 cap, const u8* src, u32 size, const FseCell* cells, u32 table_log)
{
    BitReader b;
    int e = bits_init(b, src, size);
    if (e < 0) return e;
    u32 s1 = (u32)bits_read(b, table_log); bits_reload(b);
    u32 s2 = (u32)bits_read(b, table_log); bits_reload(b);
    u8* op = dst;
    u8* const omax = dst + cap;
    u8* const olimit = omax - 3;
#define LZB_FSE_STEP(S) { FseCell c = cells[S]; *op++ = c.sym; S = (u32)c.next + (u32)bits_read(b, c.nbits); }
    while (bits_reload(b) == kBitsUnfinished && op < olimit) {
        LZB_FSE_STEP(s1) LZB_FSE_STEP(s2) LZB_FSE_STEP(s1) LZB_FSE_STEP(s2)
    }
    for (;;) {
        if (op > omax - 2) return kErrDstSmall;
        LZB_FSE_STEP(s1)
        if (bits_reload(b) == kBitsOverflow) { LZB_FSE_STEP(s2) break; }
        if (op > omax - 2) return kErrDstSmall;
        LZB_FSE_STEP(s2)
        if (bits_reload(b) == kBitsOverflow) { LZB_FSE_STEP(s1) break; }
    }
#undef LZB_FSE_STEP
    return (int)(op - dst);
}

// ------------------------------------------------------------------------------------------
// Huffman weight header -> weights[0..nsym), rank_count[0..12], table_log.
// Returns header bytes consumed (>0) or negative.  weights[] needs 256 entries.
// ------------------------------------------------------------------------------------------
struct HufStatsScratch {          // explicit so a device lane can place it where it wants
    short   norm[256];
    u16     next[256];
    FseCell cells[1u << kHufHeaderFseLog];
};

LZ_HD_COLD int huf_read_stats(u8* weights, u32* rank_count, u32* nsym, u32* table_log,
                         const u8* src, u32 size, HufStatsScratch* ws)
{
    if (size == 0) return kErrSrcSize;
    u32 isize = src[0];
    u32 osize;
    if (isize >= 128) {                      // raw 4-bit weights
        osize = isize - 127;
        isize = (osize + 1) / 2;
        if (isize + 1 > size) return kErrSrcSize;
        if (osize >= 256) return kErrCorrupt;
        for (u32 n = 0; n < osize; n += 2) {
            weights[n] = src[1 + n / 2] >> 4;
            if (n + 1 < 256) weights[n + 1] = src[1 + n / 2] & 15;
        }
    } else {                                 // FSE-coded weights (tableLog <= 6)
        if (isize + 1 > size) return kErrSrcSize;
        u32 max_sv = 255, fse_log = 0;
        int h = fse_read_ncount(ws->norm, &max_sv, &fse_log, src + 1, isize);
        if (h < 0) return h;
        if (fse_log > kHufHeaderFseLog) return kErrTableLog;
        int e = fse_build_dtable(ws->cells, ws->next, ws->norm, max_sv, fse_log);
        if (e < 0) return e;
        int n = fse_decode_stream(weights, 255, src + 1 + h, isize - (u32)h, ws->cells, fse_log);
        if (n < 0) return n;
        osize = (u32)n;
    }

    for (u32 r = 0; r <= kHufTableLogMax; ++r) rank_count[r] = 0;
    u32 total = 0;
    for (u32 n = 0; n < osize; ++n) {
        if (weights[n] >= kHufTableLogMax) return kErrCorrupt;
        rank_count[weights[n]]++;
        total += (1u << weights[n]) >> 1;
    }
    if (total == 0) return kErrCorrupt;
    u32 tl = highbit32(total) + 1;
    if (tl > kHufTableLogMax) return kErrCorrupt;
    *table_log = tl;
    {   // the last symbol's weight is implied: the sum must complete to a power of two
        u32 rest = (1u << tl) - total;
        u32 hb = highbit32(rest);
        if ((1u << hb) != rest) return kErrCorrupt;
        weights[osize] = (u8)(hb + 1);
        rank_count[hb + 1]++;
    }
    if (rank_count[1] < 2 || (rank_count[1] & 1)) return kErrCorrupt;
    *nsym = osize + 1;
    return (int)(isize + 1);
}

// Single-symbol decode table: entry = symbol | nbBits << 8, 1 << table_log entries.
// rank_count[] is consumed (turned into running start positions).
LZ_HD_COLD void huf_fill_dtable(u16* table, const u8* weights, u32* rank_count, u32 nsym, u32 table_log)
{
    u32 start = 0;
    for (u32 w = 1; w <= table_log; ++w) { u32 cur = start; start += rank_count[w] << (w - 1); rank_count[w] = cur; }
    for (u32 s = 0; s < nsym; ++s) {
        u32 w = weights[s];
        u32 len = (1u << w) >> 1;
        u16 e = (u16)(s | ((table_log + 1 - w) << 8));
        u32 at = rank_count[w];
        for (u32 i = 0; i < len; ++i) table[at + i] = e;
        rank_count[w] = at + len;
    }
}

// ---- decode tables -------------------------------------------------------------------------------------------------
// A table type answers look(hi): `hi` = the next 32 bits of the stream, left aligned; result = symbol | nbBits << 8
// (what HUF_DEltX2 holds, huf_decompress.c:85).
struct HufFull {                 // the reference's layout: 1 << tableLog entries (HUF_readDTableX2, huf_decompress.c:87-133)
    const u16* t; u32 down;      // down = 32 - tableLog
    LZ_HDM u32 look(u32 hi) const { return t[hi >> down]; }
    LZ_HDM void look2(u32 hi, u32* sym, u32* nbits) const { const u32 e = t[hi >> down]; *sym = e & 255u; *nbits = e >> 8; }
};
// Compact form of the same table for the pre-pass (3 KiB instead of 4 KiB at tableLog 11, so that the 56 tables an SM keeps
// in shared memory leave it an L1): one byte per entry for the symbol, one NIBBLE per entry for the code length.  Both are
// indexed by the same table index, so the two loads of a lookup are independent of each other -- a decoder's dependency chain
// runs through the length only (window -> index -> length -> next window), one shared-memory latency per symbol -- and the
// lookup is the same for every code length: the lanes of a warp decode 32 different streams, and any scheme with a separate
// path for the long codes runs that path for all lanes almost every time (round 1's two-level table spent 37 % of the expand
// kernel's instructions there; a symbol table with a per-SYMBOL length side table made the length load wait for the symbol
// load: profiles/r02_SUMMARY.md).
struct alignas(16) HufCompact {
    u8  sym[1u << 11];           // tableLog <= 11 (tableLog 12 streams are left to the in-kernel path)
    u8  len[1u << 10];           // nbBits of entry i in nibble i & 1 of byte i >> 1
    u32 tl, pad[3];
};
struct HufCompactView {          // what a segment decoder keeps in registers
    const HufCompact* t; u32 down;      // 32 - tl
#if defined(__CUDA_ARCH__)
    // The tables live in shared memory (prepass.cuh).  Their two base addresses are kept as opaque 32-bit shared-space
    // registers: left to itself the compiler re-derives `warp's block + job * sizeof(table)` with a multiply-add in front
    // of every one of the eight loads of a reload, two of them on the length chain.
    u32 sym_at, len_at;
    LZ_HDM u32 look(u32 hi) const
    {
        const u32 idx = hi >> down;
        u32 s, b;
        asm("ld.shared.u8 %0, [%1];" : "=r"(s) : "r"(sym_at + idx));
        asm("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(len_at + (idx >> 1)));
        const u32 n = (b >> ((idx & 1u) * 4u)) & 15u;
        return s | (n << 8);
    }
    LZ_HDM void look2(u32 hi, u32* sym, u32* nbits) const      // symbol and length apart: the length is the decoder's chain
    {
        const u32 idx = hi >> down;
        u32 s, b;
        asm("ld.shared.u8 %0, [%1];" : "=r"(s) : "r"(sym_at + idx));
        asm("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(len_at + (idx >> 1)));
        *sym = s; *nbits = (b >> ((idx & 1u) * 4u)) & 15u;
    }
#else
    LZ_HDM void look2(u32 hi, u32* sym, u32* nbits) const { const u32 e = look(hi); *sym = e & 255u; *nbits = e >> 8; }
    LZ_HDM u32 look(u32 hi) const
    {
        const u32 idx = hi >> down;
        const u32 s = t->sym[idx];
        const u32 n = ((u32)t->len[idx >> 1] >> ((idx & 1u) * 4u)) & 15u;
        return s | (n << 8);
    }
#endif
};
LZ_HD HufCompactView huf_view(const HufCompact* t)
{
    HufCompactView v; v.t = t; v.down = 32 - t->tl;
#if defined(__CUDA_ARCH__)
    v.sym_at = (u32)__cvta_generic_to_shared(t->sym);
    v.len_at = (u32)__cvta_generic_to_shared(t->len);
    asm volatile("" : "+r"(v.sym_at), "+r"(v.len_at));
#endif
    return v;
}

// rank_count[] is consumed, like huf_fill_dtable does.  table_log <= 11.
LZ_HD_COLD void huf_fill_compact(HufCompact* t, const u8* weights, u32* rank_count, u32 nsym, u32 table_log)
{
    t->tl = table_log;
    u32 start = 0;
    for (u32 w = 1; w <= table_log; ++w) { const u32 cur = start; start += rank_count[w] << (w - 1); rank_count[w] = cur; }
    for (u32 s = 0; s < nsym; ++s) {
        const u32 w = weights[s];
        if (w == 0) continue;
        const u32 n = (1u << w) >> 1, at = rank_count[w], nb = table_log + 1 - w;
        for (u32 i = 0; i < n; ++i) {
            const u32 e = at + i;
            t->sym[e] = (u8)s;
            if (e & 1u) t->len[e >> 1] = (u8)((t->len[e >> 1] & 0x0fu) | (nb << 4));
            else        t->len[e >> 1] = (u8)((t->len[e >> 1] & 0xf0u) | nb);
        }
        rank_count[w] = at + n;
    }
}

// HUF_selectDecoder (lib/entropy/huf_decompress.c:771-812): 0 = single-symbol, 1 = double-symbol.
// Both give identical bytes on valid input; they differ only in which corrupt inputs they reject.
LZ_HD u32 huf_select_decoder(u32 dst_size, u32 src_size)
{
    // {tableTime, decode256Time} for single and double symbol decoders, by compression-ratio bucket
    const u16 t0[16][2] = {{0,0},{0,0},{38,130},{448,128},{556,128},{714,128},{883,128},{897,128},
                           {926,128},{947,128},{1107,128},{1177,128},{1242,128},{1349,128},{1455,128},{722,128}};
    const u16 t1[16][2] = {{1,1},{1,1},{1313,74},{1353,74},{1353,74},{1418,74},{1437,74},{1515,75},
                           {1613,75},{1729,77},{2083,81},{2379,87},{2415,93},{2644,106},{2422,124},{1891,145}};
    u32 q = (u32)(((u64)src_size * 16) / dst_size);
    u32 d256 = dst_size >> 8;
    u32 a = t0[q][0] + t0[q][1] * d256;
    u32 b = t1[q][0] + t1[q][1] * d256;
    b += b >> 3;
    return b < a;
}

}  // namespace lzb

// ==========================================================================================
// Single-thread, exactly-as-the-reference Huff0 block decoder.  Used (a) by the CPU tests to pin
// the helpers above against the reference library, (b) on the device only as the slow path that
// decides accept/reject for streams the fast single-symbol path rejected while the reference would
// have used its double-symbol decoder (lib/entropy/huf_decompress.c:562-585 accepts a few corrupt
// tails that the single-symbol decoder refuses; see DESIGN.md "Huffman accept/reject parity").
// ==========================================================================================
namespace lzb {

LZ_HD u32 hufx_sym(BitReader& b, const u16* t, u32 tl)          // single-symbol step
{
    u32 idx = (u32)((b.win << (b.used & 63)) >> ((64 - tl) & 63));
    u16 e = t[idx];
    b.used += e >> 8;
    return e & 255;
}

// one step of the double-symbol decoder, emulated on the single-symbol table: the reference's
// 12-bit table holds a pair whenever both code lengths fit in 12 bits (huf_decompress.c:399-436)
LZ_HD u32 hufx4_pair(const BitReader& b, const u16* t, u32 tl, u32* s1, u32* s2, u32* n1, u32* n2)
{
    u32 v = (u32)((b.win << (b.used & 63)) >> 52);
    u16 e1 = t[v >> (12 - tl)];
    *s1 = e1 & 255; *n1 = e1 >> 8;
    u32 rest = (v << *n1) & 0xFFF;
    u16 e2 = t[rest >> (12 - tl)];
    *s2 = e2 & 255; *n2 = e2 >> 8;
    return (*n1 + *n2 <= 12) ? 2u : 1u;
}
LZ_HD u32 hufx4_step(BitReader& b, const u16* t, u32 tl, u8* p)
{
    u32 s1, s2, n1, n2;
    u32 len = hufx4_pair(b, t, tl, &s1, &s2, &n1, &n2);
    p[0] = (u8)s1;
    p[1] = (len == 2) ? (u8)s2 : 0;
    b.used += (len == 2) ? n1 + n2 : n1;
    return len;
}
LZ_HD void hufx4_last(BitReader& b, const u16* t, u32 tl, u8* p)
{
    u32 s1, s2, n1, n2;
    u32 len = hufx4_pair(b, t, tl, &s1, &s2, &n1, &n2);
    p[0] = (u8)s1;
    if (len == 1) b.used += n1;
    else if (b.used < 64) { b.used += n1 + n2; if (b.used > 64) b.used = 64; }
}

// finish one segment with the double-symbol loop structure (huf_decompress.c:562-585)
LZ_HD void hufx4_finish(u8* base, long p, long pend, BitReader& b, const u16* t, u32 tl)
{
    while (bits_reload(b) == kBitsUnfinished && p < pend - 7) {
        p += hufx4_step(b, t, tl, base + p); p += hufx4_step(b, t, tl, base + p);
        p += hufx4_step(b, t, tl, base + p); p += hufx4_step(b, t, tl, base + p);
    }
    while (bits_reload(b) == kBitsUnfinished && p <= pend - 2) p += hufx4_step(b, t, tl, base + p);
    while (p <= pend - 2) p += hufx4_step(b, t, tl, base + p);
    if (p < pend) hufx4_last(b, t, tl, base + p);
}
// finish one segment with the single-symbol loop structure (huf_decompress.c:155-176)
LZ_HD void hufx2_finish(u8* base, long p, long pend, BitReader& b, const u16* t, u32 tl)
{
    while (bits_reload(b) == kBitsUnfinished && p <= pend - 4) {
        base[p++] = (u8)hufx_sym(b, t, tl); base[p++] = (u8)hufx_sym(b, t, tl);
        base[p++] = (u8)hufx_sym(b, t, tl); base[p++] = (u8)hufx_sym(b, t, tl);
    }
    while (bits_reload(b) == kBitsUnfinished && p < pend) base[p++] = (u8)hufx_sym(b, t, tl);
    while (p < pend) base[p++] = (u8)hufx_sym(b, t, tl);
}

// 4-segment payload after the weight header.  dst must have 8 bytes of slack past n (the
// reference writes up to 2 bytes past tiny outputs and the pair decoder stores 2 bytes at a time).
LZ_HD_COLD int huf_decode4_serial(u8* dst, u32 n, const u8* src, u32 c, const u16* t, u32 tl, u32 algo)
{
    if (c < 10) return kErrCorrupt;
    u32 l1 = rd_le16(src), l2 = rd_le16(src + 2), l3 = rd_le16(src + 4);
    if (l1 + l2 + l3 + 6 > c) return kErrCorrupt;
    u32 l4 = c - (l1 + l2 + l3 + 6);
    const u8* i1 = src + 6; const u8* i2 = i1 + l1; const u8* i3 = i2 + l2; const u8* i4 = i3 + l3;
    long seg = (long)((n + 3) / 4);
    long e1 = seg, e2 = 2 * seg, e3 = 3 * seg, e4 = (long)n;
    long p1 = 0, p2 = e1, p3 = e2, p4 = e3;
    BitReader b1, b2, b3, b4;
    int e;
    if ((e = bits_init(b1, i1, l1)) < 0) return e;
    if ((e = bits_init(b2, i2, l2)) < 0) return e;
    if ((e = bits_init(b3, i3, l3)) < 0) return e;
    if ((e = bits_init(b4, i4, l4)) < 0) return e;
    int sig = bits_reload(b1) | bits_reload(b2) | bits_reload(b3) | bits_reload(b4);
    while (sig == kBitsUnfinished && p4 < e4 - 7) {
        for (int r = 0; r < 4; ++r) {
            if (algo) {
                p1 += hufx4_step(b1, t, tl, dst + p1); p2 += hufx4_step(b2, t, tl, dst + p2);
                p3 += hufx4_step(b3, t, tl, dst + p3); p4 += hufx4_step(b4, t, tl, dst + p4);
            } else {
                dst[p1++] = (u8)hufx_sym(b1, t, tl); dst[p2++] = (u8)hufx_sym(b2, t, tl);
                dst[p3++] = (u8)hufx_sym(b3, t, tl); dst[p4++] = (u8)hufx_sym(b4, t, tl);
            }
        }
        sig = bits_reload(b1) | bits_reload(b2) | bits_reload(b3) | bits_reload(b4);
    }
    if (p1 > e1 || p2 > e2 || p3 > e3) return kErrCorrupt;
    if (algo) {
        hufx4_finish(dst, p1, e1, b1, t, tl); hufx4_finish(dst, p2, e2, b2, t, tl);
        hufx4_finish(dst, p3, e3, b3, t, tl); hufx4_finish(dst, p4, e4, b4, t, tl);
    } else {
        hufx2_finish(dst, p1, e1, b1, t, tl); hufx2_finish(dst, p2, e2, b2, t, tl);
        hufx2_finish(dst, p3, e3, b3, t, tl); hufx2_finish(dst, p4, e4, b4, t, tl);
    }
    if (!(bits_done(b1) && bits_done(b2) && bits_done(b3) && bits_done(b4))) return kErrCorrupt;
    return (int)n;
}

struct HufDecScratch {
    HufStatsScratch stats;
    u8  weights[256];
    u32 rank[kHufTableLogMax + 1];
    u16 table[1u << kHufTableLogMax];
};

// HUF_decompress (huf_decompress.c:817-845): returns n on success, negative on error.
LZ_HD int huf_decompress_serial(u8* dst, u32 n, const u8* src, u32 c, HufDecScratch* ws)
{
    if (n == 0) return kErrDstSmall;
    if (c > n) return kErrCorrupt;
    if (c == n) { for (u32 i = 0; i < n; ++i) dst[i] = src[i]; return (int)n; }
    if (c == 1) { for (u32 i = 0; i < n; ++i) dst[i] = src[0]; return (int)n; }
    u32 algo = huf_select_decoder(n, c);
    u32 nsym = 0, tl = 0;
    int h = huf_read_stats(ws->weights, ws->rank, &nsym, &tl, src, c, &ws->stats);
    if (h < 0) return h;
    if ((u32)h >= c) return kErrSrcSize;
    huf_fill_dtable(ws->table, ws->weights, ws->rank, nsym, tl);
    return huf_decode4_serial(dst, n, src + h, c - (u32)h, ws->table, tl, algo);
}

}  // namespace lzb
