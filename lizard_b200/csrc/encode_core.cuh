// encode_core.cuh -- the Lizard block encoder written once for "a warp": every lane runs the same
// control flow on the same (uniform) cursor state; byte movement, histograms and bit-packing are split
// across lanes through the policy type W.  With W = HostLanes (one lane) the very same code builds with
// g++, which is how the CPU test-suite pins it byte-for-byte against the reference (-DLIZARD_RESET_MEM).
//
// Reference functions restated here:
//   lib/lizard_compress.c:75-109         Lizard_hash5 / Lizard_hashPtr
//   lib/lizard_common.h:475-490          Lizard_count
//   lib/lizard_parser_fastsmall.h:34-189 Lizard_compress_fastSmall   (levels 10, 30)
//   lib/lizard_parser_fast.h:41-196      Lizard_compress_fast        (levels 11, 31; same loop, hashLog 18)
//   lib/lizard_parser_pricefast.h:3-249  Lizard_FindMatchFast / _Faster / Lizard_compress_priceFast (21,22,41,42)
//   lib/lizard_compress_lz4.h:3-86       Lizard_encodeSequence_LZ4 / LastLiterals
//   lib/lizard_compress_liz.h:43-179     Lizard_encodeSequence_LIZv1 / LastLiterals
//   lib/lizard_compress.c:141-250        Lizard_writeStream / Lizard_writeBlock
//   lib/lizard_compress.c:472-547        Lizard_compress_generic
#pragma once
#include "common.cuh"
#include "entropy_enc.cuh"
#include "lanes.cuh"
#if !defined(__CUDACC__)
#include <string.h>
#endif

namespace lzb {

// ---- unaligned little-endian loads from the source block ----------------------------------------
LZ_HD u32 ld32(const u8* p)
{
#if defined(__CUDA_ARCH__)
    const size_t a = (size_t)p;
    const u32* q = (const u32*)(a & ~(size_t)3);
    const u32 sh = (u32)(a & 3) * 8;
    const u32 lo = q[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, q[1], sh);
#else
    u32 v; memcpy(&v, p, 4); return v;
#endif
}
LZ_HD u64 ld64(const u8* p)
{
#if defined(__CUDA_ARCH__)
    const size_t a = (size_t)p;
    const u64* q = (const u64*)(a & ~(size_t)7);
    const u32 sh = (u32)(a & 7) * 8;
    const u64 lo = q[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (q[1] << (64 - sh));
#else
    u64 v; memcpy(&v, p, 8); return v;
#endif
}

LZ_HD u32 hash5(u64 v, u32 hbits) { return (u32)(((v * 889523592379ULL) << 24) >> (64 - hbits)); }

// common prefix length of a[..] and b[..], a bounded by `limit` (Lizard_count)
LZ_HD u32 count_match(const u8* a, const u8* b, const u8* limit)
{
    const u8* const a0 = a;
    while (a + 7 < limit) {
        u64 d = ld64(a) ^ ld64(b);
        if (d) {
#if defined(__CUDA_ARCH__)
            return (u32)(a - a0) + ((u32)(__ffsll((long long)d) - 1) >> 3);
#else
            return (u32)(a - a0) + ((u32)__builtin_ctzll(d) >> 3);
#endif
        }
        a += 8; b += 8;
    }
    while (a < limit && *a == *b) { a++; b++; }
    return (u32)(a - a0);
}

// ---- the five output streams of one inner block -------------------------------------------------------
struct EncStreams {
    u8* lits;  u8* flags;  u8* off16;  u8* off24;
    u32 nl, nf, n16, n24;
};

// length extension: b<254 | 254,LE16 | 255,LE24
LZ_HD void put_ext(u8* lits, u32& nl, u32 v, bool writer)
{
    if (v >= (1u << 16)) { if (writer) { lits[nl] = 255; wr_le24(lits + nl + 1, v); } nl += 4; }
    else if (v >= 254)   { if (writer) { lits[nl] = 254; wr_le16(lits + nl + 1, v); } nl += 3; }
    else                 { if (writer) lits[nl] = (u8)v; nl += 1; }
}

// Lizard_encodeSequence_LZ4: literals src[anchor..ip) then a match of `ml` bytes at distance `off`
template <class W> LZ_HD void emit_lz4(EncStreams& s, const u8* src, u32 anchor, u32 ip, u32 ml, u32 off)
{
    const bool wr = W::lane() == 0;
    const u32 lit = ip - anchor;
    u32 tok;
    if (lit >= 15) { tok = 15; put_ext(s.lits, s.nl, lit - 15, wr); } else tok = lit;
    lanes_copy<W>(s.lits + s.nl, src + anchor, lit);
    s.nl += lit;
    if (wr) wr_le16(s.lits + s.nl, off);
    s.nl += 2;
    const u32 m = ml - kMinMatch;
    if (m >= 15) { tok += 15u << 4; put_ext(s.lits, s.nl, m - 15, wr); } else tok += m << 4;
    if (wr) s.flags[s.nf] = (u8)tok;
    s.nf++;
}

// Lizard_encodeSequence_LIZv1.  off == 0 means "repeat last offset".  Updates last_off.
template <class W> LZ_HD void emit_lizv1(EncStreams& s, const u8* src, u32 anchor, u32 ip, u32 ml, u32 off, u32& last_off)
{
    const bool wr = W::lane() == 0;
    const u32 lit = ip - anchor;
    u32 tok = 0;
    if (lit > 0 || off < kMax16BitOffset) {
        if (lit >= 7) { tok = 7; put_ext(s.lits, s.nl, lit - 7, wr); } else tok = lit;
        lanes_copy<W>(s.lits + s.nl, src + anchor, lit);
        s.nl += lit;
        if (off >= kMax16BitOffset) {          // literals before a 24-bit-offset match ride on a zero-length repeat token
            tok += 1u << 7;
            if (wr) s.flags[s.nf] = (u8)tok;
            s.nf++;
            tok = 0;
        }
    }
    if (off >= kMax16BitOffset) {
        if (ml - kMmLongOff >= kLastLongOff) { tok = kLastLongOff; put_ext(s.lits, s.nl, ml - kMmLongOff - kLastLongOff, wr); }
        else tok = ml - kMmLongOff;
        if (wr) wr_le24(s.off24 + s.n24, off);
        s.n24 += 3;
        last_off = off;
    } else {
        if (off == 0) tok += 1u << 7;
        else { last_off = off; if (wr) wr_le16(s.off16 + s.n16, off); s.n16 += 2; }
        if (ml >= 15) { tok += 15u << 3; put_ext(s.lits, s.nl, ml - 15, wr); } else tok += ml << 3;
    }
    if (wr) s.flags[s.nf] = (u8)tok;
    s.nf++;
}

template <class W> LZ_HD void emit_last_literals(EncStreams& s, const u8* src, u32 anchor, u32 end)
{
    lanes_copy<W>(s.lits + s.nl, src + anchor, end - anchor);
    s.nl += end - anchor;
}

// ---- parser state shared by the inner blocks of one unit -----------------------------------------------
struct ParseCtx {
    const u8* src;        // unit start (position 0); table entries are position + kDictSize, 0 = empty
    u32*      table;
    u32       hash_log;
    u32       window_log;
};

// Lizard_compress_fastSmall / Lizard_compress_fast on src[b0..b1)
template <class W> LZ_HD void parse_fast(const ParseCtx& c, u32 b0, u32 b1, EncStreams& s)
{
    const u8* const src = c.src;
    u32* const table = c.table;
    const u32 hl = c.hash_log;
    const bool wr = W::lane() == 0;
    const u32 max_dist = (1u << c.window_log) - 1;
    const u32 bias = kDictSize;
    const u32 low_limit = (bias + max_dist >= b0 + bias) ? bias : b0 + bias - max_dist;
    u32 anchor = b0, ip = b0;
    u32 ml = 0, mpos = 0;
    if (b1 - b0 < kMinInputForLz) goto last_literals;
    {
        const u32 mflimit = b1 - kMfLimit;
        const u8* const matchlimit = src + b1 - kLastLiterals;
        if (wr) table[hash5(ld64(src + ip), hl)] = ip + bias;
        W::sync();
        ip++;
        for (;;) {
            {   // search forward with growing stride
                u32 fwd = ip, step = 1, tries = 1u << kSkipTrigger;
                for (;;) {
                    ip = fwd;
                    fwd += step;
                    step = tries++ >> kSkipTrigger;
                    if (fwd > mflimit) goto last_literals;
                    const u32 h = hash5(ld64(src + ip), hl);
                    const u32 cand = table[h];
                    W::sync();
                    if (wr) table[h] = ip + bias;
                    W::sync();
                    const u32 cur = ip + bias;
                    if (cand < low_limit || cand >= cur || cand + max_dist < cur) continue;
                    if (cur - cand < kMinOffset) continue;
                    mpos = cand - bias;
                    if (ld32(src + mpos) != ld32(src + ip)) continue;
                    ml = count_match(src + ip + kMinMatch, src + mpos + kMinMatch, matchlimit);
                    while (ip > anchor && mpos > 0 && src[ip - 1] == src[mpos - 1]) { ip--; mpos--; ml++; }
                    break;
                }
            }
            for (;;) {   // _next_match
                emit_lz4<W>(s, src, anchor, ip, ml + kMinMatch, ip - mpos);
                ip += ml + kMinMatch;
                anchor = ip;
                if (ip > mflimit) goto last_literals;
                if (wr) table[hash5(ld64(src + ip - 2), hl)] = ip - 2 + bias;
                W::sync();
                const u32 h = hash5(ld64(src + ip), hl);
                const u32 cand = table[h];
                W::sync();
                if (wr) table[h] = ip + bias;
                W::sync();
                const u32 cur = ip + bias;
                if (cand >= low_limit && cand < cur && cand + max_dist >= cur && cur - cand >= kMinOffset) {
                    mpos = cand - bias;
                    if (ld32(src + mpos) == ld32(src + ip)) {
                        ml = count_match(src + ip + kMinMatch, src + mpos + kMinMatch, matchlimit);
                        continue;
                    }
                }
                break;
            }
            ip++;
        }
    }
last_literals:
    emit_last_literals<W>(s, src, anchor, b1);
}

// ---- lane-parallel building blocks ---------------------------------------------------------------------
// Position of the j-th probe of one fastSmall search relative to its first probe: the stride grows by one
// every 64 probes (step = searchMatchNb++ >> 6, lizard_parser_fastsmall.h:69-75), so the offsets are
// 0,1,2,...,65,67,69,... independent of the data.
LZ_HD u32 probe_offset(u32 j)
{
    if (j == 0) return 0;
    const u32 m = 62 + j;
    if (m < 64) return 1;
    const u32 q = m >> 6;
    return 1 + 32 * q * (q - 1) + q * (m - 64 * q + 1);
}

// Lizard_count with the lanes comparing consecutive 8-byte groups
template <class W> LZ_HD u32 count_match_par(const u8* a, const u8* b, const u8* limit)
{
    u32 total = 0;
    for (;;) {
        const u32 off = total + 8 * W::lane();
        const u8* pa = a + off;
        u32 n = 0; bool full = false;
        if (pa + 8 <= limit) {
            const u64 d = ld64(pa) ^ ld64(b + off);
            if (d == 0) { n = 8; full = true; }
            else {
#if defined(__CUDA_ARCH__)
                n = (u32)(__ffsll((long long)d) - 1) >> 3;
#else
                n = (u32)__builtin_ctzll(d) >> 3;
#endif
            }
        } else {
            while (pa + n < limit && pa[n] == b[off + n]) n++;
        }
        const u32 stop = W::ballot(!full);
        if (stop) { const u32 f = ctz32(stop); return total + 8 * f + W::shfl(n, f); }
        total += 8 * W::lanes();
    }
}

// backward extension: how many bytes before (ip, mpos) are equal, not crossing anchor / position 0
template <class W> LZ_HD u32 extend_back_par(const u8* src, u32 ip, u32 mpos, u32 anchor)
{
    u32 done = 0;
    for (;;) {
        const u32 k = done + W::lane() + 1;
        const bool can = ip >= anchor + k && mpos >= k;
        const bool eq = can && src[ip - k] == src[mpos - k];
        const u32 bad = W::ballot(!eq);
        if (bad) return done + ctz32(bad);
        done += W::lanes();
    }
}

// Lizard_compress_fastSmall / Lizard_compress_fast with the no-match run probed W::lanes() positions at a
// time.  Exactness argument: inside one search the probe positions do not depend on the data, each probe
// reads its bucket after all earlier probes wrote theirs, and the search stops at the first probe that
// matches.  So lanes evaluate probes j0..j0+L-1 together; a lane's candidate is the latest earlier lane of
// the same batch with the same bucket, else the table; the lowest hitting lane wins and only buckets of
// lanes up to the winner are committed (last writer per bucket).
template <class W> LZ_HD void parse_fast_par(const ParseCtx& c, u32 b0, u32 b1, EncStreams& s)
{
    const u8* const src = c.src;
    u32* const table = c.table;
    const u32 hl = c.hash_log;
    const u32 lane = W::lane(), NL = W::lanes();
    const bool wr = lane == 0;
    const u32 max_dist = (1u << c.window_log) - 1;
    const u32 bias = kDictSize;
    const u32 low_limit = (bias + max_dist >= b0 + bias) ? bias : b0 + bias - max_dist;
    u32 anchor = b0, ip = b0;
    u32 ml = 0, mpos = 0;
    if (b1 - b0 < kMinInputForLz) goto last_literals;
    {
        const u32 mflimit = b1 - kMfLimit;
        const u8* const matchlimit = src + b1 - kLastLiterals;
        if (wr) table[hash5(ld64(src + ip), hl)] = ip + bias;
        W::sync();
        ip++;
        for (;;) {
            {   // ---- search: batches of NL probes ----
                const u32 ip0 = ip;
                u32 j0 = 0;
                for (;;) {
                    const u32 j = j0 + lane;
                    const u32 P = ip0 + probe_offset(j);
                    const bool valid = ip0 + probe_offset(j + 1) <= mflimit;   // else this probe ends the block
                    u64 v = 0; u32 h = 0x80000000u | lane;                     // unique key: matches nobody
                    if (valid) { v = ld64(src + P); h = hash5(v, hl); }
                    const u32 peers = W::match_any(h);
                    const u32 below = peers & ((1u << lane) - 1);
                    const u32 pl = below ? highbit32(below) : lane;
                    const u32 prevP = W::shfl(P, pl);
                    const u32 cur = P + bias;
                    u32 cand = 0;
                    if (valid) cand = below ? prevP + bias : table[h];
                    bool hit = false;
                    if (valid && cand >= low_limit && cand < cur && cand + max_dist >= cur && cur - cand >= kMinOffset)
                        hit = ld32(src + (cand - bias)) == (u32)v;
                    const u32 hits = W::ballot(hit);
                    const u32 term = W::ballot(!valid);
                    const u32 w_lane = hits ? ctz32(hits) : 32;
                    const u32 t_lane = term ? ctz32(term) : 32;
                    const bool matched = w_lane < t_lane;
                    // lanes whose table write happens in program order before the search stops
                    u32 commit;
                    if (matched) commit = (w_lane >= 31) ? 0xffffffffu : ((2u << w_lane) - 1);
                    else commit = (t_lane >= 32) ? 0xffffffffu : ((1u << t_lane) - 1);
                    W::sync();
                    if ((commit >> lane) & 1) {
                        const u32 grp = peers & commit;
                        if (highbit32(grp) == lane) table[h] = cur;
                    }
                    W::sync();
                    if (matched) { ip = W::shfl(P, w_lane); mpos = W::shfl(cand, w_lane) - bias; break; }
                    if (t_lane < 32) { goto last_literals; }
                    j0 += NL;
                }
                ml = count_match_par<W>(src + ip + kMinMatch, src + mpos + kMinMatch, matchlimit);
                const u32 back = extend_back_par<W>(src, ip, mpos, anchor);
                ip -= back; mpos -= back; ml += back;
            }
            for (;;) {   // _next_match
                emit_lz4<W>(s, src, anchor, ip, ml + kMinMatch, ip - mpos);
                ip += ml + kMinMatch;
                anchor = ip;
                if (ip > mflimit) goto last_literals;
                if (wr) table[hash5(ld64(src + ip - 2), hl)] = ip - 2 + bias;
                W::sync();
                const u64 v = ld64(src + ip);
                const u32 h = hash5(v, hl);
                const u32 cand = table[h];
                W::sync();
                if (wr) table[h] = ip + bias;
                W::sync();
                const u32 cur = ip + bias;
                if (cand >= low_limit && cand < cur && cand + max_dist >= cur && cur - cand >= kMinOffset) {
                    mpos = cand - bias;
                    if (ld32(src + mpos) == (u32)v) {
                        ml = count_match_par<W>(src + ip + kMinMatch, src + mpos + kMinMatch, matchlimit);
                        continue;
                    }
                }
                break;
            }
            ip++;
        }
    }
last_literals:
    emit_last_literals<W>(s, src, anchor, b1);
}

// conditional table update of the priceFast family (lizard_parser_pricefast.h:170-171)
LZ_HD void pf_update(u32* slot, u32 cur, bool wr)
{
    const u32 old = *slot;
    if (old >= cur || cur >= old + kMinOffset) { if (wr) *slot = cur; }
}

// Lizard_compress_priceFast on src[b0..b1)
template <class W> LZ_HD void parse_price_fast(const ParseCtx& c, u32 b0, u32 b1, EncStreams& s, u32 min_match_long)
{
    const u8* const src = c.src;
    u32* const table = c.table;
    const u32 hl = c.hash_log;
    const bool wr = W::lane() == 0;
    const u32 bias = kDictSize;
    const u32 max_dist = (1u << c.window_log) - 1;
    u32 anchor = b0, ip = b0 + 1;
    u32 last_off = 0;
    if (b1 - b0 >= kMfLimit) {     // iend - MFLIMIT must not wrap
    const u32 mflimit = b1 - kMfLimit;
    const u8* const matchlimit = src + b1 - kLastLiterals;
    while (ip < mflimit) {
        u32 ml = 0, ref = 0;
        {   // Lizard_FindMatchFast
            const u32 cur = ip + bias;
            const u32 low = (bias + max_dist >= cur) ? bias : cur - max_dist;
            u32* slot = &table[hash5(ld64(src + ip), hl)];
            const u32 cand = *slot;
            bool found = false;
            if (last_off >= kMinOffset && cur - last_off >= low && last_off <= ip) {
                const u32 m = ip - last_off;
                if (ld32(src + m) == ld32(src + ip)) {
                    ml = count_match(src + ip + kMinMatch, src + m + kMinMatch, matchlimit) + kMinMatch;
                    ref = m; found = true;
                }
            }
            if (!found && cand < cur && cand >= low) {
                const u32 m = cand - bias;
                if (ip - m >= kMinOffset && ld32(src + m) == ld32(src + ip)) {
                    const u32 mlt = count_match(src + ip + kMinMatch, src + m + kMinMatch, matchlimit) + kMinMatch;
                    if (mlt >= min_match_long || ip - m < kMax16BitOffset) { ml = mlt; ref = m; }
                }
            }
            W::sync();
            pf_update(slot, cur, wr);
            W::sync();
        }
        if (!ml) { ip++; continue; }

        u32 ml2 = 0, start2 = 0, ref2 = 0;
        bool encode_now = false;
        if (ip - ref == last_off) { ref = ip; encode_now = true; }   // repeat offset: encoded as distance 0
        else { while (ip > anchor && ref > 0 && src[ip - 1] == src[ref - 1]) { ip--; ref--; ml++; } }

        for (;;) {
            if (!encode_now) {     // _Search: look ahead at the tail of the current match
                while (true) {
                    if (ip + ml >= mflimit) break;
                    start2 = ip + ml - 2;
                    {   // Lizard_FindMatchFaster
                        const u32 cur2 = start2 + bias;
                        const u32 low2 = (bias + max_dist >= cur2) ? bias : cur2 - max_dist;
                        u32* slot2 = &table[hash5(ld64(src + start2), hl)];
                        const u32 cand2 = *slot2;
                        ml2 = 0;
                        if (cand2 < cur2 && cand2 >= low2) {
                            const u32 m = cand2 - bias;
                            if (start2 - m >= kMinOffset && ld32(src + m) == ld32(src + start2)) {
                                const u32 mlt = count_match(src + start2 + kMinMatch, src + m + kMinMatch, matchlimit) + kMinMatch;
                                if (mlt >= min_match_long || start2 - m < kMax16BitOffset) { ml2 = mlt; ref2 = m; }
                            }
                        }
                        W::sync();
                        pf_update(slot2, cur2, wr);
                        W::sync();
                    }
                    if (!ml2) break;
                    while (start2 > ip && ref2 > 0 && src[start2 - 1] == src[ref2 - 1]) { start2--; ref2--; ml2++; }
                    if (ml2 <= ml) { ml2 = 0; break; }
                    if (start2 <= ip) { ip = start2; ref = ref2; ml = ml2; ml2 = 0; break; }
                    if (start2 - ip < 3) { ip = start2; ref = ref2; ml = ml2; ml2 = 0; continue; }
                    if (start2 < ip + ml) {
                        const u32 corr = ml - (start2 - ip);
                        start2 += corr; ref2 += corr; ml2 -= corr;
                        if (ml2 < 3) ml2 = 0;
                        if (ml2 < min_match_long && start2 - ref2 >= kMax16BitOffset) ml2 = 0;
                    }
                    break;
                }
            }
            // _Encode
            emit_lizv1<W>(s, src, anchor, ip, ml, ip - ref, last_off);
            ip += ml;
            anchor = ip;
            if (!ml2) break;
            ip = start2; ref = ref2; ml = ml2; ml2 = 0;
            encode_now = false;
        }
    }
    }
    emit_last_literals<W>(s, src, anchor, b1);
}

// Lizard_compress_priceFast with the no-match run probed W::lanes() consecutive positions at a time.
// Per position the reference (lizard_parser_pricefast.h:158-173) tests the repeat offset first, then the
// bucket's candidate, then conditionally refreshes the bucket.  last_off is constant during a no-match run,
// positions advance by one, so L lanes evaluate L consecutive positions; the only cross-lane dependency is
// the bucket value, which each lane reconstructs by replaying the conditional updates of the earlier lanes
// that share its bucket.  The lowest hitting lane wins; buckets of lanes up to the winner are committed.
template <class W> LZ_HD void parse_price_fast_par(const ParseCtx& c, u32 b0, u32 b1, EncStreams& s, u32 min_match_long)
{
    const u8* const src = c.src;
    u32* const table = c.table;
    const u32 hl = c.hash_log;
    const u32 lane = W::lane(), NL = W::lanes();
    const bool wr = lane == 0;
    const u32 bias = kDictSize;
    const u32 max_dist = (1u << c.window_log) - 1;
    u32 anchor = b0, ip = b0 + 1;
    u32 last_off = 0;
    if (b1 - b0 >= kMfLimit) {
    const u32 mflimit = b1 - kMfLimit;
    const u8* const matchlimit = src + b1 - kLastLiterals;
    while (ip < mflimit) {
        u32 ml = 0, ref = 0;
        {   // ---- batch of NL positions ----
            const u32 P = ip + lane;
            const bool valid = P < mflimit;
            const u32 cur = P + bias;
            const u32 low = (bias + max_dist >= cur) ? bias : cur - max_dist;
            u64 v = 0; u32 h = 0x80000000u | lane;
            if (valid) { v = ld64(src + P); h = hash5(v, hl); }
            const u32 peers = W::match_any(h);
            u32 below = peers & ((1u << lane) - 1);
            u32 seen = valid ? table[h] : 0;
            while (below) {                                   // replay earlier same-bucket lanes, in order
                const u32 bl = ctz32(below); below &= below - 1;
                const u32 pb = ip + bl + bias;
                if (seen >= pb || pb >= seen + kMinOffset) seen = pb;
            }
            bool rep_hit = false, hash_hit = false;
            if (valid) {
                if (last_off >= kMinOffset && last_off <= P && cur - last_off >= low)
                    rep_hit = ld32(src + (P - last_off)) == (u32)v;
                if (!rep_hit && seen < cur && seen >= low) {
                    const u32 m = seen - bias;
                    if (P - m >= kMinOffset && ld32(src + m) == (u32)v) {
                        if (P - m < kMax16BitOffset) hash_hit = true;
                        else hash_hit = count_match(src + P + kMinMatch, src + m + kMinMatch, matchlimit) + kMinMatch >= min_match_long;
                    }
                }
            }
            const u32 newval = (seen >= cur || cur >= seen + kMinOffset) ? cur : seen;
            const u32 hits = W::ballot(rep_hit || hash_hit);
            const u32 vmask = W::ballot(valid);
            const u32 w_lane = hits ? ctz32(hits) : 32;
            const u32 commit = (w_lane < 32) ? ((w_lane >= 31) ? 0xffffffffu : ((2u << w_lane) - 1)) : vmask;
            W::sync();
            if ((commit >> lane) & 1) {
                const u32 grp = peers & commit;
                if (highbit32(grp) == lane) table[h] = newval;
            }
            W::sync();
            if (w_lane == 32) { ip += NL; continue; }
            ip = W::shfl(P, w_lane);
            const u32 is_rep = W::shfl(rep_hit ? 1u : 0u, w_lane);
            ref = is_rep ? ip - last_off : W::shfl(seen, w_lane) - bias;
            ml = count_match_par<W>(src + ip + kMinMatch, src + ref + kMinMatch, matchlimit) + kMinMatch;
        }

        u32 ml2 = 0, start2 = 0, ref2 = 0;
        bool encode_now = false;
        if (ip - ref == last_off) { ref = ip; encode_now = true; }
        else { const u32 back = extend_back_par<W>(src, ip, ref, anchor); ip -= back; ref -= back; ml += back; }

        for (;;) {
            if (!encode_now) {
                while (true) {
                    if (ip + ml >= mflimit) break;
                    start2 = ip + ml - 2;
                    {   // Lizard_FindMatchFaster (uniform: one position)
                        const u32 cur2 = start2 + bias;
                        const u32 low2 = (bias + max_dist >= cur2) ? bias : cur2 - max_dist;
                        const u64 v2 = ld64(src + start2);
                        u32* slot2 = &table[hash5(v2, hl)];
                        const u32 cand2 = *slot2;
                        ml2 = 0;
                        bool ok = false; u32 m = 0;
                        if (cand2 < cur2 && cand2 >= low2) {
                            m = cand2 - bias;
                            ok = start2 - m >= kMinOffset && ld32(src + m) == (u32)v2;
                        }
                        W::sync();
                        pf_update(slot2, cur2, wr);
                        W::sync();
                        if (ok) {
                            const u32 mlt = count_match_par<W>(src + start2 + kMinMatch, src + m + kMinMatch, matchlimit) + kMinMatch;
                            if (mlt >= min_match_long || start2 - m < kMax16BitOffset) { ml2 = mlt; ref2 = m; }
                        }
                    }
                    if (!ml2) break;
                    {   const u32 back = extend_back_par<W>(src, start2, ref2, ip); start2 -= back; ref2 -= back; ml2 += back; }
                    if (ml2 <= ml) { ml2 = 0; break; }
                    if (start2 <= ip) { ip = start2; ref = ref2; ml = ml2; ml2 = 0; break; }
                    if (start2 - ip < 3) { ip = start2; ref = ref2; ml = ml2; ml2 = 0; continue; }
                    if (start2 < ip + ml) {
                        const u32 corr = ml - (start2 - ip);
                        start2 += corr; ref2 += corr; ml2 -= corr;
                        if (ml2 < 3) ml2 = 0;
                        if (ml2 < min_match_long && start2 - ref2 >= kMax16BitOffset) ml2 = 0;
                    }
                    break;
                }
            }
            emit_lizv1<W>(s, src, anchor, ip, ml, ip - ref, last_off);
            ip += ml;
            anchor = ip;
            if (!ml2) break;
            ip = start2; ref = ref2; ml = ml2; ml2 = 0;
            encode_now = false;
        }
    }
    }
    emit_last_literals<W>(s, src, anchor, b1);
}

// ---- Huffman stage of one stream --------------------------------------------------------------------------
struct EncHufWork {               // per-warp scratch for the entropy stage
    HufEncScratch ws;
    u32 (*seg_count)[256];        // [4][256] per-segment byte histograms (shared memory on the device)
    u32 count[256];
    HufPlan plan;
    u32 pack[(kBlockSizePad + 512) / 4 + 64];   // aligned staging of packed segments
};

template <class W> LZ_HD void histogram4(const u8* p, u32 n, u32 seg, u32 (*seg_count)[256], u32* count)
{
    for (u32 i = W::lane(); i < 1024; i += W::lanes()) (&seg_count[0][0])[i] = 0;
    W::sync();
    for (u32 i = W::lane(); i < n; i += W::lanes()) {
        u32 k = i / seg;
#if defined(__CUDA_ARCH__)
        atomicAdd(&seg_count[k][p[i]], 1u);
#else
        seg_count[k][p[i]]++;
#endif
    }
    W::sync();
    for (u32 i = W::lane(); i < 256; i += W::lanes()) count[i] = seg_count[0][i] + seg_count[1][i] + seg_count[2][i] + seg_count[3][i];
    W::sync();
}

// Pack symbols p[0..m) last-to-first, LSB-first, then a 1 bit (HUF_compress1X_usingCTable).  `out` is a
// 4-byte aligned staging area; returns nothing, size is already known from the plan.
template <class W> LZ_HD void huf_pack_segment(u32* out, const u8* p, u32 m, const HufCode* codes)
{
    const u32 L = W::lanes(), lane = W::lane();
    const u32 chunk = (m + L - 1) / L;
    // lane l owns symbol indices [hi - chunk, hi) with hi = m - l*chunk, walked downward
    const long hi = (long)m - (long)lane * chunk;
    long lo = hi - (long)chunk; if (lo < 0) lo = 0;
    u32 mybits = 0;
    for (long i = hi - 1; i >= lo; --i) mybits += codes[p[i]].nbits;
    const bool owns_end = (lo == 0 && hi > 0) || (m == 0 && lane == 0);
    if (owns_end) mybits += 1;
    u32 total = 0;
    const u32 start = W::excl_scan(mybits, &total);
    if (mybits) { out[start >> 5] = 0; out[(start + mybits - 1) >> 5] = 0; }
    W::sync();
    if (mybits) {
        u64 acc = 0; u32 nacc = 0;          // bits waiting to be stored, aligned so that bit 0 is stream bit `pos`
        u32 pos = start;
        const u32 last_word = (start + mybits - 1) >> 5;
        // bring the accumulator to a word boundary view: keep (pos & 31) zero bits in front
        nacc = pos & 31; pos &= ~31u;
        for (long i = hi - 1; i >= lo; --i) {
            const HufCode c = codes[p[i]];
            acc |= (u64)c.val << nacc; nacc += c.nbits;
            if (nacc >= 32) {
                const u32 w = pos >> 5;
                const u32 v = (u32)acc;
                if (w == (start >> 5) || w == last_word) {
#if defined(__CUDA_ARCH__)
                    atomicOr(&out[w], v);
#else
                    out[w] |= v;
#endif
                } else out[w] = v;
                acc >>= 32; nacc -= 32; pos += 32;
            }
        }
        if (owns_end) { acc |= 1ull << nacc; nacc += 1; }
        while (nacc > 0) {
            const u32 w = pos >> 5;
            const u32 v = (u32)acc;
            if (w == (start >> 5) || w == last_word) {
#if defined(__CUDA_ARCH__)
                atomicOr(&out[w], v);
#else
                out[w] |= v;
#endif
            } else out[w] = v;
            acc >>= 32; nacc = nacc > 32 ? nacc - 32 : 0; pos += 32;
        }
    }
    W::sync();
}

// Lizard_writeStream: returns 1 (Huffman'd), 0 (raw) or -1 (does not fit)
template <class W> LZ_HD int write_stream(bool use_huff, const u8* p, u32 n, u8* dst, long& op, long oend, EncHufWork* hw)
{
    const bool wr = W::lane() == 0;
    if (use_huff && n > 1024) {
        if (op + 6 > oend) return -1;
        const u32 seg = (n + 3) / 4;
        histogram4<W>(p, n, seg, hw->seg_count, hw->count);
        if (wr) huf_plan(hw->plan, hw->count, hw->seg_count, n, p[0], &hw->ws);
        W::sync();
        const int status = W::bcast(hw->plan.status);
        const u32 c = (u32)W::bcast((int)hw->plan.total);
        if (status != kHufPlanRaw && c > 0 && c + c / 8 + 512 < n) {
            if (oend - (op + 6) < (long)c) return -1;
            if (wr) { wr_le24(dst + op, n); wr_le24(dst + op + 3, c); }
            long o = op + 6;
            if (status == kHufPlanRle) { if (wr) dst[o] = hw->plan.rle_byte; }
            else {
                const u32 hs = (u32)W::bcast((int)hw->plan.header_size);
                lanes_copy<W>(dst + o, hw->ws.header, hs);
                o += hs;
                if (wr) { wr_le16(dst + o, hw->plan.seg_bytes[0]); wr_le16(dst + o + 2, hw->plan.seg_bytes[1]); wr_le16(dst + o + 4, hw->plan.seg_bytes[2]); }
                o += 6;
                for (u32 k = 0; k < 4; ++k) {
                    const u32 m = k < 3 ? seg : n - 3 * seg;
                    huf_pack_segment<W>(hw->pack, p + k * seg, m, hw->ws.codes);
                    const u32 sb = (u32)W::bcast((int)hw->plan.seg_bytes[k]);
                    lanes_copy<W>(dst + o, (const u8*)hw->pack, sb);
                    W::sync();
                    o += sb;
                }
            }
            op += (long)c + 6;
            return 1;
        }
    }
    if (op + 3 + (long)n > oend) return -1;
    if (wr) wr_le24(dst + op, n);
    lanes_copy<W>(dst + op + 3, p, n);
    op += 3 + (long)n;
    return 0;
}

// Lizard_writeBlock: 0 ok, 1 output error
template <class W> LZ_HD int write_block(const EncStreams& s, const u8* in, u32 in_size, u8* dst, long& op, long oend,
                                        bool huffman, EncHufWork* hw)
{
    const bool wr = W::lane() == 0;
    const long start = op;
    const u32 sum = s.nf + s.nl + s.n16 + s.n24;
    bool raw = (s.nl < kWildCopy) || (sum + 5 * 3 + 1 > in_size);
    if (!raw) {
        u32 hdr = 0;
        int r;
        op += 1;
        r = write_stream<W>(false, s.lits, 0, dst, op, oend, hw);       if (r < 0) return 1;   // (empty) lengths stream
        r = write_stream<W>(false, s.off16, s.n16, dst, op, oend, hw);  if (r < 0) return 1;
        r = write_stream<W>(false, s.off24, s.n24, dst, op, oend, hw);  if (r < 0) return 1;
        r = write_stream<W>(huffman, s.flags, s.nf, dst, op, oend, hw); if (r < 0) return 1;  hdr += (u32)r * kFlagFlags;
        r = write_stream<W>(huffman, s.lits, s.nl, dst, op, oend, hw);  if (r < 0) return 1;  hdr += (u32)r * kFlagLiterals;
        if (wr) dst[start] = (u8)hdr;
        const u32 out = (u32)(op - start);
        if (out + out / 32 + 512 > in_size) raw = true;
        else return 0;
    }
    if ((u32)(oend - start) < in_size + 4 || oend - start < 0) return 1;
    W::sync();      // the abandoned stream bytes (written by lane 0) must not land after the raw copy
    if (wr) { dst[start] = (u8)kFlagRaw; wr_le24(dst + start + 1, in_size); }
    lanes_copy<W>(dst + start + 4, in, in_size);
    op = start + 4 + (long)in_size;
    return 0;
}

struct EncWork {                 // per-warp global scratch
    u8 lits[kBlockSizePad];
    u8 flags[kBlockSizePad];
    u8 off16[kBlockSizePad];
    u8 off24[kBlockSizePad];
    EncHufWork huf;
};

// Lizard_compress_extState with a clean table: returns compressed size or 0
template <class W> LZ_HD int encode_unit(const u8* src, u32 src_size, u8* dst, u32 cap, int level,
                                        u32* table, EncWork* work)
{
    const LevelParams lp = level_params(level);
    if (lp.parser == kParserUnsupported) return 0;
    if (src_size > kMaxInputSize) return 0;
    for (u32 i = W::lane(); i < (1u << lp.hashLog); i += W::lanes()) table[i] = 0;
    W::sync();
    const bool wr = W::lane() == 0;
    long op = 0;
    const long oend = (long)cap;
    if (cap < 1) return 0;                          // the reference would write the level byte regardless
    if (wr) dst[0] = (u8)level;
    op = 1;
    ParseCtx pc; pc.src = src; pc.table = table; pc.hash_log = lp.hashLog; pc.window_log = lp.windowLog;
    u32 pos = 0;
    while (pos < src_size) {
        const u32 part = src_size - pos < kBlockSize ? src_size - pos : kBlockSize;
        EncStreams s;
        s.lits = work->lits; s.flags = work->flags; s.off16 = work->off16; s.off24 = work->off24;
        s.nl = s.nf = s.n16 = s.n24 = 0;
        if (lp.parser == kParserPriceFast) parse_price_fast_par<W>(pc, pos, pos + part, s, lp.minMatchLongOff);
        else parse_fast_par<W>(pc, pos, pos + part, s);
        W::sync();
        if (write_block<W>(s, src + pos, part, dst, op, oend, lp.huffman != 0, &work->huf)) return 0;
        W::sync();
        pos += part;
    }
    return (int)op;
}

}  // namespace lzb
