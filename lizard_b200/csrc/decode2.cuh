// decode2.cuh -- second-generation block decoder: one CTA per unit, a PARSER warp and a COPIER warp.
//
// The first-generation kernel (decode.cuh) gives a whole unit to one warp, which alternates between resolving the token
// chain of a batch (a serial chain of dependent loads: a token's place in the literals stream depends on every earlier
// length-extension byte) and moving the batch's bytes.  Here the two halves run on different warps of one CTA:
//
//   parser warp  -- walks the token stream exactly as decode_tokens_* does (same checks, same error codes; reference:
//                   lib/lizard_decompress_lz4.h:7-163, lib/lizard_decompress_liz.h:14-220) but moves no byte: every token
//                   becomes a 16-byte RECORD {output position, literals-stream position, match position, offset} in a
//                   shared-memory ring.  It reads the literals stream (length-extension bytes, inline LZ4 offsets) from a
//                   shared-memory ring that TMA bulk copies (cp.async.bulk + mbarrier) keep filled ahead of it.
//   copier warp  -- consumes published batches of records DESTINATION-FIRST: the output is cut into aligned 16-byte chunks,
//                   lane i of a step owns chunk i; a binary search over the batch's record positions tells it which run its
//                   chunk starts in; it fetches the two aligned 16-byte vectors that hold the source bytes (literals from
//                   the TMA ring, match sources from the already written output), realigns them with funnel shifts and
//                   stores one full vector into a 2 KiB shared-memory TILE of the output.  Chunks that contain a run
//                   boundary are finished by a second, compacted pass (one lane per such chunk walks its remaining pieces
//                   and merges them in registers); matches that read bytes produced in the same tile span ("late" matches)
//                   are resolved last from the tile.  Complete chunks leave the tile as coalesced 128-bit stores.
//
// The copier never sees a token and the parser never touches the output, so the chain latency of batch j+1 overlaps the
// byte traffic of batch j on the same SM without relying on other units' warps.  Everything a lane does is written against
// the lane policy W (lanes.cuh), so the CPU test-suite runs the same parser and copier through an in-line sink (the copier
// is called as soon as a batch is published) with one lane and with the 32-lane emulator; the device-only part is the
// plumbing between the two warps (mbarrier pipeline, TMA ring), which lives at the bottom of this file.
#pragma once
#include "decode.cuh"

namespace lzb {

// A sequence in ALIGNED OUTPUT SPACE: position x of the unit's output is x + (dst & 15), so that chunk c = bytes
// [16c, 16c+16) of that space is a 16-byte aligned vector of global memory.  The literal run is [opos, mdst) and comes from
// the literals stream at lsrc; the match is [mdst, next record's opos) and copies from `off` bytes back (off == 0 with an
// empty match for literal-only records: last literals, split long runs).
struct alignas(16) OutRec { u32 opos, lsrc, mdst, off; };

enum : u32 {
    kBatchSlots = 4,            // batches in flight between parser and copier
    kSlotRecs   = 40,           // record slots per batch: 32 records + the end marker, padded (the search reads up to index 31)
    kTileBytes  = 6144,         // output window of the copier: one span = at most this many bytes
    kTileChunks = kTileBytes / 16,
    kExtraCap   = 128,          // chunks waiting for the merge pass
    kGroup      = 4,            // chunks a lane works on at once (all their loads in flight together)
};

struct CopyShared {                       // shared memory of one parser / copier pair
    OutRec rec[kBatchSlots * kSlotRecs];  // records of batch slot k at [k * kSlotRecs, ...)
    u32    pos[kBatchSlots * kSlotRecs];  // opos of the records (the copier's search key); every entry behind a batch's last
                                          // record holds the batch's end position, so a search needs no bound
    alignas(16) u8 tile[kTileBytes];      // output bytes [T0, T0 + kTileBytes) under construction
    u32    extras[kExtraCap];             // chunks of the current span that need the merge pass
    u32    late_bits;                     // records (bit = index in the batch) whose match reads bytes of the current span
};

struct CopyState {                        // registers of the copier
    u8* dst_al;                           // unit's dst rounded down to 16 bytes
    u32 unit_lo;                          // first aligned-space position that belongs to the unit (= dst & 15)
};

// ---- small pieces ---------------------------------------------------------------------------------------------------
LZ_HD void lanes_or_u32(u32* p, u32 v)          // several lanes may set bits of the same word
{
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
LZ_HD Vec16 vec16_zero() { Vec16 r; r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0; return r; }

// bytes [delta, delta+16) of the 32 bytes a|b
LZ_HD Vec16 realign16(const Vec16& a, const Vec16& b, u32 delta)
{
#if defined(__CUDA_ARCH__)
    const bool w2 = (delta & 8) != 0, w1 = (delta & 4) != 0;
    const u32 bs = (delta & 3) * 8;
    const u32 y0 = w2 ? a.w[2] : a.w[0], y1 = w2 ? a.w[3] : a.w[1], y2 = w2 ? b.w[0] : a.w[2],
              y3 = w2 ? b.w[1] : a.w[3], y4 = w2 ? b.w[2] : b.w[0], y5 = w2 ? b.w[3] : b.w[1];
    const u32 x0 = w1 ? y1 : y0, x1 = w1 ? y2 : y1, x2 = w1 ? y3 : y2, x3 = w1 ? y4 : y3, x4 = w1 ? y5 : y4;
    Vec16 r;
    r.w[0] = __funnelshift_r(x0, x1, bs); r.w[1] = __funnelshift_r(x1, x2, bs);
    r.w[2] = __funnelshift_r(x2, x3, bs); r.w[3] = __funnelshift_r(x3, x4, bs);
    return r;
#else
    u8 t[32]; memcpy(t, &a, 16); memcpy(t + 16, &b, 16);
    Vec16 r; memcpy(&r, t + delta, 16); return r;
#endif
}

// acc with its bytes [lo, 16) replaced by x's
LZ_HD Vec16 merge_from(const Vec16& acc, const Vec16& x, u32 lo)
{
    Vec16 r;
#if defined(__CUDA_ARCH__)
    const int t = (int)lo * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int sh = t - 32 * i; sh = sh < 0 ? 0 : sh;
        u32 m;
        asm("shl.b32 %0, %1, %2;" : "=r"(m) : "r"(0xffffffffu), "r"((u32)sh));      // PTX shl clamps: shift >= 32 gives 0
        r.w[i] = (x.w[i] & m) | (acc.w[i] & ~m);
    }
#else
    u8 a[16], b[16]; memcpy(a, &acc, 16); memcpy(b, &x, 16);
    for (u32 i = lo; i < 16; ++i) a[i] = b[i];
    memcpy(&r, a, 16);
#endif
    return r;
}

// A 16-byte source vector "in flight": the two aligned vectors that hold it and its byte phase.  Loading and realigning are
// separate steps so that a lane can have the loads of several chunks outstanding before it touches any of the data.
struct Raw16 { Vec16 v0, v1; u32 delta; };
LZ_HD Vec16 raw_finish(const Raw16& r) { return realign16(r.v0, r.v1, r.delta); }
LZ_HD Raw16 raw_none() { Raw16 r; r.v0 = vec16_zero(); r.v1 = r.v0; r.delta = 0; return r; }

// 16 bytes X with X[i] = out[a + i] for i in [lo, hi), read from the already written output through aligned vectors.  Only
// vectors that hold at least one needed byte are touched (a + lo >= 0 is the caller's bound check), so no access leaves the
// 16-byte granules the needed bytes occupy.
LZ_HD Raw16 out_load(const u8* dst_al, int a, u32 lo, u32 hi)
{
    Raw16 r = raw_none();
#if defined(__CUDA_ARCH__)
    const int a0 = a & ~15;
    if (a + (int)lo < a0 + 16) r.v0 = ld_vec16(dst_al + a0);
    if (a + (int)hi > a0 + 16) r.v1 = ld_vec16(dst_al + a0 + 16);
    r.delta = (u32)a & 15u;
#else
    u8 t[16] = {0};
    for (u32 i = lo; i < hi; ++i) t[i] = dst_al[(long)a + (long)i];
    memcpy(&r.v0, t, 16);
#endif
    return r;
}

// Literals stream as a plain pointer (host build)
struct LitPtr {
    const u8* p;
    LZ_HDM u32 byte(long pos) const { return p[pos]; }
    LZ_HDM Raw16 load(int a, u32 lo, u32 hi) const
    {
        Raw16 r = raw_none();
        u8 t[16] = {0};
        for (u32 i = lo; i < hi; ++i) t[i] = p[(long)a + (long)i];
        memcpy(&r.v0, t, 16);
        return r;
    }
};

LZ_HD Vec16 tile_load(const u8* tile, u32 byte_off)
{
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4*>(tile + byte_off);
    Vec16 r; r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w; return r;
#else
    Vec16 r; memcpy(&r, tile + byte_off, 16); return r;
#endif
}
LZ_HD void tile_store(u8* tile, u32 byte_off, const Vec16& v)
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(tile + byte_off) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
#else
    memcpy(tile + byte_off, &v, 16);
#endif
}

// largest i in [0, 32) with pos[i] <= p: pos[0] <= p is the caller's invariant, entries behind the batch's last record hold
// the batch's end position, which lies above every p that is searched for
LZ_HD u32 rec_search(const u32* pos, u32 p)
{
    u32 lo = 0;
#pragma unroll
    for (u32 stp = 16; stp; stp >>= 1) if (pos[lo + stp] <= p) lo += stp;
    return lo;
}

// ---- the copier -----------------------------------------------------------------------------------------------------
// merge pass over the chunks listed in extras[0, nx): one lane per chunk walks the chunk's remaining pieces.  The walk
// depends only on the records, so a lane first issues the loads of up to four pieces and then merges them in order.
template <class W, class LV>
LZ_HD void copy_merge_pass(CopyShared* cs, const LV& lv, const CopyState& st, u32 r0, u32 T0, u32 nx, u32 cur_begin, u32 cur_end)
{
    const u32 lane = W::lane(), L = W::lanes();
    u8* const tile = cs->tile;
    for (u32 base = 0; base < nx; base += L) {
        const u32 i = base + lane;
        if (i < nx) {
            const u32 ex = cs->extras[i];
            const u32 toff = (ex & 511u) << 4;
            const u32 c0 = T0 + toff;
            u32 s = (ex >> 9) & 63u;
            u32 p = c0 + (ex >> 16);
            const u32 cend = c0 + 16 < cur_end ? c0 + 16 : cur_end;
            Vec16 acc = tile_load(tile, toff);
            while (p < cend) {
                Raw16 raw[4]; u32 lo[4]; bool have[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    have[k] = false; lo[k] = 0; raw[k] = raw_none();
                    // skip what lies behind p: empty matches, empty literal runs
                    while (p < cend && p >= cs->pos[r0 + s + 1]) ++s;
                    if (p < cend) {
                        const OutRec d = cs->rec[r0 + s];
                        const u32 nxt = cs->pos[r0 + s + 1];
                        lo[k] = p - c0;
                        if (p < d.mdst) {
                            const u32 e = d.mdst < cend ? d.mdst : cend;
                            raw[k] = lv.load((int)(d.lsrc + c0 - d.opos), p - c0, e - c0);
                            have[k] = true; p = e;
                        } else {
                            const u32 e = nxt < cend ? nxt : cend;
                            if (e - d.off > cur_begin) lanes_or_u32(&cs->late_bits, 1u << s);
                            else { raw[k] = out_load(st.dst_al, (int)(c0 - d.off), p - c0, e - c0); have[k] = true; }
                            p = e;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) if (have[k]) acc = merge_from(acc, raw_finish(raw[k]), lo[k]);
            }
            tile_store(tile, toff, acc);
        }
    }
}

// matches that read bytes produced in the same span, in record order (= output order): their source bytes are final by the
// time they are read.  Rare (datagen: a few per cent of the matches), kept out of line.
template <class W>
LZ_HD_COLD void copy_late_pass(CopyShared* cs, const CopyState& st, u32 r0, u32 T0, u32 cur_begin, u32 cur_end)
{
    const u32 lane = W::lane(), L = W::lanes();
    u8* const tile = cs->tile;
    for (u32 bits = cs->late_bits; bits; bits &= bits - 1) {
        const u32 s = ctz32(bits);
        const OutRec d = cs->rec[r0 + s];
        const u32 nxt = cs->pos[r0 + s + 1];
        const u32 a = d.mdst > cur_begin ? d.mdst : cur_begin;
        const u32 b = nxt < cur_end ? nxt : cur_end;
        if (a >= b) continue;
        const u32 len = b - a, off = d.off;
        if (off == 0) {                                            // no encoder emits it: defined output (zeros), as lanes_match
            for (u32 i = lane; i < len; i += L) tile[a + i - T0] = 0;
            W::sync();
        } else if (off >= L || off >= len) {
            for (u32 base = 0; base < len; base += L) {           // row by row: a row may read what the previous one wrote
                const u32 i = base + lane;
                u8 v = 0;
                if (i < len) {
                    const u32 sp = a - off + i;
                    v = sp >= cur_begin ? tile[sp - T0] : st.dst_al[sp];
                }
                W::sync();
                if (i < len) tile[a + i - T0] = v;
                W::sync();
            }
        } else {                                                   // periodic extension of the `off` bytes before a
            for (u32 i = lane; i < len; i += L) {
                const u32 sp = a - off + (i % off);
                tile[a + i - T0] = sp >= cur_begin ? tile[sp - T0] : st.dst_al[sp];
            }
            W::sync();
        }
    }
}

// One span: output positions [cur_begin, cur_end), at most kTileBytes from cur_begin & ~15, covered by records [r0, r0 + nrec).
// The tile holds aligned-space bytes from T0 = cur_begin & ~15 on; on entry its first chunk holds the bytes of [T0, cur_begin)
// (left there by the previous span or by copy_resync).  Invariant on entry and exit: every byte below cur_begin (on exit:
// below cur_end) is in global memory.
template <class W, class LV>
LZ_HD void copy_tile_span(CopyShared* cs, const LV& lv, const CopyState& st, u32 r0, u32 nrec, u32 cur_begin, u32 cur_end)
{
    const u32 lane = W::lane(), L = W::lanes();
    const u32 T0 = cur_begin & ~15u;
    u8* const tile = cs->tile;
    (void)nrec;
    if (lane == 0) cs->late_bits = 0;
    W::sync();
    // ---- pass 1: the run a chunk starts in fills the chunk's vector; anything else in the chunk is left to the merge pass.
    //      A lane takes kGroup chunks per step (L chunks apart, so that each load instruction of the warp covers consecutive
    //      vectors): first all searches, then all loads, then the realignments and stores.
    u32 nx = 0;
    for (u32 stepb = T0; stepb < cur_end; stepb += 16 * L * kGroup) {
        u32 sI[kGroup], eI[kGroup]; Raw16 raw[kGroup];
#pragma unroll
        for (u32 k = 0; k < kGroup; ++k) {
            const u32 c0 = stepb + 16 * (lane + L * k);
            sI[k] = c0 < cur_end ? rec_search(cs->pos + r0, c0 < cur_begin ? cur_begin : c0) : 0u;
        }
#pragma unroll
        for (u32 k = 0; k < kGroup; ++k) {
            const u32 c0 = stepb + 16 * (lane + L * k);
            raw[k] = raw_none(); eI[k] = c0;
            if (c0 < cur_end && c0 >= cur_begin) {
                const u32 cend = c0 + 16 < cur_end ? c0 + 16 : cur_end;
                const u32 s = sI[k];
                const OutRec d = cs->rec[r0 + s];
                if (c0 < d.mdst) {
                    eI[k] = d.mdst < cend ? d.mdst : cend;
                    raw[k] = lv.load((int)(d.lsrc + (c0 - d.opos)), 0, eI[k] - c0);
                } else {
                    const u32 nxt = cs->pos[r0 + s + 1];
                    eI[k] = nxt < cend ? nxt : cend;
                    if (eI[k] - d.off > cur_begin) lanes_or_u32(&cs->late_bits, 1u << s);
                    else raw[k] = out_load(st.dst_al, (int)(c0 - d.off), 0, eI[k] - c0);
                }
            }
        }
#pragma unroll
        for (u32 k = 0; k < kGroup; ++k) {
            const u32 c0 = stepb + 16 * (lane + L * k);
            bool want = false; u32 ex = 0;
            if (c0 < cur_end) {
                const u32 cend = c0 + 16 < cur_end ? c0 + 16 : cur_end;
                if (c0 >= cur_begin) tile_store(tile, c0 - T0, raw_finish(raw[k]));
                // c0 < cur_begin: the span's first chunk, partly filled earlier -- all its new bytes go through the merge pass
                if (eI[k] < cend || c0 < cur_begin) {
                    const u32 p = c0 < cur_begin ? cur_begin : eI[k];
                    ex = ((c0 - T0) >> 4) | (sI[k] << 9) | ((p - c0) << 16); want = true;
                }
            }
            const u32 wm = W::ballot(want);
            if (nx + popc32(wm) > kExtraCap) {               // list full: merge what is listed (those chunks are stored), start over
                W::sync();
                copy_merge_pass<W>(cs, lv, st, r0, T0, nx, cur_begin, cur_end);
                W::sync();
                nx = 0;
            }
            if (want) cs->extras[nx + popc32(wm & ((1u << lane) - 1))] = ex;
            nx += popc32(wm);
        }
    }
    W::sync();
    copy_merge_pass<W>(cs, lv, st, r0, T0, nx, cur_begin, cur_end);
    W::sync();
    if (cs->late_bits) copy_late_pass<W>(cs, st, r0, T0, cur_begin, cur_end);
    W::sync();
    // ---- flush: complete chunks as aligned 16-byte stores (consecutive lanes, consecutive vectors); the bytes of a trailing
    //      partial chunk go out one by one so that the invariant holds, and the chunk itself moves to the front of the tile,
    //      where the next span expects it
    const u32 fe = cur_end & ~15u;
    for (u32 cb = T0; cb < fe; cb += 16 * L * kGroup) {
        Vec16 v[kGroup];
#pragma unroll
        for (u32 k = 0; k < kGroup; ++k) {
            const u32 c = cb + 16 * (lane + L * k);
            if (c < fe) v[k] = tile_load(tile, c - T0);
        }
#pragma unroll
        for (u32 k = 0; k < kGroup; ++k) {
            const u32 c = cb + 16 * (lane + L * k);
            if (c < fe) {
                if (c >= st.unit_lo) st_vec16(st.dst_al + c, v[k].w[0], v[k].w[1], v[k].w[2], v[k].w[3]);
                else for (u32 q = st.unit_lo; q < c + 16; ++q) st.dst_al[q] = tile[q - T0];     // the unit starts inside this chunk
            }
        }
    }
    if (cur_end & 15u) {
        const u32 from = fe > st.unit_lo ? fe : st.unit_lo;
        for (u32 q = from + lane; q < cur_end; q += L) st.dst_al[q] = tile[q - T0];
        W::sync();
        if (fe != T0) {
            u32 w = 0;
            if (lane < 4) w = *reinterpret_cast<const u32*>(tile + (fe - T0) + 4 * lane);
            if (L < 4) { Vec16 t = tile_load(tile, fe - T0); tile_store(tile, 0, t); }
            W::sync();
            if (lane < 4 && L >= 4) *reinterpret_cast<u32*>(tile + 4 * lane) = w;
        }
    }
    W::sync();
}

// Published records [r0, r0 + nrec) produce output [B0, B1): in spans of at most one tile.
template <class W, class LV>
LZ_HD void copy_span(CopyShared* cs, const LV& lv, CopyState& st, u32 r0, u32 nrec, u32 B0, u32 B1)
{
    u32 cur = B0;
    while (cur < B1) {
        const u32 lim = (cur & ~15u) + kTileBytes;
        const u32 tend = B1 < lim ? B1 : lim;
        copy_tile_span<W>(cs, lv, st, r0, nrec, cur, tend);
        cur = tend;
    }
}

// (Re)start of the copier at aligned-space position `apos` (start of a unit; after the parser has written output itself):
// the bytes of apos's chunk that lie below apos are taken over from global memory into the front of the tile.
template <class W>
LZ_HD void copy_resync(CopyShared* cs, CopyState& st, u32 apos)
{
    const u32 c = apos & ~15u;
    for (u32 q = c + W::lane(); q < apos; q += W::lanes()) cs->tile[q - c] = q >= st.unit_lo ? st.dst_al[q] : (u8)0;
    W::sync();
}

// ---- sinks -----------------------------------------------------------------------------------------------------------
// What the parser talks to.  The in-line sink (host build, emulator) runs the copier at once on the same lanes; the device
// sink (PairSink, below) hands the batch to the copier warp.
//   span()                         most literals-stream bytes one published batch may cover
//   require(lp)                    the literals stream is readable from lp to lp + span() (or its end)
//   publish(act, r, n, B1, lp0)    lanes < n hold records of consecutive sequences ending at output position B1; lp0 = where their
//                                  literals-stream bytes start
//   drain()                        every published byte is in global memory when this returns
//   resync(apos)                   the parser wrote output itself up to apos (only after drain())
template <class W>
struct InlineSink {
    CopyShared* cs; CopyState st; LitPtr lits; u32 nrec_total; u32 span_limit; u32 out_pos;
    LZ_HDM u32 span() const { return span_limit; }
    LZ_HDM void require(long) {}
    LZ_HDM void set_stream(const u8* p, u32) { lits.p = p; }
    LZ_HDM const LitPtr& view() const { return lits; }
    LZ_HDM void publish(bool act, const OutRec& r, u32 n, u32 B1, u32)
    {
        const u32 r0 = 0;
        for (u32 i = W::lane(); i < kSlotRecs; i += W::lanes()) cs->pos[r0 + i] = B1;     // end marker and padding
        W::sync();
        if (act) { cs->rec[r0 + W::lane()] = r; cs->pos[r0 + W::lane()] = r.opos; }
        W::sync();
        copy_span<W>(cs, lits, st, r0, n, out_pos, B1);
        nrec_total += n; out_pos = B1;
    }
    LZ_HDM void drain() {}
    LZ_HDM void resync(u32 apos) { copy_resync<W>(cs, st, apos); out_pos = apos; }
};

// ---- the parser --------------------------------------------------------------------------------------------------------
// Same token walk as decode_tokens_lz4 / decode_tokens_lizv1 (decode.cuh), with the literals stream read through `lv` and the
// copies replaced by records.  A batch that the batch logic cannot take (a check fails, a length field is cut off) goes
// down the first generation's serial path after a drain, which reproduces the reference's verdict and error code.
template <class LV> LZ_HD bool ext_field_lv(const LV& lv, long nl, long p, u32* v, u32* size)
{
    if (p >= nl) return false;
    const u32 b = lv.byte(p);
    if (b < 254) { *v = b; *size = 1; return true; }
    const u32 sz = b == 254 ? 3u : 4u;
    if (p + (long)sz > nl) return false;
    *v = b == 254 ? (lv.byte(p + 1) | (lv.byte(p + 2) << 8)) : (lv.byte(p + 1) | (lv.byte(p + 2) << 8) | (lv.byte(p + 3) << 16));
    *size = sz;
    return true;
}
// `limit`: fields that begin more than this many bytes behind lp are not read (the staged part of the stream ends there);
// their tokens get positions far outside every bound, so that the caller sees them as "do not fit this batch".
template <class W, class LV> LZ_HD bool ext_chain_lv(const LV& lv, long nl, long lp, u32 npend, const u32* ent, u32* epre,
                                                     u32 lbias, long room, u32 gap, u32 limit, u32* total)
{
    u32 E = 0;
    for (u32 j = 0; j < npend; ++j) {
        const u32 e = ent[j];
        const long base = lp + (long)(e & 0xffffu) + (long)E;
        if (base - lp > (long)limit) {
            if (W::lane() == 0) for (u32 i = j; i < npend; ++i) epre[i] = 0x40000000u;
            *total = 0x40000000u;
            return true;
        }
        if (W::lane() == 0) epre[j] = E;
        long pm;
        if (e & (1u << 24)) {
            u32 v, sz;
            if (base > nl - room || !ext_field_lv(lv, nl, base, &v, &sz)) return false;
            E += lbias + v + (sz - 1);
            pm = base + (long)sz + (long)(lbias + v) + (long)gap;
        } else pm = base + (long)((e >> 16) & 255u) + (long)gap;
        if (e & (1u << 25)) {
            u32 v, sz;
            if (pm > nl - room || !ext_field_lv(lv, nl, pm, &v, &sz)) return false;
            E += sz - 1;
        }
    }
    *total = E;
    return true;
}

// the block's last literals (and any other literal-only stretch): records of at most span() bytes each
template <class W, class SK> LZ_HD void publish_literals(SK& sk, u32 apos, long lp, u32 n)
{
    while (n) {
        const u32 part = n < sk.span() ? n : sk.span();
        sk.require(lp);
        OutRec r; r.opos = apos; r.lsrc = (u32)lp; r.mdst = apos + part; r.off = 0;
        sk.publish(W::lane() == 0, r, 1, apos + part, (u32)lp);
        apos += part; lp += part; n -= part;
    }
}

// the first generation's serial token loops, out of line (a batch reaches them only when a check fails or a run is longer
// than a batch may cover): the parser's hot loop should not share its instruction-cache footprint with them
template <class W> LZ_HD_COLD int lz4_serial_cold(const Streams& s, u8* dst, long oend, TokCursor& c, u32 count) { return lz4_serial<W>(s, dst, oend, c, count); }
template <class W> LZ_HD_COLD int lizv1_serial_cold(const Streams& s, u8* dst, long oend, TokCursor& c, u32 count) { return lizv1_serial<W>(s, dst, oend, c, count); }
template <class W> LZ_HD_COLD int read_stream_cold(bool huff, const u8* src, long csize, long& ip, u8* scratch, const u8** ptr, u32* len,
                                                   DecWarpCore* sh, const u8* expanded)
{
    return read_stream<W>(huff, src, csize, ip, scratch, ptr, len, sh, expanded);
}

template <class W, class LV, class SK>
LZ_HD int parse_tokens_lz4(const Streams& s, const LV& lv, SK& sk, u8* dst, u32 op0, u32 oend_u, u32 skew, DecWarpCore* sh)
{
    const long nl = (long)s.nlits, oend = (long)oend_u;
    const u32 NL = W::lanes(), lane = W::lane();
    if (oend_u - op0 == 0) return (s.nflags == 1 && s.flags[0] == 0) ? 0 : -1;
    TokCursor c; c.fp = 0; c.lp = 0; c.op = op0; c.p16 = c.p24 = 0; c.last_off = 0;
    while (c.fp < s.nflags) {
        const u32 nb = s.nflags - c.fp < NL ? s.nflags - c.fp : NL;
        const bool act = lane < nb;
        sk.require(c.lp);
        if ((c.fp & 127u) < NL && c.fp + 256 + 4 * lane < s.nflags) W::prefetch(s.flags + c.fp + 256 + 4 * lane);
        const u32 tok = act ? s.flags[c.fp + lane] : 0;
        const u32 litn = tok & 15, mln = tok >> 4;
        const bool need = act && litn == 15, needm = act && mln == 15;
        const u32 adv = act ? ((need ? 1 : litn) + 2 + (needm ? 1 : 0)) : 0;
        u32 tot_adv = 0;
        const u32 A = W::excl_scan(adv, &tot_adv);
        const u32 pendmask = W::ballot(need || needm);
        const u32 npend = popc32(pendmask), myidx = popc32(pendmask & ((1u << lane) - 1));
        if (need || needm) sh->chain.ent[myidx] = A | (litn << 16) | (need ? 1u << 24 : 0u) | (needm ? 1u << 25 : 0u);
        W::sync();
        u32 tot_ext = 0;
        bool slow = !ext_chain_lv<W>(lv, nl, c.lp, npend, sh->chain.ent, sh->chain.epre, 15, 5, 2, sk.span(), &tot_ext);
        bool taken = false;
        if (!slow) {
            W::sync();
            const long tokpos = c.lp + (long)A + (long)(myidx < npend ? sh->chain.epre[myidx] : tot_ext);
            u32 my_lit = 0, my_lx = 0, my_mlv = 0, my_mx = 0;
            // a batch longer than span() may have walked off the staged part of the stream: only fields inside it are read
            const bool near = tokpos - c.lp <= (long)sk.span();
            if (need && near) { u32 v = 0, sz = 1; ext_field_lv(lv, nl, tokpos, &v, &sz); my_lit = 15 + v; my_lx = sz; }
            const long mpos = tokpos + (need ? (long)(my_lx + my_lit) : (long)litn) + 2;
            const bool nearm = near && mpos - c.lp <= (long)sk.span();
            if (needm && nearm) { u32 v = 0, sz = 1; ext_field_lv(lv, nl, mpos, &v, &sz); my_mlv = v; my_mx = sz; }
            const u32 lit_len = need ? my_lit : (act ? litn : 0);
            const long lit_src = tokpos + (need ? (long)my_lx : 0);
            const long off_pos = lit_src + lit_len;
            const long tok_end = off_pos + 2 + (needm ? (long)my_mx : 0);       // first stream byte behind this token
            const bool fits = act && near && nearm && tok_end - c.lp <= (long)sk.span();
            const u32 fitmask = W::ballot(fits);
            u32 nfit = 0;                                                        // tokens of the leading run that fit
            { const u32 inv = ~fitmask; nfit = inv ? ctz32(inv) : 32u; if (nfit > nb) nfit = nb; }
            const bool mine = lane < nfit;
            bool bad = false;
            u32 ml = 0, off = 0;
            if (mine) {
                if (lit_src + (long)lit_len > nl - 18) bad = true;
                else {
                    off = lv.byte(off_pos) | (lv.byte(off_pos + 1) << 8);
                    ml = (needm ? 15 + my_mlv : mln) + kMinMatch;
                }
            }
            u32 tot_out = 0;
            const u32 O = W::excl_scan((mine && !bad) ? lit_len + ml : 0, &tot_out);
            const long opos = c.op + (long)O;
            if (mine && !bad) {
                if (opos + (long)lit_len > oend - 16) bad = true;
                else if ((long)off > opos + (long)lit_len) bad = true;
                else if (opos + (long)lit_len + (long)ml > oend - 16) bad = true;
            }
            if (nfit > 0 && W::ballot(bad) == 0) {
                OutRec r;
                r.opos = (u32)opos + skew; r.lsrc = (u32)lit_src; r.mdst = (u32)opos + lit_len + skew; r.off = off;
                const long lp_end = (long)W::shfl((u32)tok_end, nfit - 1);      // first stream byte behind the last token taken
                sk.publish(mine, r, nfit, (u32)(c.op + (long)tot_out) + skew, (u32)c.lp);
                c.fp += nfit; c.lp = lp_end; c.op += (long)tot_out;
                taken = true;
            }
        }
        if (!taken) {            // the reference's loop, one token at a time, on the output itself
            sk.drain();
            const int e = lz4_serial_cold<W>(s, dst, oend, c, slow ? nb : 1);
            if (e < 0) return e;
            sk.resync((u32)c.op + skew);
        }
    }
    const long rest = nl - c.lp;
    if (rest < 0 || c.op + rest > oend) return -(int)c.fp - 1;
    publish_literals<W>(sk, (u32)c.op + skew, c.lp, (u32)rest);
    c.op += rest;
    return (int)(c.op - (long)op0);
}

template <class W, class LV, class SK>
LZ_HD int parse_tokens_lizv1(const Streams& s, const LV& lv, SK& sk, u8* dst, u32 op0, u32 oend_u, u32 skew, DecWarpCore* sh)
{
    const long nl = (long)s.nlits, oend = (long)oend_u;
    const u32 NL = W::lanes(), lane = W::lane();
    if (oend_u - op0 == 0) return (s.nflags == 1 && s.flags[0] == 0) ? 0 : -1;
    TokCursor c; c.fp = 0; c.lp = 0; c.op = op0; c.p16 = c.p24 = 0; c.last_off = 0;
    while (c.fp < s.nflags) {
        const u32 nb = s.nflags - c.fp < NL ? s.nflags - c.fp : NL;
        const bool act = lane < nb;
        sk.require(c.lp);
        if ((c.fp & 127u) < NL && c.fp + 256 + 4 * lane < s.nflags) W::prefetch(s.flags + c.fp + 256 + 4 * lane);
        if ((c.fp & 63u) < NL && c.p16 + 512 + 8 * lane < s.noff16) W::prefetch(s.off16 + c.p16 + 512 + 8 * lane);
        const u32 tok = act ? s.flags[c.fp + lane] : 32;
        const bool shortf = tok >= 32;
        const u32 litn = shortf ? (tok & 7) : 0;
        const u32 mln = shortf ? ((tok >> 3) & 15) : tok;
        const bool need = act && shortf && litn == 7;
        const bool mlext = act && ((shortf && mln == 15) || (!shortf && tok == kLastLongOff));
        const bool new16 = act && shortf && (tok >> 7) == 0;
        const u32 adv = act ? ((need ? 1 : litn) + (mlext ? 1 : 0)) : 0;
        u32 tot_adv = 0, tot16 = 0, tot24 = 0;
        const u32 A = W::excl_scan(adv, &tot_adv);
        const u32 P16 = W::excl_scan(new16 ? 2u : 0u, &tot16);
        const u32 P24 = W::excl_scan((act && !shortf) ? 3u : 0u, &tot24);
        const u32 pendmask = W::ballot(need || mlext);
        const u32 npend = popc32(pendmask), myidx = popc32(pendmask & ((1u << lane) - 1));
        if (need || mlext) sh->chain.ent[myidx] = A | (litn << 16) | (need ? 1u << 24 : 0u) | (mlext ? 1u << 25 : 0u);
        W::sync();
        u32 tot_ext = 0;
        bool slow = !ext_chain_lv<W>(lv, nl, c.lp, npend, sh->chain.ent, sh->chain.epre, 7, 1, 0, sk.span(), &tot_ext);
        bool taken = false;
        if (!slow) {
            W::sync();
            const long tokpos = c.lp + (long)A + (long)(myidx < npend ? sh->chain.epre[myidx] : tot_ext);
            u32 my_lit = 0, my_lx = 0, my_mlv = 0, my_mx = 0;
            const bool near = tokpos - c.lp <= (long)sk.span();
            if (need && near) { u32 v = 0, sz = 1; ext_field_lv(lv, nl, tokpos, &v, &sz); my_lit = 7 + v; my_lx = sz; }
            const long mpos = tokpos + (need ? (long)(my_lx + my_lit) : (long)litn);
            const bool nearm = near && mpos - c.lp <= (long)sk.span();
            if (mlext && nearm) { u32 v = 0, sz = 1; ext_field_lv(lv, nl, mpos, &v, &sz); my_mlv = v; my_mx = sz; }
            const u32 lit_len = need ? my_lit : (act ? litn : 0);
            const long lit_src = tokpos + (need ? (long)my_lx : 0);
            const long tok_end = lit_src + (long)lit_len + (mlext ? (long)my_mx : 0);
            const bool fits = act && near && nearm && tok_end - c.lp <= (long)sk.span();
            const u32 fitmask = W::ballot(fits);
            u32 nfit = 0;
            { const u32 inv = ~fitmask; nfit = inv ? ctz32(inv) : 32u; if (nfit > nb) nfit = nb; }
            const bool mine = lane < nfit;
            bool bad = false;
            u32 ml = 0, off = 0;
            if (mine) {
                if (shortf) {
                    if (lit_src > nl - 16 || lit_src + (long)lit_len > nl) bad = true;
                    else if (c.p16 + P16 + (new16 ? 2u : 0u) > s.noff16) bad = true;
                    else {
                        if (new16) off = rd_le16(s.off16 + c.p16 + P16);
                        ml = mlext ? 15 + my_mlv : mln;
                    }
                } else {
                    ml = (tok == kLastLongOff) ? my_mlv + kLastLongOff + kMmLongOff : tok + kMmLongOff;
                    if ((long)(c.p24 + P24) > (long)s.noff24 - 3) bad = true;
                    else off = rd_le24(s.off24 + c.p24 + P24);
                }
            }
            // repeat-offset tokens take the offset of the closest earlier token that carried one
            const bool has_off = mine && (new16 || !shortf);
            const u32 carriers = W::ballot(has_off);
            const u32 before = carriers & ((lane == 0) ? 0u : (0xffffffffu >> (32 - lane)));
            const u32 src_lane = before ? highbit32(before) : lane;
            const u32 inherited = W::shfl(off, src_lane);
            if (mine && !has_off) off = before ? inherited : c.last_off;
            u32 tot_out = 0;
            const u32 O = W::excl_scan((mine && !bad) ? lit_len + ml : 0, &tot_out);
            const long opos = c.op + (long)O;
            if (mine && !bad) {
                if (shortf && opos + (long)lit_len > oend - 16) bad = true;
                else if ((long)off > opos + (long)lit_len) bad = true;
                else if (opos + (long)lit_len + (long)ml > oend - 16) bad = true;
            }
            if (nfit > 0 && W::ballot(bad) == 0) {
                OutRec r;
                r.opos = (u32)opos + skew; r.lsrc = (u32)lit_src; r.mdst = (u32)opos + lit_len + skew; r.off = off;
                const bool all = nfit == nb;
                const long lp_end = (long)W::shfl((u32)tok_end, nfit - 1);
                const u32 n16 = all ? tot16 : W::shfl(P16, nfit & (NL - 1));
                const u32 n24 = all ? tot24 : W::shfl(P24, nfit & (NL - 1));
                const u32 last = W::shfl(off, nfit - 1);
                sk.publish(mine, r, nfit, (u32)(c.op + (long)tot_out) + skew, (u32)c.lp);
                c.fp += nfit; c.lp = lp_end; c.op += (long)tot_out;
                c.p16 += n16; c.p24 += n24; c.last_off = last;
                taken = true;
            }
        }
        if (!taken) {
            sk.drain();
            const int e = lizv1_serial_cold<W>(s, dst, oend, c, slow ? nb : 1);
            if (e < 0) return e;
            sk.resync((u32)c.op + skew);
        }
    }
    const long rest = nl - c.lp;
    if (rest < 0 || c.op + rest > oend) return -(int)c.fp - 1;
    publish_literals<W>(sk, (u32)c.op + skew, c.lp, (u32)rest);
    c.op += rest;
    return (int)(c.op - (long)op0);
}

// Lizard_decompress_safe for one unit (decode_unit of decode.cuh with the token loops above); every lane of the parser
// returns the same value.  `dst` is the unit's real output pointer; the sink was set up for this unit by the caller.
template <class W, class SK>
LZ_HD int decode_unit2(const u8* src, u32 csize_u, u8* dst, u32 cap, u8* scratch, DecWarpCore* sh, SK& sk,
                       const UnitPre* up = nullptr, const u8* arena = nullptr)
{
    const long csize = (long)csize_u;
    if (csize < 1) return 0;
    const int level = src[0];
    if (level < (int)kMinLevel || level > (int)kMaxLevel) return -1;
    const bool lizv1 = level_is_lizv1(level);
    const u32 skew = (u32)((size_t)dst & 15);
    long ip = 1;
    long op = 0;
    while (ip < csize) {
        const long ip0 = ip;
        const u32 hdr = src[ip++];
        if (hdr == kFlagRaw) {
            if (ip > csize - 3) return -1;
            const u32 len = rd_le24(src + ip); ip += 3;
            if (ip + (long)len > csize || op + (long)len > (long)cap) return -1;
            sk.drain();                                                  // stored block: the parser copies it itself
            if (len >= kWideMinBytes) lanes_copy_wide<W>(dst + op, src + ip, len, false);
            else lanes_copy<W>(dst + op, src + ip, len);
            W::sync();
            op += len; ip += len;
            sk.resync((u32)op + skew);
            continue;
        }
        if (hdr & kFlagLen) return -1;
        if (ip > csize - 15) return -1;
        {
            const long len_len = (long)rd_le24(src + ip);
            const long len_end = ip + 3 + len_len;
            if (len_end > csize - 3) return -1;
            ip = len_end;
        }
        Streams s;
        s.src_begin = src; s.src_end = src + csize;
        if (!read_stream_cold<W>(hdr & kFlagOff16, src, csize, ip, scratch + 3 * kDecStreamScratch, &s.off16, &s.noff16, sh, nullptr)) return -1;
        if (!read_stream_cold<W>(hdr & kFlagOff24, src, csize, ip, scratch + 2 * kDecStreamScratch, &s.off24, &s.noff24, sh, nullptr)) return -1;
        const bool first = up != nullptr && ip0 == 1;
        const u8* const pre_flags = (first && up->state[kSlotFlags] == kPreDone) ? arena + up->off[kSlotFlags] : nullptr;
        const u8* const pre_lits = (first && up->state[kSlotLiterals] == kPreDone) ? arena + up->off[kSlotLiterals] : nullptr;
        if (!read_stream_cold<W>(hdr & kFlagFlags, src, csize, ip, scratch + 1 * kDecStreamScratch, &s.flags, &s.nflags, sh, pre_flags)) return -1;
        if (!read_stream_cold<W>(hdr & kFlagLiterals, src, csize, ip, scratch, &s.lits, &s.nlits, sh, pre_lits)) return -1;
        if (ip > csize) return -1;
        sk.drain();                                                      // the previous inner block's stream is still being read
        sk.set_stream(s.lits, s.nlits);
        const int res = lizv1 ? parse_tokens_lizv1<W>(s, sk.view(), sk, dst, (u32)op, cap, skew, sh)
                              : parse_tokens_lz4<W>(s, sk.view(), sk, dst, (u32)op, cap, skew, sh);
        if (res <= 0) { sk.drain(); return res; }
        op += res;
    }
    sk.drain();
    return (int)op;
}

// =====================================================================================================================
// Device plumbing between the two warps of a pair.
//   * literals ring: kStages stages of kLitStage bytes; stage s of the current stream (aligned-space bytes [s*kLitStage, ..))
//     is fetched by ONE cp.async.bulk (TMA 1-D bulk copy, global -> shared) that completes on the stage's mbarrier
//     (expect_tx / complete_tx).  The parser issues loads as far ahead as the ring allows; a slot is reused once the copier
//     has finished every batch whose literals begin in it.  Both warps address the ring as (stream position + skew) & mask.
//   * batches: kBatchSlots message slots, each with a `pub` mbarrier (parser -> copier) and a `free` mbarrier
//     (copier -> parser): the classic full/empty pipeline, one arrival per phase.
// All waits are mbarrier try_wait loops (the hardware parks the warp) with a watchdog that traps instead of hanging.
// =====================================================================================================================
#if defined(__CUDACC__)
enum : u32 { kLitStage = 2048 };
enum : u32 { kMsgData = 0, kMsgSync = 1, kMsgExit = 2 };

struct PairMsg { u32 kind, r0, nrec, B0, B1, ls; u8* dst_al; u32 unit_lo, apos; };

template <u32 kStages> struct PairShared {
    CopyShared cs;
    alignas(16) u8 ring[kStages * kLitStage];
    alignas(8) unsigned long long full_bar[kStages];
    unsigned long long pub_bar[kBatchSlots], free_bar[kBatchSlots];
    PairMsg msg[kBatchSlots];
    DecWarpCore dws;
};

__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, u32 parity)
{
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, u32 parity)
{
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity)
{
    if (mbar_try(bar, parity)) return;
    // the partner is behind: wait without taking its issue slots (try_wait parks the warp for a while by itself; the sleep
    // keeps the retry rate down), and trap instead of hanging if a protocol bug ever leaves nothing to wait for
    for (u32 spins = 1;; ++spins) {
        __nanosleep(64);
        if (mbar_try(bar, parity)) return;
        if ((spins & 0x3fffffu) == 0) __trap();               // ~4 M retries of >= 64 ns: seconds
    }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, u32 bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

template <u32 kStages> struct LitRingView {
    static constexpr u32 kMask = kStages * kLitStage - 1;
    const u8* ring; u32 ls;
    __device__ __forceinline__ u32 byte(long pos) const { return ring[((u32)pos + ls) & kMask]; }
    __device__ __forceinline__ Raw16 load(int a, u32, u32) const
    {
        const u32 x = (u32)a + ls;                       // wraps consistently for the (masked-out) bytes in front of a run
        const u32 a0 = (x & ~15u) & kMask;
        Raw16 r; r.v0 = tile_load(ring, a0); r.v1 = tile_load(ring, (a0 + 16) & kMask); r.delta = x & 15u;
        return r;
    }
};

// parser side of the pair
template <u32 kStages> struct PairSink {
    typedef LitRingView<kStages> View;
    static constexpr u32 kSpan = (kStages >= 8 ? kStages - 3 : 2) * kLitStage - 64;
    PairShared<kStages>* ps;
    View lv;
    u32 msg_count, acked, nrec_total, out_pos;
    u32 lit_stage;                    // one byte per message slot: stage in which a DATA message's literals begin, 0xff otherwise
                                      // (a 128 KiB stream has at most 65 stages)
    const u8* g_al; u32 nstages, abytes, issued, waited, q_base;
    u8* dst_al; u32 unit_lo;

    __device__ __forceinline__ void init(PairShared<kStages>* p)
    {
        ps = p; lv.ring = p->ring; lv.ls = 0;
        msg_count = acked = nrec_total = out_pos = 0;
        lit_stage = 0xffffffffu;
        g_al = nullptr; nstages = abytes = issued = waited = q_base = 0;
        dst_al = nullptr; unit_lo = 0;
    }
    __device__ __forceinline__ u32 span() const { return kSpan; }
    __device__ __forceinline__ const View& view() const { return lv; }

    __device__ __forceinline__ void poll_acks()
    {
        while (acked < msg_count && mbar_test(&ps->free_bar[acked % kBatchSlots], (acked / kBatchSlots) & 1u)) ++acked;
    }
    __device__ __forceinline__ void wait_ack()            // at least one more message consumed
    {
        mbar_wait(&ps->free_bar[acked % kBatchSlots], (acked / kBatchSlots) & 1u);
        ++acked;
    }
    __device__ __forceinline__ void drain() { while (acked < msg_count) wait_ack(); }

    // stages below this one are no longer needed by anybody
    __device__ __forceinline__ u32 free_stage(long lp) const
    {
        u32 f = ((u32)lp + lv.ls) / kLitStage;
        for (u32 m = acked; m < msg_count; ++m) { const u32 t = (lit_stage >> (8 * (m % kBatchSlots))) & 0xffu; f = t < f ? t : f; }
        return f;
    }
    __device__ __forceinline__ void issue(u32 s)
    {
        if ((threadIdx.x & 31) == 0) {
            const u32 slot = s % kStages;
            const u32 off = s * kLitStage;
            const u32 bytes = abytes - off < kLitStage ? abytes - off : kLitStage;
            mbar_expect_tx(&ps->full_bar[slot], bytes);
            bulk_load(ps->ring + slot * kLitStage, g_al + off, bytes, &ps->full_bar[slot]);
        }
    }
    __device__ __forceinline__ void wait_stage(u32 s)
    {
        mbar_wait(&ps->full_bar[s % kStages], ((q_base + s) / kStages) & 1u);
    }
    // literals stream readable on [lp, lp + span + 32) (or to its end); loads are issued as far ahead as the ring allows
    __device__ __forceinline__ void require(long lp)
    {
        if (nstages == 0) return;
        u32 need = ((u32)lp + lv.ls + kSpan + 32u) / kLitStage;
        if (need > nstages - 1) need = nstages - 1;
        for (;;) {
            poll_acks();
            const u32 fs = free_stage(lp);
            // a slot takes its next stage only when the parser has seen the previous one land (one phase per mbarrier at a time)
            // and nobody needs the previous one any more
            while (issued < nstages && issued < fs + kStages && issued < waited + kStages) issue(issued++);
            if (waited > need) break;
            if (waited < issued) { wait_stage(waited); ++waited; continue; }
            if (acked == msg_count) __trap();             // nothing in flight, nothing to wait for: span() is too large
            wait_ack();
        }
    }
    // every issued load has landed; the global load counter is brought to a multiple of kStages so that the next stream's
    // stage s again lives in slot s % kStages
    __device__ __forceinline__ void finish_stream()
    {
        while (waited < issued) { wait_stage(waited); ++waited; }
        u32 q = q_base + issued;
        __syncwarp();
        while (q % kStages) {
            if ((threadIdx.x & 31) == 0) mbar_arrive(&ps->full_bar[q % kStages]);      // empty phase: completes at once
            ++q;
        }
        __syncwarp();
        q_base = q; nstages = issued = waited = 0;
    }
    __device__ __forceinline__ void set_stream(const u8* p, u32 n)
    {
        finish_stream();
        // the stream may have been written by this warp (in-kernel Huffman expansion): order those generic-proxy writes
        // before the async-proxy reads of the bulk copies
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        const u32 sk = (u32)((size_t)p & 15);
        g_al = p - sk; lv.ls = sk;
        abytes = (sk + n + 15u) & ~15u;
        nstages = (abytes + kLitStage - 1) / kLitStage;
    }
    __device__ __forceinline__ void send(const PairMsg& m, u32 lit_st)
    {
        while (msg_count - acked > kBatchSlots - 1) wait_ack();      // a message slot (and its record rows) is reused only when consumed
        const u32 slot = msg_count % kBatchSlots;
        if ((threadIdx.x & 31) == 0) ps->msg[slot] = m;
        lit_stage = (lit_stage & ~(0xffu << (8 * slot))) | ((lit_st & 0xffu) << (8 * slot));
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&ps->pub_bar[slot]);
        ++msg_count;
    }
    __device__ __forceinline__ void publish(bool act, const OutRec& r, u32 n, u32 B1, u32 lp0)
    {
        while (msg_count - acked > kBatchSlots - 1) wait_ack();
        const u32 lane = threadIdx.x & 31;
        const u32 r0 = (msg_count % kBatchSlots) * kSlotRecs;
        // the whole position row is written by this warp before the batch is published: opos for the records, the batch's end
        // position for the end marker and the padding behind it
        { const u32 i = lane; ps->cs.pos[r0 + i] = act ? r.opos : B1; if (i < kSlotRecs - 32) ps->cs.pos[r0 + 32 + i] = B1; }
        if (act) ps->cs.rec[r0 + lane] = r;
        PairMsg m; m.kind = kMsgData; m.r0 = r0; m.nrec = n; m.B0 = out_pos; m.B1 = B1; m.ls = lv.ls;
        m.dst_al = dst_al; m.unit_lo = unit_lo; m.apos = 0;
        send(m, (lp0 + lv.ls) / kLitStage);
        nrec_total += n; out_pos = B1;
    }
    __device__ __forceinline__ void resync(u32 apos)
    {
        PairMsg m; m.kind = kMsgSync; m.r0 = m.nrec = m.B0 = m.B1 = m.ls = 0; m.dst_al = dst_al; m.unit_lo = unit_lo; m.apos = apos;
        send(m, 0xffffffffu);
        out_pos = apos;
    }
    __device__ __forceinline__ void begin_unit(u8* dst)
    {
        unit_lo = (u32)((size_t)dst & 15);
        dst_al = dst - unit_lo;
        resync(unit_lo);
    }
    __device__ __forceinline__ void exit_copier()
    {
        PairMsg m; m.kind = kMsgExit; m.r0 = m.nrec = m.B0 = m.B1 = m.ls = 0; m.dst_al = nullptr; m.unit_lo = m.apos = 0;
        send(m, 0xffffffffu);
    }
};

// the copier warp's life
template <u32 kStages> __device__ __forceinline__ void copier_loop(PairShared<kStages>* ps)
{
    CopyState st; st.dst_al = nullptr; st.unit_lo = 0;
    LitRingView<kStages> lv; lv.ring = ps->ring; lv.ls = 0;
    for (u32 m = 0;; ++m) {
        const u32 slot = m % kBatchSlots;
        mbar_wait(&ps->pub_bar[slot], (m / kBatchSlots) & 1u);
        const PairMsg g = ps->msg[slot];
        if (g.kind == kMsgExit) break;
        if (g.kind == kMsgSync) { st.dst_al = g.dst_al; st.unit_lo = g.unit_lo; copy_resync<WarpLanes>(&ps->cs, st, g.apos); }
        else { lv.ls = g.ls; copy_span<WarpLanes>(&ps->cs, lv, st, g.r0, g.nrec, g.B0, g.B1); }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&ps->free_bar[slot]);
    }
}
#endif  // __CUDACC__

}  // namespace lzb
