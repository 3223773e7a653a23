// huf_expand.cuh -- Huffman pre-pass of the block decoder: expand the Huffman-coded literals / flags streams of MANY
// units at once, one lane per Huff0 segment.
//
// Inside the token kernel (decode.cuh) a warp owns one unit, and a Huffman-coded stream has only four independent
// bitstreams (HUF_compress4X, huf_compress.c:473-513): 4 of 32 lanes work while the other 28 execute the same
// instructions for nothing.  At levels 30-49 that phase was ~60 % of the kernel.  The pre-pass turns the loop around:
//   1. plan   -- one thread per unit walks the headers of the unit's FIRST inner block exactly as decode_unit /
//                read_stream do (lizard_decompress.c:115-264, :72-112) and appends one job per Huffman-coded literals /
//                flags stream: {payload, c, n, slot in the expansion arena};
//   2. expand -- a warp takes 8 jobs: lanes 0-7 read the weight headers and fill 8 single-symbol tables in shared memory
//                (HUF_readStats / HUF_readDTableX2; two-level form, HufCompact), then lane l decodes segment l%4 of
//                job l/4: 32 lanes busy;
//   3. the token kernel finds `state == kPreDone` for the stream and uses the expanded bytes from the arena.
// A job is only marked done when the stream decoded the regular way (all four bitstreams end exactly).  Anything else
// -- stored / RLE streams, damaged headers, tableLog 12, arena exhausted, later inner blocks of a multi-block unit --
// is left to the in-kernel path, which reproduces the reference's accept/reject verdicts, so results and error codes
// do not depend on whether the pre-pass ran.
#pragma once
#include "common.cuh"
#include "entropy_dec.cuh"

namespace lzb {

enum : u32 { kPreNone = 0, kPrePlanned = 1, kPreDone = 2 };
enum : u32 { kSlotLiterals = 0, kSlotFlags = 1 };

struct HufJob {
    u64 src;        // offset of the stream's payload (behind its 6-byte [n][c] header) in the batch's source arena
    u64 dst;        // offset of the expanded bytes in the expansion arena
    u32 c, n;       // compressed / expanded size
    u32 unit, slot;
};
struct UnitPre {            // per unit: where its pre-expanded streams are
    u64 off[2];
    u32 state[2];
};
struct PreHeader {          // first bytes of the pre-pass workspace, zeroed before every launch
    u32 count[2];           // jobs appended per slot (plan kernel)
    u32 next[2];            // jobs handed out per slot (expand kernel)
    unsigned long long cursor;   // bump allocator of the expansion arena
    u32 pad[2];
};

// ---- token pre-pass (decode.cuh: parse_first_block) ------------------------------------------------------------------
// The token loop of a block is a serial chain as well (a token's place in the literals stream depends on every earlier
// length-extension byte).  Inside the token kernel a warp resolves it with ~28 warp instructions per token; here ONE LANE
// walks one unit's first inner block the way the reference's loop does, all checks included, and writes one 16-byte
// sequence record per token.  The token kernel then only moves bytes: lane i of a batch loads record i.
struct UnitSeq {
    u64 off;            // first record of the unit in the sequence arena (in records)
    u32 nseq;           // records written = tokens of the first inner block
    u32 state;          // kPreDone when the whole block parsed cleanly
    u32 final_lp;       // literals-stream position behind the last token (the rest are the last literals)
    u32 final_op;       // output position behind the last match
    u32 pad[2];
};
struct SeqHeader { unsigned long long cursor; u32 pad[6]; };

// room a stream of n bytes takes in the arena: the token loops may look a few bytes past the end of a stream
LZ_HD u64 pre_slot_bytes(u32 n) { return ((u64)n + 64 + 127) & ~(u64)127; }

// Walks the first inner block of one unit; returns the number of jobs (0..2) written to out[].  Mirrors decode_unit():
// every early return there that happens before the stream in question is reached is an early return here.
LZ_HD u32 plan_unit(const u8* src, u32 csize_u, HufJob* out)
{
    const long csize = (long)csize_u;
    if (csize < 2) return 0;
    const int level = src[0];
    if (level < (int)kMinLevel || level > (int)kMaxLevel) return 0;
    long ip = 1;
    const u32 hdr = src[ip++];
    if (hdr == kFlagRaw || (hdr & kFlagLen)) return 0;
    if ((hdr & (kFlagLiterals | kFlagFlags)) == 0) return 0;
    if (ip > csize - 15) return 0;
    {
        const long len_end = ip + 3 + (long)rd_le24(src + ip);
        if (len_end > csize - 3) return 0;
        ip = len_end;
    }
    u32 nj = 0;
    // stream order in a block: off16, off24, flags, literals (lizard_decompress.c:214-226)
    const u32 bit[4] = { kFlagOff16, kFlagOff24, kFlagFlags, kFlagLiterals };
    for (int k = 0; k < 4; ++k) {
        if (hdr & bit[k]) {
            if (ip > csize - 6) return nj;
            const u32 n = rd_le24(src + ip), c = rd_le24(src + ip + 3);
            if (n > kBlockSize || ip + (long)c > csize - 6) return nj;
            if (k >= 2) {
                HufJob j; j.src = (u64)(ip + 6); j.dst = 0; j.c = c; j.n = n; j.unit = 0;
                j.slot = k == 3 ? kSlotLiterals : kSlotFlags;
                out[nj++] = j;
            }
            ip += (long)c + 6;
        } else {
            if (ip > csize - 3) return nj;
            ip += 3 + (long)rd_le24(src + ip);
        }
    }
    return nj;
}

struct HufJobScratch {      // per table-building lane; lives in global memory (cold, a few hundred bytes touched)
    HufStatsScratch stats;
    u8  weights[256];
    u32 rank[kHufTableLogMax + 1];
    u32 pad[3];
};

// Weight header -> two-level table (HufCompact, decode.cuh).  True when the stream is of the regular kind the
// pre-pass handles; *hdr_len = bytes of weight header.  Same conditions, in the same order, as the regular path of
// huf_decompress_lanes (decode.cuh) / HUF_decompress4X2 (huf_decompress.c:231-351).
LZ_HD bool huf_job_prepare(const u8* src, u32 c, u32 n, HufCompact* table, HufJobScratch* ws, u32* hdr_len)
{
    if (n == 0 || c >= n || c == 1) return false;       // error / stored / RLE: the in-kernel path deals with them
    u32 nsym = 0, tl = 0;
    const int h = huf_read_stats(ws->weights, ws->rank, &nsym, &tl, src, c, &ws->stats);
    if (h < 0 || (u32)h >= c || tl > 11) return false;
    const u32 pc = c - (u32)h;
    if (pc < 10) return false;
    const u8* pay = src + h;
    if (rd_le16(pay) + rd_le16(pay + 2) + rd_le16(pay + 4) + 6 > pc) return false;
    huf_fill_compact(table, ws->weights, ws->rank, nsym, tl);
    *hdr_len = (u32)h;
    return true;
}

}  // namespace lzb
