// entropy_enc.cuh -- serial (single-lane) pieces of the Huff0 encoder: length-limited code
// construction, the weight header (raw nibbles or FSE-compressed), and the size plan of the
// 4-segment payload.  The payload bit-packing itself is data-parallel and lives in encode.cuh.
//
// Everything here must match the reference to the bit, tie-breaks included:
//   lib/entropy/huf_compress.c:305-325   HUF_sort            (bucket by highbit(count+1), stable insertion)
//   lib/entropy/huf_compress.c:335-401   HUF_buildCTable_wksp (two-queue merge, canonical values by rank)
//   lib/entropy/huf_compress.c:223-297   HUF_setMaxHeight    (depth-limit repair)
//   lib/entropy/huf_compress.c:81-165    HUF_compressWeights / HUF_writeCTable
//   lib/entropy/fse_compress.c:477-641   FSE_optimalTableLog / FSE_normalizeCount / FSE_normalizeM2
//   lib/entropy/fse_compress.c:204-301   FSE_writeNCount
//   lib/entropy/fse_compress.c:103-182   FSE_buildCTable_wksp
//   lib/entropy/fse_compress.c:701-770   FSE_compress_usingCTable (2 interleaved states)
//   lib/entropy/huf_compress.c:517-574   HUF_compress_internal (accept / reject thresholds)
#pragma once
#include "common.cuh"
#include "entropy_dec.cuh"   // error codes

namespace lzb {

struct HufNode { u32 count; u16 parent; u8 byte; u8 nbits; };
struct HufCode { u16 val; u8 nbits; u8 pad; };

struct HufEncScratch {
    HufNode nodes[2 * 256 + 2];          // [0] is the sentinel in front of the sorted leaves
    HufCode codes[256];
    u8   weights[256];
    u8   header[256 + 8];                // serialized weight header
    // FSE over <= 13 weight symbols, tableLog <= 6
    u32  wcount[16];
    short wnorm[16];
    u16  fse_state[1u << kHufHeaderFseLog];
    u8   fse_spread[1u << kHufHeaderFseLog];
    int  fse_delta_state[16];
    u32  fse_delta_bits[16];
};

// ---- forward bit writer (lib/entropy/bitstream.h:185-248), capacity assumed ample -----------
struct BitWriter { u64 acc; u32 nbits; u8* ptr; u8* start; };
LZ_HD void bw_init(BitWriter& w, u8* dst) { w.acc = 0; w.nbits = 0; w.ptr = dst; w.start = dst; }
LZ_HD void bw_add(BitWriter& w, u64 v, u32 n) { w.acc |= (v & ((1ull << n) - 1)) << w.nbits; w.nbits += n; }
LZ_HD void bw_flush(BitWriter& w)
{
    u32 nb = w.nbits >> 3;
    for (u32 i = 0; i < 8; ++i) w.ptr[i] = (u8)(w.acc >> (8 * i));   // the reference stores the whole word too
    w.ptr += nb;
    w.nbits &= 7;
    w.acc = nb >= 8 ? 0 : w.acc >> (nb * 8);
}
LZ_HD u32 bw_close(BitWriter& w)
{
    bw_add(w, 1, 1);
    bw_flush(w);
    return (u32)(w.ptr - w.start) + (w.nbits > 0);
}

// ---- FSE table-log / normalization ---------------------------------------------------------------
LZ_HD u32 fse_min_table_log(u32 src_size, u32 max_sv)
{
    u32 a = highbit32(src_size - 1) + 1;
    u32 b = highbit32(max_sv) + 2;
    return a < b ? a : b;
}
LZ_HD u32 fse_optimal_table_log(u32 max_log, u32 src_size, u32 max_sv, u32 minus)
{
    u32 max_bits_src = highbit32(src_size - 1) - minus;
    u32 tl = max_log;
    u32 min_bits = fse_min_table_log(src_size, max_sv);
    if (tl == 0) tl = 11;
    if (max_bits_src < tl) tl = max_bits_src;
    if (min_bits > tl) tl = min_bits;
    if (tl < kFseMinTableLog) tl = kFseMinTableLog;
    if (tl > kFseMaxTableLog) tl = kFseMaxTableLog;
    return tl;
}

LZ_HD int fse_normalize_m2(short* norm, u32 tl, const u32* count, u32 total_in, u32 max_sv)
{
    u64 total = total_in;
    u32 distributed = 0;
    const u32 low_thresh = (u32)(total >> tl);
    u32 low_one = (u32)((total * 3) >> (tl + 1));
    for (u32 s = 0; s <= max_sv; ++s) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low_thresh) { norm[s] = -1; distributed++; total -= count[s]; continue; }
        if (count[s] <= low_one)    { norm[s] = 1;  distributed++; total -= count[s]; continue; }
        norm[s] = -2;
    }
    u32 to_dist = (1u << tl) - distributed;
    if ((total / to_dist) > low_one) {
        low_one = (u32)((total * 3) / (to_dist * 2));
        for (u32 s = 0; s <= max_sv; ++s)
            if (norm[s] == -2 && count[s] <= low_one) { norm[s] = 1; distributed++; total -= count[s]; }
        to_dist = (1u << tl) - distributed;
    }
    if (distributed == max_sv + 1) {
        u32 maxv = 0, maxc = 0;
        for (u32 s = 0; s <= max_sv; ++s) if (count[s] > maxc) { maxv = s; maxc = count[s]; }
        norm[maxv] += (short)to_dist;
        return 0;
    }
    {
        const u64 vstep_log = 62 - tl;
        const u64 mid = (1ull << (vstep_log - 1)) - 1;
        const u64 rstep = (((1ull << vstep_log) * to_dist) + mid) / total;
        u64 tmp = mid;
        for (u32 s = 0; s <= max_sv; ++s) {
            if (norm[s] == -2) {
                u64 end = tmp + (u64)count[s] * rstep;
                u32 s_start = (u32)(tmp >> vstep_log), s_end = (u32)(end >> vstep_log);
                u32 weight = s_end - s_start;
                if (weight < 1) return kErrGeneric;
                norm[s] = (short)weight;
                tmp = end;
            }
        }
    }
    return 0;
}

// returns table_log, 0 for the "one symbol fills everything" case, or negative
LZ_HD int fse_normalize_count(short* norm, u32 tl, const u32* count, u32 total, u32 max_sv)
{
    if (tl == 0) tl = 11;
    if (tl < kFseMinTableLog) return kErrGeneric;
    if (tl > kFseMaxTableLog) return kErrTableLog;
    if (tl < fse_min_table_log(total, max_sv)) return kErrGeneric;
    const u32 rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    const u64 scale = 62 - tl;
    const u64 step = (1ull << 62) / total;
    const u64 vstep = 1ull << (scale - 20);
    int still = 1 << tl;
    u32 largest = 0; short largest_p = 0;
    const u32 low_thresh = total >> tl;
    for (u32 s = 0; s <= max_sv; ++s) {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low_thresh) { norm[s] = -1; still--; }
        else {
            short proba = (short)(((u64)count[s] * step) >> scale);
            if (proba < 8) {
                u64 rest_to_beat = vstep * rtb[proba];
                proba += ((u64)count[s] * step) - ((u64)proba << scale) > rest_to_beat;
            }
            if (proba > largest_p) { largest_p = proba; largest = s; }
            norm[s] = proba;
            still -= proba;
        }
    }
    if (-still >= (norm[largest] >> 1)) {
        int e = fse_normalize_m2(norm, tl, count, total, max_sv);
        if (e < 0) return e;
    } else norm[largest] += (short)still;
    return (int)tl;
}

// FSE_writeNCount with a destination known to be large enough; returns bytes written or negative
LZ_HD int fse_write_ncount(u8* out0, const short* norm, u32 max_sv, u32 tl)
{
    if (tl > kFseMaxTableLog || tl < kFseMinTableLog) return kErrGeneric;
    u8* out = out0;
    const int table_size = 1 << tl;
    int nb = (int)tl + 1;
    int remaining = table_size + 1;
    int threshold = table_size;
    u32 stream = tl - kFseMinTableLog;
    int bit_count = 4;
    u32 sym = 0;
    bool prev_zero = false;
    while (remaining > 1) {
        if (prev_zero) {
            u32 start = sym;
            while (!norm[sym]) sym++;
            while (sym >= start + 24) {
                start += 24;
                stream += 0xFFFFu << bit_count;
                out[0] = (u8)stream; out[1] = (u8)(stream >> 8); out += 2;
                stream >>= 16;
            }
            while (sym >= start + 3) { start += 3; stream += 3u << bit_count; bit_count += 2; }
            stream += (sym - start) << bit_count;
            bit_count += 2;
            if (bit_count > 16) {
                out[0] = (u8)stream; out[1] = (u8)(stream >> 8); out += 2;
                stream >>= 16; bit_count -= 16;
            }
        }
        {
            int count = norm[sym++];
            const int max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += max;
            stream += (u32)count << bit_count;
            bit_count += nb;
            bit_count -= (count < max);
            prev_zero = (count == 1);
            if (remaining < 1) return kErrGeneric;
            while (remaining < threshold) { nb--; threshold >>= 1; }
        }
        if (bit_count > 16) {
            out[0] = (u8)stream; out[1] = (u8)(stream >> 8); out += 2;
            stream >>= 16; bit_count -= 16;
        }
    }
    out[0] = (u8)stream; out[1] = (u8)(stream >> 8);
    out += (bit_count + 7) / 8;
    if (sym > max_sv + 1) return kErrGeneric;
    return (int)(out - out0);
}

// HUF_compressWeights: FSE-compress the weight list.  0 = not compressible, 1 = single symbol,
// negative = error, else bytes written to dst.
LZ_HD int huf_compress_weights(u8* dst, const u8* w, u32 n, HufEncScratch* ws)
{
    if (n <= 1) return 0;
    u32 max_sv = kHufTableLogMax;
    for (u32 s = 0; s <= max_sv; ++s) ws->wcount[s] = 0;
    for (u32 i = 0; i < n; ++i) ws->wcount[w[i]]++;
    while (!ws->wcount[max_sv]) max_sv--;
    u32 max_count = 0;
    for (u32 s = 0; s <= max_sv; ++s) if (ws->wcount[s] > max_count) max_count = ws->wcount[s];
    if (max_count == n) return 1;
    if (max_count == 1) return 0;

    const u32 tl = fse_optimal_table_log(kHufHeaderFseLog, n, max_sv, 2);
    {   int e = fse_normalize_count(ws->wnorm, tl, ws->wcount, n, max_sv);
        if (e < 0) return e; }
    u8* op = dst;
    {   int h = fse_write_ncount(op, ws->wnorm, max_sv, tl);
        if (h < 0) return h;
        op += h; }

    // ---- compression table (FSE_buildCTable_wksp) ----
    const u32 size = 1u << tl, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 cumul[16];
    u32 high = size - 1;
    cumul[0] = 0;
    for (u32 u = 1; u <= max_sv + 1; ++u) {
        if (ws->wnorm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; ws->fse_spread[high--] = (u8)(u - 1); }
        else cumul[u] = cumul[u - 1] + (u32)ws->wnorm[u - 1];
    }
    cumul[max_sv + 1] = size + 1;
    {   u32 pos = 0;
        for (u32 s = 0; s <= max_sv; ++s)
            for (int i = 0; i < ws->wnorm[s]; ++i) {
                ws->fse_spread[pos] = (u8)s;
                do { pos = (pos + step) & mask; } while (pos > high);
            }
        if (pos != 0) return kErrGeneric; }
    for (u32 u = 0; u < size; ++u) { u32 s = ws->fse_spread[u]; ws->fse_state[cumul[s]++] = (u16)(size + u); }
    {   u32 total = 0;
        for (u32 s = 0; s <= max_sv; ++s) {
            int nc = ws->wnorm[s];
            if (nc == 0) continue;
            if (nc == -1 || nc == 1) {
                ws->fse_delta_bits[s] = (tl << 16) - (1u << tl);
                ws->fse_delta_state[s] = (int)total - 1;
                total++;
            } else {
                u32 max_bits_out = tl - highbit32((u32)nc - 1);
                u32 min_state_plus = (u32)nc << max_bits_out;
                ws->fse_delta_bits[s] = (max_bits_out << 16) - min_state_plus;
                ws->fse_delta_state[s] = (int)total - nc;
                total += (u32)nc;
            }
        } }

    // ---- encode, last symbol first, two alternating states (FSE_compress_usingCTable_generic) ----
    if (n <= 2) return 0;
    BitWriter bw; bw_init(bw, op);
    const u8* ip = w + n;
    u32 st1, st2;
#define LZB_FSE_INIT(ST, SYM) { u32 s_ = (SYM); u32 nbo = (ws->fse_delta_bits[s_] + (1u << 15)) >> 16; \
        u32 v = (nbo << 16) - ws->fse_delta_bits[s_]; ST = ws->fse_state[(int)(v >> nbo) + ws->fse_delta_state[s_]]; }
#define LZB_FSE_ENC(ST, SYM) { u32 s_ = (SYM); u32 nbo = (ST + ws->fse_delta_bits[s_]) >> 16; \
        bw_add(bw, ST, nbo); ST = ws->fse_state[(int)(ST >> nbo) + ws->fse_delta_state[s_]]; }
    u32 left = n;
    if (left & 1) {
        LZB_FSE_INIT(st1, *--ip) LZB_FSE_INIT(st2, *--ip) LZB_FSE_ENC(st1, *--ip)
        bw_flush(bw);
    } else {
        LZB_FSE_INIT(st2, *--ip) LZB_FSE_INIT(st1, *--ip)
    }
    left -= 2;
    if (left & 2) { LZB_FSE_ENC(st2, *--ip) LZB_FSE_ENC(st1, *--ip) bw_flush(bw); }
    while (ip > w) {
        LZB_FSE_ENC(st2, *--ip) LZB_FSE_ENC(st1, *--ip) LZB_FSE_ENC(st2, *--ip) LZB_FSE_ENC(st1, *--ip)
        bw_flush(bw);
    }
    bw_add(bw, st2, tl); bw_flush(bw);
    bw_add(bw, st1, tl); bw_flush(bw);
#undef LZB_FSE_INIT
#undef LZB_FSE_ENC
    u32 c = bw_close(bw);
    if (c == 0) return 0;
    op += c;
    return (int)(op - dst);
}

// HUF_writeCTable: header bytes for codes[0..max_sv], returns size or negative
LZ_HD int huf_write_ctable(u8* dst, const HufCode* codes, u32 max_sv, u32 huff_log, HufEncScratch* ws)
{
    u8 bits_to_weight[kHufTableLogMax + 2];
    bits_to_weight[0] = 0;
    for (u32 n = 1; n < huff_log + 1; ++n) bits_to_weight[n] = (u8)(huff_log + 1 - n);
    for (u32 n = 0; n < max_sv; ++n) ws->weights[n] = bits_to_weight[codes[n].nbits];
    {   int h = huf_compress_weights(dst + 1, ws->weights, max_sv, ws);
        if (h < 0) return h;
        if (h > 1 && (u32)h < max_sv / 2) { dst[0] = (u8)h; return h + 1; } }
    if (max_sv > 128) return kErrGeneric;
    dst[0] = (u8)(128 + (max_sv - 1));
    ws->weights[max_sv] = 0;
    for (u32 n = 0; n < max_sv; n += 2) dst[n / 2 + 1] = (u8)((ws->weights[n] << 4) + ws->weights[n + 1]);
    return (int)((max_sv + 1) / 2 + 1);
}

// ---- code construction --------------------------------------------------------------------------
LZ_HD void huf_sort(HufNode* node, const u32* count, u32 max_sv)
{
    u32 rank_base[32], rank_cur[32];
    for (u32 n = 0; n < 32; ++n) rank_base[n] = 0;
    for (u32 n = 0; n <= max_sv; ++n) rank_base[highbit32(count[n] + 1)]++;
    for (u32 n = 30; n > 0; --n) rank_base[n - 1] += rank_base[n];
    for (u32 n = 0; n < 32; ++n) rank_cur[n] = rank_base[n];
    for (u32 n = 0; n <= max_sv; ++n) {
        const u32 c = count[n];
        const u32 r = highbit32(c + 1) + 1;
        u32 pos = rank_cur[r]++;
        while (pos > rank_base[r] && c > node[pos - 1].count) { node[pos] = node[pos - 1]; pos--; }
        node[pos].count = c;
        node[pos].byte = (u8)n;
    }
}

LZ_HD u32 huf_set_max_height(HufNode* node, u32 last_non_null, u32 max_bits)
{
    const u32 largest = node[last_non_null].nbits;
    if (largest <= max_bits) return largest;
    int total_cost = 0;
    const u32 base_cost = 1u << (largest - max_bits);
    u32 n = last_non_null;
    while (node[n].nbits > max_bits) {
        total_cost += (int)(base_cost - (1u << (largest - node[n].nbits)));
        node[n].nbits = (u8)max_bits;
        n--;
    }
    while (node[n].nbits == max_bits) n--;
    total_cost >>= (largest - max_bits);

    const u32 none = 0xF0F0F0F0u;
    u32 rank_last[kHufTableLogMax + 2];
    for (u32 i = 0; i < kHufTableLogMax + 2; ++i) rank_last[i] = none;
    {   u32 cur_bits = max_bits;
        for (int pos = (int)n; pos >= 0; --pos) {
            if (node[pos].nbits >= cur_bits) continue;
            cur_bits = node[pos].nbits;
            rank_last[max_bits - cur_bits] = (u32)pos;
        } }
    while (total_cost > 0) {
        u32 dec = highbit32((u32)total_cost) + 1;
        for (; dec > 1; --dec) {
            u32 high_pos = rank_last[dec], low_pos = rank_last[dec - 1];
            if (high_pos == none) continue;
            if (low_pos == none) break;
            if (node[high_pos].count <= 2 * node[low_pos].count) break;
        }
        while (dec <= kHufTableLogMax && rank_last[dec] == none) dec++;
        total_cost -= 1 << (dec - 1);
        if (rank_last[dec - 1] == none) rank_last[dec - 1] = rank_last[dec];
        node[rank_last[dec]].nbits++;
        if (rank_last[dec] == 0) rank_last[dec] = none;
        else {
            rank_last[dec]--;
            if (node[rank_last[dec]].nbits != max_bits - dec) rank_last[dec] = none;
        }
    }
    while (total_cost < 0) {
        if (rank_last[1] == none) {
            while (node[n].nbits == max_bits) n--;
            node[n + 1].nbits--;
            rank_last[1] = n + 1;
            total_cost++;
            continue;
        }
        node[rank_last[1] + 1].nbits--;
        rank_last[1]++;
        total_cost++;
    }
    return max_bits;
}

// returns the final max code length, or negative
LZ_HD int huf_build_ctable(HufCode* codes, const u32* count, u32 max_sv, u32 max_bits, HufEncScratch* ws)
{
    HufNode* const node0 = ws->nodes;
    HufNode* const node = node0 + 1;
    const u32 kStart = 256;
    if (max_bits == 0) max_bits = kHufTableLogDefault;
    if (max_sv > 255) return kErrGeneric;
    for (u32 i = 0; i < 2 * 256 + 2; ++i) { node0[i].count = 0; node0[i].parent = 0; node0[i].byte = 0; node0[i].nbits = 0; }
    huf_sort(node, count, max_sv);

    u32 non_null = max_sv;
    while (node[non_null].count == 0) non_null--;
    int low_s = (int)non_null;
    u32 node_nb = kStart;
    const u32 root = node_nb + (u32)low_s - 1;
    int low_n = (int)node_nb;
    node[node_nb].count = node[low_s].count + node[low_s - 1].count;
    node[low_s].parent = node[low_s - 1].parent = (u16)node_nb;
    node_nb++; low_s -= 2;
    for (u32 n = node_nb; n <= root; ++n) node[n].count = 1u << 30;
    node0[0].count = 1u << 31;                   // sentinel below the leaves: node[-1]

    while (node_nb <= root) {
        u32 n1 = (node[low_s].count < node[low_n].count) ? (u32)low_s-- : (u32)low_n++;
        u32 n2 = (node[low_s].count < node[low_n].count) ? (u32)low_s-- : (u32)low_n++;
        node[node_nb].count = node[n1].count + node[n2].count;
        node[n1].parent = node[n2].parent = (u16)node_nb;
        node_nb++;
    }
    node[root].nbits = 0;
    for (u32 n = root - 1; n >= kStart; --n) node[n].nbits = (u8)(node[node[n].parent].nbits + 1);
    for (u32 n = 0; n <= non_null; ++n) node[n].nbits = (u8)(node[node[n].parent].nbits + 1);

    max_bits = huf_set_max_height(node, non_null, max_bits);
    if (max_bits > kHufTableLogMax) return kErrGeneric;

    u16 nb_per_rank[kHufTableLogMax + 1], val_per_rank[kHufTableLogMax + 1];
    for (u32 i = 0; i <= kHufTableLogMax; ++i) { nb_per_rank[i] = 0; val_per_rank[i] = 0; }
    for (u32 n = 0; n <= non_null; ++n) nb_per_rank[node[n].nbits]++;
    {   u16 min = 0;
        for (u32 n = max_bits; n > 0; --n) { val_per_rank[n] = min; min = (u16)(min + nb_per_rank[n]); min >>= 1; } }
    for (u32 n = 0; n <= max_sv; ++n) codes[node[n].byte].nbits = node[n].nbits;
    for (u32 n = 0; n <= max_sv; ++n) codes[n].val = val_per_rank[codes[n].nbits]++;
    return (int)max_bits;
}

// ---- plan of one Huffman-compressed stream ---------------------------------------------------------
struct HufPlan {
    int  status;         // kHufPlanRaw: keep stream uncompressed; kHufPlanRle; kHufPlanCoded
    u32  header_size;    // bytes of ws->header
    u32  seg_bytes[4];   // coded size of the 4 segments
    u32  total;          // header + 6 + segments  (what HUF_compress would return)
    u8   rle_byte;
};
enum : int { kHufPlanRaw = 0, kHufPlanRle = 1, kHufPlanCoded = 2 };

// Given the byte histogram of the whole stream and of its first three segments, decide what
// HUF_compress(dst, >= HUF_compressBound(n), src, n) returns and how big each piece is.
// seg_count[k][s] = occurrences of symbol s in segment k (k = 0..3).
LZ_HD void huf_plan(HufPlan& plan, const u32* count, const u32 (*seg_count)[256], u32 n, u8 first_byte, HufEncScratch* ws)
{
    plan.status = kHufPlanRaw; plan.header_size = 0; plan.total = 0; plan.rle_byte = first_byte;
    if (n == 0 || n > kHufBlockSizeMax) return;
    u32 max_sv = 255;
    while (!count[max_sv]) max_sv--;
    u32 largest = 0;
    for (u32 s = 0; s <= max_sv; ++s) if (count[s] > largest) largest = count[s];
    if (largest == n) { plan.status = kHufPlanRle; plan.total = 1; return; }
    if (largest <= (n >> 7) + 1) return;                         // not compressible enough
    u32 huff_log = fse_optimal_table_log(kHufTableLogDefault, n, max_sv, 1);
    {   int mb = huf_build_ctable(ws->codes, count, max_sv, huff_log, ws);
        if (mb < 0) return;                                      // HUF error -> stream stays raw
        huff_log = (u32)mb; }
    int h = huf_write_ctable(ws->header, ws->codes, max_sv, huff_log, ws);
    if (h < 0) return;
    if ((u32)h + 12 >= n) return;
    if (n < 12) return;
    u32 total = (u32)h + 6;
    for (int k = 0; k < 4; ++k) {
        u64 bits = 1;                                            // end mark
        for (u32 s = 0; s <= max_sv; ++s) bits += (u64)seg_count[k][s] * ws->codes[s].nbits;
        plan.seg_bytes[k] = (u32)((bits + 7) >> 3);
        total += plan.seg_bytes[k];
    }
    if (total >= n - 1) return;
    plan.status = kHufPlanCoded; plan.header_size = (u32)h; plan.total = total;
}

}  // namespace lzb
