/* lizard_oracle.c -- TEST INFRASTRUCTURE: a plain-C restatement of the reference's block codec.
 *
 * This file is the checker, never the thing measured or shipped.  It restates, function by function,
 * what inikep/lizard (commit af8518cc) computes on the hot path, in the simplest serial form:
 *
 *   o_huf_decompress / o_huf_compress      lib/entropy/huf_decompress.c:832-845, huf_compress.c:517-612
 *   o_fse_*                                lib/entropy/fse_decompress.c, fse_compress.c, entropy_common.c
 *   o_decode_lz4 / o_decode_lizv1          lib/lizard_decompress_lz4.h:7-163, lizard_decompress_liz.h:14-220
 *   oracle_Lizard_decompress_safe          lib/lizard_decompress.c:115-270
 *   o_parse_fast                           lib/lizard_parser_fastsmall.h:34-189, lizard_parser_fast.h:41-196, lizard_parser_fastbig.h:35-175
 *   o_parse_pricefast                      lib/lizard_parser_pricefast.h:3-249
 *   o_emit_lz4 / o_emit_lizv1              lib/lizard_compress_lz4.h:3-86, lizard_compress_liz.h:43-179
 *   o_write_block / oracle_Lizard_compress lib/lizard_compress.c:141-250, 472-606
 *
 * Compression uses CLEAN-STATE semantics: the hash table is empty at the start of every call, which is
 * the reference built with -DLIZARD_RESET_MEM.  64-bit little-endian only.
 *
 * Parity: pinned by tests/test_oracle.py against oracle/_ref (the compiled reference) and tests/golden.
 */
#include "lizard_oracle.h"
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64;

#define O_BLOCK      (1u << 17)
#define O_BLOCK_PAD  (O_BLOCK + 32)
#define O_BIAS       (1u << 24)          /* LIZARD_DICT_SIZE: index of the first byte of a one-shot call */
#define O_ERR        ((size_t)-1)

static u32 rd16(const u8* p) { return p[0] | (p[1] << 8); }
static u32 rd24(const u8* p) { return p[0] | (p[1] << 8) | ((u32)p[2] << 16); }
static u32 rd32(const u8* p) { u32 v; memcpy(&v, p, 4); return v; }
static u64 rd64(const u8* p) { u64 v; memcpy(&v, p, 8); return v; }
static void wr16(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); }
static void wr24(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); }
static u32 hb32(u32 v) { return 31 - (u32)__builtin_clz(v); }

int oracle_Lizard_compressBound(int isize)
{   /* lib/lizard_compress.h:124 */
    return ((unsigned)isize > 0x7E000000u) ? 0 : isize + 1 + 1 + ((isize / (int)O_BLOCK) + 1) * 4;
}

/* =====================================================================================================
 * Backward bit reader (lib/entropy/bitstream.h:260-408)
 * =================================================================================================== */
typedef struct { const u8* base; const u8* at; u64 bits; u32 used; } o_bitin;
enum { O_MORE = 0, O_ENDBUF = 1, O_DONE = 2, O_OVER = 3 };

static int o_bitin_open(o_bitin* b, const u8* p, size_t n)
{
    memset(b, 0, sizeof *b);
    if (n < 1) return -1;
    b->base = p;
    if (n >= 8) { b->at = p + n - 8; b->bits = rd64(b->at); b->used = p[n - 1] ? 8 - hb32(p[n - 1]) : 0; }
    else {
        b->at = p;
        for (size_t i = 0; i < n; ++i) b->bits |= (u64)p[i] << (8 * i);
        b->used = (p[n - 1] ? 8 - hb32(p[n - 1]) : 0) + (u32)(8 - n) * 8;
    }
    return p[n - 1] ? 0 : -1;
}
static u64 o_peek(const o_bitin* b, u32 n) { return ((b->bits << (b->used & 63)) >> 1) >> ((63 - n) & 63); }
static u64 o_peek_fast(const o_bitin* b, u32 n) { return (b->bits << (b->used & 63)) >> ((64 - n) & 63); }
static u64 o_take(o_bitin* b, u32 n) { u64 v = o_peek(b, n); b->used += n; return v; }
static int o_refill(o_bitin* b)
{
    if (b->used > 64) return O_OVER;
    if (b->at >= b->base + 8) { b->at -= b->used >> 3; b->used &= 7; b->bits = rd64(b->at); return O_MORE; }
    if (b->at == b->base) return b->used < 64 ? O_ENDBUF : O_DONE;
    {   u32 nb = b->used >> 3; int st = O_MORE;
        if (b->at - nb < b->base) { nb = (u32)(b->at - b->base); st = O_ENDBUF; }
        b->at -= nb; b->used -= nb * 8; b->bits = rd64(b->at);
        return st; }
}
static int o_bitin_finished(const o_bitin* b) { return b->at == b->base && b->used == 64; }

/* =====================================================================================================
 * FSE: only what the Huffman weight header needs (alphabet <= 13 symbols, tableLog <= 6)
 * =================================================================================================== */
/* lib/entropy/entropy_common.c:71-160 */
static long o_fse_read_ncount(short* norm, u32* maxsv, u32* tlog, const u8* p, size_t n)
{
    const u8* const end = p + n; const u8* ip = p;
    if (n < 4) return -1;
    u32 bs = rd32(ip);
    int nb = (int)(bs & 15) + 5;
    if (nb > 15) return -1;
    bs >>= 4; int bc = 4;
    *tlog = (u32)nb;
    int remaining = (1 << nb) + 1, threshold = 1 << nb;
    nb++;
    u32 sym = 0; int prev0 = 0;
    while (remaining > 1 && sym <= *maxsv) {
        if (prev0) {
            u32 n0 = sym;
            while ((bs & 0xFFFF) == 0xFFFF) {
                n0 += 24;
                if (ip < end - 5) { ip += 2; bs = rd32(ip) >> bc; } else { bs >>= 16; bc += 16; }
            }
            while ((bs & 3) == 3) { n0 += 3; bs >>= 2; bc += 2; }
            n0 += bs & 3; bc += 2;
            if (n0 > *maxsv) return -1;
            while (sym < n0) norm[sym++] = 0;
            if (ip <= end - 7 || ip + (bc >> 3) <= end - 4) { ip += bc >> 3; bc &= 7; bs = rd32(ip) >> bc; }
            else bs >>= 2;
        }
        {   int max = (2 * threshold - 1) - remaining, cnt;
            if ((bs & (u32)(threshold - 1)) < (u32)max) { cnt = (int)(bs & (u32)(threshold - 1)); bc += nb - 1; }
            else { cnt = (int)(bs & (u32)(2 * threshold - 1)); if (cnt >= threshold) cnt -= max; bc += nb; }
            cnt--;
            remaining -= cnt < 0 ? -cnt : cnt;
            norm[sym++] = (short)cnt;
            prev0 = !cnt;
            while (remaining < threshold) { nb--; threshold >>= 1; }
            if (ip <= end - 7 || ip + (bc >> 3) <= end - 4) { ip += bc >> 3; bc &= 7; }
            else { bc -= (int)(8 * (end - 4 - ip)); ip = end - 4; }
            bs = rd32(ip) >> (bc & 31);
        }
    }
    if (remaining != 1 || bc > 32) return -1;
    *maxsv = sym - 1;
    ip += (bc + 7) >> 3;
    return ip - p;
}

typedef struct { u16 base; u8 sym; u8 nb; } o_fse_dcell;

/* lib/entropy/fse_decompress.c:113-168 */
static int o_fse_build_dtable(o_fse_dcell* t, const short* norm, u32 maxsv, u32 tlog)
{
    u16 next[256];
    u32 size = 1u << tlog, high = size - 1, mask = size - 1, step = (size >> 1) + (size >> 3) + 3, pos = 0;
    if (tlog > 12) return -1;
    for (u32 s = 0; s <= maxsv; ++s) {
        if (norm[s] == -1) { t[high--].sym = (u8)s; next[s] = 1; } else next[s] = (u16)norm[s];
    }
    for (u32 s = 0; s <= maxsv; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            t[pos].sym = (u8)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    if (pos) return -1;
    for (u32 u = 0; u < size; ++u) {
        u32 nx = next[t[u].sym]++;
        t[u].nb = (u8)(tlog - hb32(nx));
        t[u].base = (u16)((nx << t[u].nb) - size);
    }
    return 0;
}

/* lib/entropy/fse_decompress.c:220-294 (two states sharing one stream) */
static long o_fse_decode(u8* out, size_t cap, const u8* p, size_t n, const o_fse_dcell* t, u32 tlog)
{
    o_bitin b; u8* op = out; u8* const omax = out + cap; u8* const olimit = omax - 3;
    if (o_bitin_open(&b, p, n)) return -1;
    u32 s1 = (u32)o_take(&b, tlog); o_refill(&b);
    u32 s2 = (u32)o_take(&b, tlog); o_refill(&b);
#define O_FSE_SYM(S) (*op++ = t[S].sym, S = t[S].base + (u32)o_take(&b, t[S].nb))
    for (; (o_refill(&b) == O_MORE) & (op < olimit);) { O_FSE_SYM(s1); O_FSE_SYM(s2); O_FSE_SYM(s1); O_FSE_SYM(s2); }
    for (;;) {
        if (op > omax - 2) return -1;
        O_FSE_SYM(s1);
        if (o_refill(&b) == O_OVER) { O_FSE_SYM(s2); break; }
        if (op > omax - 2) return -1;
        O_FSE_SYM(s2);
        if (o_refill(&b) == O_OVER) { O_FSE_SYM(s1); break; }
    }
#undef O_FSE_SYM
    return op - out;
}

/* =====================================================================================================
 * Huff0 decoder
 * =================================================================================================== */
/* lib/entropy/entropy_common.c:170-231 : weights[], rank counts, table log; returns header size */
static long o_huf_read_stats(u8* w, u32* rank, u32* nsym, u32* tlog, const u8* p, size_t n)
{
    size_t isz, osz;
    if (!n) return -1;
    isz = p[0];
    if (isz >= 128) {
        osz = isz - 127; isz = (osz + 1) / 2;
        if (isz + 1 > n) return -1;
        for (size_t k = 0; k < osz; k += 2) { w[k] = p[1 + k / 2] >> 4; w[k + 1] = p[1 + k / 2] & 15; }
    } else {
        short norm[256]; o_fse_dcell cells[64]; u32 maxsv = 255, flog = 0; long h, r;
        if (isz + 1 > n) return -1;
        h = o_fse_read_ncount(norm, &maxsv, &flog, p + 1, isz);
        if (h < 0 || flog > 6) return -1;
        if (o_fse_build_dtable(cells, norm, maxsv, flog)) return -1;
        r = o_fse_decode(w, 255, p + 1 + h, isz - (size_t)h, cells, flog);
        if (r < 0) return -1;
        osz = (size_t)r;
    }
    memset(rank, 0, 13 * sizeof(u32));
    u32 total = 0;
    for (size_t k = 0; k < osz; ++k) { if (w[k] >= 12) return -1; rank[w[k]]++; total += (1u << w[k]) >> 1; }
    if (!total) return -1;
    u32 tl = hb32(total) + 1;
    if (tl > 12) return -1;
    *tlog = tl;
    {   u32 rest = (1u << tl) - total, h = hb32(rest);
        if ((1u << h) != rest) return -1;
        w[osz] = (u8)(h + 1); rank[h + 1]++; }
    if (rank[1] < 2 || (rank[1] & 1)) return -1;
    *nsym = (u32)osz + 1;
    return (long)(isz + 1);
}

typedef struct { u8 sym; u8 nb; } o_hcell;

/* the reference has a single-symbol decoder (X2) and a double-symbol one (X4); the choice is a speed
 * heuristic (huf_decompress.c:771-812) and both produce the same bytes on valid input, but X4 accepts a
 * few malformed tails X2 rejects (its last-symbol hack, :562-585).  The oracle keeps both behaviours on
 * one single-symbol table: an X4 step is "two symbols if both codes fit in 12 bits". */
static u32 o_huf_select(size_t dst, size_t src)
{
    static const u16 a[16][2] = {{0,0},{0,0},{38,130},{448,128},{556,128},{714,128},{883,128},{897,128},{926,128},{947,128},{1107,128},{1177,128},{1242,128},{1349,128},{1455,128},{722,128}};
    static const u16 b[16][2] = {{1,1},{1,1},{1313,74},{1353,74},{1353,74},{1418,74},{1437,74},{1515,75},{1613,75},{1729,77},{2083,81},{2379,87},{2415,93},{2644,106},{2422,124},{1891,145}};
    u32 q = (u32)(src * 16 / dst), d = (u32)(dst >> 8);
    u32 t0 = a[q][0] + a[q][1] * d, t1 = b[q][0] + b[q][1] * d;
    t1 += t1 >> 3;
    return t1 < t0;
}
static u8 o_h1(o_bitin* b, const o_hcell* t, u32 tl) { o_hcell c = t[o_peek_fast(b, tl)]; b->used += c.nb; return c.sym; }
static u32 o_h2_look(const o_bitin* b, const o_hcell* t, u32 tl, o_hcell* c1, o_hcell* c2)
{
    u32 v = (u32)o_peek_fast(b, 12);
    *c1 = t[v >> (12 - tl)];
    *c2 = t[((v << c1->nb) & 0xFFF) >> (12 - tl)];
    return (c1->nb + c2->nb <= 12) ? 2 : 1;
}
static u32 o_h2(o_bitin* b, const o_hcell* t, u32 tl, u8* op)
{
    o_hcell c1, c2; u32 len = o_h2_look(b, t, tl, &c1, &c2);
    op[0] = c1.sym; op[1] = len == 2 ? c2.sym : 0;
    b->used += len == 2 ? c1.nb + c2.nb : c1.nb;
    return len;
}
static void o_h2_last(o_bitin* b, const o_hcell* t, u32 tl, u8* op)
{
    o_hcell c1, c2; u32 len = o_h2_look(b, t, tl, &c1, &c2);
    op[0] = c1.sym;
    if (len == 1) b->used += c1.nb;
    else if (b->used < 64) { b->used += c1.nb + c2.nb; if (b->used > 64) b->used = 64; }
}
static void o_huf_tail(u8* d, long p, long e, o_bitin* b, const o_hcell* t, u32 tl, u32 x4)
{
    if (x4) {
        while ((o_refill(b) == O_MORE) & (p < e - 7)) { p += o_h2(b, t, tl, d + p); p += o_h2(b, t, tl, d + p); p += o_h2(b, t, tl, d + p); p += o_h2(b, t, tl, d + p); }
        while ((o_refill(b) == O_MORE) & (p <= e - 2)) p += o_h2(b, t, tl, d + p);
        while (p <= e - 2) p += o_h2(b, t, tl, d + p);
        if (p < e) o_h2_last(b, t, tl, d + p);
    } else {
        while ((o_refill(b) == O_MORE) && (p <= e - 4)) { d[p++] = o_h1(b, t, tl); d[p++] = o_h1(b, t, tl); d[p++] = o_h1(b, t, tl); d[p++] = o_h1(b, t, tl); }
        while ((o_refill(b) == O_MORE) && (p < e)) d[p++] = o_h1(b, t, tl);
        while (p < e) d[p++] = o_h1(b, t, tl);
    }
}

/* lib/entropy/huf_decompress.c:832-845 + 231-351 / 644-764.  NOTE: like the reference this may write up
 * to 2 bytes past dst+dstSize for dstSize in {1,2,5}; Lizard always calls it on a 128 KiB scratch. */
size_t oracle_HUF_decompress(void* dstv, size_t n, const void* srcv, size_t c)
{
    u8* d = (u8*)dstv; const u8* s = (const u8*)srcv;
    u8 w[257]; u32 rank[13], nsym = 0, tl = 0;
    static __thread o_hcell table[4096];
    if (!n || c > n) return O_ERR;
    if (c == n) { memcpy(d, s, n); return n; }
    if (c == 1) { memset(d, s[0], n); return n; }
    u32 x4 = o_huf_select(n, c);
    long h = o_huf_read_stats(w, rank, &nsym, &tl, s, c);
    if (h < 0 || (size_t)h >= c) return O_ERR;
    {   u32 start = 0;
        for (u32 r = 1; r <= tl; ++r) { u32 cur = start; start += rank[r] << (r - 1); rank[r] = cur; }
        for (u32 k = 0; k < nsym; ++k) {
            u32 len = (1u << w[k]) >> 1; o_hcell cell; cell.sym = (u8)k; cell.nb = (u8)(tl + 1 - w[k]);
            for (u32 i = 0; i < len; ++i) table[rank[w[k]] + i] = cell;
            rank[w[k]] += len;
        } }
    s += h; c -= (size_t)h;
    if (c < 10) return O_ERR;
    size_t l1 = rd16(s), l2 = rd16(s + 2), l3 = rd16(s + 4);
    if (l1 + l2 + l3 + 6 > c) return O_ERR;
    size_t l4 = c - (l1 + l2 + l3 + 6);
    long seg = (long)((n + 3) / 4), e1 = seg, e2 = 2 * seg, e3 = 3 * seg, e4 = (long)n;
    long p1 = 0, p2 = e1, p3 = e2, p4 = e3;
    o_bitin b1, b2, b3, b4;
    if (o_bitin_open(&b1, s + 6, l1) || o_bitin_open(&b2, s + 6 + l1, l2) ||
        o_bitin_open(&b3, s + 6 + l1 + l2, l3) || o_bitin_open(&b4, s + 6 + l1 + l2 + l3, l4)) return O_ERR;
    int sig = o_refill(&b1) | o_refill(&b2) | o_refill(&b3) | o_refill(&b4);
    while (sig == O_MORE && p4 < e4 - 7) {
        for (int r = 0; r < 4; ++r) {
            if (x4) { p1 += o_h2(&b1, table, tl, d + p1); p2 += o_h2(&b2, table, tl, d + p2); p3 += o_h2(&b3, table, tl, d + p3); p4 += o_h2(&b4, table, tl, d + p4); }
            else { d[p1++] = o_h1(&b1, table, tl); d[p2++] = o_h1(&b2, table, tl); d[p3++] = o_h1(&b3, table, tl); d[p4++] = o_h1(&b4, table, tl); }
        }
        sig = o_refill(&b1) | o_refill(&b2) | o_refill(&b3) | o_refill(&b4);
    }
    if (p1 > e1 || p2 > e2 || p3 > e3) return O_ERR;
    o_huf_tail(d, p1, e1, &b1, table, tl, x4); o_huf_tail(d, p2, e2, &b2, table, tl, x4);
    o_huf_tail(d, p3, e3, &b3, table, tl, x4); o_huf_tail(d, p4, e4, &b4, table, tl, x4);
    if (!(o_bitin_finished(&b1) && o_bitin_finished(&b2) && o_bitin_finished(&b3) && o_bitin_finished(&b4))) return O_ERR;
    return n;
}

/* =====================================================================================================
 * Block decoder
 * =================================================================================================== */
typedef struct { const u8 *flags, *flags_end, *lits, *lits_end, *o16, *o16_end, *o24, *o24_end; } o_dstreams;

/* diagnostic for the tests: smallest match offset seen by the last oracle_Lizard_decompress_safe call of this
 * thread.  Every Lizard parser enforces offsets >= 8 (LIZARD_*_MIN_OFFSET); below that the reference's 8-byte
 * granule copies (lizard_common.h:347-377) make its output depend on stale bytes past the write cursor. */
static __thread unsigned o_min_offset_seen = 0xFFFFFFFFu;
unsigned oracle_last_min_offset(void) { return o_min_offset_seen; }

static void o_copy8(u8* d, const u8* s) { u64 v; memcpy(&v, s, 8); memcpy(d, &v, 8); }   /* 8-byte granule, as compiled */
static void o_wild16(u8* d, const u8* s, u8* e) { do { o_copy8(d, s); o_copy8(d + 8, s + 8); d += 16; s += 16; } while (d < e); }

/* ext length byte(s): b<254 | 254,LE16 | 255,LE24 */
static long o_ext(const u8** lp)
{
    long v = **lp;
    if (v >= 254) { if (v == 254) { v = (long)rd16(*lp + 1); *lp += 2; } else { v = (long)rd24(*lp + 1); *lp += 3; } }
    (*lp)++;
    return v;
}

/* lib/lizard_decompress_lz4.h:7-163 (noDict, full) */
static int o_decode_lz4(o_dstreams* s, u8* dest, int out_size, const u8* low)
{
    const u8* const base = s->flags; const u8* const iend = s->lits_end;
    u8* op = dest; u8* const oend = op + out_size; u8* cpy;
    if (out_size == 0) return ((s->flags_end - s->flags) == 1 && *s->flags == 0) ? 0 : -1;
    while (s->flags < s->flags_end) {
        u32 tok = *s->flags++;
        long len = tok & 15;
        if (len == 15) { if (s->lits > iend - 5) goto err; len = o_ext(&s->lits) + 15; }
        cpy = op + len;
        if (cpy > oend - 16 || s->lits + len > iend - 18) goto err;
        o_wild16(op, s->lits, cpy); op = cpy; s->lits += len;
        {   u32 off = rd16(s->lits); const u8* m; s->lits += 2;
            m = op - off;
            if (m < low) goto err;
            if (off < o_min_offset_seen) o_min_offset_seen = off;
            len = tok >> 4;
            if (len == 15) { if (s->lits > iend - 5) goto err; len = o_ext(&s->lits) + 15; }
            len += 4;
            cpy = op + len;
            if (cpy > oend - 16) goto err;
            o_copy8(op, m); o_copy8(op + 8, m + 8);
            if (len > 16) o_wild16(op + 16, m + 16, cpy);
            op = cpy; }
    }
    {   long rest = s->lits_end - s->lits;
        if (rest < 0 || op + rest > oend) goto err;
        memcpy(op, s->lits, (size_t)rest); op += rest; }
    return (int)(op - dest);
err:
    return (int)(-(s->flags - base)) - 1;
}

/* lib/lizard_decompress_liz.h:14-220 (noDict, full) */
static int o_decode_lizv1(o_dstreams* s, u8* dest, int out_size, const u8* low)
{
    const u8* const base = s->flags; const u8* const iend = s->lits_end;
    u8* op = dest; u8* const oend = op + out_size; u8* cpy;
    intptr_t last = 0; long len;
    if (out_size == 0) return ((s->flags_end - s->flags) == 1 && *s->flags == 0) ? 0 : -1;
    while (s->flags < s->flags_end) {
        u32 tok = *s->flags++;
        if (tok >= 32) {
            len = tok & 7;
            if (len == 7) { if (s->lits > iend - 1) goto err; len = o_ext(&s->lits) + 7; }
            cpy = op + len;
            if (cpy > oend - 16 || s->lits > iend - 16) goto err;
            o_wild16(op, s->lits, cpy); op = cpy; s->lits += len;
            if (s->o16 > s->o16_end) goto err;
            if (!(tok >> 7)) { last = -(intptr_t)rd16(s->o16); s->o16 += 2; }
            len = (tok >> 3) & 15;
            if (len == 15) { if (s->lits > iend - 1) goto err; len = o_ext(&s->lits) + 15; }
        } else if (tok < 31) {
            if (s->o24 > s->o24_end - 3) goto err;
            len = tok + 16; last = -(intptr_t)rd24(s->o24); s->o24 += 3;
        } else {
            if (s->lits > iend - 1) goto err;
            len = o_ext(&s->lits) + 47;
            if (s->o24 > s->o24_end - 3) goto err;
            last = -(intptr_t)rd24(s->o24); s->o24 += 3;
        }
        {   const u8* m = op + last;
            if (m < low) goto err;
            if ((unsigned)(-last) < o_min_offset_seen && len > 0) o_min_offset_seen = (unsigned)(-last);
            cpy = op + len;
            if (cpy > oend - 16) goto err;
            o_copy8(op, m); o_copy8(op + 8, m + 8);
            if (len > 16) o_wild16(op + 16, m + 16, cpy);
            op = cpy; }
    }
    {   long rest = s->lits_end - s->lits;
        if (rest < 0 || op + rest > oend) goto err;
        memcpy(op, s->lits, (size_t)rest); op += rest; }
    return (int)(op - dest);
err:
    return (int)(-(s->flags - base)) - 1;
}

/* lib/lizard_decompress.c:72-112 */
static int o_read_stream(int huff, const u8** ip, const u8* iend, u8* scratch, const u8** p, const u8** e)
{
    if (!huff) {
        if (*ip > iend - 3) return 0;
        *p = *ip + 3; *e = *p + rd24(*ip); *ip = *e;
        return 1;
    }
    if (*ip > iend - 6) return 0;
    {   size_t n = rd24(*ip), c = rd24(*ip + 3);
        if (n > O_BLOCK || *ip + c > iend - 6) return 0;
        if (oracle_HUF_decompress(scratch, n, *ip + 6, c) != n) return 0;
        *ip += c + 6; *p = scratch; *e = scratch + n; }
    return 1;
}

/* lib/lizard_decompress.c:115-270 */
int oracle_Lizard_decompress_safe(const char* source, char* dest, int csize, int max_out)
{
    const u8* ip = (const u8*)source; const u8* const iend = ip + csize;
    u8* op = (u8*)dest; u8* const oend = op + max_out;
    int out_left = max_out, level, res;
    u8* scratch;
    o_min_offset_seen = 0xFFFFFFFFu;
    if (csize < 1) return 0;
    level = *ip++;
    if (level < 10 || level > 49) return -1;
    scratch = (u8*)malloc(4 * (size_t)O_BLOCK + 64);
    if (!scratch) return -1;
    while (ip < iend) {
        o_dstreams s; u32 hdr = *ip++;
        if (hdr == 128) {
            u32 n;
            if (ip > iend - 3) goto bad;
            n = rd24(ip); ip += 3;
            if (ip + n > iend || op + n > oend) goto bad;
            memcpy(op, ip, n); op += n; ip += n;
            continue;
        }
        if (hdr & 16) goto bad;
        if (ip > iend - 15) goto bad;
        {   const u8* lend = ip + 3 + rd24(ip);
            if (lend > iend - 3) goto bad;
            ip = lend; }
        if (!o_read_stream(hdr & 4, &ip, iend, scratch + 3 * O_BLOCK, &s.o16, &s.o16_end)) goto bad;
        if (!o_read_stream(hdr & 8, &ip, iend, scratch + 2 * O_BLOCK, &s.o24, &s.o24_end)) goto bad;
        if (!o_read_stream(hdr & 2, &ip, iend, scratch + 1 * O_BLOCK, &s.flags, &s.flags_end)) goto bad;
        if (!o_read_stream(hdr & 1, &ip, iend, scratch, &s.lits, &s.lits_end)) goto bad;
        if (ip > iend) goto bad;
        if ((level >= 20 && level <= 29) || level >= 40) res = o_decode_lizv1(&s, op, out_left, (const u8*)dest);
        else res = o_decode_lz4(&s, op, out_left, (const u8*)dest);
        if (res <= 0) { free(scratch); return res; }
        op += res; out_left -= res;
    }
    free(scratch);
    return (int)(op - (u8*)dest);
bad:
    free(scratch);
    return -1;
}

/* =====================================================================================================
 * Huff0 encoder
 * =================================================================================================== */
typedef struct { u64 acc; u32 n; u8* p; u8* p0; } o_bitout;
static void o_put(o_bitout* b, u64 v, u32 n) { b->acc |= (v & ((1ull << n) - 1)) << b->n; b->n += n; }
static void o_flush(o_bitout* b) { u32 k = b->n >> 3; memcpy(b->p, &b->acc, 8); b->p += k; b->n &= 7; b->acc = k >= 8 ? 0 : b->acc >> (8 * k); }
static size_t o_close(o_bitout* b) { o_put(b, 1, 1); o_flush(b); return (size_t)(b->p - b->p0) + (b->n > 0); }

/* lib/entropy/fse_compress.c:470-496 */
static u32 o_fse_min_log(size_t n, u32 maxsv) { u32 a = hb32((u32)(n - 1)) + 1, b = hb32(maxsv) + 2; return a < b ? a : b; }
static u32 o_fse_opt_log(u32 maxlog, size_t n, u32 maxsv, u32 minus)
{
    u32 src_bits = hb32((u32)(n - 1)) - minus, tl = maxlog ? maxlog : 11, mn = o_fse_min_log(n, maxsv);
    if (src_bits < tl) tl = src_bits;
    if (mn > tl) tl = mn;
    if (tl < 5) tl = 5;
    if (tl > 12) tl = 12;
    return tl;
}

/* lib/entropy/fse_compress.c:499-641 */
static int o_fse_normalize(short* norm, u32 tl, const u32* cnt, size_t total, u32 maxsv)
{
    static const u32 rtb[] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    u64 scale = 62 - tl, step = ((u64)1 << 62) / total, vstep = 1ull << (scale - 20);
    int still = 1 << tl; u32 largest = 0; short largest_p = 0; u32 low = (u32)(total >> tl);
    if (tl < 5 || tl > 12 || tl < o_fse_min_log(total, maxsv)) return -1;
    for (u32 s = 0; s <= maxsv; ++s) {
        if (cnt[s] == total) return 0;
        if (!cnt[s]) { norm[s] = 0; continue; }
        if (cnt[s] <= low) { norm[s] = -1; still--; continue; }
        {   short p = (short)((cnt[s] * step) >> scale);
            if (p < 8) p += (cnt[s] * step) - ((u64)p << scale) > vstep * rtb[p];
            if (p > largest_p) { largest_p = p; largest = s; }
            norm[s] = p; still -= p; }
    }
    if (-still < (norm[largest] >> 1)) { norm[largest] += (short)still; return (int)tl; }
    {   /* secondary method (FSE_normalizeM2) */
        u32 dist = 0, low1 = (u32)((total * 3) >> (tl + 1)), todo;
        for (u32 s = 0; s <= maxsv; ++s) {
            if (!cnt[s]) { norm[s] = 0; continue; }
            if (cnt[s] <= low) { norm[s] = -1; dist++; total -= cnt[s]; continue; }
            if (cnt[s] <= low1) { norm[s] = 1; dist++; total -= cnt[s]; continue; }
            norm[s] = -2;
        }
        todo = (1u << tl) - dist;
        if ((total / todo) > low1) {
            low1 = (u32)((total * 3) / (todo * 2));
            for (u32 s = 0; s <= maxsv; ++s) if (norm[s] == -2 && cnt[s] <= low1) { norm[s] = 1; dist++; total -= cnt[s]; }
            todo = (1u << tl) - dist;
        }
        if (dist == maxsv + 1) {
            u32 mv = 0, mc = 0;
            for (u32 s = 0; s <= maxsv; ++s) if (cnt[s] > mc) { mv = s; mc = cnt[s]; }
            norm[mv] += (short)todo;
            return (int)tl;
        }
        {   u64 vlog = 62 - tl, mid = (1ull << (vlog - 1)) - 1, rstep = ((((u64)1 << vlog) * todo) + mid) / total, t = mid;
            for (u32 s = 0; s <= maxsv; ++s) if (norm[s] == -2) {
                u64 e = t + cnt[s] * rstep; u32 wgt = (u32)(e >> vlog) - (u32)(t >> vlog);
                if (wgt < 1) return -1;
                norm[s] = (short)wgt; t = e;
            } }
    }
    return (int)tl;
}

/* lib/entropy/fse_compress.c:204-301 (destination large enough) */
static long o_fse_write_ncount(u8* out0, const short* norm, u32 maxsv, u32 tl)
{
    u8* out = out0; int size = 1 << tl, nb = (int)tl + 1, remaining = size + 1, threshold = size, bc = 4, prev0 = 0;
    u32 bs = tl - 5, sym = 0;
    while (remaining > 1) {
        if (prev0) {
            u32 start = sym;
            while (!norm[sym]) sym++;
            while (sym >= start + 24) { start += 24; bs += 0xFFFFu << bc; out[0] = (u8)bs; out[1] = (u8)(bs >> 8); out += 2; bs >>= 16; }
            while (sym >= start + 3) { start += 3; bs += 3u << bc; bc += 2; }
            bs += (sym - start) << bc; bc += 2;
            if (bc > 16) { out[0] = (u8)bs; out[1] = (u8)(bs >> 8); out += 2; bs >>= 16; bc -= 16; }
        }
        {   int c = norm[sym++], max = (2 * threshold - 1) - remaining;
            remaining -= c < 0 ? -c : c;
            c++;
            if (c >= threshold) c += max;
            bs += (u32)c << bc; bc += nb; bc -= (c < max);
            prev0 = (c == 1);
            if (remaining < 1) return -1;
            while (remaining < threshold) { nb--; threshold >>= 1; } }
        if (bc > 16) { out[0] = (u8)bs; out[1] = (u8)(bs >> 8); out += 2; bs >>= 16; bc -= 16; }
    }
    out[0] = (u8)bs; out[1] = (u8)(bs >> 8); out += (bc + 7) / 8;
    if (sym > maxsv + 1) return -1;
    return out - out0;
}

/* lib/entropy/huf_compress.c:81-121 (+ fse_compress.c:103-182, 701-770): FSE-compress the weights */
static long o_huf_compress_weights(u8* dst, const u8* w, size_t n)
{
    u32 cnt[13] = {0}, maxsv = 12, maxc = 0, tl; short norm[13]; long h; u8* op = dst;
    if (n <= 1) return 0;
    for (size_t i = 0; i < n; ++i) cnt[w[i]]++;
    while (!cnt[maxsv]) maxsv--;
    for (u32 s = 0; s <= maxsv; ++s) if (cnt[s] > maxc) maxc = cnt[s];
    if (maxc == n) return 1;
    if (maxc == 1) return 0;
    tl = o_fse_opt_log(6, n, maxsv, 2);
    if (o_fse_normalize(norm, tl, cnt, n, maxsv) < 0) return -1;
    h = o_fse_write_ncount(op, norm, maxsv, tl);
    if (h < 0) return -1;
    op += h;
    {   /* compression table */
        u32 size = 1u << tl, mask = size - 1, step = (size >> 1) + (size >> 3) + 3, high = size - 1, pos = 0, total = 0;
        u32 cumul[15]; u8 spread[64]; u16 state_tab[64]; int dstate[13]; u32 dbits[13];
        cumul[0] = 0;
        for (u32 u = 1; u <= maxsv + 1; ++u) {
            if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; spread[high--] = (u8)(u - 1); }
            else cumul[u] = cumul[u - 1] + (u32)norm[u - 1];
        }
        for (u32 s = 0; s <= maxsv; ++s) for (int i = 0; i < norm[s]; ++i) {
            spread[pos] = (u8)s; pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
        if (pos) return -1;
        for (u32 u = 0; u < size; ++u) state_tab[cumul[spread[u]]++] = (u16)(size + u);
        for (u32 s = 0; s <= maxsv; ++s) {
            if (!norm[s]) continue;
            if (norm[s] == -1 || norm[s] == 1) { dbits[s] = (tl << 16) - (1u << tl); dstate[s] = (int)total - 1; total++; }
            else { u32 mbo = tl - hb32((u32)norm[s] - 1); dbits[s] = (mbo << 16) - ((u32)norm[s] << mbo); dstate[s] = (int)total - norm[s]; total += (u32)norm[s]; }
        }
        if (n <= 2) return 0;
        {   o_bitout b = {0, 0, op, op}; const u8* ip = w + n; u32 s1, s2; size_t left = n;
#define O_INIT(S, C) { u32 c_ = (C), nbo = (dbits[c_] + (1u << 15)) >> 16, v = (nbo << 16) - dbits[c_]; S = state_tab[(int)(v >> nbo) + dstate[c_]]; }
#define O_ENC(S, C)  { u32 c_ = (C), nbo = (S + dbits[c_]) >> 16; o_put(&b, S, nbo); S = state_tab[(int)(S >> nbo) + dstate[c_]]; }
            if (left & 1) { O_INIT(s1, *--ip) O_INIT(s2, *--ip) O_ENC(s1, *--ip) o_flush(&b); }
            else { O_INIT(s2, *--ip) O_INIT(s1, *--ip) }
            left -= 2;
            if (left & 2) { O_ENC(s2, *--ip) O_ENC(s1, *--ip) o_flush(&b); }
            while (ip > w) { O_ENC(s2, *--ip) O_ENC(s1, *--ip) O_ENC(s2, *--ip) O_ENC(s1, *--ip) o_flush(&b); }
            o_put(&b, s2, tl); o_flush(&b); o_put(&b, s1, tl); o_flush(&b);
#undef O_INIT
#undef O_ENC
            op += o_close(&b); }
    }
    return op - dst;
}

typedef struct { u32 count; u16 parent; u8 byte; u8 nb; } o_node;
typedef struct { u16 val; u8 nb; } o_code;

/* lib/entropy/huf_compress.c:223-297 */
static u32 o_huf_limit_depth(o_node* nd, u32 last, u32 maxb)
{
    u32 largest = nd[last].nb, n = last, none = 0xF0F0F0F0u, rl[14]; int cost = 0;
    if (largest <= maxb) return largest;
    while (nd[n].nb > maxb) { cost += (int)((1u << (largest - maxb)) - (1u << (largest - nd[n].nb))); nd[n].nb = (u8)maxb; n--; }
    while (nd[n].nb == maxb) n--;
    cost >>= (largest - maxb);
    for (int i = 0; i < 14; ++i) rl[i] = none;
    {   u32 cur = maxb;
        for (int pos = (int)n; pos >= 0; --pos) { if (nd[pos].nb >= cur) continue; cur = nd[pos].nb; rl[maxb - cur] = (u32)pos; } }
    while (cost > 0) {
        u32 d = hb32((u32)cost) + 1;
        for (; d > 1; --d) {
            u32 hp = rl[d], lp = rl[d - 1];
            if (hp == none) continue;
            if (lp == none) break;
            if (nd[hp].count <= 2 * nd[lp].count) break;
        }
        while (d <= 12 && rl[d] == none) d++;
        cost -= 1 << (d - 1);
        if (rl[d - 1] == none) rl[d - 1] = rl[d];
        nd[rl[d]].nb++;
        if (rl[d] == 0) rl[d] = none;
        else { rl[d]--; if (nd[rl[d]].nb != maxb - d) rl[d] = none; }
    }
    while (cost < 0) {
        if (rl[1] == none) { while (nd[n].nb == maxb) n--; nd[n + 1].nb--; rl[1] = n + 1; cost++; continue; }
        nd[rl[1] + 1].nb--; rl[1]++; cost++;
    }
    return maxb;
}

/* lib/entropy/huf_compress.c:305-401 */
static int o_huf_build(o_code* code, const u32* cnt, u32 maxsv, u32 maxb)
{
    o_node store[514]; o_node* nd = store + 1;
    u32 base[32] = {0}, cur[32];
    memset(store, 0, sizeof store);
    for (u32 s = 0; s <= maxsv; ++s) base[hb32(cnt[s] + 1)]++;
    for (u32 r = 30; r > 0; --r) base[r - 1] += base[r];
    memcpy(cur, base, sizeof cur);
    for (u32 s = 0; s <= maxsv; ++s) {
        u32 c = cnt[s], r = hb32(c + 1) + 1, pos = cur[r]++;
        while (pos > base[r] && c > nd[pos - 1].count) { nd[pos] = nd[pos - 1]; pos--; }
        nd[pos].count = c; nd[pos].byte = (u8)s;
    }
    u32 last = maxsv;
    while (!nd[last].count) last--;
    int ls = (int)last, ln = 256; u32 nn = 256, root = 256 + last - 1;
    nd[nn].count = nd[ls].count + nd[ls - 1].count; nd[ls].parent = nd[ls - 1].parent = (u16)nn; nn++; ls -= 2;
    for (u32 k = nn; k <= root; ++k) nd[k].count = 1u << 30;
    store[0].count = 1u << 31;
    while (nn <= root) {
        u32 a = nd[ls].count < nd[ln].count ? (u32)ls-- : (u32)ln++;
        u32 b = nd[ls].count < nd[ln].count ? (u32)ls-- : (u32)ln++;
        nd[nn].count = nd[a].count + nd[b].count; nd[a].parent = nd[b].parent = (u16)nn; nn++;
    }
    nd[root].nb = 0;
    for (u32 k = root - 1; k >= 256; --k) nd[k].nb = (u8)(nd[nd[k].parent].nb + 1);
    for (u32 k = 0; k <= last; ++k) nd[k].nb = (u8)(nd[nd[k].parent].nb + 1);
    maxb = o_huf_limit_depth(nd, last, maxb);
    if (maxb > 12) return -1;
    {   u16 per[13] = {0}, val[13] = {0}, mn = 0;
        for (u32 k = 0; k <= last; ++k) per[nd[k].nb]++;
        for (u32 r = maxb; r > 0; --r) { val[r] = mn; mn = (u16)(mn + per[r]); mn >>= 1; }
        for (u32 k = 0; k <= maxsv; ++k) code[nd[k].byte].nb = nd[k].nb;
        for (u32 s = 0; s <= maxsv; ++s) code[s].val = val[code[s].nb]++; }
    return (int)maxb;
}

/* lib/entropy/huf_compress.c:427-470: one segment, symbols last to first */
static size_t o_huf_pack(u8* dst, const u8* p, size_t n, const o_code* code)
{
    o_bitout b = {0, 0, dst, dst};
    for (size_t i = n; i-- > 0;) { o_put(&b, code[p[i]].val, code[p[i]].nb); if ((i & 3) == 0 || b.n > 40) o_flush(&b); }
    return o_close(&b);
}

/* HUF_compress(dst, cap >= HUF_compressBound(n), src, n): 0 = not compressible, 1 = RLE, else size */
size_t oracle_HUF_compress(void* dstv, size_t cap, const void* srcv, size_t n)
{
    u8* dst = (u8*)dstv; const u8* src = (const u8*)srcv; u8* op = dst;
    u32 cnt[256] = {0}, maxsv = 255, largest = 0, hlog; o_code code[256]; u8 w[256], b2w[14]; long h;
    (void)cap;
    if (!n) return 0;
    if (n > 128 * 1024) return O_ERR;
    for (size_t i = 0; i < n; ++i) cnt[src[i]]++;
    while (!cnt[maxsv]) maxsv--;
    for (u32 s = 0; s <= maxsv; ++s) if (cnt[s] > largest) largest = cnt[s];
    if (largest == n) { dst[0] = src[0]; return 1; }
    if (largest <= (n >> 7) + 1) return 0;
    hlog = o_fse_opt_log(11, n, maxsv, 1);
    memset(code, 0, sizeof code);
    {   int mb = o_huf_build(code, cnt, maxsv, hlog); if (mb < 0) return O_ERR; hlog = (u32)mb; }
    /* header (HUF_writeCTable, huf_compress.c:132-165) */
    b2w[0] = 0;
    for (u32 k = 1; k <= hlog; ++k) b2w[k] = (u8)(hlog + 1 - k);
    for (u32 s = 0; s < maxsv; ++s) w[s] = b2w[code[s].nb];
    h = o_huf_compress_weights(op + 1, w, maxsv);
    if (h < 0) return O_ERR;
    if (h > 1 && (u32)h < maxsv / 2) { op[0] = (u8)h; h += 1; }
    else {
        if (maxsv > 128) return O_ERR;
        op[0] = (u8)(128 + (maxsv - 1)); w[maxsv] = 0;
        for (u32 s = 0; s < maxsv; s += 2) op[s / 2 + 1] = (u8)((w[s] << 4) + w[s + 1]);
        h = (long)((maxsv + 1) / 2 + 1);
    }
    if ((size_t)h + 12 >= n) return 0;
    op += h;
    if (n < 12) return 0;
    {   size_t seg = (n + 3) / 4; u8* jump = op; op += 6;
        for (int k = 0; k < 4; ++k) {
            size_t m = k < 3 ? seg : n - 3 * seg, c = o_huf_pack(op, src + (size_t)k * seg, m, code);
            if (k < 3) wr16(jump + 2 * k, (u32)c);
            op += c;
        } }
    if ((size_t)(op - dst) >= n - 1) return 0;
    return (size_t)(op - dst);
}

/* =====================================================================================================
 * Block encoder
 * =================================================================================================== */
typedef struct { int window_log, hash_log, pricefast, fastbig, lizv1, huffman; u32 mm_long; int chain_log, search_num, search_len; } o_level;

static int o_level_get(int level, o_level* L)
{   /* lib/lizard_common.h:234-284, rows on the hot path */
    int b = level >= 30 ? level - 20 : level;
    if (level >= 34 && level <= 38) b = level - 21;          /* rows 34-38 mirror 13-17 (32 is an extra noChain row) */
    memset(L, 0, sizeof *L); L->huffman = level >= 30;
    switch (b) {
    case 10: L->window_log = 16; L->hash_log = 12; return 1;
    case 11: L->window_log = 16; L->hash_log = 18; return 1;
    case 13: case 14: case 15: case 16: case 17: {   /* hashChain rows: searchNum 2,4,8,16,256; searchLength 5,5,5,4,4 */
        static const int snum[5] = { 2, 4, 8, 16, 256 }, slen[5] = { 5, 5, 5, 4, 4 };
        L->window_log = 16; L->hash_log = 18; L->chain_log = 16; L->search_num = snum[b - 13]; L->search_len = slen[b - 13];
        return 1; }
    case 20: L->window_log = 22; L->hash_log = 14; L->fastbig = 1; L->lizv1 = 1; L->mm_long = 16; return 1;   /* fastBig */
    case 21: L->window_log = 22; L->hash_log = 14; L->pricefast = 1; L->lizv1 = 1; L->mm_long = 16; return 1;
    case 22: L->window_log = 22; L->hash_log = 18; L->pricefast = 1; L->lizv1 = 1; L->mm_long = 16; return 1;
    default: return 0;
    }
}

typedef struct {
    const u8* base;                 /* base[0] is the first byte of the call */
    u32* table; o_level L;
    u32* chain; u32 next_insert;    /* hashChain: distance to the previous position of the same bucket; first index not yet inserted */
    u8 *lits, *flags, *o16, *o24;   /* stream write cursors */
    u8 *lits0, *flags0, *o160, *o240;
    u32 last_off;
    u8* huf_tmp;
} o_enc;

static u32 o_hash(const u8* p, u32 bits) { return (u32)(((rd64(p) * 889523592379ULL) << 24) >> (64 - bits)); }   /* lizard_compress.c:90 */
static u32 o_count(const u8* a, const u8* b, const u8* lim)
{   /* lib/lizard_common.h:475-490 */
    const u8* a0 = a;
    while (a < lim - 7) { u64 d = rd64(a) ^ rd64(b); if (d) return (u32)(a - a0) + ((u32)__builtin_ctzll(d) >> 3); a += 8; b += 8; }
    if (a < lim - 3 && rd32(a) == rd32(b)) { a += 4; b += 4; }
    if (a < lim - 1 && rd16(a) == rd16(b)) { a += 2; b += 2; }
    if (a < lim && *a == *b) a++;
    return (u32)(a - a0);
}
static void o_len_ext(u8** lp, size_t v)
{
    if (v >= (1 << 16)) { **lp = 255; wr24(*lp + 1, (u32)v); *lp += 4; }
    else if (v >= 254) { **lp = 254; wr16(*lp + 1, (u32)v); *lp += 3; }
    else *(*lp)++ = (u8)v;
}

/* lib/lizard_compress_lz4.h:3-71 */
static void o_emit_lz4(o_enc* e, const u8** ip, const u8** anchor, size_t ml, const u8* match)
{
    size_t lit = (size_t)(*ip - *anchor); u8* tok = e->flags++;
    if (lit >= 15) { *tok = 15; o_len_ext(&e->lits, lit - 15); } else *tok = (u8)lit;
    memcpy(e->lits, *anchor, lit); e->lits += lit;
    wr16(e->lits, (u32)(*ip - match)); e->lits += 2;
    ml -= 4;
    if (ml >= 15) { *tok += 15 << 4; o_len_ext(&e->lits, ml - 15); } else *tok += (u8)(ml << 4);
    *ip += ml + 4; *anchor = *ip;
}

/* lib/lizard_compress_liz.h:43-165 */
static void o_emit_lizv1(o_enc* e, const u8** ip, const u8** anchor, size_t ml, const u8* match)
{
    u32 off = (u32)(*ip - match); size_t lit = (size_t)(*ip - *anchor); u8* tok = e->flags++;
    if (lit > 0 || off < 65536) {
        if (lit >= 7) { *tok = 7; o_len_ext(&e->lits, lit - 7); } else *tok = (u8)lit;
        memcpy(e->lits, *anchor, lit); e->lits += lit;
        if (off >= 65536) { *tok += 1 << 7; tok = e->flags++; }
    }
    if (off >= 65536) {
        if (ml - 16 >= 31) { *tok = 31; o_len_ext(&e->lits, ml - 16 - 31); } else *tok = (u8)(ml - 16);
        wr24(e->o24, off); e->o24 += 3; e->last_off = off;
    } else {
        if (off == 0) *tok += 1 << 7;
        else { e->last_off = off; wr16(e->o16, off); e->o16 += 2; }
        if (ml >= 15) { *tok += 15 << 3; o_len_ext(&e->lits, ml - 15); } else *tok += (u8)(ml << 3);
    }
    *ip += ml; *anchor = *ip;
}

/* lib/lizard_parser_fastsmall.h:34-189 == lizard_parser_fast.h:41-196 (noDict path); with L.fastbig also
 * lib/lizard_parser_fastbig.h:35-175 (levels 20 / 40): LIZv1 codewords, and a candidate 65536 or more bytes back is only
 * taken when the match is at least MM_LONGOFF long beyond MINMATCH (:99 with the backward extension counted, :142 without) --
 * a refused candidate is a plain miss, the search goes on */
static void o_parse_fast(o_enc* e, const u8* ip, const u8* const iend)
{
    const int big = e->L.fastbig;
    const u8* const base = e->base - O_BIAS;     /* indices are relative to this virtual base */
    const u8* const low_prefix = e->base;
    const u8* const mflimit = iend - 20; const u8* const matchlimit = iend - 16; const u8* anchor = ip;
    const u32 hl = (u32)e->L.hash_log, maxd = (1u << e->L.window_log) - 1;
    const u32 low = (O_BIAS + maxd >= (u32)(ip - base)) ? O_BIAS : (u32)(ip - base) - maxd;
    u32* T = e->table; const u8* match; size_t ml;
    if ((u32)(iend - ip) < 21) goto tail;
    T[o_hash(ip, hl)] = (u32)(ip - base);
    ip++;
    for (;;) {
        {   const u8* fwd = ip; u32 step = 1, tries = 1u << 6;
            for (;;) {
                u32 h, idx;
                ip = fwd; fwd += step; step = tries++ >> 6;
                if (fwd > mflimit) goto tail;
                h = o_hash(ip, hl); idx = T[h]; T[h] = (u32)(ip - base);
                if (idx < low || idx >= (u32)(ip - base) || base + idx + maxd < ip) continue;
                match = base + idx;
                if ((u32)(ip - match) < 8 || rd32(match) != rd32(ip)) continue;
                ml = o_count(ip + 4, match + 4, matchlimit);
                {   const u8* bi = ip; const u8* bm = match; size_t bl = ml;
                    while (bi > anchor && bm > low_prefix && bi[-1] == bm[-1]) { bi--; bm--; bl++; }
                    if (big && bl < 16 && (u32)(ip - match) >= 65536) continue;
                    ip = bi; match = bm; ml = bl; }
                break;
            } }
        for (;;) {
            u32 h, idx;
            if (big) o_emit_lizv1(e, &ip, &anchor, ml + 4, match); else o_emit_lz4(e, &ip, &anchor, ml + 4, match);
            if (ip > mflimit) goto tail;
            T[o_hash(ip - 2, hl)] = (u32)(ip - 2 - base);
            h = o_hash(ip, hl); idx = T[h]; T[h] = (u32)(ip - base);
            if (idx >= low && idx < (u32)(ip - base) && base + idx + maxd >= ip) {
                match = base + idx;
                if ((u32)(ip - match) >= 8 && rd32(match) == rd32(ip)) {
                    ml = o_count(ip + 4, match + 4, matchlimit);
                    if (!big || ml >= 16 || (u32)(ip - match) < 65536) continue;
                }
            }
            break;
        }
        ip++;
    }
tail:
    memcpy(e->lits, anchor, (size_t)(iend - anchor)); e->lits += iend - anchor;
}

/* ---- hashChain parser (levels 13-17 / 34-38) ---------------------------------------------------------------
 * lib/lizard_parser_hashchain.h.  Indices are relative to the virtual base (first byte of the call = O_BIAS).
 * noDict only: dictLimit == lowLimit == O_BIAS, so the external-dictionary branches are never taken. */
static u32 o_hc_hash(const o_enc* e, const u8* p)
{   /* Lizard_hashPtr, lizard_compress.c:99-109: searchLength 5 -> hash5, 4 -> hash4 (:85-89) */
    if (e->L.search_len == 5) return o_hash(p, (u32)e->L.hash_log);
    return (u32)(rd32(p) * 2654435761U) >> (32 - e->L.hash_log);
}
/* lizard_parser_hashchain.h:13-41: enter every position below `upto` into bucket + chain */
static void o_hc_insert(o_enc* e, u32 upto)
{
    const u8* const vbase = e->base - O_BIAS;
    const u32 cmask = (1u << e->L.chain_log) - 1, maxd = (1u << e->L.window_log) - 1;
    u32 i;
    for (i = e->next_insert; i < upto; ++i) {
        u32* slot = &e->table[o_hc_hash(e, vbase + i)];
        u32 dist = i - *slot;
        e->chain[i & cmask] = dist > maxd ? maxd : dist;
        if (*slot >= i || i >= *slot + 8) *slot = i;
    }
    e->next_insert = upto;               /* assigned unconditionally (:40) */
}
/* lizard_parser_hashchain.h:45-106 */
static size_t o_hc_best(o_enc* e, const u8* ip, const u8* lim, const u8** ref)
{
    const u8* const vbase = e->base - O_BIAS;
    const u32 cmask = (1u << e->L.chain_log) - 1, maxd = (1u << e->L.window_log) - 1;
    const u32 cur = (u32)(ip - vbase);
    const u32 low = (O_BIAS + maxd >= cur) ? O_BIAS : cur - maxd;
    int tries = e->L.search_num; size_t best = 0; u32 m, d;
    o_hc_insert(e, cur);
    m = e->table[o_hc_hash(e, ip)];
    while (m < cur && m >= low && tries) {
        const u8* c = vbase + m;
        tries--;
        if ((u32)(ip - c) >= 8 && c[best] == ip[best] && rd32(c) == rd32(ip)) {
            size_t len = o_count(ip + 4, c + 4, lim) + 4;
            if (len > best) { best = len; *ref = c; }
        }
        d = e->chain[m & cmask];
        if (d > m) break;
        m -= d;
    }
    return best;
}
/* lizard_parser_hashchain.h:109-185: candidates may also grow backwards, down to `floor` */
static int o_hc_wider(o_enc* e, const u8* ip, const u8* floor, const u8* lim, int longest, const u8** ref, const u8** start)
{
    const u8* const vbase = e->base - O_BIAS;
    const u8* const first = e->base;
    const u32 cmask = (1u << e->L.chain_log) - 1, maxd = (1u << e->L.window_log) - 1;
    const u32 cur = (u32)(ip - vbase);
    const u32 low = (O_BIAS + maxd >= cur) ? O_BIAS : cur - maxd;
    const long lead = (long)(ip - floor);
    int tries = e->L.search_num; u32 m, d;
    o_hc_insert(e, cur);
    m = e->table[o_hc_hash(e, ip)];
    while (m < cur && m >= low && tries) {
        const u8* c = vbase + m;
        tries--;
        if ((u32)(ip - c) >= 8 && floor[longest] == (c - lead)[longest] && rd32(c) == rd32(ip)) {
            int len = 4 + (int)o_count(ip + 4, c + 4, lim), back = 0;
            while (ip + back > floor && c + back > first && ip[back - 1] == c[back - 1]) back--;
            len -= back;
            if (len > longest) { longest = len; *ref = c + back; *start = ip + back; }
        }
        d = e->chain[m & cmask];
        if (d > m) break;
        m -= d;
    }
    return longest;
}
/* lizard_parser_hashchain.h:188-369: up to three overlapping candidates (a, b, c) are kept in flight and trimmed
 * against each other before the first one is written */
#define O_HC_OPT 18          /* OPTIMAL_ML = (ML_MASK_LZ4 - 1) + MINMATCH, :3 */
static void o_parse_hashchain(o_enc* e, const u8* ip, const u8* const iend)
{
    const u8* const mflimit = iend - 20; const u8* const matchlimit = iend - 16; const u8* anchor = ip;
    /* MFLIMIT = WILDCOPYLENGTH + MINMATCH = 20, LASTLITERALS = 16 (lizard_common.h:72-79) */
    int la, lb, lc, l0; const u8 *ra = 0, *sb = 0, *rb = 0, *sc = 0, *rc = 0, *s0, *r0;
    ip++;
    while (ip < mflimit) {
        la = (int)o_hc_best(e, ip, matchlimit, &ra);
        if (!la) { ip++; continue; }
        s0 = ip; r0 = ra; l0 = la;
    second:
        lb = (ip + la < mflimit) ? o_hc_wider(e, ip + la - 2, ip + 1, matchlimit, la, &rb, &sb) : la;
        if (lb == la) { o_emit_lz4(e, &ip, &anchor, (size_t)la, ra); continue; }
        if (s0 < ip && sb < ip + l0) { ip = s0; ra = r0; la = l0; }
        if (sb - ip < 3) { la = lb; ip = sb; ra = rb; goto second; }
    third:
        if (sb - ip < O_HC_OPT) {
            int keep = la > O_HC_OPT ? O_HC_OPT : la, shift;
            if (ip + keep > sb + lb - 4) {
                keep = (int)(sb - ip) + lb - 4;
                if (keep < 4) { o_emit_lz4(e, &ip, &anchor, (size_t)la, ra); continue; }
            }
            shift = keep - (int)(sb - ip);
            if (shift > 0) { sb += shift; rb += shift; lb -= shift; }
        }
        lc = (sb + lb < mflimit) ? o_hc_wider(e, sb + lb - 3, sb, matchlimit, lb, &rc, &sc) : lb;
        if (lc == lb) {
            if (sb < ip + la) la = (int)(sb - ip);
            o_emit_lz4(e, &ip, &anchor, (size_t)la, ra);
            ip = sb;
            o_emit_lz4(e, &ip, &anchor, (size_t)lb, rb);
            continue;
        }
        if (sc < ip + la + 3) {
            if (sc >= ip + la) {
                if (sb < ip + la) {
                    int shift = (int)(ip + la - sb);
                    sb += shift; rb += shift; lb -= shift;
                    if (lb < 4) { sb = sc; rb = rc; lb = lc; }
                }
                o_emit_lz4(e, &ip, &anchor, (size_t)la, ra);
                ip = sc; ra = rc; la = lc;
                s0 = sb; r0 = rb; l0 = lb;
                goto second;
            }
            sb = sc; rb = rc; lb = lc;
            goto third;
        }
        if (sb < ip + la) {
            if (sb - ip < 15) {
                int shift;
                if (la > O_HC_OPT) la = O_HC_OPT;
                if (ip + la > sb + lb - 4) {
                    la = (int)(sb - ip) + lb - 4;
                    if (la < 4) {
                        o_emit_lz4(e, &ip, &anchor, (size_t)la, ra);
                        ip = sc; ra = rc; la = lc;
                        s0 = sb; r0 = rb; l0 = lb;
                        goto second;
                    }
                }
                shift = la - (int)(sb - ip);
                if (shift > 0) { sb += shift; rb += shift; lb -= shift; }
            } else la = (int)(sb - ip);
        }
        o_emit_lz4(e, &ip, &anchor, (size_t)la, ra);
        ip = sb; ra = rb; la = lb;
        sb = sc; rb = rc; lb = lc;
        goto third;
    }
    memcpy(e->lits, anchor, (size_t)(iend - anchor)); e->lits += iend - anchor;
}

/* lib/lizard_parser_pricefast.h:90-128 */
static size_t o_find_faster(o_enc* e, u32 idx, const u8* ip, const u8* lim, const u8** ref)
{
    const u8* base = e->base - O_BIAS; u32 maxd = (1u << e->L.window_log) - 1, cur = (u32)(ip - base);
    u32 low = (O_BIAS + maxd >= cur) ? O_BIAS : cur - maxd; size_t ml = 0;
    if (idx < cur && idx >= low) {
        const u8* m = base + idx;
        if ((u32)(ip - m) >= 8 && rd32(m) == rd32(ip)) {
            size_t t = o_count(ip + 4, m + 4, lim) + 4;
            if (t >= e->L.mm_long || (u32)(ip - m) < 65536) { ml = t; *ref = m; }
        }
    }
    return ml;
}
/* lib/lizard_parser_pricefast.h:3-87 */
static size_t o_find_fast(o_enc* e, u32 idx, const u8* ip, const u8* lim, const u8** ref)
{
    const u8* base = e->base - O_BIAS; intptr_t maxd = ((intptr_t)1 << e->L.window_log) - 1, cur = (u32)(ip - base);
    intptr_t low = ((intptr_t)O_BIAS + maxd >= cur) ? (intptr_t)O_BIAS : cur - maxd;
    if (e->last_off >= 8) {
        intptr_t lo_idx = (ip - e->last_off) - base;
        if (lo_idx >= low) {
            const u8* m = base + lo_idx;
            if (rd32(m) == rd32(ip)) { *ref = m; return o_count(ip + 4, m + 4, lim) + 4; }
        }
    }
    return o_find_faster(e, idx, ip, lim, ref);
}

/* lib/lizard_parser_pricefast.h:132-249 */
static void o_parse_pricefast(o_enc* e, const u8* ip, const u8* const iend)
{
    const u8* const base = e->base - O_BIAS; const u8* const low_prefix = e->base;
    const u8* anchor = ip; const u8* const mflimit = iend - 20; const u8* const matchlimit = iend - 16;
    const u32 hl = (u32)e->L.hash_log; u32* T = e->table;
    size_t ml, ml2 = 0; const u8 *ref = 0, *start2 = 0, *ref2 = 0; u32* slot;
    ip++;
    while (ip < mflimit) {
        slot = &T[o_hash(ip, hl)];
        ml = o_find_fast(e, *slot, ip, matchlimit, &ref);
        if (*slot >= (u32)(ip - base) || (u32)(ip - base) >= *slot + 8) *slot = (u32)(ip - base);
        if (!ml) { ip++; continue; }
        if ((u32)(ip - ref) == e->last_off) { ml2 = 0; ref = ip; goto encode; }
        while (ip > anchor && ref > low_prefix && ip[-1] == ref[-1]) { ip--; ref--; ml++; }
    search:
        if (ip + ml >= mflimit) goto encode;
        start2 = ip + ml - 2;
        slot = &T[o_hash(start2, hl)];
        ml2 = o_find_faster(e, *slot, start2, matchlimit, &ref2);
        if (*slot >= (u32)(start2 - base) || (u32)(start2 - base) >= *slot + 8) *slot = (u32)(start2 - base);
        if (!ml2) goto encode;
        while (start2 > ip && ref2 > low_prefix && start2[-1] == ref2[-1]) { start2--; ref2--; ml2++; }
        if (ml2 <= ml) { ml2 = 0; goto encode; }
        if (start2 <= ip) { ip = start2; ref = ref2; ml = ml2; ml2 = 0; goto encode; }
        if (start2 - ip < 3) { ip = start2; ref = ref2; ml = ml2; ml2 = 0; goto search; }
        if (start2 < ip + ml) {
            size_t corr = ml - (size_t)(start2 - ip);
            start2 += corr; ref2 += corr; ml2 -= corr;
            if (ml2 < 3) ml2 = 0;
            if (ml2 < e->L.mm_long && (u32)(start2 - ref2) >= 65536) ml2 = 0;
        }
    encode:
        o_emit_lizv1(e, &ip, &anchor, ml, ref);
        if (ml2) { ip = start2; ref = ref2; ml = ml2; ml2 = 0; goto search; }
    }
    memcpy(e->lits, anchor, (size_t)(iend - anchor)); e->lits += iend - anchor;
}

/* lib/lizard_compress.c:141-183 */
static int o_write_stream(o_enc* e, int huff, const u8* p, u32 n, u8** op, u8* oend)
{
    if (huff && n > 1024) {
        size_t c;
        if (*op + 6 > oend) return -1;
        c = oracle_HUF_compress(e->huf_tmp, 0, p, n);
        if (c != O_ERR && c > 0 && c + c / 8 + 512 < n) {
            wr24(*op, n); wr24(*op + 3, (u32)c);
            if ((size_t)(oend - (*op + 6)) < c) return -1;
            memcpy(*op + 6, e->huf_tmp, c); *op += c + 6;
            return 1;
        }
    }
    if (*op + 3 + n > oend) return -1;
    wr24(*op, n); *op += 3; memcpy(*op, p, n); *op += n;
    return 0;
}

/* lib/lizard_compress.c:186-250 */
static int o_write_block(o_enc* e, const u8* in, u32 in_size, u8** op, u8* oend)
{
    u32 nf = (u32)(e->flags - e->flags0), nl = (u32)(e->lits - e->lits0), n16 = (u32)(e->o16 - e->o160), n24 = (u32)(e->o24 - e->o240);
    u8* start = *op; int r, hf = e->L.huffman;
    if (nl < 16 || nf + nl + n16 + n24 + 16 > in_size) goto raw;
    *start = 0; *op += 1;
    if ((r = o_write_stream(e, 0, e->lits0, 0, op, oend)) < 0) return 1;
    if ((r = o_write_stream(e, 0, e->o160, n16, op, oend)) < 0) return 1;
    if ((r = o_write_stream(e, 0, e->o240, n24, op, oend)) < 0) return 1;
    if ((r = o_write_stream(e, hf, e->flags0, nf, op, oend)) < 0) return 1;
    *start += (u8)(r * 2);
    if ((r = o_write_stream(e, hf, e->lits0, nl, op, oend)) < 0) return 1;
    *start += (u8)(r * 1);
    {   u32 out = (u32)(*op - start); if (out + out / 32 + 512 > in_size) goto raw; }
    return 0;
raw:
    if ((u32)(oend - start) < in_size + 4) return 1;
    *start = 128; wr24(start + 1, in_size); memcpy(start + 4, in, in_size); *op = start + 4 + in_size;
    return 0;
}

/* lib/lizard_compress.c:472-606 with a zeroed table (== -DLIZARD_RESET_MEM) */
int oracle_Lizard_compress(const char* source, char* dest, int src_size, int max_dst, int level)
{
    o_enc e; const u8* ip = (const u8*)source; u8* op = (u8*)dest; u8* const oend = op + max_dst;
    int left = src_size, ok = 1;
    if (level > 49) level = 49;
    if (level < 10) level = 17;
    memset(&e, 0, sizeof e);
    if (!o_level_get(level, &e.L) || src_size < 0 || max_dst < 1) return 0;
    e.base = ip;
    e.table = (u32*)calloc((size_t)1 << e.L.hash_log, 4);
    if (e.L.search_num) { e.chain = (u32*)malloc((size_t)4 << e.L.chain_log); memset(e.chain, 1, (size_t)4 << e.L.chain_log); e.next_insert = O_BIAS; }
    e.lits0 = (u8*)malloc(4 * (size_t)O_BLOCK_PAD + O_BLOCK_PAD + 1024);
    e.flags0 = e.lits0 + O_BLOCK_PAD; e.o160 = e.flags0 + O_BLOCK_PAD; e.o240 = e.o160 + O_BLOCK_PAD; e.huf_tmp = e.o240 + O_BLOCK_PAD;
    *op++ = (u8)level;
    while (left > 0 && ok) {
        int part = left < (int)O_BLOCK ? left : (int)O_BLOCK;
        e.lits = e.lits0; e.flags = e.flags0; e.o16 = e.o160; e.o24 = e.o240; e.last_off = 0;
        if (e.L.search_num) o_parse_hashchain(&e, ip, ip + part);
        else if (e.L.pricefast) o_parse_pricefast(&e, ip, ip + part); else o_parse_fast(&e, ip, ip + part);
        if (o_write_block(&e, ip, (u32)part, &op, oend)) ok = 0;
        ip += part; left -= part;
    }
    free(e.table); free(e.lits0); free(e.chain);
    return ok ? (int)(op - (u8*)dest) : 0;
}

/* =====================================================================================================
 * pthread timing harness for the CPU baseline (bench.py).  A pool of `threads` workers is created ONCE per
 * call, outside every timed pass; a pass is bracketed by two pthread barriers and timed with CLOCK_MONOTONIC
 * on the calling thread (which works as worker 0).  Blocks are handed out through an atomic counter.  The
 * fastest pass is returned (the reference's own bench keeps the fastest loop too, programs/bench.c:231-286);
 * the MEAN over the passes of the last call is available from oracle_time_last_mean(), so a caller can quote
 * the same statistic as for the GPU arm.
 * =================================================================================================== */
typedef struct {
    oracle_compress_fn cfn; oracle_decompress_fn dfn;
    const char* src; size_t src_size; int block; int level; char* dst; size_t stride; int* sizes;
    const int* csizes; size_t nblocks; size_t next;
    pthread_barrier_t start, stop; pthread_mutex_t gate; int iters;
} o_job;

static __thread double o_last_mean = 0.0;
double oracle_time_last_mean(void) { return o_last_mean; }

static void o_pass(o_job* j)
{
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->nblocks) break;
        if (j->cfn) {
            size_t off = i * (size_t)j->block; int n = (int)(j->src_size - off < (size_t)j->block ? j->src_size - off : (size_t)j->block);
            j->sizes[i] = j->cfn(j->src + off, j->dst + i * j->stride, n, (int)j->stride, j->level);
        } else {
            j->sizes[i] = j->dfn(j->src + i * j->stride, j->dst + i * (size_t)j->block, j->csizes[i], j->block);
        }
    }
}
static void* o_worker(void* arg)
{
    o_job* j = (o_job*)arg;
    pthread_mutex_lock(&j->gate); pthread_mutex_unlock(&j->gate);      /* the barriers exist once the gate opens */
    for (int it = 0; it < j->iters; ++it) {
        pthread_barrier_wait(&j->start);
        o_pass(j);
        pthread_barrier_wait(&j->stop);
    }
    return 0;
}
static double o_run(o_job* j, int threads, int iters)
{
    double best = 1e30, sum = 0.0;
    if (threads < 1) threads = 1;
    if (iters < 1) iters = 1;
    j->iters = iters;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    pthread_mutex_init(&j->gate, 0);
    pthread_mutex_lock(&j->gate);
    int made = 1;
    for (int t = 1; t < threads; ++t) { if (pthread_create(&th[made], 0, o_worker, j) != 0) break; made++; }
    pthread_barrier_init(&j->start, 0, (unsigned)made);        /* sized for the workers that really exist */
    pthread_barrier_init(&j->stop, 0, (unsigned)made);
    pthread_mutex_unlock(&j->gate);
    for (int it = 0; it < iters; ++it) {
        struct timespec a, b;
        j->next = 0;                                  /* workers are parked at `start`: nobody reads it yet */
        clock_gettime(CLOCK_MONOTONIC, &a);
        pthread_barrier_wait(&j->start);
        o_pass(j);
        pthread_barrier_wait(&j->stop);
        clock_gettime(CLOCK_MONOTONIC, &b);
        {   double s = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec); if (s < best) best = s; sum += s; }
    }
    for (int t = 1; t < made; ++t) pthread_join(th[t], 0);
    pthread_barrier_destroy(&j->start); pthread_barrier_destroy(&j->stop);
    pthread_mutex_destroy(&j->gate);
    free(th);
    o_last_mean = sum / iters;
    return best;
}
double oracle_time_compress(oracle_compress_fn fn, const char* src, size_t src_size, int block, int level,
                            char* dst, size_t stride, int* sizes, int threads, int iters)
{
    o_job j; memset(&j, 0, sizeof j);
    j.cfn = fn; j.src = src; j.src_size = src_size; j.block = block; j.level = level; j.dst = dst; j.stride = stride; j.sizes = sizes;
    j.nblocks = (src_size + (size_t)block - 1) / (size_t)block;
    return o_run(&j, threads, iters);
}
double oracle_time_decompress(oracle_decompress_fn fn, const char* comp, size_t stride, const int* csizes, size_t nblocks,
                              char* dst, int block, int threads, int iters)
{
    o_job j; memset(&j, 0, sizeof j);
    int* out = (int*)malloc(sizeof(int) * (nblocks ? nblocks : 1));
    j.dfn = fn; j.src = comp; j.stride = stride; j.csizes = csizes; j.nblocks = nblocks; j.dst = dst; j.block = block; j.sizes = out;
    double s = o_run(&j, threads, iters);
    free(out);
    return s;
}
