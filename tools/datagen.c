/* datagen.c -- synthetic benchmark input generator (host utility, plain C).
 *
 * Produces byte-for-byte the stream that the reference's `datagen -g<size> -P<pct> -s<seed>`
 * writes to stdout (reference: programs/datagen.c:54-188 RDG_genOut/RDG_genBlock,
 * tests/datagencli.c:95-191 for the option -> parameter mapping).  BASELINE.json's configs are
 * all quoted on "datagen -P50" buffers, so bench.py and the tests need this generator on the
 * GPU box, where /root/reference does not exist.
 *
 * The generator is a 32-bit LCG-with-rotate driving a two-way choice per step:
 *   - with probability `match_pct` copy 4..19 (7/8 of the time) or 19..530 bytes from up to 32 KiB back,
 *   - otherwise emit 0..15 / 15..526 literal bytes drawn from a skewed table over '(' .. '}'.
 * State is kept as a 32 KiB history + 128 KiB chunk so arbitrarily large outputs stream out.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DG_TABLE_BITS 13
#define DG_TABLE_SIZE (1u << DG_TABLE_BITS)
#define DG_HISTORY    (32u << 10)
#define DG_CHUNK      (128u << 10)

static uint32_t dg_next(uint32_t* state)
{
    uint32_t r = *state * 2654435761u;
    r ^= 2246822519u;
    r = (r << 13) | (r >> 19);
    *state = r;
    return r;
}

/* literal byte table: byte value c occupies a run whose length shrinks geometrically */
static void dg_fill_table(uint8_t* table, double lit_proba)
{
    const int flat = !(lit_proba > 0.0);
    const uint8_t lo = flat ? 0 : (uint8_t)'(';
    const uint8_t hi = flat ? 255 : (uint8_t)'}';
    uint8_t c = flat ? 0 : (uint8_t)'0';
    uint32_t u = 0;
    while (u < DG_TABLE_SIZE) {
        uint32_t run = (uint32_t)((double)(DG_TABLE_SIZE - u) * lit_proba) + 1;
        uint32_t stop = u + run < DG_TABLE_SIZE ? u + run : DG_TABLE_SIZE;
        while (u < stop) table[u++] = c;
        c = (c >= hi) ? lo : (uint8_t)(c + 1);   /* wrap after the last character */
    }
}

static uint8_t dg_char(uint32_t* s, const uint8_t* table) { return table[dg_next(s) & (DG_TABLE_SIZE - 1)]; }
static uint32_t dg_rand15(uint32_t* s) { return (dg_next(s) >> 3) & 32767u; }
static uint32_t dg_length(uint32_t* s)
{
    if ((dg_next(s) >> 7) & 7u) return dg_next(s) & 15u;
    return (dg_next(s) & 511u) + 15u;
}

/* fill buf[start..size) given that buf[0..start) already holds history */
static void dg_block(uint8_t* buf, size_t size, size_t start, double match_proba,
                     const uint8_t* table, uint32_t* seed)
{
    const uint32_t match_thresh = (uint32_t)(32768 * match_proba);
    size_t pos = start;

    while (match_proba >= 1.0) {       /* degenerate "all zeros with sparse noise" mode */
        size_t z = dg_next(seed) & 3u;
        z = (size_t)1 << (16 + z * 2);
        z += dg_next(seed) & (z - 1);
        if (size < pos + z) { memset(buf + pos, 0, size - pos); return; }
        memset(buf + pos, 0, z);
        pos += z;
        buf[pos - 1] = dg_char(seed, table);
    }

    if (pos == 0) { buf[0] = dg_char(seed, table); pos = 1; }

    while (pos < size) {
        if (dg_rand15(seed) < match_thresh) {
            size_t len = (size_t)dg_length(seed) + 4;
            uint32_t dist = dg_rand15(seed) + 1;
            size_t from, stop;
            if (dist > pos) dist = (uint32_t)pos;
            from = pos - dist;
            stop = pos + len < size ? pos + len : size;
            while (pos < stop) buf[pos++] = buf[from++];
        } else {
            size_t len = dg_length(seed);
            size_t stop = pos + len < size ? pos + len : size;
            while (pos < stop) buf[pos++] = dg_char(seed, table);
        }
    }
}

/* Write `size` bytes of `datagen -g<size> -P<match_pct> -s<seed>` output into `out`.
 * lit_pct <= 0 selects the generator's default literal skew (match probability / 4.5). */
int lizb200_datagen(void* out, unsigned long long size, double match_pct, double lit_pct, unsigned seed)
{
    uint8_t* table = (uint8_t*)malloc(DG_TABLE_SIZE);
    uint8_t* work  = (uint8_t*)malloc(DG_HISTORY + DG_CHUNK);
    uint8_t* dst   = (uint8_t*)out;
    double mp = match_pct / 100.0, lp = lit_pct / 100.0;
    unsigned long long done = 0;
    uint32_t s = seed;
    if (!table || !work) { free(table); free(work); return -1; }
    if (mp > 1.0) mp = 1.0;
    if (!(lp > 0.0)) lp = mp / 4.5;
    dg_fill_table(table, lp);

    dg_block(work, DG_HISTORY, 0, mp, table, &s);
    while (done < size) {
        size_t emit = DG_CHUNK;
        dg_block(work, DG_HISTORY + DG_CHUNK, DG_HISTORY, mp, table, &s);
        if (size - done < DG_CHUNK) emit = (size_t)(size - done);
        memcpy(dst + done, work, emit);      /* the stream starts with the 32 KiB history itself */
        done += emit;
        memmove(work, work + DG_CHUNK, DG_HISTORY);
    }
    free(table); free(work);
    return 0;
}
