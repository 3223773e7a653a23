/* block_bench.c -- what `lizard -b# -B131072` (programs/bench.c:151-337) becomes on the batch entry points: the buffer is
 * cut into 128 KiB blocks, ALL blocks go to the GPU in one LizardB200_compress_blocks call (bench.c:231-246 calls
 * Lizard_compress once per block), then one LizardB200_decompress_blocks call (bench.c:266-286), fastest of N loops,
 * result verified.  Plain C99 host code over include/lizard_b200.h; buffers are ordinary host memory, so the numbers
 * include both PCIe directions (bench.py's `e2e` measures the same thing through the frame API with pinned memory).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/block_bench.c tools/datagen.c \
 *       -Llizard_b200 -llizard_b200 -Wl,-rpath,$PWD/lizard_b200 -lm -o block_bench
 *   ./block_bench [level=10] [MiB=256] [loops=3]
 *
 * Exit status: 0 ok, 1 = the library reported an error (no B200: there is no CPU fallback), 2 = round trip differs. */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "lizard_b200.h"

int lizb200_datagen(void* out, unsigned long long size, double match_pct, double lit_pct, unsigned seed);

static double now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec * 1e3 + (double)t.tv_nsec / 1e6;
}

int main(int argc, char** argv)
{
    const int level = argc > 1 ? atoi(argv[1]) : 10;
    const size_t n = (size_t)(argc > 2 ? atoi(argv[2]) : 256) << 20;
    const int loops = argc > 3 ? atoi(argv[3]) : 3;
    const int block = 1 << 17;
    const size_t nblocks = (n + (size_t)block - 1) / (size_t)block;
    const size_t stride = (size_t)Lizard_compressBound(block);       /* bench.c:179 sizes every slot with the bound */
    char *src, *comp, *back;
    int *csize, *dsize;
    double best_c = 1e30, best_d = 1e30;
    unsigned long long total = 0;
    size_t i;
    int loop;

    src = (char*)malloc(n ? n : 1); back = (char*)malloc(n ? n : 1); comp = (char*)malloc(nblocks * stride + 1);
    csize = (int*)malloc(nblocks * sizeof(int) + 1); dsize = (int*)malloc(nblocks * sizeof(int) + 1);
    if (!src || !back || !comp || !csize || !dsize) { fprintf(stderr, "out of memory\n"); return 1; }
    lizb200_datagen(src, n, 50.0, 0.0, 0);

    for (loop = 0; loop < loops + 1; loop++) {                  /* loop 0 warms the device context up */
        double t0 = now_ms(), t1, t2;
        int st = LizardB200_compress_blocks(src, n, block, comp, stride, (int)stride, csize, level);
        if (st != LIZARDB200_OK) { fprintf(stderr, "LizardB200_compress_blocks: %d (%s)\n", st, LizardB200_lastError()); return 1; }
        t1 = now_ms();
        st = LizardB200_decompress_blocks(comp, stride, csize, nblocks, back, block, dsize);
        if (st != LIZARDB200_OK) { fprintf(stderr, "LizardB200_decompress_blocks: %d (%s)\n", st, LizardB200_lastError()); return 1; }
        t2 = now_ms();
        if (loop && t1 - t0 < best_c) best_c = t1 - t0;
        if (loop && t2 - t1 < best_d) best_d = t2 - t1;
    }
    for (i = 0; i < nblocks; i++) {
        const size_t want = (i + 1 < nblocks) ? (size_t)block : n - i * (size_t)block;
        if (csize[i] <= 0 || (size_t)dsize[i] != want) { fprintf(stderr, "block %lu: compressed %d, decoded %d\n", (unsigned long)i, csize[i], dsize[i]); return 2; }
        total += (unsigned long long)csize[i];
    }
    if (memcmp(src, back, n) != 0) { fprintf(stderr, "round trip differs\n"); return 2; }
    /* bench.c:253-255 prints MB/s of uncompressed data, MB = 10^6 B */
    printf("level %2d: %lu -> %llu (%.3f), %7.1f MB/s, %7.1f MB/s   [%lu blocks of 128 KiB, host buffers, best of %d]\n",
           level, (unsigned long)n, total, total ? (double)n / (double)total : 0.0,
           (double)n / 1e3 / best_c, (double)n / 1e3 / best_d, (unsigned long)nblocks, loops);
    free(src); free(back); free(comp); free(csize); free(dsize);
    return 0;
}
