/* frame_roundtrip.c -- a plain-C99 host of liblizard_b200.so, written the way programs/lizardio.c drives the reference
 * (lib/lizard_frame.h): compress a buffer into one Lizard frame of independent 128 KiB blocks with LizardF_compressFrame,
 * decode it with LizardF_decompress, compare.  Only include/lizard_b200.h is needed; all buffers are ordinary host memory,
 * the library moves them to the B200 and back.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/frame_roundtrip.c tools/datagen.c \
 *       -Llizard_b200 -llizard_b200 -Wl,-rpath,$PWD/lizard_b200 -lm -o frame_roundtrip
 *   ./frame_roundtrip [MiB=64] [level=10]
 *
 * Exit status: 0 = round trip equal, 1 = the library reported an error (e.g. no B200: there is no CPU fallback),
 * 2 = bytes differ. */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "lizard_b200.h"

/* synthetic input, same generator as the reference's `datagen -P50` (tools/datagen.c) */
int lizb200_datagen(void* out, unsigned long long size, double match_pct, double lit_pct, unsigned seed);

static double now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec * 1e3 + (double)t.tv_nsec / 1e6;
}

int main(int argc, char** argv)
{
    const size_t n = (size_t)(argc > 1 ? atoi(argv[1]) : 64) << 20;
    const int level = argc > 2 ? atoi(argv[2]) : 10;
    LizardF_preferences_t prefs;
    LizardF_decompressionContext_t dctx = NULL;
    char *src, *frame, *back;
    size_t bound, csize, produced = 0, consumed = 0;
    double t0, t1, t2;
    int pass;

    memset(&prefs, 0, sizeof prefs);
    prefs.frameInfo.blockSizeID = LizardF_max128KB;
    prefs.frameInfo.blockMode = LizardF_blockIndependent;      /* a zeroed struct means linked blocks (lizard_frame.h:84-86) */
    prefs.compressionLevel = level;

    bound = LizardF_compressFrameBound(n, &prefs);
    src = (char*)malloc(n ? n : 1); frame = (char*)malloc(bound); back = (char*)malloc(n ? n : 1);
    if (!src || !frame || !back) { fprintf(stderr, "out of memory\n"); return 1; }
    lizb200_datagen(src, n, 50.0, 0.0, 0);

    for (pass = 0; pass < 2; pass++) {                        /* pass 0 warms the device context up */
        t0 = now_ms();
        csize = LizardF_compressFrame(frame, bound, src, n, &prefs);
        if (LizardF_isError(csize)) {
            fprintf(stderr, "LizardF_compressFrame: %s (%s)\n", LizardF_getErrorName(csize), LizardB200_lastError());
            return 1;
        }
        t1 = now_ms();
        if (LizardF_isError(LizardF_createDecompressionContext(&dctx, LIZARDF_VERSION))) return 1;
        produced = n; consumed = csize;
        {
            const size_t hint = LizardF_decompress(dctx, back, &produced, frame, &consumed, NULL);
            if (LizardF_isError(hint) || hint != 0) {
                fprintf(stderr, "LizardF_decompress: %s\n", LizardF_isError(hint) ? LizardF_getErrorName(hint) : "frame not finished");
                return 1;
            }
        }
        LizardF_freeDecompressionContext(dctx);
        t2 = now_ms();
    }
    if (produced != n || consumed != csize || memcmp(src, back, n) != 0) { fprintf(stderr, "round trip differs\n"); return 2; }
    printf("level %d: %lu -> %lu bytes (ratio %.3f); compress %.1f ms = %.0f MB/s, decompress %.1f ms = %.0f MB/s, host buffers\n",
           level, (unsigned long)n, (unsigned long)csize, csize ? (double)n / (double)csize : 0.0,
           t1 - t0, (double)n / 1e3 / (t1 - t0), t2 - t1, (double)n / 1e3 / (t2 - t1));
    printf("round trip ok\n");
    free(src); free(frame); free(back);
    return 0;
}
