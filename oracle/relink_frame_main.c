/* relink_frame_main.c -- TEST INFRASTRUCTURE.  INTEGRATION.md section 1 says a maintainer gets the GPU codec under the
 * reference's own callers by relinking: this program IS that relink.  oracle/Makefile compiles the UNMODIFIED
 * lib/lizard_frame.c + lib/xxhash/xxhash.c from the reference tree together with this file and links the result against
 * liblizard_b200.so, so the reference's frame layer (header, block size words, XXH32, per-block dispatch:
 * lizard_frame.c:456-483, 1148-1169) runs over OUR Lizard_createStream / Lizard_compress_extState /
 * Lizard_decompress_safe.  The frame it writes must equal the frame of the pure reference (tests/test_gpu_relink.py).
 *
 *   relinked_frame <level> <MiB> <frame-out-path> [content-checksum 0|1]
 *   exit 0 = frame written and decoded back to the input, 1 = a frame call failed, 2 = round trip differs
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lizard_frame.h"          /* the reference's header (-I$(REF)/lib) */

int lizb200_datagen(void* out, unsigned long long size, double match_pct, double lit_pct, unsigned seed);

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s level MiB frame-out [checksum]\n", argv[0]); return 1; }
    const int level = atoi(argv[1]);
    const size_t n = (size_t)atoi(argv[2]) << 20;
    LizardF_preferences_t prefs;
    memset(&prefs, 0, sizeof prefs);
    prefs.frameInfo.blockSizeID = LizardF_max128KB;
    prefs.frameInfo.blockMode = LizardF_blockIndependent;
    prefs.frameInfo.contentChecksumFlag = (argc > 4 && atoi(argv[4])) ? LizardF_contentChecksumEnabled : LizardF_noContentChecksum;
    prefs.compressionLevel = level;

    char* src = (char*)malloc(n), *back = (char*)malloc(n);
    const size_t cap = LizardF_compressFrameBound(n, &prefs);
    char* frame = (char*)malloc(cap);
    if (!src || !back || !frame) return 1;
    lizb200_datagen(src, n, 50.0, 0.0, 0);

    const size_t fsize = LizardF_compressFrame(frame, cap, src, n, &prefs);
    if (LizardF_isError(fsize)) { fprintf(stderr, "LizardF_compressFrame: %s\n", LizardF_getErrorName(fsize)); return 1; }
    FILE* f = fopen(argv[3], "wb");
    if (!f || fwrite(frame, 1, fsize, f) != fsize) return 1;
    fclose(f);

    /* blocks the frame layer had to store raw (bit 31 of the size word): with a working codec datagen has none */
    unsigned raw = 0, blocks = 0;
    for (size_t p = 7; p + 4 <= fsize; ) {
        const unsigned w = (unsigned char)frame[p] | (unsigned char)frame[p + 1] << 8 | (unsigned char)frame[p + 2] << 16 | (unsigned)(unsigned char)frame[p + 3] << 24;
        if (w == 0) break;
        blocks++; raw += w >> 31;
        p += 4 + (w & 0x7fffffffu);
    }

    LizardF_decompressionContext_t d;
    if (LizardF_isError(LizardF_createDecompressionContext(&d, LIZARDF_VERSION))) return 1;
    size_t ip = 0, op = 0, hint = 1;
    while (hint != 0 && ip < fsize) {
        size_t si = fsize - ip, so = n - op;
        hint = LizardF_decompress(d, back + op, &so, frame + ip, &si, NULL);
        if (LizardF_isError(hint)) { fprintf(stderr, "LizardF_decompress: %s\n", LizardF_getErrorName(hint)); return 1; }
        ip += si; op += so;
        if (si == 0 && so == 0) break;
    }
    LizardF_freeDecompressionContext(d);
    if (op != n || memcmp(src, back, n) != 0) { fprintf(stderr, "round trip differs (%lu of %lu bytes)\n", (unsigned long)op, (unsigned long)n); return 2; }
    printf("level %d: %lu -> %lu bytes, %u blocks, %u stored raw; round trip ok\n", level, (unsigned long)n, (unsigned long)fsize, blocks, raw);
    return 0;
}
