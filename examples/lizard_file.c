/* lizard_file.c -- whole-file batching for the frame API: what programs/lizardio.c does for `lizard -B1 FILE` / `lizard -d`,
 * with the feed sized for a batch codec instead of a streaming one.
 *
 * The reference CLI reads 128 KiB..4 MiB at a time and hands ONE block per LizardF_compressUpdate call to the codec
 * (programs/lizardio.c:397-441), and feeds the decoder 64 KiB per LizardF_decompress call (:617, :657-677): a per-call
 * pattern that gives a batch library one block of work per launch.  Here the unit of I/O is a large chunk (default
 * 256 MiB): every LizardF_compressUpdate / LizardF_decompress call carries thousands of independent 128 KiB blocks, which
 * the library turns into one launch with H2D / kernel / D2H overlapped chunk by chunk.  File format, header, block records,
 * end mark and content checksum are exactly the reference's (doc/lizard_Frame_format.md): the output of
 *     lizard_file c 10 in out.liz
 * is byte-identical to `lizard -10 -B1 -BD in out.liz` of the reference built with -DLIZARD_RESET_MEM (and to the default
 * build at levels whose parser keeps no stale state), and `lizard -d` reads it.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/lizard_file.c -Llizard_b200 -llizard_b200 -Wl,-rpath,$PWD/lizard_b200 -o lizard_file
 *   ./lizard_file c <level> <in> <out.liz> [chunk MiB] [checksum 0|1]
 *   ./lizard_file d <in.liz> <out>          [chunk MiB]
 *
 * Exit status: 0 ok, 1 library / I/O error (no B200: there is no CPU fallback). */
#define _POSIX_C_SOURCE 200112L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lizard_b200.h"

static int fail(const char* what, const char* detail)
{
    fprintf(stderr, "lizard_file: %s: %s\n", what, detail);
    return 1;
}

static int compress_file(int level, const char* in_name, const char* out_name, size_t chunk, int checksum)
{
    FILE* in = fopen(in_name, "rb");
    FILE* out = in ? fopen(out_name, "wb") : NULL;
    LizardF_compressionContext_t cctx = NULL;
    LizardF_preferences_t prefs;
    char *src = NULL, *dst = NULL;
    size_t cap, n, r;
    unsigned long long total_in = 0, total_out = 0;
    int rc = 1;

    if (!in || !out) { fail("cannot open", in ? out_name : in_name); goto done; }
    memset(&prefs, 0, sizeof prefs);
    prefs.frameInfo.blockSizeID = LizardF_max128KB;
    prefs.frameInfo.blockMode = LizardF_blockIndependent;          /* programs/lizardio.c:109 sets the same */
    prefs.frameInfo.contentChecksumFlag = checksum ? LizardF_contentChecksumEnabled : LizardF_noContentChecksum;
    prefs.compressionLevel = level;
    cap = LizardF_compressBound(chunk, &prefs) + 32;              /* room for one update + header / end mark */
    src = (char*)malloc(chunk); dst = (char*)malloc(cap);
    if (!src || !dst) { fail("out of memory", ""); goto done; }
    if (LizardF_isError(LizardF_createCompressionContext(&cctx, LIZARDF_VERSION))) { fail("createCompressionContext", ""); goto done; }
    r = LizardF_compressBegin(cctx, dst, cap, &prefs);
    if (LizardF_isError(r)) { fail("LizardF_compressBegin", LizardF_getErrorName(r)); goto done; }
    if (fwrite(dst, 1, r, out) != r) { fail("write", out_name); goto done; }
    total_out += r;
    while ((n = fread(src, 1, chunk, in)) > 0) {                   /* one call = thousands of independent blocks = one batch */
        r = LizardF_compressUpdate(cctx, dst, cap, src, n, NULL);
        if (LizardF_isError(r)) { fail("LizardF_compressUpdate", LizardF_getErrorName(r)); fprintf(stderr, "  %s\n", LizardB200_lastError()); goto done; }
        if (fwrite(dst, 1, r, out) != r) { fail("write", out_name); goto done; }
        total_in += n; total_out += r;
    }
    r = LizardF_compressEnd(cctx, dst, cap, NULL);                 /* last partial block, end mark, checksum */
    if (LizardF_isError(r)) { fail("LizardF_compressEnd", LizardF_getErrorName(r)); goto done; }
    if (fwrite(dst, 1, r, out) != r) { fail("write", out_name); goto done; }
    total_out += r;
    printf("compressed %llu -> %llu bytes, level %d, 128 KiB independent blocks, chunk %lu MiB\n", total_in, total_out, level,
           (unsigned long)(chunk >> 20));
    rc = 0;
done:
    if (cctx) LizardF_freeCompressionContext(cctx);
    free(src); free(dst);
    if (in) fclose(in);
    if (out) fclose(out);
    return rc;
}

static int decompress_file(const char* in_name, const char* out_name, size_t chunk)
{
    FILE* in = fopen(in_name, "rb");
    FILE* out = in ? fopen(out_name, "wb") : NULL;
    LizardF_decompressionContext_t dctx = NULL;
    char *src = NULL, *dst = NULL;
    const size_t out_cap = 2 * chunk + (256u << 10);              /* whatever one input chunk may expand to is drained in a loop */
    size_t have = 0, hint = 1;
    unsigned long long total_in = 0, total_out = 0;
    int rc = 1, eof = 0;

    if (!in || !out) { fail("cannot open", in ? out_name : in_name); goto done; }
    src = (char*)malloc(chunk); dst = (char*)malloc(out_cap);
    if (!src || !dst) { fail("out of memory", ""); goto done; }
    if (LizardF_isError(LizardF_createDecompressionContext(&dctx, LIZARDF_VERSION))) { fail("createDecompressionContext", ""); goto done; }
    for (;;) {
        size_t pos = 0;
        if (!eof) {
            const size_t n = fread(src + have, 1, chunk - have, in);
            if (n == 0) eof = 1;
            have += n; total_in += n;
        }
        if (have == 0) break;
        while (pos < have) {                                      /* every call takes all complete blocks it finds: one batch */
            size_t consumed = have - pos, produced = out_cap;
            hint = LizardF_decompress(dctx, dst, &produced, src + pos, &consumed, NULL);
            if (LizardF_isError(hint)) { fail("LizardF_decompress", LizardF_getErrorName(hint)); fprintf(stderr, "  %s\n", LizardB200_lastError()); goto done; }
            if (produced && fwrite(dst, 1, produced, out) != produced) { fail("write", out_name); goto done; }
            total_out += produced; pos += consumed;
            if (consumed == 0 && produced == 0) break;            /* needs more input than this chunk holds */
        }
        memmove(src, src + pos, have - pos);                      /* an incomplete block waits for the next read */
        have -= pos;
        if (eof && (have == 0 || pos == 0)) break;
    }
    if (hint != 0 || have != 0) { fail("truncated frame", in_name); goto done; }
    printf("decompressed %llu -> %llu bytes, chunk %lu MiB\n", total_in, total_out, (unsigned long)(chunk >> 20));
    rc = 0;
done:
    if (dctx) LizardF_freeDecompressionContext(dctx);
    free(src); free(dst);
    if (in) fclose(in);
    if (out) fclose(out);
    return rc;
}

int main(int argc, char** argv)
{
    if (argc >= 5 && argv[1][0] == 'c') {
        const size_t chunk = (size_t)(argc > 5 ? atoi(argv[5]) : 256) << 20;
        return compress_file(atoi(argv[2]), argv[3], argv[4], chunk ? chunk : (size_t)1 << 20, argc > 6 ? atoi(argv[6]) : 1);
    }
    if (argc >= 4 && argv[1][0] == 'd') {
        const size_t chunk = (size_t)(argc > 4 ? atoi(argv[4]) : 256) << 20;
        return decompress_file(argv[2], argv[3], chunk ? chunk : (size_t)1 << 20);
    }
    fprintf(stderr, "usage: lizard_file c <level> <in> <out.liz> [chunk MiB] [checksum 0|1]\n       lizard_file d <in.liz> <out> [chunk MiB]\n");
    return 1;
}
