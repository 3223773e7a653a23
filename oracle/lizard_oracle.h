/* lizard_oracle.h -- TEST INFRASTRUCTURE.  Plain-C, single-threaded restatement of the reference's
 * block codec hot path (see lizard_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; the product library never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against the unmodified
 * reference compiled from /root/reference (oracle/_ref, built by oracle/Makefile) and against the
 * golden fixtures in tests/golden/ that were generated from that build. */
#ifndef LIZARD_ORACLE_H
#define LIZARD_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int oracle_Lizard_compressBound(int isize);
/* clean-state semantics (hash table empty at the start of the call) == reference with -DLIZARD_RESET_MEM.
 * Levels 10,11,30,31 (fastSmall/fast) and 21,22,41,42 (priceFast); other levels return 0. */
int oracle_Lizard_compress(const char* src, char* dst, int srcSize, int maxDstSize, int level);
int oracle_Lizard_decompress_safe(const char* src, char* dst, int compressedSize, int maxDecompressedSize);
/* smallest match offset met by this thread's last oracle_Lizard_decompress_safe call (0xFFFFFFFF if none) */
unsigned oracle_last_min_offset(void);

/* Huff0 stage on its own (what Lizard_writeStream / Lizard_readStream call) */
size_t oracle_HUF_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize);   /* 0, 1 or size; (size_t)-1 on error */
size_t oracle_HUF_decompress(void* dst, size_t dstSize, const void* src, size_t srcSize);    /* dstSize or (size_t)-1 */

/* CPU timing harness used by bench.py: runs `fn` (Lizard_compress-shaped or Lizard_decompress_safe-shaped
 * function pointer taken from oracle/_ref or from this library) over n independent blocks on `threads`
 * pthreads, `iters` passes, and returns the best wall-clock seconds of one pass. */
typedef int (*oracle_compress_fn)(const char*, char*, int, int, int);
typedef int (*oracle_decompress_fn)(const char*, char*, int, int);
double oracle_time_compress(oracle_compress_fn fn, const char* src, size_t srcSize, int blockSize, int level,
                            char* dst, size_t dstStride, int* outSizes, int threads, int iters);
double oracle_time_decompress(oracle_decompress_fn fn, const char* comp, size_t compStride, const int* compSizes,
                              size_t nBlocks, char* dst, int blockSize, int threads, int iters);
/* mean seconds per pass of the calling thread's last oracle_time_* call (the worker pool is created outside the timed
 * passes, so best and mean both exclude thread creation) */
double oracle_time_last_mean(void);

#ifdef __cplusplus
}
#endif
#endif
