/* lizard_b200.h -- C ABI of liblizard_b200.so: the Lizard block codec hot path on NVIDIA B200 (sm_100a).
 *
 * Two groups of entry points:
 *
 *  (1) DROP-IN symbols: same names, argument meaning, return values and error conventions as the
 *      reference library, so an application (or the reference's own frame layer / bench / CLI) can be
 *      relinked against this library unchanged.  All pointers are HOST pointers, as in the reference.
 *        reference declaration                              replaced implementation
 *        lib/lizard_compress.h:82    Lizard_versionNumber
 *        lib/lizard_compress.h:97    Lizard_compress          lib/lizard_compress.c:596-606
 *        lib/lizard_compress.h:136   Lizard_compressBound     lib/lizard_compress.c:67
 *        lib/lizard_compress.h:146   Lizard_sizeofState       lib/lizard_compress.c:311-323
 *        lib/lizard_compress.h:147   Lizard_compress_extState lib/lizard_compress.c:583-593
 *        lib/lizard_decompress.h:73  Lizard_decompress_safe   lib/lizard_decompress.c:267-270
 *      Compression output is byte-identical to the reference built with -DLIZARD_RESET_MEM
 *      (hash table empty at the start of every call), for the levels whose parsers are implemented on
 *      the GPU: 10, 11, 30, 31 (fastSmall / fast), 13-17, 34-38 (hashChain), 20, 40 (fastBig) and 21, 22, 41, 42 (priceFast).
 *      Any other level makes the compress entry points return 0 ("failed"), never a CPU fallback.
 *      Decompression accepts every level 10..49 (the block format only has two codeword flavours).
 *
 *  (2) BATCH symbols (LizardB200_*): what the reference's per-block loops
 *      (lib/lizard_frame.c:544-556 compressUpdate, :1148-1169 decodeCBlock, programs/bench.c:231-286)
 *      turn into: n independent units per call, one GPU launch.  Host-pointer and device-pointer variants.
 *
 * A unit's compressed form is exactly what one Lizard_compress call returns; a unit's decoded size is
 * exactly what Lizard_decompress_safe returns (negative values are the reference's error codes).
 *
 * There is no CPU fallback anywhere: if no usable CUDA device exists every entry point fails
 * (compress -> 0, decompress -> LIZARDB200_ERR_NO_DEVICE, batch calls -> negative status).
 */
#ifndef LIZARD_B200_H
#define LIZARD_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define LIZARD_B200_VERSION_NUMBER 10000          /* same numbering as LIZARD_VERSION_NUMBER 1.0.0 */
#define LIZARD_MIN_CLEVEL   10
#define LIZARD_MAX_CLEVEL   49
#define LIZARD_BLOCK_SIZE   (1 << 17)
#define LIZARD_MAX_INPUT_SIZE 0x7E000000
#define LIZARD_COMPRESSBOUND(isize) \
    ((unsigned)(isize) > (unsigned)LIZARD_MAX_INPUT_SIZE ? 0 : (isize) + 1 + 1 + (((isize) / LIZARD_BLOCK_SIZE) + 1) * 4)

/* status codes of the batch API (all negative); per-unit results use the reference's conventions */
#define LIZARDB200_OK               0
#define LIZARDB200_ERR_NO_DEVICE   (-1001)   /* no CUDA device / driver, or kernels not built for this GPU */
#define LIZARDB200_ERR_CUDA        (-1002)   /* a CUDA call failed; see LizardB200_lastError() */
#define LIZARDB200_ERR_ARGUMENT    (-1003)
#define LIZARDB200_ERR_LEVEL       (-1004)   /* compression level whose parser is not implemented on the GPU */
#define LIZARDB200_ERR_MEMORY      (-1005)

/* ---------------------------------------------------------------------------------------------
 * (1) drop-in symbols
 * ------------------------------------------------------------------------------------------- */
int Lizard_versionNumber(void);
int Lizard_compressBound(int inputSize);
int Lizard_sizeofState(int compressionLevel);
/* returns compressed size, or 0 when it failed / did not fit maxDstSize */
int Lizard_compress(const char* src, char* dst, int srcSize, int maxDstSize, int compressionLevel);
/* `state` is accepted for signature compatibility (must be pointer-aligned, else 0); the device keeps its own */
int Lizard_compress_extState(void* state, const char* src, char* dst, int srcSize, int maxDstSize, int compressionLevel);
/* returns decoded size (>= 0) or a negative error exactly as the reference:
 * -1 for a bad level byte / block header / stream, -(tokenIndex)-1 from the token loops.
 * Two deliberate differences, both on streams no encoder produces: a match offset below 8 is decoded with
 * byte-serial semantics (the reference's 8-byte granule copies give bytes that depend on stale memory), and a call whose
 * raw inner block is followed by a compressed one, decoded into less room than it needs, returns a negative value -- the
 * reference does not charge raw inner blocks against maxDecompressedSize (lib/lizard_decompress.c:164-180), decodes the
 * following block past dst + maxDecompressedSize and reports success.  This library never writes outside
 * [dst, dst + maxDecompressedSize): it is the safer of the two. */
int Lizard_decompress_safe(const char* src, char* dst, int compressedSize, int maxDecompressedSize);

/* ---------------------------------------------------------------------------------------------
 * (1a) the rest of the reference's export list (lib/dll/liblizard.def:3-19), so that its own callers link:
 *      lib/lizard_frame.c and the sources under programs/ reference every one of these.
 *      Stream objects are functional: lib/lizard_frame.c:379-401 creates one per compression context and hands it to
 *      Lizard_compress_extState.  The streaming / dictionary family (linked blocks, cross-call windows;
 *      lib/lizard_compress.h:178-198, lib/lizard_decompress.h:89-145) is OUT OF SCOPE and fails with the reference's
 *      failure values: 0 from Lizard_loadDict / Lizard_saveDict / Lizard_compress_continue, -1 from
 *      Lizard_decompress_safe_continue / _partial; Lizard_decompress_safe_usingDict is Lizard_decompress_safe when
 *      dictSize == 0 (lib/lizard_decompress.c:353-355) and -1 with a dictionary.  No CPU code path behind any of them.
 * ------------------------------------------------------------------------------------------- */
typedef struct Lizard_stream_s Lizard_stream_t;                 /* lib/lizard_compress.h:72 */
typedef struct Lizard_streamDecode_s Lizard_streamDecode_t;     /* lib/lizard_decompress.h:100 */
Lizard_stream_t* Lizard_createStream(int compressionLevel);
int              Lizard_freeStream(Lizard_stream_t* streamPtr);
Lizard_stream_t* Lizard_resetStream(Lizard_stream_t* streamPtr, int compressionLevel);
int Lizard_loadDict(Lizard_stream_t* streamPtr, const char* dictionary, int dictSize);
int Lizard_saveDict(Lizard_stream_t* streamPtr, char* safeBuffer, int dictSize);
int Lizard_compress_continue(Lizard_stream_t* streamPtr, const char* src, char* dst, int srcSize, int maxDstSize);
int Lizard_decompress_safe_partial(const char* source, char* dest, int compressedSize, int targetOutputSize, int maxDecompressedSize);
Lizard_streamDecode_t* Lizard_createStreamDecode(void);
int Lizard_freeStreamDecode(Lizard_streamDecode_t* streamPtr);
int Lizard_setStreamDecode(Lizard_streamDecode_t* streamPtr, const char* dictionary, int dictSize);
int Lizard_decompress_safe_continue(Lizard_streamDecode_t* streamPtr, const char* source, char* dest, int compressedSize, int maxDecompressedSize);
int Lizard_decompress_safe_usingDict(const char* source, char* dest, int compressedSize, int maxDecompressedSize,
                                     const char* dictStart, int dictSize);

/* ---------------------------------------------------------------------------------------------
 * (1b) drop-in frame layer: same names, types and error values as lib/lizard_frame.h:57-297 and
 *      lib/lizard_frame_static.h:56-67.  Frame format: doc/lizard_Frame_format.md (magic 0x184D2206).
 *      All full blocks handed to one LizardF_compressUpdate / LizardF_compressFrame / LizardF_decompress
 *      call are processed by ONE batch on the GPU (reference loops: lib/lizard_frame.c:544-556, 1010-1320).
 *      Only LizardF_blockIndependent is supported (linked blocks need the out-of-scope streaming dictionary
 *      API): a linked request / frame returns -LizardF_ERROR_blockMode_invalid.  NOTE that, as in the
 *      reference, a zeroed LizardF_preferences_t means blockLinked: set blockMode = LizardF_blockIndependent
 *      (the reference CLI does, programs/lizardio.c:109).
 * ------------------------------------------------------------------------------------------- */
typedef size_t LizardF_errorCode_t;
typedef enum { LizardF_default = 0, LizardF_max128KB = 1, LizardF_max256KB = 2, LizardF_max1MB = 3, LizardF_max4MB = 4,
               LizardF_max16MB = 5, LizardF_max64MB = 6, LizardF_max256MB = 7 } LizardF_blockSizeID_t;
typedef enum { LizardF_blockLinked = 0, LizardF_blockIndependent } LizardF_blockMode_t;
typedef enum { LizardF_noContentChecksum = 0, LizardF_contentChecksumEnabled } LizardF_contentChecksum_t;
typedef enum { LizardF_frame = 0, LizardF_skippableFrame } LizardF_frameType_t;
typedef struct {
    LizardF_blockSizeID_t     blockSizeID;
    LizardF_blockMode_t       blockMode;
    LizardF_contentChecksum_t contentChecksumFlag;
    LizardF_frameType_t       frameType;
    unsigned long long        contentSize;
    unsigned                  reserved[2];
} LizardF_frameInfo_t;
typedef struct {
    LizardF_frameInfo_t frameInfo;
    int      compressionLevel;
    unsigned autoFlush;
    unsigned reserved[4];
} LizardF_preferences_t;
typedef struct { unsigned stableSrc; unsigned reserved[3]; } LizardF_compressOptions_t;
typedef struct { unsigned stableDst; unsigned reserved[3]; } LizardF_decompressOptions_t;
typedef struct LizardF_cctx_s* LizardF_compressionContext_t;
typedef struct LizardF_dctx_s* LizardF_decompressionContext_t;
#define LIZARDF_VERSION 100

unsigned    LizardF_isError(LizardF_errorCode_t code);
const char* LizardF_getErrorName(LizardF_errorCode_t code);
size_t LizardF_compressFrameBound(size_t srcSize, const LizardF_preferences_t* preferencesPtr);
size_t LizardF_compressFrame(void* dstBuffer, size_t dstMaxSize, const void* srcBuffer, size_t srcSize,
                             const LizardF_preferences_t* preferencesPtr);
LizardF_errorCode_t LizardF_createCompressionContext(LizardF_compressionContext_t* cctxPtr, unsigned version);
LizardF_errorCode_t LizardF_freeCompressionContext(LizardF_compressionContext_t cctx);
size_t LizardF_compressBegin(LizardF_compressionContext_t cctx, void* dstBuffer, size_t dstMaxSize, const LizardF_preferences_t* prefsPtr);
size_t LizardF_compressBound(size_t srcSize, const LizardF_preferences_t* prefsPtr);
size_t LizardF_compressUpdate(LizardF_compressionContext_t cctx, void* dstBuffer, size_t dstMaxSize, const void* srcBuffer,
                              size_t srcSize, const LizardF_compressOptions_t* cOptPtr);
size_t LizardF_flush(LizardF_compressionContext_t cctx, void* dstBuffer, size_t dstMaxSize, const LizardF_compressOptions_t* cOptPtr);
size_t LizardF_compressEnd(LizardF_compressionContext_t cctx, void* dstBuffer, size_t dstMaxSize, const LizardF_compressOptions_t* cOptPtr);
LizardF_errorCode_t LizardF_createDecompressionContext(LizardF_decompressionContext_t* dctxPtr, unsigned version);
LizardF_errorCode_t LizardF_freeDecompressionContext(LizardF_decompressionContext_t dctx);
size_t LizardF_getFrameInfo(LizardF_decompressionContext_t dctx, LizardF_frameInfo_t* frameInfoPtr,
                            const void* srcBuffer, size_t* srcSizePtr);
size_t LizardF_decompress(LizardF_decompressionContext_t dctx, void* dstBuffer, size_t* dstSizePtr,
                          const void* srcBuffer, size_t* srcSizePtr, const LizardF_decompressOptions_t* dOptPtr);

/* ---------------------------------------------------------------------------------------------
 * (2) batch symbols
 * ------------------------------------------------------------------------------------------- */
/* Select the CUDA device used by this thread's subsequent calls (default 0). Returns LIZARDB200_OK or error. */
int LizardB200_setDevice(int device);
/* 1 if a usable sm_100 device is present and the context could be created, else 0 */
int LizardB200_available(void);
const char* LizardB200_lastError(void);

/* Host-pointer batch: unit i = src[i][0..srcSize[i]) -> dst[i][0..dstCapacity[i]); result[i] as Lizard_compress
 * (0 = failed / did not fit).  Inputs are staged through pinned memory, one launch for the whole batch. */
int LizardB200_compress_batch(const void* const* src, const int* srcSize,
                              void* const* dst, const int* dstCapacity, int* result,
                              int nUnits, int compressionLevel);
/* result[i] as Lizard_decompress_safe */
int LizardB200_decompress_batch(const void* const* src, const int* compressedSize,
                                void* const* dst, const int* dstCapacity, int* result, int nUnits);

/* Contiguous host buffers, units described by offset/size arrays (what a frame or a file splitter has).
 * dstStride: unit i is written at dst + i*dstStride with capacity dstCapacityEach. */
int LizardB200_compress_blocks(const void* src, size_t srcSize, int blockSize,
                               void* dst, size_t dstStride, int dstCapacityEach, int* result,
                               int compressionLevel);
/* Inverse: unit i = src + i*srcStride, compressedSize[i] bytes -> dst + i*blockSize (capacity blockSize). */
int LizardB200_decompress_blocks(const void* src, size_t srcStride, const int* compressedSize, size_t nUnits,
                                 void* dst, int blockSize, int* result);

/* Device-pointer variants: everything (payload, offset/size tables, results) already lives in device memory
 * of the current device; the call only enqueues kernels on `cudaStream` (a cudaStream_t, may be NULL) and
 * returns without synchronising.  Workspace is owned by the library, shared by all calls on a device and grown on demand
 * (growing it -- the first call, or a batch larger than any before -- is the one case in which these calls synchronise the
 * stream and allocate; do not capture that call in a CUDA graph).  Calls on DIFFERENT streams are serialised against each
 * other on the device (each launch waits for the previous launch's completion event when the stream changes), so results
 * do not depend on how the caller spreads calls over streams; calls on one stream run in stream order. */
int LizardB200_decompress_device(const void* dSrc, const uint64_t* dSrcOff, const uint32_t* dSrcLen,
                                 void* dDst, const uint64_t* dDstOff, const uint32_t* dDstCap,
                                 int* dResult, unsigned nUnits, void* cudaStream);
int LizardB200_compress_device(const void* dSrc, const uint64_t* dSrcOff, const uint32_t* dSrcLen,
                               void* dDst, const uint64_t* dDstOff, const uint32_t* dDstCap,
                               int* dResult, unsigned nUnits, int compressionLevel, void* cudaStream);
/* Concatenation step of a block writer on the device (what lib/lizard_frame.c:544-549 does by advancing dstPtr block by
 * block): segment i = dSrc + dSrcOff[i], dLen[i] bytes (entries <= 0 are skipped, e.g. failed units) is copied to
 * dDst + dDstOff[i].  With dLen = the result array of LizardB200_compress_device and dDstOff = its exclusive prefix sum this
 * packs the units of a batch back to back.  Enqueue-only, like the calls above. */
int LizardB200_gather_device(const void* dSrc, const uint64_t* dSrcOff, const int* dLen,
                             void* dDst, const uint64_t* dDstOff, unsigned nUnits, void* cudaStream);
/* diagnostics: default launch shape of the encode kernel for a level (no device needed): warps per CTA, how many of them
 * keep their hash table in shared memory, CTAs per SM, dynamic shared memory per CTA.  LIZARDB200_ERR_LEVEL for levels
 * whose parser is not implemented. */
int LizardB200_encodeShape(int compressionLevel, int* warpsPerCta, int* smemTables, int* ctasPerSM, int* smemBytes);
/* diagnostics (no device needed): pipeline chunk of a unit in a host-buffer call of nUnits units with unitsPerChunk units per
 * chunk (ramp != 0: the decoder's doubling ramp of small first chunks), computed by the host code and by the kernels'
 * arithmetic; returns the number of chunks, -1 if the two disagree. */
int LizardB200_chunkPlan(unsigned nUnits, unsigned unitsPerChunk, int ramp, unsigned unit, unsigned* chunk, unsigned* first, unsigned* count);
/* number of kernel launches issued by this library since load (bench.py reports it as gpu_launches) */
unsigned long long LizardB200_launchCount(void);
/* diagnostics: how this thread's device decodes, four bits: 1 = pooled copy sweeps, 2 = compact length-extension chain,
 * 4 = Huffman pre-pass kernels, 8 = token pre-pass kernel ahead of the token kernel, 16 = second-generation kernel (one CTA
 * per unit: a parser warp and a copier warp, literals stream staged through shared memory by TMA bulk copies; slower than the
 * first generation on the B200 so far, see profiles/r02_SUMMARY.md).  Default 7.  Results are identical for every value;
 * tools/dec_bench.py times them against each other. */
int LizardB200_setDecodeVariant(int variant);

#if defined(__cplusplus)
}
#endif
#endif /* LIZARD_B200_H */
