// common.cuh -- shared constants and tiny helpers for the Lizard B200 kernels.
//
// Everything here is format-level knowledge restated from the reference:
//   lib/lizard_common.h:72-123 (block/stream constants), lib/lizard_compress.h:118-124 (limits),
//   lib/lizard_common.h:234-284 (level table; only the rows on the hot path are kept).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define LZ_HD __host__ __device__ __forceinline__
#define LZ_D  __device__ __forceinline__
#define LZ_HDM __host__ __device__ __forceinline__
#define LZ_HD_COLD __host__ __device__ __noinline__      /* once-per-stream serial code: keep it out of the hot loops' register budget */
#else
#define LZ_HD_COLD static
#define LZ_HDM inline
#define LZ_HD static inline
#define LZ_D  static inline
#endif

namespace lzb {

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// ---- block format (lib/lizard_common.h:72-123) ----
enum : u32 {
    kMinMatch        = 4,
    kWildCopy        = 16,          // WILDCOPYLENGTH
    kLastLiterals    = 16,          // LASTLITERALS
    kMfLimit         = 20,          // MFLIMIT = WILDCOPYLENGTH + MINMATCH
    kMinInputForLz   = 21,          // Lizard_minLength (lizard_parser_fast.h:38)
    kDictSize        = 1u << 24,    // LIZARD_DICT_SIZE: index bias of position 0 of a one-shot call
    kMax16BitOffset  = 1u << 16,
    kMmLongOff       = 16,          // MM_LONGOFF
    kLastLongOff     = 31,          // LIZARD_LAST_LONG_OFF
    kBlockSize       = 1u << 17,    // LIZARD_BLOCK_SIZE (inner block)
    kBlockSizePad    = kBlockSize + 32,
    kMaxInputSize    = 0x7E000000u, // LIZARD_MAX_INPUT_SIZE
    kMinLevel        = 10,
    kMaxLevel        = 49,
    kDefaultLevel    = 17,
    kMinOffset       = 8,           // LIZARD_*_MIN_OFFSET
    kSkipTrigger     = 6,           // Lizard_skipTrigger
};

// header byte flags (lib/lizard_common.h:110-115)
enum : u32 {
    kFlagLiterals = 1, kFlagFlags = 2, kFlagOff16 = 4, kFlagOff24 = 8, kFlagLen = 16, kFlagRaw = 128,
};

// Huff0 / FSE limits (lib/entropy/huf.h:117-132, fse.h:673-679)
enum : u32 {
    kHufTableLogMax     = 12,
    kHufTableLogDefault = 11,
    kHufSymbolMax       = 255,
    kHufBlockSizeMax    = 128 * 1024,
    kFseMinTableLog     = 5,
    kFseMaxTableLog     = 12,
    kFseAbsMaxTableLog  = 15,
    kHufHeaderFseLog    = 6,        // MAX_FSE_TABLELOG_FOR_HUFF_HEADER
};

// Parsers on the hot path.  Levels 10/30 fastSmall, 11/31 fast, 13-17/34-38 hashChain, 20/40 fastBig, 21/22/41/42 priceFast.
enum Parser : int { kParserFastSmall = 0, kParserFast = 1, kParserFastBig = 2, kParserHashChain = 3, kParserPriceFast = 5, kParserUnsupported = -1 };

struct LevelParams {
    u32 windowLog;
    u32 hashLog;
    u32 searchLength;     // hash input width in bytes (mls)
    u32 minMatchLongOff;
    int parser;
    int lizv1;            // 1: LIZv1 codewords, 0: LZ4 codewords
    int huffman;          // 1: Huffman on flags+literals (level >= 30)
    u32 searchNum;        // hashChain: candidates walked per position
    u32 chainLog;         // hashChain: log2 of the chain table (contentLog)
};

// lib/lizard_common.h:234-284 -- rows for the levels this library implements on the GPU.
LZ_HD LevelParams level_params(int level)
{
    LevelParams p = {0, 0, 0, 0, kParserUnsupported, 0, 0, 0, 0};
    int base = level >= 30 ? level - 20 : level;          // rows 30..49 mirror 10..29 with Huffman on ...
    if (level >= 34 && level <= 38) base = level - 21;    // ... except 32-38: 32 is an extra noChain row, 34-38 = 13-17
    else if (level >= 32 && level <= 33) base = -1;
    p.huffman = level >= 30;
    switch (base) {
    case 10: p.windowLog = 16; p.hashLog = 12; p.parser = kParserFastSmall; break;
    case 11: p.windowLog = 16; p.hashLog = 18; p.parser = kParserFast;      break;
    case 13: case 14: case 15: case 16: case 17:          // searchNum 2,4,8,16,256; searchLength 5,5,5,4,4
             p.windowLog = 16; p.hashLog = 18; p.chainLog = 16; p.parser = kParserHashChain;
             p.searchNum = base == 17 ? 256u : (2u << (base - 13)); p.searchLength = base >= 16 ? 4 : 5; break;
    case 20: p.windowLog = 22; p.hashLog = 14; p.searchLength = 5; p.minMatchLongOff = kMmLongOff;
             p.parser = kParserFastBig; p.lizv1 = 1; break;
    case 21: p.windowLog = 22; p.hashLog = 14; p.searchLength = 5; p.minMatchLongOff = kMmLongOff;
             p.parser = kParserPriceFast; p.lizv1 = 1; break;
    case 22: p.windowLog = 22; p.hashLog = 18; p.searchLength = 5; p.minMatchLongOff = kMmLongOff;
             p.parser = kParserPriceFast; p.lizv1 = 1; break;
    default: break;
    }
    return p;
}

// decoder side only needs the codeword flavour: levels 10-19 and 30-39 are LZ4 codewords
// (lib/lizard_decompress.c:234-241 via Lizard_defaultParameters[].decompressType)
LZ_HD int level_is_lizv1(int level) { return (level >= 20 && level <= 29) || (level >= 40 && level <= 49); }

// LIZARD_COMPRESSBOUND (lib/lizard_compress.h:124)
LZ_HD int compress_bound(int isize)
{
    return ((unsigned)isize > (unsigned)kMaxInputSize) ? 0 : isize + 1 + 1 + ((isize / (int)kBlockSize) + 1) * 4;
}

LZ_HD u32 rd_le16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
LZ_HD u32 rd_le24(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16); }
LZ_HD u32 rd_le32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
LZ_HD u64 rd_le64(const u8* p) { return (u64)rd_le32(p) | ((u64)rd_le32(p + 4) << 32); }
LZ_HD void wr_le16(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); }
LZ_HD void wr_le24(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); }
LZ_HD void wr_le32(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }

LZ_HD u32 highbit32(u32 v)   // position of the highest set bit; v != 0
{
#if defined(__CUDA_ARCH__)
    return 31u - (u32)__clz((int)v);
#else
    return 31u - (u32)__builtin_clz(v);
#endif
}

}  // namespace lzb
