// host_shim.cpp -- TEST-ONLY host build of the lane-generic codec code.
// Compiled with g++ into lizard_b200/libhostshim.so so the CPU test-suite can pin the
// __host__ __device__ code (entropy_dec.cuh / entropy_enc.cuh / encode_core.cuh, instantiated with the
// one-lane policy HostLanes) against the reference library without a GPU.  Nothing in the product
// path links or loads this file; liblizard_b200.so has no CPU code path.
#include "entropy_dec.cuh"
#include "encode_core.cuh"
#include <stdlib.h>

extern "C" int lzb_host_huf_decompress(unsigned char* dst, unsigned n, const unsigned char* src, unsigned c)
{
    lzb::HufDecScratch* ws = (lzb::HufDecScratch*)malloc(sizeof(lzb::HufDecScratch));
    int r = lzb::huf_decompress_serial(dst, n, src, c, ws);
    free(ws);
    return r;
}

extern "C" int lzb_host_compress(const unsigned char* src, int n, unsigned char* dst, int cap, int level)
{
    if (n < 0 || cap < 0) return 0;
    if (level > 49) level = 49;
    if (level < 10) level = 17;
    lzb::LevelParams lp = lzb::level_params(level);
    if (lp.parser == lzb::kParserUnsupported) return 0;
    lzb::u32* table = (lzb::u32*)malloc(sizeof(lzb::u32) << lp.hashLog);
    lzb::EncWork* work = (lzb::EncWork*)malloc(sizeof(lzb::EncWork));
    work->huf.seg_count = (lzb::u32 (*)[256])malloc(4 * 256 * sizeof(lzb::u32));
    int r = lzb::encode_unit<lzb::HostLanes>(src, (lzb::u32)n, dst, (lzb::u32)cap, level, table, work);
    free(work->huf.seg_count); free(table); free(work);
    return r;
}
