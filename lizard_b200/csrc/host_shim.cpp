// host_shim.cpp -- TEST-ONLY host build of the serial (single-lane) device helpers.
// Compiled with g++ into lizard_b200/libhostshim.so so the CPU test-suite can pin the
// __host__ __device__ code in entropy_dec.cuh / entropy_enc.cuh against the reference library
// without a GPU.  Nothing in the product path links or loads this file.
#include "entropy_dec.cuh"
#include <stdlib.h>

extern "C" int lzb_host_huf_decompress(unsigned char* dst, unsigned n, const unsigned char* src, unsigned c)
{
    lzb::HufDecScratch* ws = (lzb::HufDecScratch*)malloc(sizeof(lzb::HufDecScratch));
    int r = lzb::huf_decompress_serial(dst, n, src, c, ws);
    free(ws);
    return r;
}
