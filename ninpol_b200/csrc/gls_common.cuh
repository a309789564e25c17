// gls_common.cuh — argument block and warp helpers shared by the GLS kernels.
#pragma once
#include "common.cuh"

struct GlsArgs {
    const int32_t *esup_ptr, *esup, *fsup_ptr, *fsup;
    const int2 *esuf2;
    const uint8_t *bpoint, *nflag;
    const double *coords, *cent, *fcent, *fnormal, *perm, *diff_mag;
    double *wbuf;
    int32_t *rowcnt;
    double *neumann;
    i64 wbase;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dense fallback (k2_gls_dense.cu): Householder QR of the whole system in a global-memory workspace
int npb_gls_dense(npb_ctx *c, const GlsArgs &a, const int32_t *list, int count);
