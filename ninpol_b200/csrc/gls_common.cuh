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

// Matrix entries exactly as the reference's host arithmetic rounds them (x86-64 without FMA contraction, C order;
// SURVEY.md App. C), because an ulp of difference in an entry is amplified by cond(A) ~ 2e4 in the weights:
//   * K N (gls.pyx:320-321,397: DGEMV 'T' on the row-major 3x3 tensor): scipy's OpenBLAS kernel evaluates row q as
//     fma(K[q][2], N2, fma(K[q][0], N0, K[q][1] * N1)) - established by comparing every entry of the oracle's
//     dumped systems (the very dgemv the reference binds) with the candidate orders: 100 % of 3240 entries;
//   * T2 = N x T1 as a*b - c*d, |T2| = sqrt((x^2 + y^2) + z^2), tau * T2 (gls.pyx:306-318,358-372): no FMA.
// pow() itself is CUDA's (<= 2 ulp from glibc's correctly rounded result).
__device__ __forceinline__ double gls_kn(const double *__restrict__ Kq, double N0, double N1, double N2)
{
    return __fma_rn(Kq[2], N2, __fma_rn(Kq[0], N0, __dmul_rn(Kq[1], N1)));
}
__device__ __forceinline__ double gls_cross(double a, double b, double c, double d)   // a*b - c*d
{
    return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d));
}
__device__ __forceinline__ double gls_norm3(double x, double y, double z)
{
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// dense fallback (k2_gls_dense.cu): Householder QR of the whole system in a global-memory workspace
int npb_gls_dense(npb_ctx *c, const GlsArgs &a, const int32_t *list, int count);
