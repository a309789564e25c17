// k1_geometry.cu — centroids, face centres, float32 face normals and areas (kernel group K1, geometry).
//
// Replaces Grid.calculate_centroids (ninpol/_interpolator/grid.pyx:669-719) and
// Grid.calculate_normal_faces (:721-809).  Bit-exactness rules:
//   * no FMA contraction anywhere: this file is compiled with -fmad=false AND uses the _rn intrinsics;
//   * centroid = sum_j (x_j / n) in local-node order (divide, then add), n an int;
//   * face centre = (sum_j x_j) / n;
//   * normals: differences in double rounded to float, cross product / norm / divide in float
//     (grid.pyx:732-736 declares the locals `float`); grid.pyx is built as C++ so sqrt(float) is the
//     float overload, hence the quad area is ((float)(norm + sqrtf(..))) / 2.0 in double.
#include "common.cuh"

__global__ void k_centroids(ElemTables tab, const int32_t *__restrict__ inpoel, const uint8_t *__restrict__ etype,
                            const double *__restrict__ coords, i64 n_elems, int spe, int dim, double *__restrict__ cent)
{
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    int npe = tab.npoel[etype[e]];
    double dn = (double)npe;
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int j = 0; j < npe; j++) {
        const double *x = coords + (i64)inpoel[e * spe + j] * 3;
        cx = __dadd_rn(cx, __ddiv_rn(x[0], dn));
        if (dim > 1) cy = __dadd_rn(cy, __ddiv_rn(x[1], dn));
        if (dim > 2) cz = __dadd_rn(cz, __ddiv_rn(x[2], dn));
    }
    cent[e * NPB_CSTRIDE + 0] = cx;
    cent[e * NPB_CSTRIDE + 1] = cy;
    cent[e * NPB_CSTRIDE + 2] = cz;
#if NPB_CSTRIDE > 3
    cent[e * NPB_CSTRIDE + 3] = 0.0;
#endif
}

__device__ __forceinline__ float cross_norm_sq(float v1x, float v1y, float v1z, float v2x, float v2y, float v2z,
                                               float &nx, float &ny, float &nz)
{
    nx = __fsub_rn(__fmul_rn(v1y, v2z), __fmul_rn(v1z, v2y));
    ny = __fsub_rn(__fmul_rn(v1z, v2x), __fmul_rn(v1x, v2z));
    nz = __fsub_rn(__fmul_rn(v1x, v2y), __fmul_rn(v1y, v2x));
    return __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
}

__global__ void k_face_geometry(const int32_t *__restrict__ inpofa, const double *__restrict__ coords, i64 n_faces,
                                int dim, double *__restrict__ fcent, double *__restrict__ fnormal,
                                double *__restrict__ farea)
{
    i64 f = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    int4 fn = reinterpret_cast<const int4 *>(inpofa)[f];
    int ids[4] = {fn.x, fn.y, fn.z, fn.w};
    // face centre (grid.pyx:708-717)
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int npofa = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (ids[j] >= 0 && npofa == j) {
            npofa = j + 1;
            const double *x = coords + (i64)ids[j] * 3;
            sx = __dadd_rn(sx, x[0]);
            if (dim > 1) sy = __dadd_rn(sy, x[1]);
            if (dim > 2) sz = __dadd_rn(sz, x[2]);
        }
    }
    double dn = (double)npofa;
    fcent[f * 3 + 0] = __ddiv_rn(sx, dn);
    fcent[f * 3 + 1] = dim > 1 ? __ddiv_rn(sy, dn) : 0.0;
    fcent[f * 3 + 2] = dim > 2 ? __ddiv_rn(sz, dn) : 0.0;
    const double *p1 = coords + (i64)ids[0] * 3;
    const double *p2 = coords + (i64)ids[1] * 3;
    if (dim == 3) {
        const double *p3 = coords + (i64)ids[2] * 3;
        float v1x = __double2float_rn(__dsub_rn(p1[0], p2[0]));
        float v1y = __double2float_rn(__dsub_rn(p1[1], p2[1]));
        float v1z = __double2float_rn(__dsub_rn(p1[2], p2[2]));
        float v2x = __double2float_rn(__dsub_rn(p3[0], p2[0]));
        float v2y = __double2float_rn(__dsub_rn(p3[1], p2[1]));
        float v2z = __double2float_rn(__dsub_rn(p3[2], p2[2]));
        float nx, ny, nz;
        float norm = fabsf(__fsqrt_rn(cross_norm_sq(v1x, v1y, v1z, v2x, v2y, v2z, nx, ny, nz)));
        fnormal[f * 3 + 0] = (double)__fdiv_rn(nx, norm);
        fnormal[f * 3 + 1] = (double)__fdiv_rn(ny, norm);
        fnormal[f * 3 + 2] = (double)__fdiv_rn(nz, norm);
        if (ids[3] < 0) {
            farea[f] = __ddiv_rn((double)norm, 2.0);
        } else {
            const double *p4 = coords + (i64)ids[3] * 3;
            v1x = __double2float_rn(__dsub_rn(p1[0], p4[0]));
            v1y = __double2float_rn(__dsub_rn(p1[1], p4[1]));
            v1z = __double2float_rn(__dsub_rn(p1[2], p4[2]));
            v2x = __double2float_rn(__dsub_rn(p3[0], p4[0]));
            v2y = __double2float_rn(__dsub_rn(p3[1], p4[1]));
            v2z = __double2float_rn(__dsub_rn(p3[2], p4[2]));
            float both = __fadd_rn(norm, __fsqrt_rn(cross_norm_sq(v1x, v1y, v1z, v2x, v2y, v2z, nx, ny, nz)));
            farea[f] = __ddiv_rn((double)both, 2.0);
        }
    } else {
        // 2-D: faces are edges (grid.pyx:787-806)
        float v1x = __double2float_rn(__dsub_rn(p1[0], p2[0]));
        float v1y = __double2float_rn(__dsub_rn(p1[1], p2[1]));
        float nx = -v1y, ny = v1x;
        float norm = fabsf(__fsqrt_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny))));
        fnormal[f * 3 + 0] = (double)__fdiv_rn(nx, norm);
        fnormal[f * 3 + 1] = (double)__fdiv_rn(ny, norm);
        fnormal[f * 3 + 2] = 0.0;
        farea[f] = (double)norm;
    }
}

int npb_k1_geometry(npb_ctx *c)
{
    const int T = 256;
    cudaStream_t s = c->stream;
    NpbTimer tm(c, "k1_geom");   // outputs were allocated by npb_k1_build
    k_centroids<<<npb_blocks(c->n_elems, T), T, 0, s>>>(c->tab, c->inpoel, c->etype, c->coords, c->n_elems, c->spe, c->dim,
                                                       c->centroids);
    NPB_LAUNCH(c);
    if (c->n_faces > 0) {
        k_face_geometry<<<npb_blocks(c->n_faces, T), T, 0, s>>>(c->inpofa, c->coords, c->n_faces, c->dim, c->fcent,
                                                               c->fnormal, c->farea);
        NPB_LAUNCH(c);
    }
    tm.stop();
    return NPB_OK;
}
